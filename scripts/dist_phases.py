"""Per-phase timings of the inner iteration on a row-partitioned run (torchrun)."""
import os
import sys
sys.path[:0] = ['.', 'ip-nonlinear-solver_b200']
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from ipsolver_b200 import synthetic, device, distributed as D  # noqa: E402

rank = int(os.environ.get("RANK", "0")); size = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
gloo = dist.new_group(backend="gloo")
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
H, A, c, b = synthetic.make_qp(n, 'window', 1e4, 0)
S = D.DistSolver(H, A, rank, size, group=gloo, device=local)
names = {0: "update", 3: "w = M r", 4: "grid barrier", 6: "grid barrier + all-reduce", 7: "grid barrier + halo push"}
only = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else list(names)
for ph in only:
    name = names[ph]
    dist.barrier()
    torch.cuda.synchronize()
    print("rank", rank, "phase", ph, "warm-up", flush=True)
    device.bench_phase(S.proj, ph, 5)
    print("rank", rank, "phase", ph, "warm-up done", flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    ms = min(device.bench_phase(S.proj, ph, 40) for _ in range(3)) / 40
    if rank == 0:
        print("%-28s %9.2f us" % (name, 1e3 * ms), flush=True)
dc = S.vector(S.local(c, 0))
dist.barrier()
device.bench_inner_iterations(S.proj, dc, 20)
ms = min(device.bench_inner_iterations(S.proj, dc, 60) for _ in range(3)) / 61
if rank == 0:
    print("%-28s %9.2f us" % ("inner iteration", 1e3 * ms), flush=True)
dist.barrier()
dist.destroy_process_group()
