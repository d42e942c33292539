#!/usr/bin/env python
"""Headline benchmark: projected-CG iterations/s on the synthetic sparse
equality-constrained QP of BASELINE.json (n = 1e7, m = 5e6, banded SPD H,
windowed 3-nnz/row A), one B200 (or N ranks under torchrun).

A *step* is one ``projected_cg`` solve with exactly ``--iters`` CG iterations
forced (``tol=1e-30, trust_radius=inf``) on a matrix
already resident in HBM -- i.e. one pass of the hot path: 2 + iters
null-space projections and 1 row-space solve (each one (A A')^-1 solve: a
sweep of the banded direct solve, or an inner CG), 2 + iters Hessian products
and the fused vector updates, in ONE persistent kernel launch.

  value  iterations/s with c, b, x resident in HBM (CUDA events on the
         library's stream, around the K steps)
  e2e    the same through the reference-facing Python API ``projected_cg(H, c,
         Z, Y, b, ...)`` with pinned HOST vectors: H2D of c and b, D2H of x and
         the info struct inside the timed region (wall clock, synchronous API);
         ``e2e.pageable_inputs``: the same with ordinary numpy arrays
  roofline   algorithmic bytes of the launch (applied to the
         launch's own counters) / the launch's device time, against the measured
         HBM copy bandwidth; ``roofline.timeline``: phase durations inside the
         kernel (%globaltimer at every grid barrier)
  cpu_baseline   the oracle (scipy SuperLU + numpy restatement of the
         reference path) on a bounded sample, host cores; ``reference_config_run``
         is this arm on that same sample config
  end_to_end_solve   BASELINE configs 2-4 through ``minimize_constrained`` next
         to the reference's own records of the same problems

``--impl reference`` times only the CPU path (oracle port; the reference is
pure Python and cannot travel to the GPU box) on the largest size it can
factorise, and says so (``same_config: false``).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "ip-nonlinear-solver_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "projected_cg_iterations_per_s"
UNIT = "it/s"


def algorithmic_bytes(n, m, nnzH, nnzA, nnzM, niter, inner_iters, n_solves, n_refin, n_hp,
                     n_sweeps=0, has_b=1, fused_hp=0):
    """Compulsory HBM bytes of one launch of the solve kernel, for the
    algorithm AS RUN.  8-byte values, 4-byte indices,
    every matrix / vector counted once per logical read or write:

      B_spmv(rows, cols, nnz) = 12 nnz + 4 (rows+1) + 8 cols + 8 rows
      outer iteration = B_H + 72 n  (+ its Z apply); 64 n when p = -g + beta p is
                formed inside H p (fused_hp: g read + p written; p read is B_H's x)
      Z apply = 2 B_A + B_At + 8 n  (A~ x, z = x - A~' v with the re-read of x,
                the orthogonality product A z)  + its (A A')^-1 solve
      (A A')^-1 by the banded direct solve (n_sweeps > 0):
          B_sweep = 160 m per sweep: the factor rows once forwards and once
          backwards (2 x 64 m), right-hand side (8 m), forward result written
          and read back (16 m), solution (8 m).  Rows read again in the overlap
          of neighbouring chunks are NOT counted (they are waste, visible as
          ncu traffic above this figure).  Z applies are verified by the
          orthogonality product they contain; the Y / LS solves by one product
          with M (B_M + 8 m).
      (A A')^-1 by CG (n_sweeps == 0), per inner iteration:
          explicit M = A A' (nnzM > 0):  B_in = B_spmv(m, m, nnzM) + 72 m
          A then A' (nnzM < 0):          B_in = B_A + B_At + 8 m + 72 m

    ``survey`` is the same launch priced with SURVEY.md section 8d's model
    of the reference formulation (an inner CG of k_in = 35 iterations per
    solve, B_in = B_A + B_At + 56 m), i.e. the traffic the path would need
    without the direct solve."""
    B_H = 12 * nnzH + 4 * (n + 1) + 16 * n
    B_A = 12 * nnzA + 4 * (m + 1) + 8 * n + 8 * m
    B_At = 12 * nnzA + 4 * (n + 1) + 8 * m + 8 * n
    B_M = 12 * nnzM + 4 * (m + 1) + 16 * m if nnzM > 0 else None
    B_in_survey = B_A + B_At + 56 * m
    n_y = 1 if has_b else 0                      # x = Y(-b): the only non-Z solve of a launch
    n_z = n_solves - n_y
    z_products = 2 * B_A + B_At + 8 * n
    parts = dict(B_H=B_H, B_A=B_A, B_At=B_At, B_M=B_M, B_in_survey=B_in_survey)
    if n_sweeps > 0:
        B_sweep = 160 * m
        solves = n_sweeps * B_sweep + n_y * (B_M + 8 * m)
        refin = n_refin * (B_At + 8 * n + B_A)
        parts.update(B_sweep=B_sweep, B_in=B_sweep, B_Z=z_products + B_sweep)
    else:
        B_in = B_M + 72 * m if nnzM > 0 else B_A + B_At + 8 * m + 72 * m
        solves = inner_iters * B_in
        refin = n_refin * z_products
        parts.update(B_in=B_in)
    iter_vec = (64 if fused_hp else 72) * n
    rest = n_z * z_products + n_y * (24 * m + B_At) + n_hp * B_H + niter * iter_vec
    total = solves + refin + rest
    survey = (n_solves + n_refin) * 35 * B_in_survey + (n_z + n_refin) * (z_products + 8 * n) \
        + n_hp * B_H + niter * 72 * n
    return total, survey, parts


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def count(self):
        """Samples written so far (nvidia-smi needs ~0.1-0.3 s to start)."""
        try:
            with open(self.path) as fh:
                return sum(1 for _ in fh)
        except Exception:
            return 0

    def keep_loaded_until_sampling(self, step, min_samples=2, max_steps=400):
        """Extra untimed warm-up steps until the sampler delivers: the timed
        region of a fast step can be shorter than nvidia-smi's start-up, and a
        line without clocks would be unusable.  Returns the steps spent."""
        spent = 0
        if self.proc is None:
            return spent
        while self.count() < min_samples and spent < max_steps:
            step()
            spent += 1
        return spent

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [],
               "window": "last warm-up steps + timed region (GPU under the same load throughout)"}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(smax))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        t = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        t = 1
    return t


def reference_l1():
    """The reference's own ``projections`` / ``projected_cg`` from the offline
    install under baseline/_ref (scripts/install_reference.sh; it travels to the
    GPU box with the snapshot), else the oracle port.  Returns
    ``(kind, build_projectors, projected_cg)``."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isfile(os.path.join(ref, "ipsolver", "_large_scale_constrained", "qp_subproblem.py")):
        try:
            if ref not in sys.path:
                sys.path.insert(0, ref)
            import importlib
            rp = importlib.import_module("ipsolver._large_scale_constrained.projections")
            rq = importlib.import_module("ipsolver._large_scale_constrained.qp_subproblem")
            return ("reference", lambda A: rp.projections(A, "AugmentedSystem"), rq.projected_cg)
        except Exception as exc:          # broken install: say so, use the port
            print("bench: baseline/_ref unusable (%r), timing the oracle port" % (exc,),
                  file=sys.stderr)
    import oracle
    return ("port", oracle.build_projectors, oracle.eqp_projected_cg)


def cpu_reference_run(n, iters, steps, warmup, pattern, kappa):
    """The reference's CPU path -- the UNMODIFIED reference from baseline/_ref
    when it is installed (``kind`` "reference": its ``projections(A,
    'AugmentedSystem')`` = scipy SuperLU, its ``projected_cg``), else the
    oracle's restatement (``kind`` "port"): factorisation outside the timed
    region, ``steps`` forced-iteration solves inside."""
    from ipsolver_b200 import synthetic
    kind, build, pcg = reference_l1()
    H, A, c, b = synthetic.make_qp(n, pattern, kappa, seed=0)
    t0 = time.perf_counter()
    Z, LS, Y = build(A)
    t_factor = time.perf_counter() - t0
    for _ in range(warmup):
        pcg(H, c, Z, Y, b, tol=1e-30, max_iter=iters)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        _, info = pcg(H, c, Z, Y, b, tol=1e-30, max_iter=iters)
        done += info["niter"]
    dt = time.perf_counter() - t0
    return done / dt, dt, t_factor, kind


def solve_problem(n, pattern="window", kappa=10.0):
    """BASELINE config 2: convex equality-constrained QP through
    ``minimize_constrained(..., method='equality_constrained_sqp')``."""
    from ipsolver_b200 import synthetic
    H, A, q, b = synthetic.make_qp(n, pattern, kappa, seed=0)
    args = (lambda x: 0.5 * x.dot(H.dot(x)) + q.dot(x), np.zeros(n),
            lambda x: H.dot(x) + q, lambda x: H)
    return args, A, b


def end_to_end_solve(n, hot_path=None):
    import ipsolver_b200 as ipb
    from threadpoolctl import threadpool_limits
    args, A, b = solve_problem(n)
    # host BLAS threads as in the run that produced the reference record (8): the tail of this
    # xtol-terminated run hangs on round-off in the host's dot / norm, for the reference as well
    # (profiles/r5_host_threads.txt)
    with threadpool_limits(limits=8):
        t0 = time.perf_counter()
        res = ipb.minimize_constrained(*args, ipb.LinearConstraint(A, ("equals", b)),
                                       method='equality_constrained_sqp', hot_path=hot_path)
        dt = time.perf_counter() - t0
    return {"n": n, "m": n // 2, "seconds": dt, "niter": int(res.niter),
            "cg_niter": int(res.cg_niter), "nfev": int(res.nfev), "status": int(res.status),
            "optimality": float(res.optimality), "fun": float(res.fun),
            "constr_violation": float(res.constr_violation), "x_checksum": x_checksum(res.x)}


def reference_record(name):
    """What the unmodified reference gave for a named end-to-end problem in the
    build container (tests/golden/e2e_large.npz, made by make_golden_large.py)."""
    try:
        g = np.load(os.path.join(ROOT, "tests", "golden", "e2e_large.npz"))
        s = g[name + "/scalars"]
        return {"niter": int(s[0]), "cg_niter": int(s[1]), "status": int(s[2]), "nfev": int(s[3]),
                "optimality": float(s[4]), "constr_violation": float(s[5]), "fun": float(s[6]),
                "seconds": float(s[8]),
                "where": "unmodified reference, build container CPU (scipy SuperLU), golden record"}
    except Exception as exc:
        return {"unavailable": str(exc)}


def named_solve(name):
    """BASELINE configs 3 / 4 through the public API (tests/e2e_problems.py)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import e2e_problems
    import ipsolver_b200 as ipb
    api = dict(minimize_constrained=ipb.minimize_constrained,
               LinearConstraint=ipb.LinearConstraint, BoxConstraint=ipb.BoxConstraint,
               NonlinearConstraint=ipb.NonlinearConstraint)
    t0 = time.perf_counter()
    res = e2e_problems.LARGE_PROBLEMS[name](api)()
    dt = time.perf_counter() - t0
    return {"problem": name, "seconds": dt, "niter": int(res.niter), "cg_niter": int(res.cg_niter),
            "nfev": int(res.nfev), "status": int(res.status), "optimality": float(res.optimality),
            "fun": float(res.fun), "constr_violation": float(res.constr_violation),
            "reference": reference_record(name)}


def run_reference(args):
    """--impl reference: the CPU path on the host cores (rank 0 only).  The
    reference factorises with SuperLU, which fails at the headline size (n =
    1e7: ``factorized`` raises, the dense fallback then needs 364 TiB --
    profiles/r3_reference_n1e7.txt), so this arm runs the same workload at
    ``--cpu-n`` (default 4e6, the largest size SuperLU factorises here) and SAYS SO: ``value`` and ``ms_per_step`` are
    what was measured at that size, ``config.n`` is that size and
    ``same_config`` is false.  The GPU arm prints its own measurement at the
    same size (``reference_config_run``) so the two can be compared on one
    config; the figure scaled to n = 1e7 is a labelled extra."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_s = args.cpu_n
    its, dt, t_factor, kind = cpu_reference_run(n_s, args.iters, args.steps, args.warmup,
                                                args.pattern, args.kappa)
    sample = (("the UNMODIFIED reference from baseline/_ref (projections(A, 'AugmentedSystem') = "
               "scipy SuperLU, its own projected_cg), " if kind == "reference" else
               "oracle port of the reference path (scipy SuperLU AugmentedSystem + numpy), ") +
              "n=%d m=%d window A, %d steps x %d forced iterations in %.1f s; factorisation "
              "%.1f s outside the timed region; SuperLU factor / solve, csr_matvec and the "
              "numpy loops are single-threaded (cores = 1), only BLAS dot / norm may use "
              "blas_threads") % (n_s, n_s // 2, args.steps, args.iters, dt, t_factor)
    ran = argparse.Namespace(**vars(args))
    ran.n = n_s
    line = {
        "impl": "reference", "metric": METRIC, "value": its, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(ran),
        "same_config": bool(n_s == args.n),
        "requested_n": args.n,
        "extrapolated": {"n": args.n, "value": its * n_s / args.n, "unit": UNIT,
                         "how": "per-iteration cost of the CPU path is linear in n for the "
                                "windowed pattern; NOT a measurement"},
        "cpu_baseline": {"value": its, "unit": UNIT, "cores": 1,
                         "blas_threads": host_threads(),
                         "kind": kind, "sample": sample, "host_cpus": os.cpu_count()},
        "e2e": {"value": its, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": "synthetic sparse convex equality-constrained QP "
                        "(BASELINE.json metric config): n=%d, m=%d, H banded SPD "
                        "5 nnz/row cond=%g, A %s 3 nnz/row; projected_cg with "
                        "%d forced iterations per step (tol=1e-30, radius=inf), "
                        "NormalEquation projector: banded direct solve of A A' "
                        "(CG fallback), residual tolerance %g"
                        % (args.n, args.n // 2, args.kappa, args.pattern,
                           args.iters, args.inner_rtol),
            "n": args.n, "m": args.n // 2, "iters_per_step": args.iters,
            "pattern": args.pattern, "inner_rtol": args.inner_rtol,
            "l2": "working set per outer iteration (~3.4 GB at n=1e7) exceeds "
                  "the 126 MB L2; no explicit flush",
            "parallelism": "1 GPU" if args.gpus == 1 else "row-partitioned x%d" % args.gpus}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def timeline_names(info):
    """Phases between the grid barriers of one outer iteration (banded path)."""
    if int(info.n_sweeps) <= 0:
        return None
    names = ["x+=alpha p; r+=alpha Hp", "rhs = A~ r", "banded sweep", "g = r - A~' v",
             "A g (orthogonality)"]
    if int(info.reserved) == 1:
        return names + ["H p with p = -g + beta p formed on the fly"]
    return names + ["p = -g + beta p", "H p"]


def x_checksum(x):
    """Order-independent fingerprint of the solution, printed by every N so
    that the driver's own records show N-GPU = 1-GPU: ||x||_2, sum(x) and a
    weighted sum with a fixed pseudo-random weight vector."""
    x = np.asarray(x, dtype=np.float64)
    w = np.cos(np.arange(x.shape[0], dtype=np.float64) * 0.618033988749895)
    return {"norm2": float(np.linalg.norm(x)), "sum": float(x.sum()),
            "weighted": float(np.dot(w, x)), "n": int(x.shape[0])}


def gpu_run_at(n, args):
    """The GPU arm on the config the reference arm can run (``--cpu-n``): the same
    measurement as the headline (device-resident it/s and e2e with host vectors),
    so that GPU and CPU can be compared on ONE config."""
    from ipsolver_b200 import synthetic, device
    from ipsolver_b200.projections import projections
    from ipsolver_b200.qp_subproblem import projected_cg, HessianOperator
    m = n // 2
    H, A, c, b = synthetic.make_qp(n, args.pattern, args.kappa, seed=0)
    ctx = device.Context.default()
    Z, LS, Y = projections(A)
    Hop = HessianOperator(n, [H])
    Hd = Hop.device
    proj = Z.projector
    dc, db = device.DeviceVector(ctx, c), device.DeviceVector(ctx, b)
    dx = device.DeviceVector(ctx, length=n)
    for _ in range(3):
        device.projected_cg_dev(proj, Hd, dc, db, dx, tol=1e-30, max_iter=args.iters)
    steps = max(args.steps, 5)
    ctx.timer_start()
    for _ in range(steps):
        info = device.projected_cg_dev(proj, Hd, dc, db, dx, tol=1e-30, max_iter=args.iters)
    ms = ctx.timer_stop()
    for _ in range(2):
        projected_cg(Hop, c, Z, Y, b, tol=1e-30, max_iter=args.iters)
    t0 = time.perf_counter()
    for _ in range(steps):
        projected_cg(Hop, c, Z, Y, b, tol=1e-30, max_iter=args.iters)
    t_e2e = time.perf_counter() - t0
    return {"n": n, "m": m, "steps": steps, "iters_per_step": args.iters,
            "value": steps * int(info.niter) / (ms * 1e-3), "unit": UNIT,
            "ms_per_step": ms / steps,
            "e2e": {"value": steps * int(info.niter) / t_e2e, "unit": UNIT,
                    "ms_per_step": 1e3 * t_e2e / steps,
                    "note": "pageable numpy c, b in, x out"},
            "what": "this arm on the config `bench.py --impl reference` runs"}


def run_b200(args):
    import torch
    from ipsolver_b200 import synthetic, device
    from ipsolver_b200.projections import projections, set_inner_options
    from ipsolver_b200.qp_subproblem import projected_cg, HessianOperator

    if args.gpus != 1:
        return run_b200_dist(args)
    rank = 0
    n, m = args.n, args.n // 2
    set_inner_options(inner_rtol=args.inner_rtol)
    if args.band_chunk:
        set_inner_options(band=(0, args.band_chunk, 0))     # tuning aid: rows per sweep chunk
    t0 = time.perf_counter()
    H, A, c, b = synthetic.make_qp(n, args.pattern, args.kappa, seed=0)
    t_gen = time.perf_counter() - t0
    ctx = device.Context.default()
    t0 = time.perf_counter()
    Z, LS, Y = projections(A)                  # upload A, build A' on device
    Hop = HessianOperator(n, [H])
    Hd = Hop.device                            # upload H
    ctx.sync()
    t_setup = time.perf_counter() - t0
    proj = Z.projector

    # ---- value: inputs resident in HBM
    dc, db = device.DeviceVector(ctx, c), device.DeviceVector(ctx, b)
    dx = device.DeviceVector(ctx, length=n)
    sampler = ClockSampler(ctx.device)
    sampler.start()
    for _ in range(args.warmup):
        device.projected_cg_dev(proj, Hd, dc, db, dx, tol=1e-30, max_iter=args.iters)
    sampler.keep_loaded_until_sampling(
        lambda: device.projected_cg_dev(proj, Hd, dc, db, dx, tol=1e-30, max_iter=args.iters))
    launches0 = ctx.launches
    kernel_ms, infos = [], []
    ctx.timer_start()
    for _ in range(args.steps):
        info = device.projected_cg_dev(proj, Hd, dc, db, dx, tol=1e-30,
                                       max_iter=args.iters)
        kernel_ms.append(info.kernel_ms)
        infos.append(info)
    total_ms = ctx.timer_stop()
    launches = ctx.launches - launches0
    clocks = sampler.stop()
    niter = sum(int(i.niter) for i in infos)
    value = niter / (total_ms * 1e-3)

    # ---- roofline of the dominant kernel (the persistent solve kernel)
    i0 = infos[-1]
    nnzM = proj.aat_nnz
    band = proj.band_info
    bytes_launch, bytes_survey, parts = algorithmic_bytes(
        n, m, H.nnz, A.nnz, nnzM, int(i0.niter), int(i0.inner_iters), int(i0.n_solves),
        int(i0.n_refinements), int(i0.n_spmv_h), int(i0.n_sweeps), fused_hp=int(i0.reserved))
    avg_ms = float(np.mean(kernel_ms))
    peak, peak_src = peaks()
    achieved = bytes_launch / (avg_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("k_projected_cg_bytes_per_launch")
            traffic_src = ("NOT measured in this run: ncu --set full capture of the same "
                           "command, committed as profiles/traffic.json (%s)" % tj.get("source", "?"))
        except Exception:
            traffic = None
    if int(i0.n_sweeps) > 0:
        model = ("as run: (A A')^-1 by the banded direct solve (half-bandwidth %d, chunks of %d "
                 "rows + %d overlap), B_sweep = 160 m; Z applies verified by their "
                 "orthogonality product, the Y solve by one product with M (nnzM = %d)"
                 % (band["bandwidth"], band["chunk_rows"], band["overlap"], nnzM))
    elif nnzM > 0:
        model = "as run: inner CG on the explicit product M = A A' (nnzM = %d), B_in = B_M + 72 m" % nnzM
    else:
        model = "as run: inner CG applies A' then A, B_in = B_A + B_At + 80 m"
    roofline = {"bound": "hbm", "kernel": "k_projected_cg (persistent cooperative)",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_launch,
                "bytes_model": model,
                "avg_launch_ms": avg_ms,
                "inner_iters_per_launch": int(i0.inner_iters),
                "inner_iters_per_projection": int(i0.inner_iters) / max(int(i0.n_solves), 1),
                "solves_per_launch": int(i0.n_solves), "sweeps_per_launch": int(i0.n_sweeps),
                "refinements_per_launch": int(i0.n_refinements),
                "cg_fallbacks_per_launch": int(i0.n_fallbacks),
                "band": band,
                "bytes_per_unit": parts,
                "survey_8d_model": {
                    "algorithmic_bytes_per_launch": bytes_survey,
                    "effective_GBs": bytes_survey / (avg_ms * 1e-3) / 1e9,
                    "effective_frac": bytes_survey / (avg_ms * 1e-3) / 1e9 / peak,
                    "note": "same launch priced with SURVEY.md 8d's formulation (an inner CG of "
                            "35 iterations per solve, B_in = B_A + B_At + 56 m): the traffic "
                            "the path would need without the direct solve -- a speed-up "
                            "figure, not a roofline fraction"}}

    # phase timeline of one launch: %globaltimer at every grid barrier (CTA 0)
    try:
        ctx.trace(4096)
        device.projected_cg_dev(proj, Hd, dc, db, dx, tol=1e-30, max_iter=args.iters)
        stamps = ctx.trace_read().astype(np.int64)
        ctx.trace(0)
        d_us = np.diff(stamps) / 1e3
        names = timeline_names(i0)
        per_it = len(names) if names else None
        tl = {"barriers": int(len(stamps)), "sum_ms": float(d_us.sum() / 1e3)}
        if per_it and len(d_us) >= per_it * int(i0.niter):
            last = d_us[-per_it * int(i0.niter):].reshape(int(i0.niter), per_it)
            tl["per_iteration_us"] = {k: float(v) for k, v in zip(names, np.median(last, axis=0))}
        roofline["timeline"] = tl
    except Exception as exc:          # measurement aid only
        roofline["timeline"] = {"error": str(exc)}

    # the (A A')^-1 sweep in isolation (same device code, repeated inside one launch)
    if band["in_use"]:
        device.bench_phase(proj, 9, 3)
        sw_ms = min(device.bench_phase(proj, 9, 20) for _ in range(3)) / 20
        roofline["sweep_us"] = 1e3 * sw_ms
        roofline["sweep_GBs"] = parts["B_sweep"] / (sw_ms * 1e-3) / 1e9
    else:
        inner_ms = device.bench_inner_iterations(proj, dc, 50)
        inner_ms = device.bench_inner_iterations(proj, dc, 50)
        roofline["inner_iteration_us"] = 1e3 * inner_ms / 51
        roofline["inner_iteration_GBs"] = parts["B_in"] / (inner_ms * 1e-3 / 51) / 1e9

    # ---- e2e: host vectors through the public API
    hc = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
    hb = torch.empty(m, dtype=torch.float64).pin_memory().numpy()
    hc[:] = c
    hb[:] = b
    for _ in range(max(args.warmup, 2)):     # warms the pinned result-buffer pool as well
        x_e2e, _ = projected_cg(Hop, hc, Z, Y, hb, tol=1e-30, max_iter=args.iters)
    t0 = time.perf_counter()
    e2e_iters = 0
    for _ in range(args.steps):
        x_e2e, inf = projected_cg(Hop, hc, Z, Y, hb, tol=1e-30, max_iter=args.iters)
        e2e_iters += inf["niter"]
    t_e2e = time.perf_counter() - t0
    e2e = {"value": e2e_iters / t_e2e, "unit": UNIT,
           "h2d_bytes_per_step": 8 * (n + m), "d2h_bytes_per_step": 8 * n + 96,
           "ms_per_step": 1e3 * t_e2e / args.steps,
           "note": "projected_cg(H, c, Z, Y, b, ...) with pinned host c, b; x returned "
                   "as a numpy array; A and H uploaded once (setup_s)"}
    # the same call with ordinary (pageable) numpy arrays, as the SQP loop passes them
    for _ in range(2):
        projected_cg(Hop, c, Z, Y, b, tol=1e-30, max_iter=args.iters)
    t0 = time.perf_counter()
    pg_iters = 0
    for _ in range(args.steps):
        _, inf = projected_cg(Hop, c, Z, Y, b, tol=1e-30, max_iter=args.iters)
        pg_iters += inf["niter"]
    t_pg = time.perf_counter() - t0
    e2e["pageable_inputs"] = {"value": pg_iters / t_pg, "unit": UNIT,
                              "ms_per_step": 1e3 * t_pg / args.steps,
                              "note": "c, b ordinary numpy arrays (pageable H2D), result in a "
                                      "pinned pooled buffer"}
    # the result must be the same as the device-resident run
    xd = dx.download()
    same = float(np.linalg.norm(xd - x_e2e) / max(np.linalg.norm(xd), 1e-300))

    # ---- CPU baseline on a bounded sample (rank 0, N=1), and the GPU on that sample
    cpu, ref_cfg = None, None
    if not args.no_cpu:
        its, dt, t_factor, kind = cpu_reference_run(args.cpu_n, args.iters, args.cpu_steps, 0,
                                                    args.pattern, args.kappa)
        cpu = {"value": its, "unit": UNIT, "cores": 1, "blas_threads": host_threads(),
               "kind": kind,
               "sample": ("the unmodified reference (baseline/_ref" if kind == "reference"
                          else "oracle (port") + "; scipy SuperLU AugmentedSystem + numpy) on the same workload "
                         "at n=%d (the reference cannot factorise n=%d: SuperLU fails, "
                         "profiles/r3_reference_n1e7.txt): %d steps x %d forced iterations, "
                         "%.1f s timed + %.1f s factorisation; value is the MEASURED it/s at "
                         "that size" % (args.cpu_n, n, args.cpu_steps, args.iters, dt, t_factor),
               "n": args.cpu_n, "host_cpus": os.cpu_count(),
               "extrapolated_to_workload": {"value": its * args.cpu_n / n, "unit": UNIT,
                                            "how": "x n_sample/n (cost per iteration linear in n); "
                                                   "not a measurement"}}
        ref_cfg = gpu_run_at(args.cpu_n, args)

    # ---- end-to-end solve (BASELINE config 2: n = 1e6 through the SQP outer loop)
    solve = {}
    if args.solve_n > 0:
        solve["b200"] = end_to_end_solve(args.solve_n)
        solve["what"] = ("BASELINE config 2: minimize_constrained(method='equality_constrained_sqp') "
                         "on the synthetic convex QP, outer loop on the host, every linear-algebra "
                         "step on the device, vectors resident across the iteration")
        if args.solve_n == 1_000_000:
            solve["reference"] = reference_record("eqp_window_1000000")
    for key, name in (("config3_interior_point", args.solve_config3),
                      ("config4_thomson", args.solve_config4)):
        if name and name != "none":
            try:
                solve[key] = named_solve(name)
            except Exception as exc:          # never lose the headline over an extra
                solve[key] = {"problem": name, "error": repr(exc)}
    if args.solve_n > 0:
        if not args.no_cpu and args.solve_cpu_n > 0:
            import oracle
            solve["cpu_oracle"] = end_to_end_solve(args.solve_cpu_n, oracle.host_hot_path())

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu,
        "reference_config_run": ref_cfg,
        "setup_s": {"generate": t_gen, "upload_and_transpose": t_setup},
        "end_to_end_solve": solve,
        "grid": dict(zip(("blocks", "threads", "sms"), ctx.grid())),
        "e2e_vs_resident_rel_diff": same,
        "result_check": {"rel_feasibility": float(np.linalg.norm(A.dot(xd) + b) / np.linalg.norm(b)),
                         "x_checksum": x_checksum(xd)},
    }
    if rank == 0:
        print(json.dumps(line), flush=True)


def dist_end_to_end_solve(n, rank, size, group, local, D):
    """``minimize_constrained(method='equality_constrained_sqp')`` on the n-variable
    QP of BASELINE config 2 with ``DistSolver.hot_path()``: the host loop replicated
    on every rank, every device step on the ranks' own blocks."""
    import ipsolver_b200 as ipb
    from threadpoolctl import threadpool_limits
    from ipsolver_b200 import synthetic
    H, A, q, b = synthetic.make_qp(n, "window", 10.0, seed=0)
    t0 = time.perf_counter()
    S = D.DistSolver(H, A, rank, size, group=group, device=local)
    t_setup = time.perf_counter() - t0
    args_ = (lambda x: 0.5 * x.dot(H.dot(x)) + q.dot(x), np.zeros(n), lambda x: H.dot(x) + q,
             lambda x: H)
    with threadpool_limits(limits=8):     # torchrun exports OMP_NUM_THREADS=1; see end_to_end_solve
        t0 = time.perf_counter()
        res = ipb.minimize_constrained(*args_, ipb.LinearConstraint(A, ("equals", b)),
                                       method='equality_constrained_sqp', hot_path=S.hot_path(H, A))
        dt = time.perf_counter() - t0
    out = {"b200_row_partitioned": {
        "n": n, "m": n // 2, "ranks": size, "seconds": dt, "setup_s": t_setup,
        "niter": int(res.niter), "cg_niter": int(res.cg_niter), "nfev": int(res.nfev),
        "status": int(res.status), "optimality": float(res.optimality), "fun": float(res.fun),
        "constr_violation": float(res.constr_violation), "x_checksum": x_checksum(res.x)},
        "what": "BASELINE config 2 through equality_constrained_sqp with DistSolver.hot_path(): "
                "dogleg, projected CG, H d, A d, LS c on the row-partitioned GPUs, host loop replicated"}
    if n == 1_000_000:
        out["reference"] = reference_record("eqp_window_1000000")
    return out


def run_b200_dist(args):
    """N ranks (torchrun), the same QP row-partitioned over the GPUs of one
    box: strong scaling.  Each rank runs the same persistent kernel on its
    block; ghost entries and reductions cross NVLink inside the kernel
    (peer-memory stores + flag barriers), torch.distributed is used for setup
    and for the max-over-ranks clock only."""
    import torch
    import torch.distributed as dist
    from ipsolver_b200 import synthetic, distributed as D

    rank = int(os.environ.get("RANK", "0"))
    size = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if size != args.gpus:
        raise SystemExit("--gpus %d needs torchrun with %d ranks (WORLD_SIZE=%d)"
                         % (args.gpus, args.gpus, size))
    torch.cuda.set_device(local)
    # keep stdout for the one JSON line: NCCL and friends may print banners
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # setup-time collectives only; a bounded timeout so that a rank that failed in an extra leg
    # (the SQP solve below) surfaces as an exception on its peers instead of a hang
    import datetime
    gloo = dist.new_group(backend="gloo", timeout=datetime.timedelta(seconds=600))
    n, m = args.n, args.n // 2
    hashed = args.pattern == "hashed"
    t0 = time.perf_counter()
    if hashed:
        # index-addressable problem: each rank generates a window around its own rows
        H = A = c = b = None
        t_gen = 0.0
        S = D.DistSolver(None, None, rank, size, group=gloo, device=local,
                         inner_rtol=args.inner_rtol,
                         generated=dict(n=n, kappa=args.kappa, seed=0))
        (n0, n1), (m0, m1) = S.part.own(rank, 0), S.part.own(rank, 1)
        c_loc, b_loc = synthetic.hashed_vectors(n0, n1, m0, m1, 0)
    else:
        H, A, c, b = synthetic.make_qp(n, args.pattern, args.kappa, seed=0)
        t_gen = time.perf_counter() - t0
        t0 = time.perf_counter()
        S = D.DistSolver(H, A, rank, size, group=gloo, device=local, inner_rtol=args.inner_rtol)
        c_loc, b_loc = S.local(c, 0), S.local(b, 1)
    t_setup = time.perf_counter() - t0
    ctx = S.ctx
    dc, db = S.vector(c_loc), S.vector(b_loc, space=1)
    dx = S.vector(space=0)

    def sync_all():
        ctx.sync()
        dist.barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        S.projected_cg(dc, db, dx, tol=1e-30, max_iter=args.iters)
    # every rank must run the same number of extra warm-up steps (the solves are
    # collective): rank 0 decides, the others follow
    more = torch.zeros(1, dtype=torch.int32, device="cuda")
    for _ in range(400):
        more.fill_(1 if (rank == 0 and sampler.proc is not None and sampler.count() < 2) else 0)
        dist.broadcast(more, src=0)
        if int(more.item()) == 0:
            break
        S.projected_cg(dc, db, dx, tol=1e-30, max_iter=args.iters)
    launches0 = ctx.launches
    sync_all()
    infos, kernel_ms = [], []
    ctx.timer_start()
    for _ in range(args.steps):
        info = S.projected_cg(dc, db, dx, tol=1e-30, max_iter=args.iters)
        infos.append(info)
        kernel_ms.append(info.kernel_ms)
    my_ms = ctx.timer_stop()
    sync_all()
    launches = ctx.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([my_ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    niter = sum(int(i.niter) for i in infos)
    value = niter / (total_ms * 1e-3)

    # e2e: host blocks in pinned memory -> device, solve, result block back
    hc = torch.from_numpy(c_loc).pin_memory().numpy()
    hb = torch.from_numpy(b_loc).pin_memory().numpy()
    hx = torch.empty(S.n_own, dtype=torch.float64).pin_memory().numpy()
    sync_all()
    t0 = time.perf_counter()
    e2e_iters = 0
    for _ in range(args.steps):
        dc.upload(hc)
        db.upload(hb)
        inf = S.projected_cg(dc, db, dx, tol=1e-30, max_iter=args.iters)
        dx.download(hx)
        e2e_iters += int(inf.niter)
    ctx.sync()
    t_my = time.perf_counter() - t0
    t = torch.tensor([t_my], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_e2e = float(t.item())

    # roofline of the solve kernel: whole-problem bytes over the max-over-ranks time
    i0 = infos[-1]
    nnz_loc = torch.tensor([S.parts["M"].nnz, S.parts["H"].nnz, S.parts["A_raw"].nnz],
                           dtype=torch.int64, device="cuda")
    dist.all_reduce(nnz_loc)
    nnzM, nnzH, nnzA = (int(v) for v in nnz_loc.tolist())
    bytes_launch, bytes_survey, parts = algorithmic_bytes(
        n, m, nnzH, nnzA, nnzM, int(i0.niter), int(i0.inner_iters), int(i0.n_solves),
        int(i0.n_refinements), int(i0.n_spmv_h), int(i0.n_sweeps), fused_hp=int(i0.reserved))
    t = torch.tensor([float(np.mean(kernel_ms))], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    avg_ms = float(t.item())
    peak, peak_src = peaks()
    achieved = bytes_launch / (avg_ms * 1e-3) / 1e9
    # phase timeline of one launch on rank 0 (%globaltimer at every grid barrier)
    timeline = None
    try:
        ctx.trace(4096)
        sync_all()
        inf_t = S.projected_cg(dc, db, dx, tol=1e-30, max_iter=args.iters)
        stamps = ctx.trace_read().astype(np.int64)
        ctx.trace(0)
        d_us = np.diff(stamps) / 1e3
        timeline = {"barriers": int(len(stamps)), "sum_ms": float(d_us.sum() / 1e3),
                    "kernel_ms": float(inf_t.kernel_ms)}
        names = timeline_names(i0)
        if names and len(d_us) >= len(names) * int(i0.niter):
            last = d_us[-len(names) * int(i0.niter):].reshape(int(i0.niter), len(names))
            timeline["per_iteration_us"] = {k: float(v) for k, v in zip(names, np.median(last, axis=0))}
    except Exception as exc:
        timeline = {"error": str(exc)}
    if hashed:
        feas = None            # nobody holds the whole A; parity is scripts/dist_check.py's job
        x_all = S.gather(dx.download())
    else:
        x_all = S.gather(dx.download())
        feas = float(np.linalg.norm(A.dot(x_all) + b) / np.linalg.norm(b))
    checksum = x_checksum(x_all)

    # ---- BASELINE config 2 through the SQP outer loop, row-partitioned over the same ranks: the
    # driver's own multi-GPU record then shows N GPUs = 1 GPU = the reference's record
    solve = None
    if args.solve_n > 0 and not hashed:
        try:
            solve = dist_end_to_end_solve(args.solve_n, rank, size, gloo, local, D)
        except Exception as exc:          # never lose the headline over an extra
            solve = {"error": repr(exc)}
    if rank == 0:
        cfg = workload_config(args)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": size,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg, "clocks": clocks,
            "e2e": {"value": e2e_iters / t_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": 8 * (n + m), "d2h_bytes_per_step": 8 * n + 96 * size,
                    "ms_per_step": 1e3 * t_e2e / args.steps,
                    "note": "per rank: pinned host blocks of c, b uploaded, solve, block of x "
                            "downloaded; bytes are summed over the ranks"},
            "gpu_launches": int(launches) * size,
            "roofline": {"bound": "hbm", "kernel": "k_projected_cg (persistent cooperative, "
                                                    "one per rank)",
                         "achieved": achieved, "peak": peak * size, "unit": "GB/s",
                         "frac": achieved / (peak * size), "traffic": None,
                         "peak_source": peak_src + " x %d GPUs" % size,
                         "algorithmic_bytes_per_launch": bytes_launch,
                         "avg_launch_ms": avg_ms,
                         "inner_iters_per_launch": int(i0.inner_iters),
                         "sweeps_per_launch": int(i0.n_sweeps),
                         "cg_fallbacks_per_launch": int(i0.n_fallbacks),
                         "banded_direct_solve": bool(S.band),
                         "timeline_rank0": timeline,
                         "bytes_per_unit": parts},
            "cpu_baseline": None,
            "comm": {"data_path": "in-kernel NVLink peer-memory stores + flag barriers "
                                  "(no NCCL call inside the timed region)",
                     "ghosts_per_rank": {"n_space": [len(g) for g in S.parts["ghosts"][0]],
                                         "m_space": [len(g) for g in S.parts["ghosts"][1]]},
                     "cross_gpu_reductions_per_outer_iteration": 3,
                     "halo_exchanges_per_outer_iteration": 5},
            "setup_s": {"generate": t_gen, "partition_upload_connect": t_setup},
            "result_check": {"rel_feasibility": feas, "x_checksum": checksum},
            "end_to_end_solve": solve,
        }
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--problem-size", dest="n", type=int, default=10_000_000)
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--pattern", default="window")
    ap.add_argument("--kappa", type=float, default=1e4)
    ap.add_argument("--inner-rtol", type=float, default=1e-13)
    ap.add_argument("--band-chunk", type=int, default=0,
                    help="rows per chunk of the banded sweep (0: the library's rule)")
    ap.add_argument("--cpu-n", type=int, default=4_000_000)
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--solve-n", type=int, default=1_000_000,
                    help="size of the end-to-end SQP solve (0: skip)")
    ap.add_argument("--solve-config3", default="ipqp_window_100000",
                    help="tr_interior_point run reported under end_to_end_solve (tests/"
                         "e2e_problems.py LARGE_PROBLEMS; 'ipqp_window_1000000' = BASELINE config "
                         "3 at full size, minutes; 'none' to skip)")
    ap.add_argument("--solve-config4", default="thomson_500_equals",
                    help="Thomson run reported under end_to_end_solve ('thomson_2000_equals' = "
                         "BASELINE config 4 at full size, minutes; 'none' to skip)")
    ap.add_argument("--solve-cpu-n", type=int, default=100_000,
                    help="size of the same solve on the CPU oracle (0: skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
