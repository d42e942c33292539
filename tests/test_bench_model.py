"""The measurement model of ``bench.py`` against the
records committed under ``profiles/``: the bytes formula must reproduce the
figures the bench lines carry, the roofline fraction must be achieved / peak,
and the multi-GPU lines must describe the same solve as the single-GPU one.
CPU only -- nothing here touches the device or the oracle."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import bench  # noqa: E402


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip("no record " + name)
    return json.loads(open(path).read().strip().splitlines()[-1])


def test_bytes_per_unit_follow_the_stated_formulas():
    # B_spmv(rows, cols, nnz) = 12 nnz + 4 (rows + 1) + 8 cols + 8 rows
    n, m, nnzH, nnzA, nnzM = 1000, 500, 4994, 1500, 1930
    total, survey, parts = bench.algorithmic_bytes(n, m, nnzH, nnzA, nnzM, niter=8, inner_iters=11,
                                                   n_solves=11, n_refin=0, n_hp=9, n_sweeps=11,
                                                   fused_hp=1)
    spmv = lambda rows, cols, nnz: 12 * nnz + 4 * (rows + 1) + 8 * cols + 8 * rows  # noqa: E731
    assert parts["B_H"] == spmv(n, n, nnzH)
    assert parts["B_A"] == spmv(m, n, nnzA)
    assert parts["B_At"] == spmv(n, m, nnzA)
    assert parts["B_M"] == spmv(m, m, nnzM)
    assert parts["B_sweep"] == 160 * m
    assert parts["B_Z"] == 2 * parts["B_A"] + parts["B_At"] + 8 * n + parts["B_sweep"]
    # one launch: 10 Z applies + the Y solve, 9 Hessian products, 8 iterations of vector work
    z = 2 * parts["B_A"] + parts["B_At"] + 8 * n
    expect = (11 * parts["B_sweep"] + (parts["B_M"] + 8 * m) + 10 * z + (24 * m + parts["B_At"])
              + 9 * parts["B_H"] + 8 * 64 * n)
    assert total == expect
    # the unfused update costs one more n-vector read per iteration
    total_unfused, _, _ = bench.algorithmic_bytes(n, m, nnzH, nnzA, nnzM, 8, 11, 11, 0, 9, 11,
                                                  fused_hp=0)
    assert total_unfused - total == 8 * 8 * n
    assert survey > total          # the 35-iteration inner CG of SURVEY 8d moves far more


def test_cg_mode_counts_inner_iterations():
    n, m, nnzH, nnzA, nnzM = 1000, 500, 4994, 1500, 1930
    t1, _, p1 = bench.algorithmic_bytes(n, m, nnzH, nnzA, nnzM, 8, 100, 11, 0, 9)
    t2, _, _ = bench.algorithmic_bytes(n, m, nnzH, nnzA, nnzM, 8, 101, 11, 0, 9)
    assert t2 - t1 == p1["B_in"] == p1["B_M"] + 72 * m
    # without the explicit product an inner iteration is A' then A
    _, _, p3 = bench.algorithmic_bytes(n, m, nnzH, nnzA, -1, 8, 100, 11, 0, 9)
    assert p3["B_in"] == p3["B_A"] + p3["B_At"] + 8 * m + 72 * m


@pytest.mark.parametrize("name", ["r6i_bench.json", "r3y_bench.json"])
def test_single_gpu_record_is_consistent_with_the_model(name):
    d = _line(name)
    r, cfg = d["roofline"], d["config"]
    n, m = cfg["n"], cfg["m"]
    assert d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["dtype"] == "f64"
    # value = iterations / time, ms_per_step = time / steps
    assert d["value"] == pytest.approx(cfg["iters_per_step"] / (d["ms_per_step"] * 1e-3), rel=1e-6)
    # the bytes of the record are the model's for the matrices of the workload
    parts = r["bytes_per_unit"]
    nnzH = (parts["B_H"] - 4 * (n + 1) - 16 * n) // 12
    nnzA = (parts["B_A"] - 4 * (m + 1) - 8 * n - 8 * m) // 12
    nnzM = (parts["B_M"] - 4 * (m + 1) - 16 * m) // 12
    # 5-diagonal H; pivot + two window entries per row of A (coinciding entries merged)
    assert nnzH == pytest.approx(5 * n, rel=1e-3) and 2.5 * m < nnzA <= 3 * m
    total, _, _ = bench.algorithmic_bytes(n, m, nnzH, nnzA, nnzM, cfg["iters_per_step"],
                                          r["inner_iters_per_launch"], r["solves_per_launch"],
                                          r["refinements_per_launch"],
                                          # H x for the first residual, H p before the loop and
                                          # one H p per iteration (qp_subproblem.py:503,511,634)
                                          cfg["iters_per_step"] + 2,
                                          r["sweeps_per_launch"], fused_hp=1)
    assert total == r["algorithmic_bytes_per_launch"]
    # achieved = bytes / launch time, frac = achieved / peak, peak = the measured copy bandwidth
    assert r["achieved"] == pytest.approx(total / (r["avg_launch_ms"] * 1e-3) / 1e9, rel=1e-9)
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-12)
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0.5 < r["frac"] < 1.0
    # the direct solve: one sweep per solve, no fallback
    assert r["inner_iters_per_projection"] == 1.0 and r["cg_fallbacks_per_launch"] == 0
    # the kernel's own barrier timeline adds up to the launch
    assert r["timeline"]["sum_ms"] == pytest.approx(r["avg_launch_ms"], rel=0.03)
    # end to end: host vectors in and out are counted, and cost time
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 8 * (n + m) and e["d2h_bytes_per_step"] >= 8 * n
    assert e["value"] < d["value"]
    assert d["gpu_launches"] == d["steps"]
    assert d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] >= 0.95 * d["clocks"]["sm_max_mhz"]
    # BASELINE config 2 in the same record: the reference's counts
    s = d["end_to_end_solve"]
    for k in ("niter", "cg_niter", "nfev", "status"):
        assert s["b200"][k] == s["reference"][k]
    assert s["b200"]["optimality"] == pytest.approx(s["reference"]["optimality"], rel=1e-6)
    assert s["b200"]["fun"] == pytest.approx(s["reference"]["fun"], rel=1e-12)


@pytest.mark.parametrize("name", ["r6h_bench2.json", "r6k_bench8.json"])
def test_multi_gpu_records_describe_the_same_solve(name):
    one, many = _line("r6i_bench.json"), _line(name)
    assert many["n_gpus"] in (2, 8) and many["scaling"] == "strong"
    assert many["config"]["n"] == one["config"]["n"]
    assert (many["roofline"]["algorithmic_bytes_per_launch"]
            == one["roofline"]["algorithmic_bytes_per_launch"])
    a, b = one["result_check"]["x_checksum"], many["result_check"]["x_checksum"]
    for k in ("norm2", "sum", "weighted"):
        assert b[k] == pytest.approx(a[k], rel=1e-12)
    assert many["result_check"]["rel_feasibility"] < 1e-12
    assert many["value"] > one["value"]
    # config 2 through the row-partitioned SQP loop: the reference's record
    s = many["end_to_end_solve"]
    got, ref = s["b200_row_partitioned"], s["reference"]
    assert got["ranks"] == many["n_gpus"]
    for k in ("niter", "cg_niter", "nfev", "status"):
        assert got[k] == ref[k]
    assert got["optimality"] == pytest.approx(ref["optimality"], rel=1e-6)
    assert got["fun"] == pytest.approx(ref["fun"], rel=1e-12)


def test_x_checksum_is_independent_of_how_x_is_split():
    rng = np.random.RandomState(0)
    x = rng.standard_normal(10007)
    whole = bench.x_checksum(x)
    again = bench.x_checksum(np.concatenate([x[:1234], x[1234:7000], x[7000:]]))
    assert whole == again and whole["n"] == x.size
    assert whole["norm2"] == pytest.approx(np.linalg.norm(x))
    other = bench.x_checksum(x[::-1].copy())
    assert other["weighted"] != whole["weighted"]      # position-dependent: a permuted x differs
