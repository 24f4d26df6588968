"""CPU tests of the oracle itself: it must reproduce every golden vector the reference ships for this
path (the 7 Examples/*.out fronts) before it is trusted as a checker."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import aira_oracle as ao
from oracle.lpformat import parse_out, read_model, synthetic_ap, synthetic_kp, write_lp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["2AP05", "3AP05", "4AP05", "3KP10", "4KP10"]
# (IP count, generator iterations, cache hits) of the -t 1 stream: SURVEY.md section 3.5 probe table
STREAM = {"2AP05": (19, 9, 0), "3AP05": (64, 57, 34), "4AP05": (464, 1283, 1147), "3KP10": (33, 16, 4),
          "4KP10": (99, 82, 41)}


@pytest.mark.parametrize("stem", SMALL)
def test_oracle_reproduces_golden_front(examples, stem):
    e = examples[stem]
    m = read_model(e["path"])
    fs = ao.FeasibleSet(m)
    trace = []
    front = ao.pareto_front(m, fs, trace=trace)
    assert front == e["rows"] and len(front) == e["count"]
    assert ao.brute_force_front(m, fs) == e["rows"]              # solver-free ground truth
    ips, iters, hits = STREAM[stem]
    assert fs.ip_calls == ips and len(trace) - 1 == iters and sum(t[1] for t in trace) == hits


@pytest.mark.parametrize("stem", SMALL)
@pytest.mark.parametrize("threads,normal", [(2, False), (2, True), (8, False)])
def test_oracle_epp_reproduces_golden_front(examples, stem, threads, normal):
    e = examples[stem]
    m = read_model(e["path"])
    fs = ao.FeasibleSet(m)
    assert ao.pareto_front(m, fs, split=True, num_threads=threads, split_normal=normal) == e["rows"]
    assert ao.pareto_front(m, fs, split=True, num_threads=threads, split_normal=normal, shared_cache=False) == e["rows"]


def test_oracle_mop_and_general_integers(examples):
    e = examples["moip_2_30_1_knapsack"]
    m = read_model(e["path"])
    assert m.sense == "MIN" and m.k == 2 and m.n == 30 and np.all(m.ub > 1e19) and np.all(m.is_int)
    o = ao.MilpOracle(m)
    assert ao.pareto_front(m, o) == e["rows"]
    assert o.ip_calls == 141                                   # "IPs solved" of the committed .out


@pytest.mark.slow
def test_oracle_2kp50(examples):
    e = examples["2KP50"]
    m = read_model(e["path"])
    o = ao.MilpOracle(m)
    assert ao.pareto_front(m, o) == e["rows"] and o.ip_calls == 87


def test_lp_reader_details(examples, tmp_path):
    m = read_model(examples["2KP50"]["path"])
    assert m.b[0] == 1917.5 and m.row_sense == ["L"] and m.k == 2 and m.sense == "MAX"
    m = read_model(examples["2AP05"]["path"])
    assert m.row_sense == ["E"] * 10 and m.C[0, 9] == 0 and m.names[0] == "X1X1"
    # round trip through the writer used for the synthetic instances
    s = synthetic_kp(12, 3, 4)
    p = str(tmp_path / "kp.lp")
    write_lp(s, p)
    r = read_model(p)
    assert np.array_equal(r.A, s.A) and np.array_equal(r.C, s.C) and np.array_equal(r.b, s.b) and r.sense == "MAX"
    s = synthetic_ap(4, 2, 1)
    write_lp(s, p)
    r = read_model(p)
    # zero-cost columns still exist because the assignment rows name them
    assert r.n == 16 and np.array_equal(r.A, s.A) and np.array_equal(r.C, s.C)


def test_parse_out_ignores_timing_lines(examples):
    rows, count = parse_out(examples["4AP05"]["out_text"])
    assert count == 33 and len(rows) == 33 and rows[0] == (60, 39, 35, 32)


def test_solutions_restatement_matches_compiled_reference():
    """oracle.Solutions.find == the reference's own Solutions::find compiled from /root/reference
    (oracle/_ref, built by oracle/Makefile); skipped where that library was not built."""
    path = os.path.join(ROOT, "oracle", "_ref", "libaira_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (no reference tree)")
    lib = C.CDLL(path)
    lib.refsol_create.restype = C.c_void_p
    lib.refsol_create.argtypes = [C.c_int]
    lib.refsol_insert.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]
    lib.refsol_find.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
    lib.refsol_sort_unique.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int]
    lib.refsol_destroy.argtypes = [C.c_void_p]
    rng = np.random.default_rng(0)
    for k, sense in ((2, 0), (3, 1), (4, 0)):
        h = lib.refsol_create(k)
        py = ao.Solutions(k)
        for _ in range(300):
            ip = rng.integers(0, 12, k).astype(float)
            ip[rng.random(k) < 0.3] = 1e20 if sense == 0 else -1e20
            res = rng.integers(0, 12, k).astype(np.int32)
            inf = bool(rng.random() < 0.3)
            lib.refsol_insert(h, ip.ctypes.data_as(C.POINTER(C.c_double)), res.ctypes.data_as(C.POINTER(C.c_int)), int(inf))
            py.insert(ip, res, inf)
        for _ in range(500):
            q = rng.integers(0, 12, k).astype(float)
            assert lib.refsol_find(h, q.ctypes.data_as(C.POINTER(C.c_double)), sense) == py.find(q, "MIN" if sense == 0 else "MAX")[0]
        rows = np.zeros((400, k), dtype=np.int32)
        n = lib.refsol_sort_unique(h, k, rows.ctypes.data_as(C.POINTER(C.c_int)), 400)
        py.sort_unique()
        assert [tuple(r) for r in rows[:n]] == [tuple(r.result) for r in py.store if not r.infeasible]
        lib.refsol_destroy(h)


def test_pdhg_port_agrees_with_highs():
    """The C restatement of K1 (oracle/pdhg_ref.c) against HiGHS: 1e-6 relative on the LP value."""
    from oracle import pdhg_oracle as po
    m = synthetic_ap(8, 3, 1)
    cost, rhs, masks = po.sample_node_batch(m, 12, seed=7)
    st, obj = po.highs_lp(m, cost, rhs, masks)
    r = po.pdhg_ref(m, cost, rhs, masks, eps=1e-9, max_iter=400000)
    ok = st == 0
    assert ok.sum() >= 3
    assert np.all(np.abs(r["primal_obj"][ok] - obj[ok]) <= 1e-6 * np.maximum(1, np.abs(obj[ok])))
    assert np.all(r["status"][st == 2] == 3)
    assert np.all(r["dual_bound"][ok] <= obj[ok] + 1e-6 * np.maximum(1, np.abs(obj[ok])))


def test_synthetic_golden_against_bruteforce():
    """tests/golden/synthetic.json (made by the HiGHS-driven restatement) agrees with solver-free enumeration
    where that is cheap: 3-objective assignment n=8 (8! permutations)."""
    import json
    from oracle import aira_oracle as ao
    from oracle.lpformat import synthetic_ap
    with open(os.path.join(os.path.dirname(__file__), "golden", "synthetic.json")) as fh:
        g = json.load(fh)
    import itertools
    model = synthetic_ap(8, 3, 1)
    C = model.C.reshape(3, 8, 8)                       # cost of assigning row i to column j, per objective
    perms = np.array(list(itertools.permutations(range(8))))
    P = np.unique(np.rint(sum(C[:, i, perms[:, i]] for i in range(8))).astype(np.int64).T, axis=0)
    keep = []                                         # np.unique sorts lexicographically: a later row never dominates an earlier one
    for p in P:
        if not any(all(q[i] <= p[i] for i in range(3)) for q in keep):
            keep.append(tuple(int(v) for v in p))
    want = sorted(keep, key=lambda r: tuple(-v for v in r))
    assert [tuple(r) for r in g["ap3_8_1"]["rows"]] == want
    for name, v in g.items():       # every stored front is sorted like the reference prints it and has no dominated row
        rows = [tuple(r) for r in v["rows"]]
        assert rows == sorted(set(rows), key=lambda r: tuple(-x for x in r)), name
        sgn = 1 if v["kind"] == "ap" else -1
        R = np.array(rows) * sgn
        for i in range(0, len(R), max(1, len(R) // 40)):
            dom = np.all(R <= R[i], axis=1) & np.any(R < R[i], axis=1)
            assert not dom.any(), name
