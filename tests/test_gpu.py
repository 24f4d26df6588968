"""GPU parity tests (run with -m gpu on a B200): every call goes through the C ABI of
include/moip_b200.h; the checker is the oracle (oracle/), the golden fronts of the reference's
Examples (tests/golden/examples.json) and -- for K3 -- the reference's own compiled Solutions::find
(oracle/_ref/libaira_ref.so)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["2AP05", "3AP05", "4AP05", "3KP10", "4KP10"]


@pytest.fixture(scope="module")
def mb(lib):
    return lib


def _oracle(path):
    from oracle import aira_oracle as ao
    from oracle.lpformat import read_model
    m = read_model(path)
    return m, ao.FeasibleSet(m)


def _ref_solutions():
    path = os.path.join(ROOT, "oracle", "_ref", "libaira_ref.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.refsol_create.restype = C.c_void_p
    lib.refsol_create.argtypes = [C.c_int]
    lib.refsol_destroy.argtypes = [C.c_void_p]
    lib.refsol_insert.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]
    lib.refsol_find_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int)]
    return lib


# ----------------------------------------------------------------------------------------- K3
def _random_store(rng, k, R, sense, lo=0, hi=60):
    recs = []
    for _ in range(R):
        ip = rng.integers(lo, hi, size=k).astype(float)
        free = rng.random(k) < 0.3
        ip[free] = 1e20 if sense == 0 else -1e20
        inf = rng.random() < 0.25
        res = rng.integers(lo, hi, size=k).astype(np.int32)
        recs.append((ip, res, inf))
    return recs


@pytest.mark.parametrize("k,R,Q,sense", [(2, 1, 5, 0), (3, 37, 64, 0), (4, 1000, 300, 1), (4, 5000, 2000, 0), (3, 129, 1, 1)])
def test_k3_scan_matches_reference_find(mb, examples, k, R, Q, sense):
    """moip_cache_find_batch == reference Solutions::find (src/solutions.cpp:11-81), bit-exact indices."""
    from oracle import aira_oracle as ao
    stem = {2: "2AP05", 3: "3AP05", 4: "4AP05"}[k]
    ctx = mb.Context(mb.Problem(examples[stem]["path"]))
    rng = np.random.default_rng(11)
    recs = _random_store(rng, k, R, sense)
    store = mb.Solutions(ctx)
    py = ao.Solutions(k)
    ref = _ref_solutions()
    h = ref.refsol_create(k) if ref else None
    for ip, res, inf in recs:
        store.insert(ip, res, inf)
        py.insert(ip, res, inf)
        if ref:
            ref.refsol_insert(h, ip.ctypes.data_as(C.POINTER(C.c_double)), res.ctypes.data_as(C.POINTER(C.c_int)), int(inf))
    qs = rng.integers(0, 60, size=(Q, k)).astype(float)
    qs[rng.random((Q, k)) < 0.2] = 1e20 if sense == 0 else -1e20
    got = store.find_batch(qs, sense)
    want = np.array([py.find(q, "MIN" if sense == 0 else "MAX")[0] for q in qs])
    assert np.array_equal(got, want)
    if ref:
        out = np.zeros(Q, dtype=np.int32)
        ref.refsol_find_batch(h, Q, k, qs.ctypes.data_as(C.POINTER(C.c_double)), sense, out.ctypes.data_as(C.POINTER(C.c_int)))
        assert np.array_equal(got, out)
        ref.refsol_destroy(h)
    assert len(store) == R
    # empty store and empty query batch
    empty = mb.Solutions(ctx)
    assert np.array_equal(empty.find_batch(qs[:1], sense), [-1])
    assert len(store.find_batch(np.zeros((0, k)), sense)) == 0
    ctx.close()


def test_k3_merge_and_sort_unique(mb, examples):
    """merge splices in front (src/solutions.h:41-44); sort_unique = descending lexicographic + dedupe."""
    from oracle import aira_oracle as ao
    ctx = mb.Context(mb.Problem(examples["3AP05"]["path"]))
    rng = np.random.default_rng(3)
    a, b = mb.Solutions(ctx), mb.Solutions(ctx)
    pa, pb = ao.Solutions(3), ao.Solutions(3)
    for s, p, cnt in ((a, pa, 40), (b, pb, 25)):
        for _ in range(cnt):
            ip = rng.integers(0, 9, 3).astype(float)
            res = rng.integers(0, 4, 3).astype(np.int32)
            inf = rng.random() < 0.2
            s.insert(ip, res, inf)
            p.insert(ip, res, inf)
    a.merge(b)
    pa.merge(pb)
    assert len(a) == 65 and len(b) == 0
    q = rng.integers(0, 9, size=(50, 3)).astype(float)
    assert np.array_equal(a.find_batch(q, 0), [pa.find(x, "MIN")[0] for x in q])
    pa.sort_unique()
    assert a.sort_unique() == [tuple(r.result) for r in pa.store if not r.infeasible]
    ctx.close()


# ----------------------------------------------------------------------------------------- K4
@pytest.mark.parametrize("stem", ["3AP05", "4KP10", "2KP50", "moip_2_30_1_knapsack"])
def test_k4_verify_exact(mb, examples, stem):
    from oracle.lpformat import read_model
    path = examples[stem]["path"]
    m = read_model(path)
    pr = mb.Problem(path)
    ctx = mb.Context(pr)
    rng = np.random.default_rng(5)
    B = 257
    ub = np.where(m.ub > 1e19, 3, m.ub).astype(int)
    x = rng.integers(0, ub + 1, size=(B, m.n)).astype(np.int32)
    if stem.endswith("AP05"):
        for b in range(0, B, 2):       # half of the points are permutation matrices (feasible)
            x[b] = np.eye(5, dtype=np.int32)[rng.permutation(5)].reshape(-1)
    obj, feas = ctx.verify(x)
    want_obj = (x.astype(object) @ m.C.astype(int).astype(object).T)
    assert np.array_equal(obj.astype(object), want_obj)
    act = x.astype(np.int64) @ m.A.T.astype(np.float64)
    ok = np.ones(B, dtype=bool)
    for i, s in enumerate(m.row_sense):
        if s in "LE":
            ok &= act[:, i] <= m.b[i] + 1e-9
        if s in "GE":
            ok &= act[:, i] >= m.b[i] - 1e-9
    assert np.array_equal(feas, ok)
    # with objective-bound rows
    rhs = np.tile(np.median(obj, axis=0).astype(float), (B, 1))
    rhs[:, 0] = 1e20 if m.sense == "MIN" else -1e20
    _, feas2 = ctx.verify(x, rhs)
    sgn = 1 if m.sense == "MIN" else -1
    ok2 = ok & np.all((sgn * obj[:, 1:] <= sgn * rhs[:, 1:]), axis=1)
    assert np.array_equal(feas2, ok2)
    ctx.close()


# ----------------------------------------------------------------------------------------- K1
def _lp_case(kind):
    from oracle.lpformat import synthetic_ap, synthetic_kp
    kind = kind.split("-")[0]
    k = 4 if kind.startswith("kp") else 3
    if kind[-1] in "234" and kind[-2] == "k":       # e.g. ap16k4: 4 objectives
        k = int(kind[-1]); kind = kind[:-2]
    size = int(kind[2:])
    return synthetic_ap(size, k, 1) if kind.startswith("ap") else synthetic_kp(size, k, 1)


# K1 code paths: k1_small (all rows dense, n <= 64: 8 lanes per node, CPT = 2 / 5 / 8 columns per lane), k1_fast (other
# models with n <= 64), the register-resident k1_reg with NT x CPT = 128x2 (n <= 256), 256x2 (n <= 512) and 256x4
# (n <= 1024), and the generic kernel.
@pytest.mark.parametrize("kind,B", [("ap8", 24), ("kp40", 48), ("ap30", 16), ("ap12", 24), ("ap20", 16),
                                    ("kp100", 32), ("ap16k4", 16), ("ap12k2", 16), ("kp12k3", 64), ("kp50k2", 32), ("kp64", 32)])
def test_k1_lp_objective_vs_highs_and_port(mb, tmp_path, kind, B, monkeypatch):
    """LP relaxation objectives within 1e-6 relative of HiGHS (stand-in: the reference pins no LP value)
    and of the C restatement; infeasible nodes are recognised; the dual bound is a valid bound."""
    from oracle import pdhg_oracle as po
    from oracle.lpformat import write_lp
    model = _lp_case(kind)
    path = str(tmp_path / f"{kind}.lp")
    write_lp(model, path)
    ctx = mb.Context(mb.Problem(path))
    cost, rhs, masks = po.sample_node_batch(model, B, seed=7)
    got = ctx.lp_batch_solve(cost, rhs, masks, mb.Context.lp_params(eps=1e-9, max_iter=400000), want_x=True)
    st, obj = po.highs_lp(model, cost, rhs, masks)
    port = po.pdhg_ref(model, cost, rhs, masks, eps=1e-9, max_iter=400000)
    sgn = 1.0 if model.sense == "MIN" else -1.0
    feas = st == 0
    assert feas.sum() >= 3
    assert np.all(got["status"][feas] == mb.LP_CONVERGED)
    rel = np.abs(got["primal_obj"][feas] - obj[feas]) / np.maximum(1.0, np.abs(obj[feas]))
    assert rel.max() <= 1e-6, rel                                   # north-star tolerance: 1e-6 relative
    relp = np.abs(got["primal_obj"][feas] - port["primal_obj"][feas]) / np.maximum(1.0, np.abs(obj[feas]))
    assert relp.max() <= 1e-6
    # valid bound: never on the wrong side of the true LP optimum
    assert np.all(sgn * got["dual_bound"][feas] <= sgn * obj[feas] + 1e-6 * np.maximum(1, np.abs(obj[feas])))
    # infeasible nodes (HiGHS status 2) must not be reported as converged
    assert np.all(got["status"][st == 2] == mb.LP_INFEASIBLE)
    assert np.array_equal(got["status"] == mb.LP_INFEASIBLE, port["status"] == 3)
    # x respects the fixings
    l, u = po.unpack_masks(model, masks)
    assert np.all(got["x"] >= l - 1e-9) and np.all(got["x"] <= u + 1e-9)
    ctx.close()


@pytest.mark.parametrize("kind", ["ap8", "ap12", "ap30", "kp100", "kp40", "kp12k3", "kp50k2"])
def test_k1_fixed_iterations_match_port(mb, tmp_path, kind, monkeypatch):
    """Same arithmetic as the C restatement: after a fixed number of iterations the iterates agree."""
    from oracle import pdhg_oracle as po
    from oracle.lpformat import write_lp
    model = _lp_case(kind)
    path = str(tmp_path / f"{kind}.lp")
    write_lp(model, path)
    ctx = mb.Context(mb.Problem(path))
    cost, rhs, masks = po.sample_node_batch(model, 12, seed=3)
    for iters in (1, 7, 40, 100):
        got = ctx.lp_batch_solve(cost, rhs, masks, mb.Context.lp_params(fixed_iters=iters), want_x=True)
        port = po.pdhg_ref(model, cost, rhs, masks, fixed_iters=iters, norm_every=int(os.environ.get("MOIP_NORM_EVERY", "16")))
        assert np.all(got["iters"] == iters)
        assert np.allclose(got["x"], port["x"], rtol=0, atol=1e-9)
        assert np.allclose(got["primal_obj"], port["primal_obj"], rtol=1e-10, atol=1e-9)
        assert np.allclose(got["dual_bound"], port["dual_bound"], rtol=1e-9, atol=1e-8)
    ctx.close()


@pytest.mark.parametrize("streaming", [False, True])
def test_k1_generic_kernel_and_streaming_mode(mb, examples, tmp_path, monkeypatch, streaming):
    """The generic kernel (models outside the specialised paths) and its streaming mode (iterate in an HBM scratch
    region, used when 5n + 8m doubles exceed shared memory) give the port's iterates and the golden front."""
    from oracle import pdhg_oracle as po
    from oracle.lpformat import write_lp
    monkeypatch.setenv("MOIP_K1_GENERIC", "1")
    if streaming:
        monkeypatch.setenv("MOIP_K1_FORCE_STREAMING", "1")
    model = _lp_case("ap8")
    path = str(tmp_path / "ap8.lp")
    write_lp(model, path)
    ctx = mb.Context(mb.Problem(path))
    cost, rhs, masks = po.sample_node_batch(model, 12, seed=3)
    got = ctx.lp_batch_solve(cost, rhs, masks, mb.Context.lp_params(fixed_iters=40), want_x=True)
    port = po.pdhg_ref(model, cost, rhs, masks, fixed_iters=40, norm_every=int(os.environ.get("MOIP_NORM_EVERY", "16")))
    assert np.allclose(got["x"], port["x"], rtol=0, atol=1e-9)
    ctx.close()
    e = examples["3AP05"]
    ctx = mb.Context(mb.Problem(e["path"]))
    assert ctx.pareto_front() == e["rows"]
    ctx.close()


def test_k1_cutoff_and_edge_cases(mb, examples):
    pr = mb.Problem(examples["2AP05"]["path"])
    ctx = mb.Context(pr)
    free = [1e20, 1e20]
    base = ctx.lp_batch_solve([0], [free], None, mb.Context.lp_params(eps=1e-9))
    assert base["status"][0] == mb.LP_CONVERGED and abs(base["primal_obj"][0] - 21.0) < 1e-5   # front row (21,55)
    cut = ctx.lp_batch_solve([0], [free], None, mb.Context.lp_params(eps=1e-9, cutoff=15.0))
    assert cut["status"][0] == mb.LP_CUTOFF and cut["dual_bound"][0] >= 15.0
    # all columns fixed to 0 violates the assignment rows
    words = pr.mask_words
    allzero = np.zeros((1, words), dtype=np.uint32)
    for j in range(pr.n):
        allzero[0, j >> 4] |= np.uint32(2 << ((j & 15) * 2))
    inf = ctx.lp_batch_solve([1], [free], allzero, mb.Context.lp_params(eps=1e-9))
    assert inf["status"][0] == mb.LP_INFEASIBLE
    assert len(ctx.lp_batch_solve([], np.zeros((0, 2)))["status"]) == 0
    ctx.close()


# ----------------------------------------------------------------------------------------- solve()
@pytest.mark.parametrize("stem", SMALL)
def test_lex_solve_matches_oracle_on_recorded_stream(mb, examples, stem):
    """moip_lex_solve == the exact lexicographic optimum for every subproblem the sequential
    generator issues on this instance (reference src/aira.cpp:452-536 semantics)."""
    from oracle import aira_oracle as ao
    path = examples[stem]["path"]
    model, fs = _oracle(path)
    trace = []
    ao.pareto_front(model, fs, trace=trace)
    ctx = mb.Context(mb.Problem(path))
    k = model.k
    seen = set()
    for rhs, hit, infeasible, result in trace:
        if hit or rhs in seen:
            continue
        seen.add(rhs)
        st, res = ctx.solve(rhs)
        if infeasible:
            assert st == mb.MIP_INFEASIBLE, (rhs, st, res)
        else:
            assert st == mb.MIP_OPTIMAL and tuple(res) == result, (rhs, res, result)
    # other permutations and partial lexicographic depth (EPP sub-levels)
    rng = np.random.default_rng(1)
    for _ in range(6):
        perm = [int(v) for v in rng.permutation(k)]
        n_obj = int(rng.integers(1, k + 1))
        rhs = list(trace[int(rng.integers(0, len(trace)))][0])
        st, res = ctx.solve(rhs, perm=perm, n_obj=n_obj)
        ost, ores = fs.lex_solve(perm, n_obj, rhs)
        assert st == ost
        if ores is not None:
            assert [res[perm[i]] for i in range(n_obj)] == [ores[perm[i]] for i in range(n_obj)]
    s = ctx.stats()
    assert s["ip_solved"] > 0 and s["node_lps"] > 0 and s["kernel_launches"] > 0
    ctx.close()


def test_get_limit(mb, examples):
    from oracle import aira_oracle as ao
    for stem in ("3AP05", "4KP10"):
        path = examples[stem]["path"]
        model, fs = _oracle(path)
        ctx = mb.Context(mb.Problem(path))
        free = [1e20 if model.sense == "MIN" else -1e20] * model.k
        for obj in range(model.k):
            st, res = ctx.get_limit(obj, free)
            want = fs.get_limit(obj, free)
            assert st == mb.MIP_OPTIMAL and res[obj] == want[obj]
        tight = list(free)
        tight[0] = -5.0 if model.sense == "MIN" else 1e7
        st, res = ctx.get_limit(1, tight)
        assert st == mb.MIP_INFEASIBLE and res is None
        ctx.close()


# ----------------------------------------------------------------------------------------- fronts
@pytest.mark.parametrize("stem", SMALL + ["2KP50", "moip_2_30_1_knapsack"])
def test_front_matches_golden_out(mb, examples, stem):
    """The reference's own acceptance test (Examples/CMakeLists.txt:4-7 + scripts/checkResults.sh:10):
    front rows and `N Solutions found` equal the committed .out, default options (-t 1)."""
    e = examples[stem]
    ctx = mb.Context(mb.Problem(e["path"]))
    front = ctx.pareto_front()
    assert front == e["rows"] and len(front) == e["count"]
    ctx.close()


@pytest.mark.parametrize("stem", SMALL)
@pytest.mark.parametrize("threads,normal", [(2, False), (2, True), (8, False)])
def test_front_epp_matches_golden_out(mb, examples, stem, threads, normal):
    """`-t 2 --split` (flat2) and `-t 2 --split --split-normal` (normal2) of Examples/CMakeLists.txt:19-28,
    plus the 8-strip configuration of BASELINE.json configs[2]."""
    e = examples[stem]
    ctx = mb.Context(mb.Problem(e["path"]))
    front = ctx.pareto_front(split=True, num_threads=threads, split_normal=normal)
    assert front == e["rows"] and len(front) == e["count"]
    ctx.close()


@pytest.mark.parametrize("stem", SMALL)
@pytest.mark.parametrize("threads,normal,workers", [(2, False, 2), (8, False, 4), (8, True, 8), (3, False, 8)])
def test_front_pool_matches_golden_out(mb, examples, stem, threads, normal, workers):
    """Concurrent strips (one solver context per host thread on one GPU, reference src/aira.cpp:1920-1933)
    give the same front as the committed .out; the pool's strip runner agrees with the sequential one."""
    e = examples[stem]
    pr = mb.Problem(e["path"])
    pool = mb.WorkerPool(pr, 0, workers)
    assert pool.workers == workers
    front = pool.pareto_front(threads, normal)
    assert front == e["rows"] and len(front) == e["count"]
    st = pool.stats()
    assert st["ip_solved"] > 0 and st["kernel_launches"] > 0
    pool.close()


def _synthetic_case(name, tmp_path):
    from moip_aira_b200 import instances
    import glob
    g = {}
    for f in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.json"))):
        if os.path.basename(f) != "examples.json":
            with open(f) as fh:
                g.update(json.load(fh))
    if name not in g:
        pytest.skip(f"no oracle front committed for {name} (tests/golden/make_ap30.py)")
    g = g[name]
    path = str(tmp_path / f"{name}.lp")
    (instances.write_ap if g["kind"] == "ap" else instances.write_kp)(path, g["n"], g["k"], g["seed"])
    return path, [tuple(r) for r in g["rows"]]


@pytest.mark.parametrize("name", ["ap3_8_1", "ap3_10_1", "ap4_7_2", "ap2_20_3", "kp4_20_1", "kp2_80_5"])
def test_front_synthetic_golden_sequential(mb, tmp_path, name):
    """Synthetic AP / KP instances of SURVEY.md 8d (items 4-5): the front of the default (-t 1) run equals the
    front the CPU oracle computed (tests/golden/make_synthetic.py: restated generator + HiGHS, knapsacks
    cross-checked by a solver-free DP)."""
    path, want = _synthetic_case(name, tmp_path)
    ctx = mb.Context(mb.Problem(path))
    assert ctx.pareto_front() == want
    ctx.close()


@pytest.mark.parametrize("name,strips,workers", [("ap3_12_1", 12, 12), ("ap3_15_1", 12, 12), ("kp4_25_1", 8, 8), ("kp3_40_1", 12, 6),
                                                 ("ap3_20_1", 12, 12), ("ap3_30_1", 12, 12), ("kp4_40_1", 12, 12)])
def test_front_synthetic_golden_pool(mb, tmp_path, name, strips, workers):
    """Same, larger instances, EPP strips solved concurrently on one GPU (--split -t strips)."""
    path, want = _synthetic_case(name, tmp_path)
    pool = mb.WorkerPool(mb.Problem(path), 0, workers)
    assert pool.pareto_front(strips) == want
    pool.close()


@pytest.mark.parametrize("name,strips,workers", [("ap3_12_1", 2, 8), ("kp4_25_1", 1, 6), ("ap3_15_1", 3, 12)])
def test_pool_work_stealing(mb, tmp_path, name, strips, workers):
    """Fewer strips than workers: the idle workers cut the busy strips' remaining ranges in two (moip_pool_run_strips_claim).
    A strip is only a range of the last objective (src/aira.cpp:1895-1916), so the front must not change."""
    path, want = _synthetic_case(name, tmp_path)
    pool = mb.WorkerPool(mb.Problem(path), 0, workers)
    assert pool.pareto_front(strips) == want
    assert pool.strips_stolen() > 0
    pool.close()


@pytest.mark.parametrize("name,env", [("ap3_12_1", {"MOIP_CHAIN": "0"}), ("ap3_12_1", {"MOIP_CHAIN_Q": "64"}),
                                      ("kp4_25_1", {"MOIP_CHAIN": "0"}), ("kp4_25_1", {"MOIP_CHAIN_Q": "64"}),
                                      ("ap3_15_1", {"MOIP_CHAIN_Q": "256", "MOIP_BB_LEVELS": "1"})])
def test_chained_rounds_and_host_loop_agree(mb, tmp_path, monkeypatch, name, env):
    """The B&B of an IP runs chained on the device (csrc/bbchain.h) or round by round through the host (MOIP_CHAIN=0); with a
    small pool (MOIP_CHAIN_Q) tree levels outgrow it and those IPs are handed back to the host loop in mid-run.  Same front."""
    path, want = _synthetic_case(name, tmp_path)
    for a, v in env.items():
        monkeypatch.setenv(a, v)
    pool = mb.WorkerPool(mb.Problem(path), 0, 8)
    assert pool.pareto_front(8) == want
    pool.close()


@pytest.mark.parametrize("name,boxes,wins", [("ap3_12_1", 16, 4), ("kp4_25_1", 16, 2), ("ap3_15_1", 32, 8), ("kp3_40_1", 12, 3)])
def test_front_boxes_golden(mb, tmp_path, monkeypatch, name, boxes, wins):
    """EPP strips crossed with windows on objective 1 (moip_worker::window, aira.epp_front): same front."""
    from moip_aira_b200 import aira
    path, want = _synthetic_case(name, tmp_path)
    monkeypatch.setenv("MOIP_WINDOWS", str(wins))
    monkeypatch.setenv("MOIP_WORKERS", "8")
    be = aira.GpuBackend(path, device=0)
    stats = []
    front = aira.epp_front(be, aira.Dist(None), boxes, False, stats)
    assert [tuple(r) for r in front] == want
    assert stats[-1]["strips"] <= boxes and stats[-1]["windows"] >= 1
    be.pool.close()


def test_front_synthetic_vs_bruteforce(mb, tmp_path):
    """Synthetic assignment / knapsack instances against the solver-free brute-force front."""
    from oracle import aira_oracle as ao
    from oracle.lpformat import synthetic_ap, synthetic_kp, write_lp
    for name, model in (("ap6", synthetic_ap(6, 3, 2)), ("kp14", synthetic_kp(14, 4, 3)), ("ap5k2", synthetic_ap(5, 2, 9))):
        path = str(tmp_path / f"{name}.lp")
        write_lp(model, path)
        fs = ao.FeasibleSet(model)
        want = ao.brute_force_front(model, fs)
        ctx = mb.Context(mb.Problem(path))
        assert ctx.pareto_front() == want
        assert ctx.pareto_front(split=True, num_threads=3) == want
        ctx.close()


def test_out_file_layout(mb, examples, tmp_path):
    """aira -p <file> -o <out>: same layout as the reference writer (src/aira.cpp:252, :336-358),
    compared with the rule of scripts/checkResults.sh:10."""
    from moip_aira_b200 import aira
    from oracle.lpformat import parse_out
    e = examples["3KP10"]
    out = str(tmp_path / "3KP10.out")
    rc = aira.main(["-p", e["path"], "-o", out])
    assert rc == 0
    text = open(out).read()
    assert parse_out(text) == (e["rows"], e["count"])
    lines = text.splitlines()
    assert lines[0] == "" and lines[1].startswith("Using improved algorithm")
    assert all(l.endswith("\t") for l in lines[2:2 + e["count"]])
    assert lines[2 + e["count"]] == "" and lines[3 + e["count"]] == "---"
    assert lines[-1].endswith("Solutions found") and lines[-2].endswith("IPs solved")
