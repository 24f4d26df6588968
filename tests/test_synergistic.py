"""Cooperative ("synergistic") workers -- SURVEY.md 8f-3: `-t W` without --split (src/aira.cpp:277-308) with a defined,
race-free replacement for the reference's bound-sharing cells (src/aira.cpp:923-1086, :1111-1552; the reference's own
protocol loses points for W >= 4, profiles/r01_seam1.md).  Protocol: csrc/generator.cpp (CoopBackend).

CPU tests drive the host C++ (moip_coop_optimise_with: W host threads, monotone limit atomics) with the oracle's exact
solver through the callback hook and compare with a brute-force Pareto filter; GPU tests run the product path
(moip_pool_synergistic_front: worker i on solver context i of one GPU) against the committed goldens."""
import json
import os
import threading
import time

import pytest

from oracle import aira_oracle as ao
from oracle.lpformat import read_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _nondominated(P, minimise):
    """Brute-force Pareto filter of the objective vectors P: every point is tested against the non-dominated points
    among its lexicographic predecessors (a dominating point always precedes it lexicographically)."""
    sgn = 1 if minimise else -1
    pts = sorted({tuple(sgn * int(v) for v in p) for p in P})
    keep = []
    for p in pts:
        if not any(all(q[i] <= p[i] for i in range(len(p))) for q in keep):
            keep.append(p)
    return sorted([tuple(sgn * v for v in p) for p in keep], reverse=True)


def _coop(lib, model, fs, workers, delay=0.0):
    """W cooperative workers on the oracle's solver: per-worker solution stores, one shared infeasible store."""
    k = model.k
    lock = threading.Lock()
    inf = ao.Solutions(k)
    sols = [ao.Solutions(k) for _ in range(workers)]

    def solve(w, perm, n_obj, rhs):
        r = fs.lex_solve(perm, n_obj, rhs)
        if delay:
            time.sleep(delay)              # equal cost per subproblem; releases the GIL so that the workers interleave
        return r

    def find(w, ip):
        with lock:
            _, r = inf.find(ip, model.sense)
        if r is None:
            _, r = sols[w].find(ip, model.sense)
        return None if r is None else (r.infeasible, r.result)

    def insert(w, ip, res, infeasible):
        if infeasible:
            with lock:
                inf.insert(ip, res, True)
        else:
            sols[w].insert(ip, res, False)

    solves, skipped = lib.coop_optimise_with(k, 0 if model.sense == "MIN" else 1, workers, solve, find, insert)
    pts = sorted({tuple(r.result) for s in sols for r in s.store if not r.infeasible}, reverse=True)
    return pts, solves, skipped


def test_coop_worker_permutations(lib):
    for k in (2, 3, 4):
        for w in range(1, k + 1):
            perms = lib.coop_workers(k, w)
            assert all(sorted(p) == list(range(k)) for p in perms)
            assert len({p[-1] for p in perms}) == w                 # every worker owns a different objective
            assert perms[0] == list(range(k))                       # worker 0 is the reference's -t 1 worker
    with pytest.raises(lib.MoipError):
        lib.coop_workers(3, 4)                                      # at most one worker per objective


def test_coop_rejects_workers_sharing_an_objective(lib):
    import ctypes as C
    ws = (lib.Worker * 2)(lib.make_worker(3, perm=(0, 1, 2)), lib.make_worker(3, perm=(1, 0, 2), wid=1))
    cb = lambda *a: 0                                               # noqa: E731  never called
    rc = lib._lib.moip_coop_optimise_with(3, 0, 2, ws, lib.SOLVE_FN(cb), lib.FIND_CB(cb), lib.INSERT_CB(cb), None, None, None)
    assert rc == 1                                                  # MOIP_ERR_ARG
    ws = (lib.Worker * 1)(lib.make_worker(3, perm=(0, 1, 2), split=True))
    assert lib._lib.moip_coop_optimise_with(3, 0, 1, ws, lib.SOLVE_FN(cb), lib.FIND_CB(cb), lib.INSERT_CB(cb), None, None,
                                            None) == 1
    assert C.sizeof(lib.Worker) > 0


@pytest.mark.parametrize("kind,k,n,seed", [("kp", 3, 11, 31), ("kp", 3, 11, 33), ("kp", 3, 11, 36), ("ap", 3, 4, 41),
                                           ("ap", 3, 4, 46), ("kp", 4, 9, 53), ("kp", 2, 12, 63), ("ap", 4, 3, 71),
                                           ("ap", 2, 5, 72)])
def test_coop_front_equals_brute_force(lib, tmp_path, kind, k, n, seed):
    """For every worker count 1..k and several (scheduler-dependent) interleavings, the union of the workers' points
    is exactly the non-dominated set of a brute-force Pareto filter; W = 1 behaves like the sequential generator."""
    from moip_aira_b200 import instances
    path = str(tmp_path / f"{kind}{k}_{n}_{seed}.lp")
    (instances.write_ap if kind == "ap" else instances.write_kp)(path, n, k, seed)
    m = read_model(path)
    fs = ao.FeasibleSet(m)
    want = _nondominated(fs.P, m.sense == "MIN")
    base = None
    for workers in range(1, k + 1):
        for delay in (0.0, 0.001):
            pts, solves, skipped = _coop(lib, m, fs, workers, delay)
            assert pts == want, (workers, delay, solves, skipped)
            if workers == 1:
                base = solves[0]
                assert skipped == [0]
    assert base >= len(want)


@pytest.mark.parametrize("kind,k,n,seed", [("ap", 3, 6, 81), ("ap", 3, 7, 82), ("kp", 3, 15, 84), ("kp", 4, 13, 85)])
def test_coop_shortens_the_critical_path(lib, tmp_path, kind, k, n, seed):
    """The point of the protocol: with equal-cost subproblems the busiest of k workers solves about half as many
    subproblems as the single worker (measured: 3AP n=6 67 -> 29, 3AP n=7 123 -> 57, 4KP n=13 81 -> 39; on the goldens
    3AP n=8 179 -> 79 with 3 workers and 4KP n=20 226 -> 91 with 4), at 25-50 % more subproblems in total."""
    from moip_aira_b200 import instances
    path = str(tmp_path / f"{kind}{k}_{n}_{seed}.lp")
    (instances.write_ap if kind == "ap" else instances.write_kp)(path, n, k, seed)
    m = read_model(path)
    fs = ao.FeasibleSet(m)
    want = _nondominated(fs.P, m.sense == "MIN")
    pts1, solves1, _ = _coop(lib, m, fs, 1)
    ptsw, solvesw, _ = _coop(lib, m, fs, k, delay=0.002)
    assert pts1 == want and ptsw == want
    assert max(solvesw) <= 0.75 * solves1[0], (solves1, solvesw)
    assert sum(solvesw) <= 2.0 * solves1[0], (solves1, solvesw)


def test_coop_finished_partner_stops_the_others(lib, examples):
    """A worker that runs to completion publishes 'done'; partners answer what is left as infeasible without solving.
    Forced here by making every worker but the first wait until the first one is through."""
    m = read_model(examples["3AP05"]["path"])
    fs = ao.FeasibleSet(m)
    k = m.k
    first_done = threading.Event()
    lock = threading.Lock()
    inf = ao.Solutions(k)
    sols = [ao.Solutions(k) for _ in range(k)]
    calls = [0] * k

    def solve(w, perm, n_obj, rhs):
        if w != 0:
            first_done.wait(timeout=60)
        calls[w] += 1
        return fs.lex_solve(perm, n_obj, rhs)

    def find(w, ip):
        with lock:
            _, r = inf.find(ip, m.sense)
        if r is None:
            _, r = sols[w].find(ip, m.sense)
        return None if r is None else (r.infeasible, r.result)

    def insert(w, ip, res, infeasible):
        if infeasible:
            with lock:
                inf.insert(ip, res, True)
        else:
            sols[w].insert(ip, res, False)

    def watcher():                      # worker 0 is done when its solve count stops moving; it never waits
        last = -1
        while True:
            time.sleep(0.2)
            if calls[0] == last and last > 0:
                first_done.set()
                return
            last = calls[0]

    threading.Thread(target=watcher, daemon=True).start()
    solves, skipped = lib.coop_optimise_with(k, 0, k, solve, find, insert)
    pts = sorted({tuple(r.result) for s in sols for r in s.store if not r.infeasible}, reverse=True)
    assert pts == [tuple(r) for r in examples["3AP05"]["rows"]]
    assert all(s <= 1 for s in solves[1:]) and all(sk >= 1 for sk in skipped[1:]), (solves, skipped)


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.timeout(180)
@pytest.mark.parametrize("stem", ["2AP05", "3KP10", "3AP05", "4KP10", "4AP05"])
def test_synergistic_front_examples_gpu(lib, examples, stem):
    pr = lib.Problem(examples[stem]["path"])
    pool = lib.WorkerPool(pr, 0, pr.objcnt)
    try:
        for workers in range(1, pr.objcnt + 1):
            assert pool.synergistic_front(workers) == examples[stem]["rows"], workers
        st = pool.stats()
        assert st["kernel_launches"] > 0 and st["node_lps"] > 0
    finally:
        pool.close()


@pytest.mark.gpu
@pytest.mark.timeout(300)
@pytest.mark.parametrize("name", ["ap3_12_1", "kp4_20_1"])
def test_synergistic_front_synthetic_gpu(lib, tmp_path, name):
    from moip_aira_b200 import instances
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "synthetic.json")))[name]
    path = str(tmp_path / (name + ".lp"))
    (instances.write_ap if g["kind"] == "ap" else instances.write_kp)(path, g["n"], g["k"], g["seed"])
    pr = lib.Problem(path)
    pool = lib.WorkerPool(pr, 0, pr.objcnt)
    try:
        assert pool.synergistic_front(pr.objcnt) == [tuple(r) for r in g["rows"]]
    finally:
        pool.close()
