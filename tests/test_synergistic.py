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


class _OracleCoopBackend:
    """aira.synergistic_front backend for host-logic tests: the oracle's exact solver plays the GPU (one cooperative
    worker per rank through moip_coop_optimise_one_with); every solve sleeps a little so that the ranks interleave."""

    def __init__(self, path, delay=0.002):
        import moip_aira_b200 as mb
        self.mb = mb
        self.model = read_model(path)
        self.fs = ao.FeasibleSet(self.model)
        self.k = self.model.k
        self.sense = 0 if self.model.sense == "MIN" else 1
        self.delay = delay
        self.solves = self.skipped = 0

    def coop_worker(self, perm, limits):
        inf, sols = ao.Solutions(self.k), ao.Solutions(self.k)

        def solve(pm, n_obj, rhs):
            time.sleep(self.delay)
            return self.fs.lex_solve(pm, n_obj, rhs)

        def find(ip):
            _, r = inf.find(ip, self.model.sense)
            if r is None:
                _, r = sols.find(ip, self.model.sense)
            return None if r is None else (r.infeasible, r.result)

        def insert(ip, res, infeasible):
            (inf if infeasible else sols).insert(ip, res, infeasible)

        self.solves, self.skipped = self.mb.coop_optimise_one_with(self.k, self.sense, perm, limits, solve, find, insert)
        return sorted({tuple(r.result) for r in sols.store if not r.infeasible}, reverse=True)


def test_coop_limits_handle(lib):
    lim = lib.CoopLimits(3, 0, [2, 0])                   # MIN: tighter = smaller
    assert lim.read(2) == (0, 0)
    lim.publish(2, 50)
    lim.publish(2, 70)                                   # a stale, looser value never loosens the limit
    assert lim.read(2) == (1, 50)
    lim.publish(2, 41)
    assert lim.read(2) == (1, 41)
    lim.publish(0, done=True)
    assert lim.read(0)[0] == 2
    lim.publish(0, 5)                                    # done is final
    assert lim.read(0)[0] == 2
    lim.close()
    lim = lib.CoopLimits(3, 1, [1])                      # MAX: tighter = larger
    lim.publish(1, 10)
    lim.publish(1, 7)
    lim.publish(1, 12)
    assert lim.read(1) == (1, 12)
    with pytest.raises(lib.MoipError):
        lim.publish(3, 1)
    lim.close()


def test_synergistic_front_single_process_host_logic(lib, examples):
    """world = 1: one worker, no exchange -- the sequential front."""
    from moip_aira_b200 import aira
    be = _OracleCoopBackend(examples["3AP05"]["path"], delay=0.0)
    assert aira.synergistic_front(be, aira.Dist(None)) == examples["3AP05"]["rows"]
    assert be.skipped == 0


_COOP_RANK_SCRIPT = r'''
import json, os, sys
sys.path.insert(0, {root!r})
sys.path.insert(0, os.path.join({root!r}, "tests"))
from moip_aira_b200 import aira
from test_synergistic import _OracleCoopBackend
dist = aira.Dist(None)
be = _OracleCoopBackend({path!r})
front = aira.synergistic_front(be, dist)
json.dump({{"front": front, "solves": be.solves, "skipped": be.skipped}}, open({out!r} + "." + os.environ["RANK"], "w"))
import torch.distributed as td
td.destroy_process_group()
'''


@pytest.mark.parametrize("stem,world", [("3AP05", 2), ("3AP05", 3), ("4KP10", 4), ("2AP05", 2), ("3KP10", 4)])
def test_synergistic_front_ranks_gloo(lib, examples, stem, world, tmp_path):
    """One cooperative worker per rank over gloo (the CPU stand-in for one rank per GPU over NCCL): limits travel through
    the job's c10d store, the points are all-gathered at the end; every rank ends up with the golden front.  A rank
    beyond k (3KP10 with 4 ranks) owns nothing and only gathers."""
    import subprocess
    import sys
    e = examples[stem]
    out = str(tmp_path / "front.json")
    script = tmp_path / "rank.py"
    script.write_text(_COOP_RANK_SCRIPT.format(root=ROOT, path=e["path"], out=out))
    port = 31500 + (os.getpid() % 2000)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    res = [json.load(open(out + "." + str(r))) for r in range(world)]
    for r in res:
        assert [tuple(x) for x in r["front"]] == e["rows"]
    k = len(e["rows"][0])
    assert all(r["solves"] == 0 for r in res[k:])        # ranks without an objective of their own do not solve


_INFEASIBLE_LP = """\\ no feasible point: x0 + x1 >= 3 with binaries
minimize 0
subject to
 x0 + x1 >= 3
 3 x0 + 2 x1 + 4 x2 > 1
 1 x0 + 5 x1 + 2 x2 > 2
binary
 x0 x1 x2
end
"""
_SINGLE_POINT_LP = """\\ exactly one feasible point (all ones): objectives 9 and 8
minimize 0
subject to
 x0 + x1 + x2 = 3
 3 x0 + 2 x1 + 4 x2 > 1
 1 x0 + 5 x1 + 2 x2 > 2
binary
 x0 x1 x2
end
"""


@pytest.mark.parametrize("text,want", [(_INFEASIBLE_LP, []), (_SINGLE_POINT_LP, [(9, 8)])])
def test_edge_models_host_logic(lib, tmp_path, text, want):
    """Empty front (infeasible model) and a one-point front through the sequential generator, the cooperative workers
    and the reference driver behind the seam."""
    import re
    import subprocess
    path = str(tmp_path / "edge.lp")
    open(path, "w").write(text)
    m = read_model(path)
    fs = ao.FeasibleSet(m)
    assert _nondominated(fs.P, True) == want
    for workers in (1, 2):
        pts, solves, skipped = _coop(lib, m, fs, workers)
        assert pts == want
    aira = os.path.join(ROOT, "oracle", "_ref", "aira_seam1")
    fake = os.path.join(ROOT, "oracle", "_build", "libfake_mip.so")
    if os.path.exists(aira) and os.path.exists(fake):
        out = str(tmp_path / "edge.out")
        r = subprocess.run([aira, "-p", path, "-o", out], env=dict(os.environ, LD_PRELOAD=fake), capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        text_out = open(out).read()
        rows = [tuple(int(v) for v in l.split()) for l in text_out.splitlines() if re.fullmatch(r"\s*-?\d+(\s+-?\d+)*\s*", l)]
        assert rows == want and f"{len(want)} Solutions found" in text_out


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
def test_synergistic_front_synthetic_gpu(lib, tmp_path):
    """3AP n=12 (263 points, committed golden) with 2 and 3 cooperative workers on one GPU."""
    from moip_aira_b200 import instances
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "synthetic.json")))["ap3_12_1"]
    path = str(tmp_path / "ap3_12_1.lp")
    instances.write_ap(path, g["n"], g["k"], g["seed"])
    pr = lib.Problem(path)
    pool = lib.WorkerPool(pr, 0, pr.objcnt)
    try:
        for workers in (2, 3):
            assert pool.synergistic_front(workers) == [tuple(r) for r in g["rows"]], workers
    finally:
        pool.close()


@pytest.mark.gpu
@pytest.mark.timeout(300)
@pytest.mark.parametrize("kind,k,n,seed", [("kp", 4, 13, 85), ("kp", 4, 14, 83), ("kp", 3, 15, 84), ("ap", 3, 6, 81)])
def test_synergistic_front_random_instances_gpu(lib, tmp_path, kind, k, n, seed):
    """Fresh small instances, k workers on one GPU, against the brute-force Pareto filter (the CPU cases above)."""
    from moip_aira_b200 import instances
    path = str(tmp_path / f"{kind}{k}_{n}_{seed}.lp")
    (instances.write_ap if kind == "ap" else instances.write_kp)(path, n, k, seed)
    m = read_model(path)
    want = _nondominated(ao.FeasibleSet(m).P, m.sense == "MIN")
    pr = lib.Problem(path)
    pool = lib.WorkerPool(pr, 0, k)
    try:
        assert pool.synergistic_front(k) == want
    finally:
        pool.close()


@pytest.mark.gpu
@pytest.mark.timeout(180)
def test_coop_worker_on_a_limits_handle_gpu(lib, examples):
    """moip_coop_optimise on the GPU: the per-rank worker of aira.synergistic_front.  (a) world = 1: no partner, the
    sequential front; (b) a partner that has published limits and then 'done': the worker only looks where the partner
    has not, and everything it finds is a point of the golden front inside that region."""
    from moip_aira_b200 import aira
    e = examples["3AP05"]
    be = aira.GpuBackend(e["path"], device=0)
    assert aira.synergistic_front(be, aira.Dist(None)) == e["rows"]
    rows = e["rows"]
    cut = sorted(r[0] for r in rows)[len(rows) // 2]           # pretend the owner of objective 0 has covered f0 > cut
    lim = lib.CoopLimits(3, 0, [2, 0])
    lim.publish(0, cut)
    found = be.coop_worker([0, 1, 2], lim)
    assert set(found) == {r for r in rows if r[0] <= cut}
    lim.publish(0, done=True)
    assert be.coop_worker([0, 1, 2], lim) == []
    lim.close()


@pytest.mark.gpu
@pytest.mark.timeout(180)
@pytest.mark.parametrize("text,want", [(_INFEASIBLE_LP, []), (_SINGLE_POINT_LP, [(9, 8)])])
def test_edge_models_gpu(lib, tmp_path, text, want):
    """Empty and one-point fronts on the GPU: sequential generator, EPP strips and cooperative workers."""
    path = str(tmp_path / "edge.lp")
    open(path, "w").write(text)
    pr = lib.Problem(path)
    ctx = lib.Context(pr, device=0)
    try:
        assert ctx.pareto_front() == want
    finally:
        ctx.close()
    pool = lib.WorkerPool(pr, 0, 2)
    try:
        assert pool.pareto_front(2) == want
        assert pool.synergistic_front(2) == want
    finally:
        pool.close()
