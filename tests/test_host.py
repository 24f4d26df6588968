"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/moip_b200.h
declares, the CPLEX-free loader matches the oracle's independent reader, and the generator state
machine (host C++) reproduces the golden fronts when driven by the oracle's exact IP solver through
the callback hook.  No compute entry point is called here (there is no GPU and no CPU fallback)."""
import os
import re
import subprocess
import sys
import time

import numpy as np
import pytest

from oracle import aira_oracle as ao
from oracle.lpformat import read_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["2AP05", "3AP05", "4AP05", "3KP10", "4KP10"]


def test_abi_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "moip_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(moip_[a-z0-9_]+)\s*\(", header)) - {"moip_solve_fn", "moip_find_cb", "moip_insert_cb"}
    assert declared, "no declarations found"
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (moip_[a-z0-9_]+)", out))
    assert declared <= exported, declared - exported
    assert set(lib.EXPORTED) == declared
    assert "sm_100a" in lib.version()


def test_no_gpu_means_loud_failure(lib, examples):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    pr = lib.Problem(examples["2AP05"]["path"])
    with pytest.raises(lib.MoipError):
        lib.Context(pr)
    with pytest.raises(lib.MoipError):
        lib.WorkerPool(pr, 0, 2)


def test_seam_primitives_validate_arguments_and_need_a_gpu(lib, examples):
    """moip_mip_solve / moip_ctx_create_own_stream / moip_device_count (the three entry points behind Seam 1's CPXmipopt):
    argument errors are reported as MOIP_ERR_ARG; without a device the context cannot be created (no CPU solve)."""
    import ctypes as C
    import torch
    L = lib._lib
    pr = lib.Problem(examples["3KP10"]["path"])
    rhs = (C.c_double * 3)(1e20, 1e20, 1e20)
    st = C.c_int(0)
    assert L.moip_mip_solve(None, 0, rhs, None, None, None, C.byref(st)) == 1          # MOIP_ERR_ARG: no context
    h = C.c_void_p()
    assert L.moip_ctx_create_own_stream(None, 0, C.byref(h)) == 1
    assert L.moip_ctx_create_own_stream(pr._h, 0, None) == 1
    n = C.c_int(0)
    ss = (C.c_double * 2)(10.0, 0.0)
    assert L.moip_pool_run_boxes_claim(None, 3, 1, ss, ss, lib.CLAIM_FN(0), None, None, 0, C.byref(n)) == 1   # no pool
    assert L.moip_pool_run_boxes_claim(None, 3, 1, ss, None, lib.CLAIM_FN(0), None, None, 0, C.byref(n)) == 1  # boxes need windows
    assert L.moip_ctx_set_ip_node_budget(None, 10) == 1
    assert L.moip_pool_boxes_postponed(None) == -1
    if torch.cuda.is_available():
        assert L.moip_device_count() >= 1
    else:
        assert L.moip_device_count() == 0
        assert L.moip_ctx_create_own_stream(pr._h, 0, C.byref(h)) == 3 and not h.value  # MOIP_ERR_CUDA


@pytest.mark.parametrize("stem", SMALL + ["2KP50", "moip_2_30_1_knapsack"])
def test_loader_matches_oracle_reader(lib, examples, stem):
    path = examples[stem]["path"]
    pr, m = lib.Problem(path), read_model(path)
    A, rs, b, lb, ub, isint = pr.dense()
    assert (pr.n, pr.ms, pr.objcnt) == (m.n, m.ms, m.k)
    assert pr.objsen == (0 if m.sense == "MIN" else 1)
    assert np.array_equal(A, m.A) and rs == m.row_sense and np.array_equal(b, m.b)
    assert np.array_equal(pr.objcoef, m.C) and np.array_equal(lb, m.lb) and np.array_equal(ub, m.ub)
    assert np.array_equal(isint, m.is_int) and pr.colnames() == m.names
    assert np.all(np.abs(pr.rhs) == 1e20)                    # src/problem.cpp:122-132


@pytest.mark.parametrize("kind,k,n,path", [("ap", 3, 30, "k1_reg"), ("ap", 3, 12, "k1_reg"), ("ap", 4, 16, "k1_reg"), ("ap", 2, 20, "k1_reg"),
                                           ("ap", 3, 5, "k1_fast"), ("ap", 2, 40, "k1_fast"), ("ap", 2, 72, "generic-streaming"),
                                           ("kp", 4, 40, "k1_small"), ("kp", 3, 10, "k1_small"), ("kp", 2, 50, "k1_small"),
                                           ("kp", 3, 100, "k1_reg"), ("kp", 2, 2000, "k1_fast")])
def test_packed_kernel_images_are_consistent(lib, tmp_path, kind, k, n, path):
    """Host half of K1 (no GPU): the packed column records reproduce the scaled matrix, every short-row nonzero owns
    exactly one in-range product slot of the register-resident kernel, and the model maps to the expected kernel."""
    from moip_aira_b200 import instances
    p = str(tmp_path / f"{kind}{k}_{n}.lp")
    (instances.write_ap if kind == "ap" else instances.write_kp)(p, n, k, 1)
    got, err = lib.Problem(p).selfcheck()
    assert err is None, err
    assert got == path


def test_packed_kernel_images_of_the_examples(lib, examples):
    want = {"2AP05": "k1_fast", "3AP05": "k1_fast", "4AP05": "k1_fast", "3KP10": "k1_small", "4KP10": "k1_small",
            "2KP50": "k1_small", "moip_2_30_1_knapsack": "k1_small"}
    for stem, path in want.items():
        got, err = lib.Problem(examples[stem]["path"]).selfcheck()
        assert err is None and got == path, (stem, got, err)


def test_loader_rejects_bad_input(lib, tmp_path):
    p = tmp_path / "bad.lp"
    p.write_text("Minimize 0\nsubject to\n x + y <= 3\n x + 2 y < 7\nBINARY\n x\n y\nEND\n")   # k = 7 > rows
    with pytest.raises(lib.MoipError):
        lib.Problem(str(p))
    p.write_text("Minimize 0\nsubject to\n x + y <= 3\n x + 2.5 y < 1\nBINARY\n x\n y\nEND\n")  # fractional objective
    with pytest.raises(lib.MoipError):
        lib.Problem(str(p))
    with pytest.raises(lib.MoipError):
        lib.Problem(str(tmp_path / "missing.lp"))
    q = tmp_path / "x.txt"
    q.write_text("")
    with pytest.raises(lib.MoipError):
        lib.Problem(str(q))


def _drive(lib, model, fs, worker):
    s, inf = ao.Solutions(model.k), ao.Solutions(model.k)

    def solve(perm, n_obj, rhs):
        return fs.lex_solve(perm, n_obj, rhs)

    def find(ip):
        _, r = inf.find(ip, model.sense)
        if r is None:
            _, r = s.find(ip, model.sense)
        return None if r is None else (r.infeasible, r.result)

    def insert(ip, res, infeasible):
        (inf if infeasible else s).insert(ip, res, infeasible)

    it, hits = lib.optimise_with(model.k, 0 if model.sense == "MIN" else 1, worker, solve, find, insert)
    return s, it, hits


@pytest.mark.parametrize("stem", SMALL)
def test_generator_state_machine_reproduces_golden_front(lib, examples, stem):
    """Host C++ generator (moip_optimise_with) + oracle IP solver == committed .out, with the same
    iteration / cache-hit counts as the Python restatement."""
    e = examples[stem]
    m = read_model(e["path"])
    fs = ao.FeasibleSet(m)
    s, it, hits = _drive(lib, m, fs, lib.make_worker(m.k))
    s.sort_unique()
    assert [tuple(r.result) for r in s.store if not r.infeasible] == e["rows"]
    trace = []
    ao.pareto_front(m, ao.FeasibleSet(m), trace=trace)
    assert it == len(trace) - 1 and hits == sum(t[1] for t in trace)


def test_generator_other_permutations(lib, examples):
    """Every objective permutation enumerates the same front (what synergistic workers rely on)."""
    import itertools
    e = examples["3AP05"]
    m = read_model(e["path"])
    fs = ao.FeasibleSet(m)
    for perm in itertools.permutations(range(3)):
        s, _, _ = _drive(lib, m, fs, lib.make_worker(3, perm=perm))
        s.sort_unique()
        assert [tuple(r.result) for r in s.store if not r.infeasible] == e["rows"]


def _nondominated(P, minimise):
    """Brute-force Pareto filter of the objective vectors P (independent of every generator under test)."""
    pts = sorted({tuple(int(v) for v in p) for p in P})
    sgn = 1 if minimise else -1
    keep = []
    for p in pts:
        dominated = any(q != p and all(sgn * q[i] <= sgn * p[i] for i in range(len(p))) for q in pts)
        if not dominated:
            keep.append(p)
    return sorted(keep, reverse=True)             # Result::operator< order: descending lexicographic (src/result.cpp:20-27)


@pytest.mark.parametrize("kind,k,n,seed", [("kp", 2, 12, 11), ("kp", 3, 11, 12), ("kp", 4, 10, 13), ("kp", 3, 9, 14),
                                           ("ap", 2, 4, 15), ("ap", 3, 4, 16), ("ap", 4, 3, 17), ("ap", 3, 3, 18)])
def test_generator_on_random_instances_against_brute_force(lib, tmp_path, kind, k, n, seed):
    """Size-independent property on fresh synthetic instances (SURVEY 8d-4/5 generators, other seeds than the goldens):
    for EVERY objective permutation the host generator enumerates exactly the non-dominated set obtained by a
    brute-force Pareto filter over all feasible points."""
    import itertools
    import random
    from moip_aira_b200 import instances
    path = str(tmp_path / f"{kind}{k}_{n}_{seed}.lp")
    (instances.write_ap if kind == "ap" else instances.write_kp)(path, n, k, seed)
    m = read_model(path)
    fs = ao.FeasibleSet(m)
    want = _nondominated(fs.P, m.sense == "MIN")
    assert len(want) >= 1
    perms = list(itertools.permutations(range(k)))
    random.Random(seed).shuffle(perms)
    for perm in perms[:4]:
        s, _, _ = _drive(lib, m, fs, lib.make_worker(k, perm=perm))
        s.sort_unique()
        assert [tuple(r.result) for r in s.store if not r.infeasible] == want, perm


@pytest.mark.parametrize("sense,big,small,t,normal", [(0, 55, 21, 2, False), (1, 474, 361, 8, False), (0, 60, 21, 3, True),
                                                     (1, 100, 7, 12, True), (0, 2147483647, 5, 2, False)])
def test_split_strips_match_oracle(lib, sense, big, small, t, normal):
    got = lib.split_strips(sense, big, small, t, normal)
    want = ao.strip_bounds(sense == 0, big, small, t, normal)
    assert np.allclose(np.array(got), np.array(want), rtol=0, atol=0)


class _OracleBackend:
    """aira.py backend for host-logic tests: the oracle plays the solver, strips run through the host
    C++ generator via the callback hook."""

    def __init__(self, path):
        import moip_aira_b200 as mb
        self.mb = mb
        self.model = read_model(path)
        self.fs = ao.FeasibleSet(self.model)
        self.k = self.model.k
        self.sense = 0 if self.model.sense == "MIN" else 1

    def get_limit(self, obj, rhs):
        return self.fs.get_limit(obj, rhs)

    exchange = True            # offer the level's stores to aira.RecordExchange (the cross-rank cache sharing)

    def run_strips(self, n_obj, strips, claim, share=0, windows=None):
        here, inf = ao.Solutions(self.k), ao.Solutions(self.k)
        self._stores, self._cursor, self._foreign = (inf, here), [0, 0], set()
        self.hits_on_foreign = getattr(self, "hits_on_foreign", 0)

        def find(ip):
            _, r = inf.find(ip, self.model.sense)
            if r is None:
                _, r = here.find(ip, self.model.sense)
            if r is not None and id(r) in self._foreign:
                self.hits_on_foreign += 1
            return None if r is None else (r.infeasible, r.result)

        def insert(ip, res, infeasible):
            (inf if infeasible else here).insert(ip, res, infeasible)

        while True:
            t = claim()                       # job-wide counter: every strip is solved by exactly one rank
            if t >= len(strips):
                break
            a, b = strips[t]
            w = self.mb.make_worker(self.k, n_obj=n_obj, split=True, split_start=a, split_stop=b, wid=t,
                                    window=windows[t] if windows is not None else None)
            self.boxes_run = getattr(self, "boxes_run", 0) + (windows is not None)
            self.mb.optimise_with(self.k, self.sense, w, self.fs.lex_solve, find, insert)
            if getattr(self, "strip_delay", 0):
                time.sleep(self.strip_delay)  # lets the exchange thread carry records between the ranks
        stores, self._stores = self._stores, None
        return [tuple(r.result) for r in here.store if not r.infeasible and id(r) not in self._foreign]

    # -- the endpoint of aira.RecordExchange (the role moip_pool_export_records / _import_records play on the GPU)
    def exchange_endpoint(self):
        return self if self.exchange else None

    def export_records(self, cap):
        ip, res, infl = [], [], []
        stores = getattr(self, "_stores", None)
        if stores:
            for w, st in enumerate(stores):
                while self._cursor[w] < len(st.store) and len(infl) < cap:
                    r = st.store[self._cursor[w]]
                    self._cursor[w] += 1
                    if id(r) in self._foreign:
                        continue
                    ip.append(list(r.ip)); res.append([0] * self.k if r.infeasible else list(r.result)); infl.append(int(r.infeasible))
        return np.array(ip, dtype=float).reshape(-1, self.k), np.array(res, dtype=np.int32).reshape(-1, self.k), np.array(infl, dtype=np.int32)

    def import_records(self, ip, res, infl):
        stores = getattr(self, "_stores", None)
        if not stores:
            return
        for a, b, f in zip(ip, res, infl):
            st = stores[0] if f else stores[1]
            st.insert([float(v) for v in a], None if f else [int(v) for v in b], bool(f))
            self._foreign.add(id(st.store[-1]))

    def sequential_front(self):
        return ao.pareto_front(self.model, self.fs)

    def synergistic_local(self, n_workers):
        """-t N without --split in one process: min(N, k) cooperative workers on host threads (oracle as the solver)"""
        import threading
        w = max(1, min(int(n_workers), self.k))
        infs, sols, lock = ao.Solutions(self.k), [ao.Solutions(self.k) for _ in range(w)], threading.Lock()

        def solve(i, pm, n_obj, rhs):
            with lock:
                return self.fs.lex_solve(pm, n_obj, rhs)

        def find(i, ip):
            _, r = infs.find(ip, self.model.sense)
            if r is None:
                _, r = sols[i].find(ip, self.model.sense)
            return None if r is None else (r.infeasible, r.result)

        def insert(i, ip, res, infeasible):
            (infs if infeasible else sols[i]).insert(ip, res, infeasible)

        self.mb.coop_optimise_with(self.k, self.sense, w, solve, find, insert)
        return sorted({tuple(r.result) for s_ in sols for r in s_.store if not r.infeasible}, key=lambda r: tuple(-v for v in r))

    def split_strips(self, biggest, smallest, num_threads, split_normal):
        return self.mb.split_strips(self.sense, biggest, smallest, num_threads, split_normal)

    def ip_count(self):
        return self.fs.ip_calls


@pytest.mark.parametrize("stem", ["3KP10", "4AP05"])
def test_aira_cli_epp_single_process(lib, examples, stem, tmp_path):
    from moip_aira_b200 import aira
    from oracle.lpformat import parse_out
    e = examples[stem]
    out = str(tmp_path / "o.out")
    assert aira.main(["-p", e["path"], "-o", out, "--split", "-t", "3"], backend_factory=_OracleBackend) == 0
    assert parse_out(open(out).read()) == (e["rows"], e["count"])


def test_aira_cli_threads_without_split_runs_cooperative_workers(lib, examples, tmp_path, monkeypatch):
    """-t N without --split (the reference's synergistic mode, src/aira.cpp:277-308): min(N, k) cooperative workers;
    MOIP_THREADS_AS_STRIPS=1 maps the N workers onto N EPP strips instead.  Same front either way."""
    from moip_aira_b200 import aira
    from oracle.lpformat import parse_out
    e = examples["3AP05"]
    out = str(tmp_path / "o.out")
    assert aira.main(["-p", e["path"], "-o", out, "-t", "4"], backend_factory=_OracleBackend) == 0
    assert parse_out(open(out).read()) == (e["rows"], e["count"])
    monkeypatch.setenv("MOIP_THREADS_AS_STRIPS", "1")
    assert aira.main(["-p", e["path"], "-o", out, "-t", "4"], backend_factory=_OracleBackend) == 0
    assert parse_out(open(out).read()) == (e["rows"], e["count"])
    monkeypatch.delenv("MOIP_THREADS_AS_STRIPS")
    assert aira.main(["-p", e["path"], "-o", out], backend_factory=_OracleBackend) == 0      # -t 1: sequential generator
    assert parse_out(open(out).read()) == (e["rows"], e["count"])


_RANK_SCRIPT = r'''
import os, sys
sys.path.insert(0, {root!r})
sys.path.insert(0, os.path.join({root!r}, "tests"))
from moip_aira_b200 import aira
from test_host import _OracleBackend
rc = aira.main(["-p", {path!r}, "-o", {out!r}, "--split", "-t", "{threads}"], backend_factory=_OracleBackend)
sys.exit(rc)
'''


@pytest.mark.parametrize("stem,threads,windows", [("4KP10", 4, 0), ("3AP05", 2, 0), ("3AP05", 6, 3), ("4KP10", 8, 2), ("4AP05", 8, 4)])
def test_aira_cli_epp_two_ranks_gloo(lib, examples, stem, threads, windows, tmp_path, monkeypatch):
    """world_size 2 over gloo: the strips of every EPP level are sharded over the ranks and the points
    all-gathered between levels; rank 0 writes the same .out.  windows > 0: the levels with three or more objectives are cut
    into boxes (strips x windows on objective 1, moip_worker::window) dealt to the ranks as a Latin square."""
    from oracle.lpformat import parse_out
    if windows:
        monkeypatch.setenv("MOIP_WINDOWS", str(windows))
        monkeypatch.setenv("MOIP_LOWER_STRIPS", "3")
    e = examples[stem]
    out = str(tmp_path / "o.out")
    script = tmp_path / "rank.py"
    script.write_text(_RANK_SCRIPT.format(root=ROOT, path=e["path"], out=out, threads=threads))
    port = 29500 + (os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    assert parse_out(open(out).read()) == (e["rows"], e["count"])


_EXCHANGE_RANK_SCRIPT = r'''
import json, os, sys
sys.path.insert(0, {root!r})
sys.path.insert(0, os.path.join({root!r}, "tests"))
from moip_aira_b200 import aira
from test_host import _OracleBackend
dist = aira.Dist(None)
be = _OracleBackend({path!r})
be.strip_delay = 0.05
stats = []
front = aira.epp_front(be, dist, {threads}, False, stats)
json.dump({{"front": front, "stats": stats, "ips": be.ip_count(), "hits_on_foreign": be.hits_on_foreign}},
          open({out!r} + "." + os.environ["RANK"], "w"))
import torch.distributed as td
td.destroy_process_group()
'''


@pytest.mark.parametrize("stem,world,threads", [("4AP05", 2, 8), ("4KP10", 3, 9), ("3AP05", 2, 6)])
def test_epp_record_exchange_ranks_gloo(lib, examples, stem, world, threads, tmp_path):
    """The cross-rank knowledge exchange (aira.RecordExchange): while a level's strips run, every rank all-gathers the
    cache records it produces; the other ranks' strips may answer their subproblems from them (Solutions::find semantics,
    reference src/solutions.cpp:11-81, so the front cannot change).  Every rank must end up with the golden front, records
    must actually have travelled, and each rank reports only the points it found itself."""
    e = examples[stem]
    out = str(tmp_path / "x.json")
    script = tmp_path / "rank.py"
    script.write_text(_EXCHANGE_RANK_SCRIPT.format(root=ROOT, path=e["path"], out=out, threads=threads))
    port = 31500 + (os.getpid() % 2000)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), MOIP_EXCHANGE_PERIOD_MS="2")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    import json
    res = [json.load(open(out + "." + str(r))) for r in range(world)]
    for r in res:
        assert [list(x) for x in r["front"]] == [list(x) for x in e["rows"]]
    top = [r["stats"][-1] for r in res]
    assert sum(t["records_sent"] for t in top) > 0
    assert sum(t["records_received"] for t in top) == (world - 1) * sum(t["records_sent"] for t in top)
    assert all(t["exchange_rounds"] == top[0]["exchange_rounds"] for t in top)      # the loop ends in the same round everywhere


def test_pool_record_exchange_is_a_noop_between_runs(lib, examples):
    """moip_pool_export_records / _import_records need no GPU to refuse bad arguments"""
    import ctypes as C
    n = C.c_int(-1)
    assert lib._lib.moip_pool_export_records(None, 0, None, None, None, C.byref(n)) == 1
    assert lib._lib.moip_pool_import_records(None, 0, None, None, None) == 1
    assert lib._lib.moip_pool_set_max_workers(None, 1) == 1


def _windows(lo, hi, count, is_min):
    """`count` windows on an objective whose front values lie in [lo, hi]: (near edge, far edge) pairs; the first is open
    towards "free", the last has no far edge.  Bounds are upper limits for MIN models, lower limits for MAX models."""
    big = 1e20
    if count <= 1 or hi <= lo:
        return [(big if is_min else -big, -big if is_min else big)]
    cuts = sorted({lo + (hi - lo) * i // count for i in range(1, count)})
    out = []
    if is_min:
        edges = [big] + [float(c) for c in reversed(cuts)]           # near edges, from the top
        for i, e in enumerate(edges):
            far = edges[i + 1] + 1 if i + 1 < len(edges) else -big   # the next window starts at cut, this one ends at cut + 1
            out.append((e, far))
    else:
        edges = [-big] + [float(c) for c in cuts]
        for i, e in enumerate(edges):
            far = edges[i + 1] - 1 if i + 1 < len(edges) else big
            out.append((e, far))
    return out


@pytest.mark.parametrize("kind,k,n,seed", [("kp", 3, 11, 21), ("kp", 4, 10, 22), ("kp", 3, 12, 23), ("ap", 3, 4, 24),
                                           ("ap", 4, 3, 25), ("kp", 4, 9, 26), ("ap", 3, 4, 27), ("kp", 3, 10, 28)])
@pytest.mark.parametrize("strips,wins,shared", [(1, 2, True), (2, 3, True), (3, 4, False), (4, 2, True), (2, 5, False)])
def test_generator_boxes_enumerate_the_front(lib, tmp_path, kind, k, n, seed, strips, wins, shared):
    """EPP strips (ranges of the last objective) crossed with windows on the objective of the innermost sweeps
    (moip_worker::window): the boxes of the top level together enumerate exactly the brute-force front, with one pair of
    stores shared by all boxes (what a pool does) or a pair per box."""
    from moip_aira_b200 import instances
    path = str(tmp_path / f"{kind}{k}_{n}_{seed}.lp")
    (instances.write_ap if kind == "ap" else instances.write_kp)(path, n, k, seed)
    m = read_model(path)
    fs = ao.FeasibleSet(m)
    is_min = m.sense == "MIN"
    want = _nondominated(fs.P, is_min)
    last = [p[k - 1] for p in want]
    ss = lib.split_strips(0 if is_min else 1, max(last), min(last), strips, False)
    w1 = [p[1] for p in want]
    found = set()
    stores = (ao.Solutions(k), ao.Solutions(k))
    solved = 0
    for t in range(strips):
        for win in _windows(min(w1), max(w1), wins, is_min):
            s, inf = stores if shared else (ao.Solutions(k), ao.Solutions(k))

            def find(ip):
                _, r = inf.find(ip, m.sense)
                if r is None:
                    _, r = s.find(ip, m.sense)
                return None if r is None else (r.infeasible, r.result)

            def insert(ip, res, infeasible):
                (inf if infeasible else s).insert(ip, res, infeasible)

            def solve(perm, n_obj, rhs):
                nonlocal solved
                solved += 1
                return fs.lex_solve(perm, n_obj, rhs)
            w = lib.make_worker(k, split=True, split_start=ss[t][0], split_stop=ss[t][1], wid=t, window=win)
            lib.optimise_with(k, 0 if is_min else 1, w, solve, find, insert)
            found |= {tuple(r.result) for r in s.store if not r.infeasible}
    assert sorted(found, reverse=True) == want


def test_window_edges_and_counts(monkeypatch):
    """aira.window_edges: the windows tile the whole axis (first one open towards "free", last one without a far edge, one
    unit between a far edge and the next near edge), for both senses; aira.windows_for: powers of two that leave 8 strips."""
    from moip_aira_b200 import aira
    vals = sorted({3 * i + (i % 5) for i in range(40)})
    for is_min in (True, False):
        ed = aira.window_edges(vals, 4, is_min)
        assert 2 <= len(ed) <= 4
        assert ed[0][0] == (1e20 if is_min else -1e20) and ed[-1][1] == (-1e20 if is_min else 1e20)
        for (n0, f0), (n1, f1) in zip(ed, ed[1:]):
            assert (f0 == n1 + 1) if is_min else (f0 == n1 - 1)
            assert (n0 > n1) if is_min else (n0 < n1)
        # every value belongs to exactly one window
        for v in vals:
            owners = [i for i, (ne, fa) in enumerate(ed) if ((v <= ne and v >= fa) if is_min else (v >= ne and v <= fa))]
            assert len(owners) == 1, (v, ed)
    assert aira.window_edges([1, 2, 3], 4, True) == [(1e20, -1e20)]            # too few values: one window
    monkeypatch.delenv("MOIP_WINDOWS", raising=False)
    assert aira.windows_for(2, 768, 8) == 1
    assert [aira.windows_for(3, t, 1) for t in (8, 16, 24, 96, 192, 768)] == [1, 2, 2, 8, 16, 16]
    monkeypatch.setenv("MOIP_WINDOWS", "1")
    assert aira.windows_for(4, 768, 8) == 1
