"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/moip_b200.h
declares, the CPLEX-free loader matches the oracle's independent reader, and the generator state
machine (host C++) reproduces the golden fronts when driven by the oracle's exact IP solver through
the callback hook.  No compute entry point is called here (there is no GPU and no CPU fallback)."""
import os
import re
import subprocess
import sys

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

    def run_strips(self, n_obj, strips, claim):
        here, inf = ao.Solutions(self.k), ao.Solutions(self.k)

        def find(ip):
            _, r = inf.find(ip, self.model.sense)
            if r is None:
                _, r = here.find(ip, self.model.sense)
            return None if r is None else (r.infeasible, r.result)

        def insert(ip, res, infeasible):
            (inf if infeasible else here).insert(ip, res, infeasible)

        while True:
            t = claim()                       # job-wide counter: every strip is solved by exactly one rank
            if t >= len(strips):
                break
            a, b = strips[t]
            w = self.mb.make_worker(self.k, n_obj=n_obj, split=True, split_start=a, split_stop=b, wid=t)
            self.mb.optimise_with(self.k, self.sense, w, self.fs.lex_solve, find, insert)
        return [tuple(r.result) for r in here.store if not r.infeasible]

    def sequential_front(self):
        return ao.pareto_front(self.model, self.fs)

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


def test_aira_cli_threads_without_split_runs_strips(lib, examples, tmp_path):
    """-t N without --split (the reference's synergistic mode): same front, computed as N EPP strips."""
    from moip_aira_b200 import aira
    from oracle.lpformat import parse_out
    e = examples["3AP05"]
    out = str(tmp_path / "o.out")
    assert aira.main(["-p", e["path"], "-o", out, "-t", "4"], backend_factory=_OracleBackend) == 0
    assert parse_out(open(out).read()) == (e["rows"], e["count"])
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


@pytest.mark.parametrize("stem,threads", [("4KP10", 4), ("3AP05", 2)])
def test_aira_cli_epp_two_ranks_gloo(lib, examples, stem, threads, tmp_path):
    """world_size 2 over gloo: the strips of every EPP level are sharded over the ranks and the points
    all-gathered between levels; rank 0 writes the same .out."""
    from oracle.lpformat import parse_out
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
