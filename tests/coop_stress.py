"""TEST UTILITY (not collected by pytest): randomised stress run of the cooperative-worker protocol against a brute-force
Pareto filter -- random small KP/AP instances, 2..k workers, jittered worker speeds.  usage: python tests/coop_stress.py SEED SECONDS"""
import sys, time, os, tempfile, random, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import moip_aira_b200 as lib
from moip_aira_b200 import instances
from oracle import aira_oracle as ao
from oracle.lpformat import read_model
import test_synergistic as ts
d = tempfile.mkdtemp()
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
def coop_jitter(model, fs, W, jit):
    k = model.k
    lock = threading.Lock(); inf = ao.Solutions(k); sols = [ao.Solutions(k) for _ in range(W)]
    speeds = [rng.choice([0.0, 0.0005, 0.002, 0.006]) for _ in range(W)]
    def solve(w, perm, n_obj, rhs):
        r = fs.lex_solve(perm, n_obj, rhs)
        if jit: time.sleep(speeds[w] * rng.random())
        return r
    def find(w, ip):
        with lock: _, r = inf.find(ip, model.sense)
        if r is None: _, r = sols[w].find(ip, model.sense)
        return None if r is None else (r.infeasible, r.result)
    def insert(w, ip, res, infeasible):
        if infeasible:
            with lock: inf.insert(ip, res, True)
        else: sols[w].insert(ip, res, False)
    solves, skipped = lib.coop_optimise_with(k, 0 if model.sense == "MIN" else 1, W, solve, find, insert)
    return sorted({tuple(r.result) for s in sols for r in s.store if not r.infeasible}, reverse=True), solves, skipped
bad = 0; runs = 0; t0 = time.time()
while time.time() - t0 < float(sys.argv[2]) if len(sys.argv) > 2 else 200:
    kind = rng.choice(["kp", "ap"]); k = rng.choice([2, 3, 3, 4])
    n = rng.randint(6, 13) if kind == "kp" else rng.randint(3, 5 if k < 4 else 4)
    seed = rng.randint(100, 10**6)
    path = os.path.join(d, f"{kind}{k}_{n}_{seed}.lp")
    (instances.write_ap if kind=="ap" else instances.write_kp)(path, n, k, seed)
    m = read_model(path); fs = ao.FeasibleSet(m)
    want = ts._nondominated(fs.P, m.sense=="MIN")
    for W in range(2, k+1):
        for jit in (0, 1, 1):
            pts, solves, skipped = coop_jitter(m, fs, W, jit); runs += 1
            if pts != want:
                bad += 1
                print("BAD", kind, k, n, seed, W, jit, len(pts), len(want), solves, skipped, flush=True)
print("runs", runs, "bad", bad)
