"""Random boxes (EPP strips x windows on objective 1, moip_worker::window) against a brute-force Pareto filter, on the CPU:
the host generator (moip_optimise_with) with the oracle's exact enumeration solver behind it.  usage: fuzz_boxes.py <runs>
Round 2: 150 + 80 runs, no mismatch (the committed test is tests/test_host.py::test_generator_boxes_enumerate_the_front)."""
import sys, os, random, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import moip_aira_b200 as lib
from moip_aira_b200 import instances
from oracle import aira_oracle as ao
from oracle.lpformat import read_model
import test_host as th
rnd = random.Random(5)
tmp = tempfile.mkdtemp()
bad = 0; total = 0; ips_plain = 0; ips_box = 0
for it in range(int(sys.argv[1])):
    kind = rnd.choice(["kp", "ap"]); k = rnd.choice([3, 4])
    n = rnd.choice([9, 10, 11, 12, 13]) if kind == "kp" else (4 if k == 3 else 3)
    seed = 1000 + it
    path = os.path.join(tmp, f"{kind}{k}_{n}_{seed}.lp")
    (instances.write_ap if kind == "ap" else instances.write_kp)(path, n, k, seed)
    m = read_model(path); fs = ao.FeasibleSet(m); is_min = m.sense == "MIN"
    want = th._nondominated(fs.P, is_min)
    strips = rnd.choice([1, 2, 3, 5]); wins = rnd.choice([2, 3, 4, 6]); shared = rnd.random() < 0.7
    last = [p[k - 1] for p in want]; w1 = [p[1] for p in want]
    ss = lib.split_strips(0 if is_min else 1, max(last), min(last), strips, False)
    found = set(); stores = (ao.Solutions(k), ao.Solutions(k)); cnt = [0]
    for t in range(strips):
        for win in th._windows(min(w1), max(w1), wins, is_min):
            s, inf = stores if shared else (ao.Solutions(k), ao.Solutions(k))
            def find(ip):
                _, r = inf.find(ip, m.sense)
                if r is None: _, r = s.find(ip, m.sense)
                return None if r is None else (r.infeasible, r.result)
            def insert(ip, res, infeasible): (inf if infeasible else s).insert(ip, res, infeasible)
            def solve(perm, n_obj, rhs):
                cnt[0] += 1
                return fs.lex_solve(perm, n_obj, rhs)
            w = lib.make_worker(k, split=True, split_start=ss[t][0], split_stop=ss[t][1], wid=t, window=win)
            lib.optimise_with(k, 0 if is_min else 1, w, solve, find, insert)
            found |= {tuple(r.result) for r in s.store if not r.infeasible}
    total += 1
    ok = sorted(found, reverse=True) == want
    if not ok:
        bad += 1
        print("MISMATCH", kind, k, n, seed, strips, wins, shared, len(want), len(found), flush=True)
print("runs", total, "bad", bad)
