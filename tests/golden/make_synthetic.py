"""Golden Pareto fronts of the synthetic instances of SURVEY.md section 8d (items 4-5) at sizes the CPU
oracle finishes in minutes.  The fronts come from oracle/aira_oracle.py (the restatement of the
reference's generator, pinned on the 7 committed .out files) driven by HiGHS (scipy.optimize.milp; NOT
CPLEX) and, for the knapsacks, cross-checked by a solver-free Pareto DP over the capacity.

    python tests/golden/make_synthetic.py        # writes tests/golden/synthetic.json (a few minutes)
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import aira_oracle as ao                      # noqa: E402
from oracle.lpformat import synthetic_ap, synthetic_kp    # noqa: E402

CASES = [("ap", 3, 8, 1), ("ap", 3, 10, 1), ("ap", 3, 12, 1), ("ap", 3, 15, 1), ("ap", 4, 7, 2), ("ap", 2, 20, 3),
         ("kp", 4, 20, 1), ("kp", 4, 25, 1), ("kp", 3, 40, 1), ("kp", 2, 80, 5)]


def kp_front_dp(w, V, cap):
    """all non-dominated value vectors of a binary knapsack (MAX), solver-free"""
    states = {0: {tuple([0] * len(V))}}          # weight -> set of non-dominated value vectors at exactly that weight
    front = {tuple([0] * len(V))}

    def prune(s):
        pts = sorted(s, reverse=True)
        keep = []
        for p in pts:
            if not any(all(q[i] >= p[i] for i in range(len(p))) and q != p for q in keep):
                keep.append(p)
        return set(keep)

    allpts = {tuple([0] * len(V))}
    cur = {(0,) + tuple([0] * len(V))}
    for j in range(len(w)):
        nxt = set(cur)
        for s in cur:
            nw = s[0] + int(w[j])
            if nw <= cap:
                nxt.add((nw,) + tuple(s[1 + i] + int(V[i][j]) for i in range(len(V))))
        # dominance pruning on (weight small, values large)
        lst = sorted(nxt, key=lambda t: (t[0],) + tuple(-v for v in t[1:]))
        keep = []
        for t in lst:
            if not any(q[0] <= t[0] and all(q[1 + i] >= t[1 + i] for i in range(len(V))) for q in keep):
                keep.append(t)
        cur = set(keep)
    return sorted(prune({t[1:] for t in cur}), key=lambda r: tuple(-v for v in r))


def main():
    out = {}
    for kind, k, n, seed in CASES:
        model = synthetic_ap(n, k, seed) if kind == "ap" else synthetic_kp(n, k, seed)
        t = time.time()
        orc = ao.MilpOracle(model)
        front = ao.pareto_front(model, orc)
        name = f"{kind}{k}_{n}_{seed}"
        if kind == "kp" and n <= 40:
            w = model.A[0]
            dp = kp_front_dp(w, model.C, int(model.b[0]))
            assert [tuple(r) for r in front] == dp, name
        out[name] = {"kind": kind, "k": k, "n": n, "seed": seed, "rows": [list(map(int, r)) for r in front]}
        print(name, len(front), "rows", f"{time.time() - t:.1f}s", flush=True)
    with open(os.path.join(os.path.dirname(__file__), "synthetic.json"), "w") as fh:
        json.dump(out, fh)


if __name__ == "__main__":
    main()
