"""Golden front of the headline workload (synthetic 3-objective assignment n=30, seed 1),
computed on the CPU by the restatement of the reference's EPP driver (oracle/aira_oracle.py, src/aira.cpp:1886-1990)
with HiGHS (scipy.optimize.milp; NOT CPLEX) as the IP solver.  The strips of each level are independent
(private caches), so they are farmed out to one process per core.

    python tests/golden/make_ap30.py [n] [k] [seed] [strips] [ap|kp]    # 88 min on 8 cores for 3AP n=30
writes tests/golden/<kind><k>_<n>_<seed>.json"""
import json
import multiprocessing as mp
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import aira_oracle as ao                      # noqa: E402
from oracle.lpformat import synthetic_ap, synthetic_kp     # noqa: E402


def make_model(kind, n, k, seed):
    return synthetic_ap(n, k, seed) if kind == "ap" else synthetic_kp(n, k, seed)


def strip(args):
    kind, n, k, seed, n_obj, a, b = args
    model = make_model(kind, n, k, seed)
    orc = ao.MilpOracle(model)
    mine, minf = ao.Solutions(k), ao.Solutions(k)
    ao.optimise(model, orc, mine, minf, ao.Worker(range(k), n_obj, a, b), split=True)
    return [list(map(int, r.result)) for r in mine.store if not r.infeasible]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    strips = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    kind = sys.argv[5] if len(sys.argv) > 5 else "ap"
    model = make_model(kind, n, k, seed)
    MIN = model.sense == "MIN"
    orc = ao.MilpOracle(model)
    free = [ao.INF if MIN else -ao.INF] * k
    t0 = time.time()
    sols = [orc.get_limit(0, free)]
    with mp.get_context("spawn").Pool(os.cpu_count()) as pool:
        for n_obj in range(2, k + 1):
            res = orc.get_limit(n_obj - 1, free)
            if MIN:
                smallest = res[n_obj - 1]
                biggest = max([ao.INT_MIN] + [s[n_obj - 1] for s in sols])
                if biggest == smallest:
                    biggest = ao.INT_MAX
            else:
                biggest = res[n_obj - 1]
                smallest = min([ao.INT_MAX] + [s[n_obj - 1] for s in sols])
                if biggest == smallest:
                    smallest = ao.INT_MIN
            bounds = ao.strip_bounds(MIN, biggest, smallest, strips, False)
            parts = pool.map(strip, [(kind, n, k, seed, n_obj, a, b) for a, b in bounds], chunksize=1)
            sols = [r for part in parts for r in part]
            print(f"level {n_obj}: {len(sols)} rows, {time.time() - t0:.0f}s", flush=True)
    store = ao.Solutions(k)
    for r in sols:
        store.insert([0.0] * k, r, False)
    store.sort_unique()
    rows = [list(map(int, r.result)) for r in store.store if not r.infeasible]
    name = f"{kind}{k}_{n}_{seed}"
    with open(os.path.join(os.path.dirname(__file__), name + ".json"), "w") as fh:
        json.dump({name: {"kind": kind, "k": k, "n": n, "seed": seed, "rows": rows}}, fh)
    print(name, len(rows), "rows", f"{time.time() - t0:.0f}s")


if __name__ == "__main__":
    main()
