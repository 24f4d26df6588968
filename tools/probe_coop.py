"""Time-to-front of the cooperative ("synergistic") workers on one GPU (run on the GPU box), next to the EPP driver.
usage: probe_coop.py ap:3:20 kp:4:30 ...   (kind:k:n[:seed]);  PROBE_EPP=T also runs moip_pool_pareto_front with T strips"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import moip_aira_b200 as mb
from moip_aira_b200 import instances
d = tempfile.mkdtemp()
for spec in sys.argv[1:]:
    parts = spec.split(":")
    kind, k, n = parts[0], int(parts[1]), int(parts[2])
    seed = int(parts[3]) if len(parts) > 3 else 1
    p = os.path.join(d, f"{kind}{k}_{n}_{seed}.lp")
    (instances.write_ap if kind == "ap" else instances.write_kp)(p, n, k, seed)
    pr = mb.Problem(p)
    T = int(os.environ.get("PROBE_EPP", "0"))
    pool = mb.WorkerPool(pr, 0, max(k, min(T, 12)))
    fronts = {}
    for w in list(range(k, 0, -1)):
        s0 = pool.stats()
        t = time.perf_counter()
        f = pool.synergistic_front(w)
        dt = time.perf_counter() - t
        s1 = pool.stats()
        fronts[w] = f
        print(f"{spec} cooperative W={w}: {dt:.3f}s front={len(f)} ips={s1['ip_solved'] - s0['ip_solved']} "
              f"nodes={s1['bb_nodes'] - s0['bb_nodes']} lps={s1['node_lps'] - s0['node_lps']} "
              f"iters/lp={(s1['lp_iterations'] - s0['lp_iterations']) / max(1, s1['node_lps'] - s0['node_lps']):.0f}", flush=True)
    if T:
        s0 = pool.stats()
        t = time.perf_counter()
        f = pool.pareto_front(T)
        dt = time.perf_counter() - t
        s1 = pool.stats()
        fronts["epp"] = f
        print(f"{spec} EPP strips={T} contexts={pool.workers}: {dt:.3f}s front={len(f)} ips={s1['ip_solved'] - s0['ip_solved']}", flush=True)
    print(f"{spec} all fronts equal: {len({tuple(v) for v in fronts.values()}) == 1}", flush=True)
    pool.close()
