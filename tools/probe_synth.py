"""Time-to-front + B&B statistics on synthetic AP / KP instances of growing size (run on the GPU box).
usage: [PROBE_SPLIT=T PROBE_WORKERS=W] probe_synth.py ap:3:8 ap:3:10 kp:4:20 ...   (kind:k:n[:seed])"""
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
    T, W = int(os.environ.get("PROBE_SPLIT", "0")), int(os.environ.get("PROBE_WORKERS", "1"))
    pr = mb.Problem(p)
    ctx = mb.WorkerPool(pr, 0, W) if (T and W > 1) else mb.Context(pr)
    t = time.perf_counter()
    f = ctx.pareto_front(T) if (T and W > 1) else (ctx.pareto_front(split=True, num_threads=T) if T else ctx.pareto_front())
    dt = time.perf_counter() - t
    s = ctx.stats()
    spec = f"{spec} split={T} workers={W}"
    print(f"{spec}: {dt:.3f}s front={len(f)} ips={s['ip_solved']} nodes={s['bb_nodes']} lps={s['node_lps']} "
          f"iters/lp={s['lp_iterations']/max(1,s['node_lps']):.0f} launches={s['kernel_launches']} "
          f"ms/ip={1e3*dt/max(1,s['ip_solved']):.2f} nodes/ip={s['bb_nodes']/max(1,s['ip_solved']):.1f} "
          f"solver_s={s['solver_seconds']:.2f}", flush=True)
    if os.environ.get("MOIP_KERNEL_TIMING"):
        kt = ctx.kernel_times()
        print("   kernel-class ms (CUDA events, summed over contexts): " + " ".join(f"{a}={v:.0f}" if isinstance(v, float) else f"{a}={v}" for a, v in kt.items()), flush=True)
    ctx.close()
