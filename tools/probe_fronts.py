"""Time-to-front + B&B statistics on the shipped Examples (run on the GPU box)."""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import moip_aira_b200 as mb
from oracle.lpformat import parse_out
ex = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "examples.json")))
d = tempfile.mkdtemp()
which = sys.argv[1:] or list(ex)
for stem in which:
    e = ex[stem]
    p = os.path.join(d, e["file"]); open(p, "w").write(e["input"])
    ctx = mb.Context(mb.Problem(p))
    t = time.perf_counter(); f = ctx.pareto_front(); dt = time.perf_counter() - t
    s = ctx.stats()
    ok = f == parse_out(e["out"])[0]
    print(f"{stem}: {dt:.3f}s ok={ok} front={len(f)} ips={s['ip_solved']} nodes={s['bb_nodes']} lps={s['node_lps']} "
          f"iters/lp={s['lp_iterations']/max(1,s['node_lps']):.0f} launches={s['kernel_launches']} "
          f"ms/ip={1e3*dt/max(1,s['ip_solved']):.2f} nodes/ip={s['bb_nodes']/max(1,s['ip_solved']):.1f}", flush=True)
    ctx.close()
