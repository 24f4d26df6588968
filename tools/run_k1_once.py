"""One warm-up + one timed launch of K1 on a fixed batch (target for ncu)."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import moip_aira_b200 as mb
from moip_aira_b200 import instances
name = sys.argv[1] if len(sys.argv) > 1 else "ap30"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
d = tempfile.mkdtemp(); p = os.path.join(d, name + ".lp")
(instances.write_ap(p, 30, 3, 1) if name == "ap30" else instances.write_kp(p, 40, 4, 1))
ctx = mb.Context(mb.Problem(p))
cost, rhs, masks = instances.sample_node_batch(ctx, B)
ctx.lp_batch_upload(cost, rhs, masks)
params = ctx.lp_params(fixed_iters=iters) if iters > 0 else ctx.lp_params(eps=1e-6, check_every=int(os.environ.get("MOIP_CHECK_EVERY", "32")))
for _ in range(2):
    ctx.lp_batch_run(params); r = ctx.lp_batch_download()
torch.cuda.synchronize(); t = time.perf_counter()
ctx.lp_batch_run(params); r = ctx.lp_batch_download()
dt = time.perf_counter() - t
print(f"{name} B={B} iters={iters}: {dt*1e3:.2f} ms, {r['iters'].sum()/dt:.3g} node-iter/s, {B/dt:.0f} LP/s, mean iters {r['iters'].mean():.0f}, "
      f"status counts {np.bincount(r['status'], minlength=4).tolist()}")
