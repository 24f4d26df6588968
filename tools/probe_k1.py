"""Perf probe for kernel K1 (run on the GPU box): LP/s and node-iterations/s for both synthetic
workloads at a few batch sizes, fixed-iteration and converged."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import moip_aira_b200 as mb
from moip_aira_b200 import instances

d = tempfile.mkdtemp()
which = sys.argv[1:] or ["ap30", "kp40"]
for name, writer, args in (("ap30", instances.write_ap, (30, 3, 1)), ("kp40", instances.write_kp, (40, 4, 1))):
    if name not in which:
        continue
    p = os.path.join(d, name + ".lp"); writer(p, *args)
    pr = mb.Problem(p); ctx = mb.Context(pr)
    bytes_it = 16 * (pr.n + pr.m) + (pr.n + 3) // 4 + 8 * pr.objcnt + 4
    for B in (1184, 8192):
        cost, rhs, masks = instances.sample_node_batch(ctx, B)
        ctx.lp_batch_upload(cost, rhs, masks)
        for label, params in (("fixed1000", ctx.lp_params(fixed_iters=1000)), ("eps1e-6", ctx.lp_params(eps=1e-6))):
            ctx.lp_batch_run(params); ctx.lp_batch_download()
            torch.cuda.synchronize(); t = time.perf_counter()
            ctx.lp_batch_run(params); r = ctx.lp_batch_download()
            dt = time.perf_counter() - t
            its = r["iters"].sum()
            print(f"{name} B={B} {label}: {dt*1e3:.2f} ms, {B/dt:.0f} LP/s, {its/dt:.3g} node-iter/s, "
                  f"roofline frac {its/dt*bytes_it/6543.4e9:.3f}, mean it {its/B:.0f}, max it {r['iters'].max()}, "
                  f"status {np.bincount(r['status'], minlength=4).tolist()}", flush=True)
    ctx.close()
