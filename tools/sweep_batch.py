"""Node-LP throughput sweep of SURVEY.md section 8d item 6 / BASELINE.json configs[4]: batch B = 1 ... 65 536 node
LPs of synthetic 3AP n=30 and 4KP n=40, fixed 1000 iterations (the roofline figure) and converged to relative KKT
1e-6 (LP/s).  CUDA events on the launching stream, L2 flushed before every timed launch.  Prints one JSON object per
(workload, B); run on the GPU box:  python tools/sweep_batch.py > gpurun_out/sweep.jsonl"""
import json, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import moip_aira_b200 as mb
from moip_aira_b200 import instances

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
d = tempfile.mkdtemp()
stream = torch.cuda.current_stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, writer, args in (("ap30", instances.write_ap, (30, 3, 1)), ("kp40", instances.write_kp, (40, 4, 1))):
    p = os.path.join(d, name + ".lp"); writer(p, *args)
    pr = mb.Problem(p); ctx = mb.Context(pr, stream=stream.cuda_stream)
    bytes_it = 16 * (pr.n + pr.m) + (pr.n + 3) // 4 + 8 * pr.objcnt + 4
    cost, rhs, masks = instances.sample_node_batch(ctx, 65536)
    B = 1
    while B <= 65536:
        ctx.lp_batch_upload(cost[:B], rhs[:B], masks[:B])
        row = {"workload": name, "B": B, "bytes_per_node_iter": bytes_it}
        for label, params in (("fixed1000", ctx.lp_params(fixed_iters=1000)), ("eps1e-6", ctx.lp_params(eps=1e-6))):
            ctx.lp_batch_run(params); ctx.lp_batch_download()
            best = None
            for rep in range(3):
                flush.fill_(rep)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); ctx.lp_batch_run(params); e1.record(stream)
                r = ctx.lp_batch_download()
                ms = e0.elapsed_time(e1)
                best = ms if best is None else min(best, ms)
            its = float(r["iters"].sum())
            row[label] = {"ms": round(best, 4), "lp_per_s": round(B / (best * 1e-3), 1), "node_iter_per_s": its / (best * 1e-3),
                          "roofline_frac": round(its * bytes_it / (best * 1e-3) / 1e9 / PEAK, 4), "mean_iters": round(its / B, 1)}
        print(json.dumps(row), flush=True)
        B *= 2
    ctx.close()
