#!/usr/bin/env python
"""bench.py -- node-LP relaxations / s (BASELINE.json's throughput metric) on the synthetic
3-objective assignment instance n=30 (900 binaries, 63 rows), plus time-to-front on the shipped
Examples reported alongside.

A "step" is one pass of the hot path (kernel K1 behind moip_lp_batch_run) over one batch of B node LPs
(root LP + random depth-d fixing + random objective bounds + random cost index, SURVEY.md section
8d-6), each solved to relative KKT 1e-6.  `value` = node LPs per second with the batch resident in
HBM; `e2e` = the same through moip_lp_batch_solve with pinned HOST buffers (H2D of cost index / rhs /
fixing masks and D2H of objective, bound, status, iterations inside the timed region).
`roofline` is SURVEY.md section 8d's HBM figure: algorithmic bytes 16(n+m)+ceil(n/4)+8k+4 per
node-iteration x the iterations the launch executed / the launch's CUDA-event duration, against the
measured copy bandwidth in MEASURED_PEAKS.json.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--workload ap30|kp40]

N > 1: launched by torchrun, one rank per GPU; every rank solves its own batch (weak scaling, no
data-path collective), timing = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

EPS = 1e-6
WORKLOADS = {"ap30": ("synthetic 3-objective assignment n=30 (900 binaries, 63 rows), seed 1", "write_ap", (30, 3, 1)),
             "kp40": ("synthetic 4-objective knapsack n=40 (40 binaries, 5 rows), seed 1", "write_kp", (40, 4, 1))}


def algorithmic_bytes_per_node_iter(n, m, k):
    return 16 * (n + m) + math.ceil(n / 4) + 8 * k + 4


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------ reference arm
def _highs_chunk(args):
    from oracle import pdhg_oracle as po
    from oracle.lpformat import read_model
    path, cost, rhs, masks = args
    model = read_model(path)
    t = time.perf_counter()
    st, obj = po.highs_lp(model, cost, rhs, masks)
    return time.perf_counter() - t, len(cost)


def cpu_reference(path, cost, rhs, masks, sample, procs):
    """The reference solves these LPs inside CPLEX (src/aira.cpp:480), which cannot be installed here
    (BASELINE.md section 2).  Stand-ins timed on the host cores: HiGHS dual simplex via scipy (one
    process per core) and the plain-C port of K1's algorithm (oracle/pdhg_ref.c, one thread per core)."""
    import multiprocessing as mp
    from oracle import pdhg_oracle as po
    from oracle.lpformat import read_model
    sample = min(sample, len(cost))
    chunks = [(path, cost[i::procs][: max(1, sample // procs)], rhs[i::procs][: max(1, sample // procs)],
               masks[i::procs][: max(1, sample // procs)]) for i in range(procs)]
    done = sum(len(c[1]) for c in chunks)
    t = time.perf_counter()
    with mp.get_context("spawn").Pool(procs) as pool:
        pool.map(_highs_chunk, chunks)
    highs_rate = done / (time.perf_counter() - t)
    model = read_model(path)
    t = time.perf_counter()
    po.pdhg_ref(model, cost[:sample], rhs[:sample], masks[:sample], eps=EPS, threads=procs)
    port_rate = sample / (time.perf_counter() - t)
    return highs_rate, port_rate, done


# ------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--workload", default="ap30", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=256)
    ap.add_argument("--no-fronts", action="store_true")
    ap.add_argument("--front-instance", default="ap3_30_1", help="synthetic instance of tests/golden/synthetic.json whose "
                    "Pareto front is timed with the EPP strips sharded over the ranks (ap3_12_1, ap3_15_1, ap3_20_1, kp4_25_1 ...)")
    ap.add_argument("--front-strips-per-gpu", type=int, default=0,
                    help="EPP strips per GPU for the synthetic front (12 solver contexts per GPU draw them dynamically); "
                         "0 = 24 on one GPU (two strips per context even out the strips' very different sizes: 3AP n=30 "
                         "21.6 s at 12 strips, 18.2 s at 24, 17.7 s at 36, profiles/r01_front_strips.md), 12 per GPU on "
                         "several GPUs (the configuration of profiles/r01_scaling.md)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    desc, writer, wargs = WORKLOADS[args.workload]
    tmp = tempfile.mkdtemp(prefix="moip_bench_")
    path = os.path.join(tmp, f"{args.workload}.lp")
    from moip_aira_b200 import instances
    getattr(instances, writer)(path, *wargs)
    config = {"workload": f"{desc}; node-LP batch B={args.batch} per GPU (depth-d fixings d~U{{0..20}}, rhs between "
                          f"ideal and nadir, random cost index, seed 7+rank), each LP to relative KKT {EPS:g}",
              "batch_per_gpu": args.batch, "eps": EPS, "l2": "flushed between timed steps (256 MiB write)",
              "parallelism": f"{world} independent shard(s), one per GPU, no data-path collective"}
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        # needs a node batch: the generator solves k root LPs; do that with HiGHS here (CPU only arm)
        from oracle import pdhg_oracle as po
        from oracle.lpformat import read_model
        model = read_model(path)
        sample = max(cores, min(args.cpu_sample, args.batch))
        cost, rhs, masks = po.sample_node_batch(model, sample, seed=7)
        rates = []
        for _ in range(args.warmup + args.steps):
            h, p, done = cpu_reference(path, cost, rhs, masks, sample, cores)
            rates.append((h, p))
        rates = rates[args.warmup:]
        h = float(np.mean([r[0] for r in rates])); p = float(np.mean([r[1] for r in rates]))
        v = max(h, p)
        line = {"impl": "reference", "metric": "node_lp_relaxations_per_sec", "value": v, "unit": "LP/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample / v,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": v, "unit": "LP/s", "cores": cores, "kind": "port",
                                 "sample": f"{sample} node LPs of the same workload per step; best of HiGHS dual simplex "
                                           f"({h:.1f} LP/s, one process per core) and the C port of K1 ({p:.1f} LP/s, one "
                                           "thread per core); reference CPLEX unavailable offline"},
                "e2e": {"value": v, "unit": "LP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import moip_aira_b200 as mb
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream()
    pr = mb.Problem(path)
    ctx = mb.Context(pr, device=local, stream=stream.cuda_stream)
    B = args.batch
    cost, rhs, masks = instances.sample_node_batch(ctx, B, seed=7 + rank)
    params = ctx.lp_params(eps=EPS)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident path (value + roofline)
    ctx.lp_batch_upload(cost, rhs, masks)
    for _ in range(args.warmup):
        ctx.lp_batch_run(params)
        ctx.lp_batch_download()
    sampler = ClockSampler(local)
    sampler.start()
    ctx.reset_stats()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    iters_total = 0
    barrier()
    for s in range(args.steps):
        flush.fill_(s)                      # L2 flush, outside the event pair
        ev[s][0].record(stream)
        ctx.lp_batch_run(params)
        ev[s][1].record(stream)
        r = ctx.lp_batch_download()         # syncs; also gives the iteration counts for the roofline
        iters_total += int(r["iters"].sum())
    barrier()
    launches = ctx.stats()["kernel_launches"]
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    it = torch.tensor([float(iters_total)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(it, op=dist.ReduceOp.SUM)
    max_ms = float(t.item())
    # ---- the unambiguous-work figure of SURVEY.md 8d-6: every LP runs exactly 1000 iterations
    fixed = ctx.lp_params(fixed_iters=1000)
    ctx.lp_batch_run(fixed); ctx.lp_batch_download()
    fe = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    flush.fill_(1)
    fe[0].record(stream); ctx.lp_batch_run(fixed); fe[1].record(stream)
    ctx.lp_batch_download()
    fixed_ms = fe[0].elapsed_time(fe[1])
    # ---- end-to-end path through the host-buffer call (pinned host memory)
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    hc, hr, hm = pin(cost), pin(rhs), pin(masks)
    ctx.lp_batch_solve(hc, hr, hm, params)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        ctx.lp_batch_solve(hc, hr, hm, params)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    clocks = sampler.stop()
    status = np.bincount(r["status"], minlength=4)

    n, m, k = pr.n, pr.m, pr.objcnt
    bytes_iter = algorithmic_bytes_per_node_iter(n, m, k)
    peak, peak_kind = measured_peak()
    # roofline of the dominant kernel (K1) on this rank: its launches are the whole timed region
    my_iters = iters_total
    achieved = my_iters * bytes_iter / (dev_ms * 1e-3) / 1e9
    value = world * B * args.steps / (max_ms * 1e-3)
    line = {} if rank != 0 else {"metric": "node_lp_relaxations_per_sec", "value": value, "unit": "LP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_kind": peak_kind,
                         "kernel": "k1_reg_kernel<256,4,3,2,2>" if args.workload == "ap30" else "k1_small_kernel<5,5,3>",
                         "algorithmic_bytes_per_node_iter": bytes_iter,
                         "node_iters_per_launch": my_iters / args.steps,
                         "note": "iterates stay in shared memory across iterations; HBM traffic is far below the "
                                 "streaming figure this fraction is defined on (SURVEY.md 8d)"},
            "e2e": {"value": world * B * args.steps / e2e_s, "unit": "LP/s",
                    "h2d_bytes_per_step": int(hc.nbytes + hr.nbytes + hm.nbytes), "d2h_bytes_per_step": int(B * (8 + 8 + 4 + 4))},
            "gpu_launches": int(launches), "clocks": clocks,
            "lp": {"mean_iters": my_iters / (args.steps * B), "status_counts": status.tolist(),
                   "node_iters_per_sec": float(it.item()) / (max_ms * 1e-3),
                   "fixed_1000_iterations": {"ms": fixed_ms, "node_iters_per_sec": B * 1000 / (fixed_ms * 1e-3),
                                             "roofline_frac": B * 1000 * bytes_iter / (fixed_ms * 1e-3) / 1e9 / peak}}}
    tr = os.path.join(ROOT, "profiles", "k1_traffic.json")      # dram bytes per launch of this command from ncu --set full
    if rank == 0 and os.path.exists(tr) and args.workload == "ap30" and B == 65536:
        line["roofline"]["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch")
    # ---- CPU baseline beside it (rank 0, N = 1 only; bounded sample)
    if world == 1:
        sample = max(cores, min(args.cpu_sample, B))
        h, p, done = cpu_reference(path, cost, rhs, masks, sample, cores)
        line["cpu_baseline"] = {"value": max(h, p), "unit": "LP/s", "cores": cores, "kind": "port",
                                "sample": f"first {sample} node LPs of the same batch; best of HiGHS dual simplex ({h:.1f} LP/s, "
                                          f"one process per core; stand-in, NOT CPLEX) and the C port of K1 ({p:.1f} LP/s)"}
        if not args.no_fronts:
            line["time_to_front_s"] = time_to_front(mb, local, stream.cuda_stream)
            line["cpu_baseline"]["front"] = cpu_front("ap3_10_1", mb, local, stream.cuda_stream, tmp)
    print_line = rank == 0
    if not args.no_fronts:
        ctx.close()
        per_gpu = args.front_strips_per_gpu if args.front_strips_per_gpu > 0 else (24 if world == 1 else 12)
        fr = synthetic_front(args.front_instance, per_gpu * world, local, tmp)   # collective: every rank
        if print_line:
            line.setdefault("time_to_front_s", {})[f"{args.front_instance} --split -t {per_gpu * world}"] = fr
    if print_line and world == 1 and not args.no_fronts:
        try:                                       # first GPU timings of the cooperative workers; never fatal for the line
            line["time_to_front_s"]["cooperative_workers"] = coop_fronts(tmp)
        except Exception as e:                     # noqa: BLE001
            line["time_to_front_s"]["cooperative_workers"] = {"error": repr(e)[:200]}
    if print_line:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def synthetic_goldens():
    """oracle fronts of the synthetic instances: tests/golden/synthetic.json + one file per large instance"""
    import glob
    g = {}
    for f in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.json"))):
        if os.path.basename(f) != "examples.json":
            with open(f) as fh:
                g.update(json.load(fh))
    return g


def parse_instance(name):
    """'ap3_20_1' -> kind, k, n, seed"""
    kind, rest = name[:2], name[2:].split("_")
    return kind, int(rest[0]), int(rest[1]), int(rest[2])


def cpu_front(name, mb, device, stream, tmp):
    """Time-to-front side by side on one small synthetic instance: the CPU restatement of the reference's
    generator driven by HiGHS (stand-in for aira + CPLEX, which cannot be installed; one core, -t 1) against
    this library's default run on the GPU."""
    from moip_aira_b200 import instances
    from oracle import aira_oracle as ao            # checker / baseline only
    from oracle.lpformat import read_model
    g = synthetic_goldens()[name]
    path = os.path.join(tmp, name + "_cpu.lp")
    (instances.write_ap if g["kind"] == "ap" else instances.write_kp)(path, g["n"], g["k"], g["seed"])
    model = read_model(path)
    t = time.perf_counter()
    cpu_rows = ao.pareto_front(model, ao.MilpOracle(model))
    cpu_s = time.perf_counter() - t
    ctx = mb.Context(mb.Problem(path), device=device, stream=stream)
    t = time.perf_counter()
    gpu_rows = ctx.pareto_front()
    gpu_s = time.perf_counter() - t
    ctx.close()
    want = [tuple(r) for r in g["rows"]]
    return {"instance": name, "cpu_seconds": cpu_s, "cpu_kind": "port (oracle generator + HiGHS milp, 1 core, -t 1; NOT CPLEX)",
            "gpu_seconds": gpu_s, "front": len(want), "cpu_matches_golden": [tuple(r) for r in cpu_rows] == want,
            "gpu_matches_golden": gpu_rows == want}


def synthetic_front(name, strips, device, tmp):
    """Pareto front of a synthetic instance with the EPP strips of every level sharded over the ranks
    (one rank per GPU, NCCL all-gather of the points between levels) and solved concurrently inside a rank;
    checked against the committed oracle front.  Time = max over ranks."""
    import torch
    import torch.distributed as dist
    from moip_aira_b200 import aira, instances
    g = synthetic_goldens().get(name)            # None: no oracle front committed for this instance
    kind, k, n, seed = parse_instance(name)
    path = os.path.join(tmp, name + ".lp")
    (instances.write_ap if kind == "ap" else instances.write_kp)(path, n, k, seed)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    d = aira.Dist(torch.device("cuda", device) if world > 1 else None)
    be = aira.GpuBackend(path, device=device)
    d.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    front = aira.epp_front(be, d, strips, False)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device="cuda")
    ips = torch.tensor([float(be.ip_count())], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ips, op=dist.ReduceOp.SUM)
    return {"seconds": float(dt.item()), "front": len(front),
            "matches_golden": ([list(r) for r in front] == g["rows"]) if g else None,
            "ips": int(ips.item()), "workers_per_gpu": be.workers, "n_gpus": world}


_COOP_SCRIPT = r"""
import json, sys, time
sys.path.insert(0, sys.argv[1])
import moip_aira_b200 as mb
from moip_aira_b200 import instances
kind, k, n, seed, path = sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
(instances.write_ap if kind == "ap" else instances.write_kp)(path, n, k, seed)
pr = mb.Problem(path)
pool = mb.WorkerPool(pr, 0, k)
pool.synergistic_front(1)                           # warm-up: CUDA start-up, module load, root LPs
out = {}
for w in range(1, k + 1):
    s0 = pool.stats()
    t = time.perf_counter()
    front = pool.synergistic_front(w)
    dt = time.perf_counter() - t
    s1 = pool.stats()
    out[str(w)] = {"seconds": dt, "front": [list(r) for r in front], "ips": s1["ip_solved"] - s0["ip_solved"],
                   "node_lps": s1["node_lps"] - s0["node_lps"]}
pool.close()
print("COOP " + json.dumps(out))
"""


def coop_fronts(tmp, name="ap3_15_1", timeout_s=150):
    """Cooperative ("synergistic") workers, -t W without --split: time-to-front of a synthetic instance
    with W = 1..k workers on one GPU, each front checked against the committed golden.  Runs in a child process with a
    time limit (the path is new on the GPU), so that nothing it does can cost the bench line."""
    import subprocess
    g = synthetic_goldens()[name]
    kind, k, n, seed = parse_instance(name)
    path = os.path.join(tmp, name + "_coop.lp")
    try:
        r = subprocess.run([sys.executable, "-c", _COOP_SCRIPT, ROOT, kind, str(k), str(n), str(seed), path],
                           capture_output=True, text=True, timeout=timeout_s)
    except subprocess.TimeoutExpired:
        return {"instance": name, "timeout_s": timeout_s}
    lines = [l for l in r.stdout.splitlines() if l.startswith("COOP ")]
    if r.returncode != 0 or not lines:
        return {"instance": name, "rc": r.returncode, "stderr": r.stderr[-300:]}
    res = json.loads(lines[-1][5:])
    out = {"instance": name, "golden_points": len(g["rows"])}
    for w, d in sorted(res.items()):
        out["-t " + w] = {"seconds": d["seconds"], "matches_golden": d["front"] == [list(x) for x in g["rows"]],
                          "ips": d["ips"], "node_lps": d["node_lps"]}
    return out


def time_to_front(mb, device, stream):
    """Second half of BASELINE.json's metric: Pareto-front time-to-solve on the shipped Examples
    (configs[0..2]); every front is checked against the committed golden .out."""
    from oracle.lpformat import parse_out          # checker only
    out = {}
    with open(os.path.join(ROOT, "tests", "golden", "examples.json")) as fh:
        ex = json.load(fh)
    tmp = tempfile.mkdtemp(prefix="moip_front_")
    for stem in ("2AP05", "3KP10", "3AP05", "4KP10", "4AP05"):
        e = ex[stem]
        p = os.path.join(tmp, e["file"])
        with open(p, "w") as fh:
            fh.write(e["input"])
        ctx = mb.Context(mb.Problem(p), device=device, stream=stream)
        t = time.perf_counter()
        front = ctx.pareto_front()
        dt = time.perf_counter() - t
        st = ctx.stats()
        out[stem] = {"seconds": dt, "front": len(front), "matches_golden": front == parse_out(e["out"])[0],
                     "ips": st["ip_solved"], "node_lps": st["node_lps"]}
        ctx.close()
    return out


if __name__ == "__main__":
    sys.exit(main())
