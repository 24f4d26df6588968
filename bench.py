#!/usr/bin/env python
"""bench.py -- BASELINE.json's two-part metric on one node of B200s.

(1) Node-LP relaxations / s (the line's `value`): a "step" is one pass of the hot path (kernel K1 behind
    moip_lp_batch_run) over one batch of B node LPs of the synthetic 3-objective assignment instance n=30 (900 binaries,
    63 rows), each solved to relative KKT 1e-6.  `value` = node LPs per second with the batch resident in HBM; `e2e` = the
    same through moip_lp_batch_solve with pinned HOST buffers (H2D of cost index / rhs / fixing masks and D2H of objective,
    bound, status, iterations inside the timed region).  `roofline` is SURVEY.md section 8d's HBM figure (algorithmic bytes
    16(n+m)+ceil(n/4)+8k+4 per node-iteration x the iterations the launch executed / the launch's CUDA-event duration,
    against the measured copy bandwidth in MEASURED_PEAKS.json) plus the resource that actually binds, the fp64 pipe.
(2) Pareto-front time-to-solve (`time_to_front_s`): the shipped Examples, synthetic 3AP n=30 and 4KP n=40 with the top
    EPP level cut into boxes (strips of the last objective x windows on objective 1; four per solver context)
    sharded over the N ranks (strong scaling), the cooperative ("synergistic") workers one per rank, and the Examples'
    `--split -t 8`; every front is checked against its committed golden.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--workload ap30|kp40] [--no-fronts]

N > 1: launched by torchrun, one rank per GPU; every rank solves its own node-LP batch (weak scaling, no data-path
collective), timing = max over ranks; the fronts are ONE job over all ranks (strong scaling).
`--impl reference`: the CPU stand-ins for the reference's CPLEX path on all host cores (CPLEX cannot be installed here,
BASELINE.md section 2) -- same config, metric and unit; rank 0 only.
"""
from __future__ import annotations

import argparse
import importlib.util
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
# fp64 warp-instructions of one K1 node-iteration (ncu, profiles/r01_k1_reg_summary.md): 2 issue cycles each on one of the
# 4 sub-partitions of an SM  =>  cycles per node-iteration and SM when the fp64 pipe is the only limit
FP64_WARP_INSTR_PER_NODE_ITER = {"ap30": 1047}
CPU_FRONT_INSTANCE = "ap3_12_1"


def algorithmic_bytes_per_node_iter(n, m, k):
    return 16 * (n + m) + math.ceil(n / 4) + 8 * k + 4


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def load_instances():
    """moip_aira_b200/instances.py by file path: the workload writers need numpy only, and importing the package would map
    libmoip_b200.so into the process -- the reference arm must not."""
    spec = importlib.util.spec_from_file_location("moip_instances", os.path.join(ROOT, "moip_aira_b200", "instances.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


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


# ------------------------------------------------------------------------------------ CPU legs (stand-ins, NOT CPLEX)
_HIGHS_MODEL = None


def _highs_init(path):
    """worker start-up, outside every timed region: parse the model, import scipy"""
    global _HIGHS_MODEL
    from oracle.lpformat import read_model
    from scipy.optimize import linprog  # noqa: F401
    _HIGHS_MODEL = read_model(path)


def _highs_chunk(args):
    from oracle import pdhg_oracle as po
    cost, rhs, masks = args
    t = time.perf_counter()
    if len(cost):
        po.highs_lp(_HIGHS_MODEL, cost, rhs, masks)
    return time.perf_counter() - t, len(cost)


class CpuLp:
    """The reference solves these LPs inside CPLEX (src/aira.cpp:480), which cannot be installed here (BASELINE.md
    section 2).  Stand-ins on the host cores: HiGHS dual simplex via scipy (one worker process per core, created and
    warmed ONCE, outside the clock) and the plain-C port of K1's algorithm (oracle/pdhg_ref.c, one thread per core)."""

    def __init__(self, path, procs):
        import multiprocessing as mp
        from oracle.lpformat import read_model
        self.procs = procs
        self.model = read_model(path)
        self.pool = mp.get_context("spawn").Pool(procs, initializer=_highs_init, initargs=(path,))
        z = np.zeros(0, dtype=np.int32)
        self.pool.map(_highs_chunk, [(z, z, z)] * (2 * procs))            # every worker is up before anything is timed

    def close(self):
        self.pool.close()
        self.pool.join()

    def rates(self, cost, rhs, masks, sample):
        from oracle import pdhg_oracle as po
        procs = self.procs
        sample = min(sample, len(cost))
        per = max(1, sample // procs)
        chunks = [(cost[i::procs][:per], rhs[i::procs][:per], masks[i::procs][:per]) for i in range(procs)]
        done = sum(len(c[0]) for c in chunks)
        self.pool.map(_highs_chunk, [(c[0][:1], c[1][:1], c[2][:1]) for c in chunks])     # warm (first LP of each chunk)
        t = time.perf_counter()
        res = self.pool.map(_highs_chunk, chunks, chunksize=1)
        wall = time.perf_counter() - t
        per_core = [cnt / sec for sec, cnt in res if sec > 0]
        t = time.perf_counter()
        po.pdhg_ref(self.model, cost[:sample], rhs[:sample], masks[:sample], eps=EPS, threads=procs)
        port_rate = sample / (time.perf_counter() - t)
        return {"highs_lp_per_s": float(sum(per_core)),                  # sum of the workers' own rates (solve time only)
                "highs_lp_per_s_per_core": float(np.mean(per_core)) if per_core else 0.0,
                "highs_lp_per_s_wall": done / wall,                      # wall clock around the map, dispatch included
                "highs_pool_overhead_s": wall - max(sec for sec, _ in res),
                "port_lp_per_s": port_rate, "highs_sample": done, "port_sample": sample}


def cpu_front_leg(tmp, cores):
    """Time-to-front on the host cores: restated generator + HiGHS milp, EPP with one strip per core (oracle/cpu_front.py)."""
    from oracle import cpu_front
    inst = load_instances()
    g = synthetic_goldens()[CPU_FRONT_INSTANCE]
    path = os.path.join(tmp, CPU_FRONT_INSTANCE + "_cpu.lp")
    (inst.write_ap if g["kind"] == "ap" else inst.write_kp)(path, g["n"], g["k"], g["seed"])
    pool = cpu_front.make_pool(path, cores)                 # workers up and warm before the clock starts
    t = time.perf_counter()
    front, ips, level_s = cpu_front.epp_front(path, cores, pool)
    sec = time.perf_counter() - t
    pool.close()
    pool.join()
    return {"instance": CPU_FRONT_INSTANCE, "seconds": sec, "ips": ips, "front": len(front), "cores": cores,
            "matches_golden": [list(r) for r in front] == g["rows"], "level_seconds": level_s,
            "kind": f"port (oracle generator + HiGHS milp, --split -t {cores}: one strip per core; NOT CPLEX)"}, path


def sampler_note():
    return ("depth-d fixings d~U{0..20} (assignment models: consistent with a random permutation, so most nodes stay "
            "feasible -- SURVEY 8d-6 says uniform {0,1}), rhs drawn in [ideal + 0.3 (nadir - ideal), nadir] of every "
            "bounded objective (SURVEY 8d-6: [ideal, nadir]), random cost index, seed 7+rank")


# ------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--workload", default="ap30", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=1024)
    ap.add_argument("--no-fronts", action="store_true")
    ap.add_argument("--front-instances", default="ap3_30_1,kp4_40_1",
                    help="synthetic instances (tests/golden) whose Pareto fronts are timed with the EPP strips sharded over the ranks")
    ap.add_argument("--front-strips-per-gpu", type=int, default=0,
                    help="boxes per GPU (0 = four per solver context): EPP strips of the last objective x windows on objective 1 "
                         "(aira.windows_for); idle workers cut busy boxes")
    ap.add_argument("--front-budget-s", type=float, default=600.0,
                    help="wall-clock budget of the time-to-front section: when it runs out the line is printed with what is there")
    ap.add_argument("--syn-instances", default="ap3_30_1,kp4_40_1",
                    help="instances for the cooperative (synergistic) workers, one per rank (N=1: min(k, contexts) on one GPU)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    desc, writer, wargs = WORKLOADS[args.workload]
    tmp = tempfile.mkdtemp(prefix="moip_bench_")
    path = os.path.join(tmp, f"{args.workload}.lp")
    inst = load_instances()
    getattr(inst, writer)(path, *wargs)
    config = {"workload": f"{desc}; node-LP batch B={args.batch} per GPU ({sampler_note()}), each LP to relative KKT {EPS:g}",
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
        cpu = CpuLp(path, cores)
        rates = [cpu.rates(cost, rhs, masks, sample) for _ in range(args.warmup + args.steps)][args.warmup:]
        cpu.close()
        h = float(np.mean([r["highs_lp_per_s"] for r in rates]))
        hw = float(np.mean([r["highs_lp_per_s_wall"] for r in rates]))
        p = float(np.mean([r["port_lp_per_s"] for r in rates]))
        v = max(h, p)
        line = {"impl": "reference", "metric": "node_lp_relaxations_per_sec", "value": v, "unit": "LP/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample / v,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": v, "unit": "LP/s", "cores": cores, "kind": "port",
                                 "sample": f"{sample} node LPs of the same workload per step; best of HiGHS dual simplex "
                                           f"({h:.1f} LP/s = sum of {cores} warmed worker processes' own rates, "
                                           f"{hw:.1f} LP/s by the wall clock around the dispatch) and the C port of K1 "
                                           f"({p:.1f} LP/s, {cores} threads); reference CPLEX unavailable offline",
                                 "detail": rates[-1]},
                "e2e": {"value": v, "unit": "LP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        if not args.no_fronts:
            line["time_to_front_s"] = {"cpu": cpu_front_leg(tmp, cores)[0]}
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
    cost, rhs, masks = inst.sample_node_batch(ctx, B, seed=7 + rank)
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
    fixed_rate = B * 1000 / (fixed_ms * 1e-3)
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    sm_hz = 1e6 * (clocks.get("sm_mhz") or 1965.0)
    fp64 = None
    if args.workload in FP64_WARP_INSTR_PER_NODE_ITER:
        cyc = FP64_WARP_INSTR_PER_NODE_ITER[args.workload] * 2 / 4          # cycles per node-iteration and SM, fp64 pipe only
        fp64 = {"fp64_warp_instr_per_node_iter": FP64_WARP_INSTR_PER_NODE_ITER[args.workload],
                "bound_node_iters_per_sec": sms * sm_hz / cyc,
                "frac_converged": (my_iters / (dev_ms * 1e-3)) / (sms * sm_hz / cyc),
                "frac_fixed_1000": fixed_rate / (sms * sm_hz / cyc),
                "note": "the binding on-chip resource: fp64 warp-instructions (ncu, profiles/) x 2 issue cycles / 4 "
                        "sub-partitions per SM at the SM clock sampled during this run"}
    line = {} if rank != 0 else {"metric": "node_lp_relaxations_per_sec", "value": value, "unit": "LP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_kind": peak_kind,
                         "kernel": "k1_reg_kernel<256,4,3,2,2>" if args.workload == "ap30" else "k1_small_kernel<5,5,3>",
                         "algorithmic_bytes_per_node_iter": bytes_iter,
                         "node_iters_per_launch": my_iters / args.steps,
                         "streaming_equivalent": True, "fp64_pipe": fp64,
                         "note": "achieved/frac are the STREAMING-EQUIVALENT figure SURVEY.md 8d defines (iterations x "
                                 "algorithmic bytes); the iterates stay in registers / shared memory, so the DRAM traffic "
                                 "(`traffic`, one ncu --set full capture of this command, per launch) is ~1e-4 of it and "
                                 "the kernel is bound by the fp64 pipe (`fp64_pipe`), not by HBM"},
            "e2e": {"value": world * B * args.steps / e2e_s, "unit": "LP/s",
                    "h2d_bytes_per_step": int(hc.nbytes + hr.nbytes + hm.nbytes), "d2h_bytes_per_step": int(B * (8 + 8 + 4 + 4))},
            "gpu_launches": int(launches), "clocks": clocks,
            "lp": {"mean_iters": my_iters / (args.steps * B), "status_counts": status.tolist(),
                   "node_iters_per_sec": float(it.item()) / (max_ms * 1e-3),
                   "fixed_1000_iterations": {"ms": fixed_ms, "node_iters_per_sec": fixed_rate,
                                             "roofline_frac": B * 1000 * bytes_iter / (fixed_ms * 1e-3) / 1e9 / peak}}}
    tr = os.path.join(ROOT, "profiles", "k1_traffic.json")      # dram bytes per launch of this command from ncu --set full
    if rank == 0 and os.path.exists(tr):
        tj = json.load(open(tr))
        if tj.get("workload", "ap30") == args.workload and int(tj.get("batch", 65536)) == B:
            line["roofline"]["traffic"] = tj.get("dram_bytes_per_launch")
            line["roofline"]["traffic_source"] = tj.get("source")
    ctx.close()
    # ---- CPU baseline beside it (rank 0, N = 1 only; bounded sample)
    cpu_path = None
    if world == 1:
        sample = max(cores, min(args.cpu_sample, B))
        cpu = CpuLp(path, cores)
        rt = cpu.rates(cost, rhs, masks, sample)
        cpu.close()
        line["cpu_baseline"] = {"value": max(rt["highs_lp_per_s"], rt["port_lp_per_s"]), "unit": "LP/s", "cores": cores, "kind": "port",
                                "sample": f"first {sample} node LPs of the same batch; best of HiGHS dual simplex "
                                          f"({rt['highs_lp_per_s']:.1f} LP/s, {cores} warmed worker processes; stand-in, NOT "
                                          f"CPLEX) and the C port of K1 ({rt['port_lp_per_s']:.1f} LP/s, {cores} threads)",
                                "detail": rt}
        if not args.no_fronts:
            line["cpu_baseline"]["front"], cpu_path = cpu_front_leg(tmp, cores)
    # ---- time-to-front (every rank takes part in the sharded runs).  A front job is one collective over all ranks and
    # cannot be cancelled half way, so the section runs against a wall-clock budget: when it is spent, rank 0 prints the
    # line with the entries finished so far (the unfinished one is named) and every rank leaves.
    ttf = {}

    def out_of_time():
        if rank == 0:
            ttf["unfinished"] = f"time-to-front budget of {args.front_budget_s:.0f} s spent"
            print(json.dumps(line), flush=True)
        os._exit(0)
    watchdog = threading.Timer(args.front_budget_s, out_of_time)
    watchdog.daemon = True
    if not args.no_fronts:
        watchdog.start()
        if rank == 0:
            line["time_to_front_s"] = ttf
        if world == 1:
            ttf["examples"] = example_fronts(mb, local)
            g = time_gpu_front(cpu_path, CPU_FRONT_INSTANCE, local)         # the CPU front leg's instance on the GPU, same mode
            line["cpu_baseline"]["front"]["gpu_seconds"] = g["seconds"]
            line["cpu_baseline"]["front"]["gpu_matches_golden"] = g["matches_golden"]
        ttf["examples --split -t 8"] = {stem: example_epp(stem, 8, local, tmp) for stem in ("4AP05", "4KP10")}
        for name in [x for x in args.front_instances.split(",") if x]:
            from moip_aira_b200 import aira
            per_gpu = args.front_strips_per_gpu if args.front_strips_per_gpu > 0 else 4 * aira.default_workers()
            ttf[f"{name} --split -t {per_gpu * world}"] = synthetic_front(name, per_gpu * world, local, tmp)
        for name in [x for x in args.syn_instances.split(",") if x]:
            ttf[f"{name} synergistic"] = synergistic(name, local, tmp)
    watchdog.cancel()
    if rank == 0:
        print(json.dumps(line), flush=True)
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


def write_instance(name, tmp):
    inst = load_instances()
    kind, k, n, seed = parse_instance(name)
    path = os.path.join(tmp, name + ".lp")
    if not os.path.exists(path):
        (inst.write_ap if kind == "ap" else inst.write_kp)(path, n, k, seed)
    return path


def _gather(obj):
    """every rank's object, in rank order (world 1: [obj])"""
    import torch.distributed as dist
    if int(os.environ.get("WORLD_SIZE", "1")) == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def _front_job(path, golden_rows, device, run):
    """One front job over all ranks: run(backend, dist) -> (front, extra); wall time = max over ranks; per-rank busy time,
    IPs and device-side kernel-class split gathered for the critical-path view."""
    import torch
    from moip_aira_b200 import aira
    world = int(os.environ.get("WORLD_SIZE", "1"))
    d = aira.Dist(torch.device("cuda", device) if world > 1 else None)
    be = aira.GpuBackend(path, device=device)
    be.get_limit(0, [1e20 if be.sense == 0 else -1e20] * be.k)           # warm-up: CUDA modules, first-touch allocations
    ips0 = be.ip_count()
    d.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    front, extra = run(be, d)
    torch.cuda.synchronize()
    mine = time.perf_counter() - t
    d.barrier()
    wall = time.perf_counter() - t
    st = be.stats()
    per_rank = _gather({"rank": d.rank, "busy_s": round(mine, 3), "ips": int(be.ip_count() - ips0),
                        "node_lps": int(st.get("node_lps", 0)), "bb_nodes": int(st.get("bb_nodes", 0)),
                        "strips_cut_by_idle_workers": be.pool.strips_stolen() if be._pool is not None else 0,
                        "boxes_postponed": be.pool.boxes_postponed() if be._pool is not None else 0,
                        "solver_s": round(float(st.get("solver_seconds", 0.0)), 2),
                        "kernel_ms": ({a: round(v) for a, v in be.pool.kernel_times().items()} if os.environ.get("MOIP_KERNEL_TIMING") and be._pool is not None else None),
                        **(extra or {})})
    walls = _gather(wall)
    return {"seconds": max(walls), "front": len(front),
            "matches_golden": ([list(r) for r in front] == golden_rows) if golden_rows is not None else None,
            "ips": sum(p["ips"] for p in per_rank), "node_lps": sum(p["node_lps"] for p in per_rank),
            "n_gpus": world, "workers_per_gpu": be.workers, "per_rank": per_rank}


def synthetic_front(name, strips, device, tmp):
    """Pareto front of a synthetic instance with the EPP strips of every level sharded over the ranks (one rank per GPU:
    one job-wide strip counter, cache records all-gathered over NCCL while the strips run, points all-gathered between
    levels) and solved concurrently inside a rank; checked against the committed oracle front."""
    from moip_aira_b200 import aira
    g = synthetic_goldens().get(name)

    def run(be, d):
        stats = []
        front = aira.epp_front(be, d, strips, False, stats)
        return front, {"levels": stats}
    out = _front_job(write_instance(name, tmp), g["rows"] if g else None, device, run)
    out["strips"] = strips
    return out


def synergistic(name, device, tmp):
    """`-t W` without --split: the cooperative workers, one per rank when there are several ranks (the
    W published limits travel through the job's store, the points are all-gathered at the end), else min(k, contexts)
    workers on this GPU's pool.  At most k workers are active: one owner per objective."""
    from moip_aira_b200 import aira
    g = synthetic_goldens().get(name)
    world = int(os.environ.get("WORLD_SIZE", "1"))

    def run(be, d):
        if world > 1:
            return aira.synergistic_front(be, d), {"mode": "one worker per rank"}
        return be.synergistic_local(be.k), {"mode": f"{be.k} workers on one GPU"}
    out = _front_job(write_instance(name, tmp), g["rows"] if g else None, device, run)
    out["active_workers"] = min(world, parse_instance(name)[1]) if world > 1 else parse_instance(name)[1]
    return out


def time_gpu_front(path, name, device):
    """the CPU front leg's instance on one GPU in the same mode (--split, 24 strips)"""
    from moip_aira_b200 import aira
    g = synthetic_goldens()[name]
    return _front_job(path, g["rows"], device, lambda be, d: (aira.epp_front(be, d, 24, False), None))


def example_epp(stem, threads, device, tmp):
    """BASELINE.json configs[2]: `--split -t 8` on a shipped 4-objective Example, strips sharded over the ranks"""
    from moip_aira_b200 import aira
    from oracle.lpformat import parse_out          # checker only
    with open(os.path.join(ROOT, "tests", "golden", "examples.json")) as fh:
        e = json.load(fh)[stem]
    p = os.path.join(tmp, e["file"])
    if not os.path.exists(p):
        with open(p, "w") as fh:
            fh.write(e["input"])
    rows = [list(r) for r in parse_out(e["out"])[0]]
    _front_job(p, rows, device, lambda be, d: (aira.epp_front(be, d, threads, False), None))        # untimed: first-touch of the pool's contexts
    out = _front_job(p, rows, device, lambda be, d: (aira.epp_front(be, d, threads, False), None))
    out.pop("per_rank")
    return out


def example_fronts(mb, device):
    """BASELINE.json configs[0..1]: time-to-front of the shipped Examples in the default mode (-t 1), each front checked
    against the committed golden .out.  Every model runs once untimed first (kernel images are loaded lazily on first
    launch, allocations are first-touch), then once timed on a fresh context."""
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
        pr = mb.Problem(p)
        warm = mb.Context(pr, device=device)
        warm.pareto_front()
        warm.close()
        ctx = mb.Context(pr, device=device)
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
