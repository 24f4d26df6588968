"""`aira` command-line mirror (reference src/aira.cpp:140-364) on top of the C ABI.

    python -m moip_aira_b200.aira -p Examples/3KP10.lp [-o out] [--split] [--split-normal] [-t N] [-c N] [-s]

Options, defaults, output-file naming and the `.out` layout follow the reference (options :159-188,
naming :213-221, writer :252, :336-358).  `-c` (CPLEX threads) is accepted and ignored: there is no
CPLEX.  -t 1 runs the sequential generator on this rank's GPU.  -t N > 1 without --split is the reference's
synergistic mode (src/aira.cpp:277-308): it runs the cooperative workers of csrc/generator.cpp -- min(N, k) workers that
each own one objective and publish a monotone limit on it (the race-free re-hosting of the bound-sharing protocol,
SURVEY.md section 8f-3) -- one per rank (one GPU each) under torchrun, else on one GPU's pool of solver contexts.
With --split the EPP strips of every level (src/aira.cpp:1886-1990) are sharded over the ranks of a torch.distributed
job (one rank per GPU, `torchrun --nproc-per-node G`): the ranks draw strips from one job-wide counter, all-gather the
cache records they produce while they solve (every strip of the job can reuse every other strip's relaxations, like
the reference's threads share `here` / `infeasibles`, src/aira.cpp:1918-1933) and all-gather the points found between
levels; inside a rank the strips run concurrently on a pool of solver contexts (MOIP_WORKERS host threads; default: up to 16, bounded by the rank's share of the host cores)
that share the GPU.  On this backend --split is the faster mode for assignment-type models (profiles/r02_fronts.md).
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

INT_MAX, INT_MIN = 2147483647, -2147483648


# ------------------------------------------------------------------------------------ distributed
class Dist:
    """Thin wrapper: rank/world from the environment, all-gather of variable-length int rows."""

    def __init__(self, device=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.device = device
        self.pg = None
        Dist._jobs = getattr(Dist, "_jobs", 0) + 1      # every rank makes the same sequence of Dist objects: names job-wide keys
        self._job = Dist._jobs
        if self.world > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                backend = "nccl" if device is not None else "gloo"
                dist.init_process_group(backend=backend)
            self.pg = dist

    def allgather_rows(self, rows, k):
        """rows: list of k-int tuples on this rank -> concatenation over ranks (rank order)."""
        if self.world == 1:
            return [tuple(r) for r in rows]
        import torch
        dev = self.device if self.device is not None else "cpu"
        cnt = torch.tensor([len(rows)], dtype=torch.int64, device=dev)
        cnts = [torch.zeros_like(cnt) for _ in range(self.world)]
        self.pg.all_gather(cnts, cnt)
        mx = max(int(c.item()) for c in cnts)
        buf = torch.zeros((max(mx, 1), k), dtype=torch.int32, device=dev)
        if rows:
            buf[:len(rows)] = torch.tensor(np.asarray(rows, dtype=np.int32).reshape(-1, k), device=dev)
        bufs = [torch.zeros_like(buf) for _ in range(self.world)]
        self.pg.all_gather(bufs, buf)
        out = []
        for c, b in zip(cnts, bufs):
            out += [tuple(int(v) for v in r) for r in b[:int(c.item())].cpu().numpy()]
        return out

    def barrier(self):
        if self.world > 1:
            self.pg.barrier()

    def counter(self, name):
        """Job-wide atomic counter (0, 1, 2, ...): the ranks draw EPP strips from it, so that no GPU idles while
        another still has strips queued.  One process: a local counter; several: the job's c10d store."""
        if self.world == 1:
            import itertools
            import threading
            it, lock = itertools.count(), threading.Lock()

            def nxt():
                with lock:
                    return next(it)
            return nxt
        from torch.distributed import distributed_c10d
        store = distributed_c10d._get_default_store()
        self._seq = getattr(self, "_seq", 0) + 1          # every rank makes the same sequence of calls
        key = "moip_strips_%d_%d_%s" % (self._job, self._seq, name)
        return lambda: store.add(key, 1) - 1


class RecordExchange:
    """All-gather of newly produced cache records between the ranks WHILE a level's strips are being solved (the
    reference's strip threads share `here` / `infeasibles` in one address space, src/aira.cpp:1918-1933; SURVEY.md 8e:
    "all-gather of cache records so strips can reuse each other's relaxations").  A background thread per rank runs the
    same sequence of fixed-size all-gathers (NCCL on device tensors under torchrun on GPUs, gloo in the CPU tests): every
    `period_s` it exports what this rank's pool has produced since the last round and imports what the others sent.  The
    loop ends in the same round on every rank: when all ranks have reported "my strips are done and drained".

        with RecordExchange(dist, endpoint, k):
            rows = backend.run_strips(...)

    `endpoint` offers export_records(cap) -> (ip[n][k], result[n][k], infeasible[n]) and import_records(ip, result,
    infeasible); None (or one rank) makes this a no-op."""

    def __init__(self, dist: Dist, endpoint, k, period_s=None, cap=1024):
        self.dist, self.ep, self.k, self.cap = dist, endpoint, int(k), int(cap)
        self.period_s = float(os.environ.get("MOIP_EXCHANGE_PERIOD_MS", "15")) * 1e-3 if period_s is None else period_s
        self.active = dist.world > 1 and endpoint is not None and not os.environ.get("MOIP_NO_EXCHANGE")
        self.rounds = self.sent = self.received = 0
        self.error = None
        self._done = None
        self._th = None

    def __enter__(self):
        if self.active:
            import threading
            self._done = threading.Event()
            self._th = threading.Thread(target=self._loop, daemon=True)
            self._th.start()
        return self

    def __exit__(self, *exc):
        if self.active:
            self._done.set()
            self._th.join()
            if self.error is not None and exc[0] is None:
                raise self.error
        return False

    def _loop(self):
        try:
            import torch
            k, cap, world, rank = self.k, self.cap, self.dist.world, self.dist.rank
            dev = self.dist.device
            if dev is not None:
                torch.cuda.set_device(dev)
            width = 2 * k + 1
            host = np.zeros(2 + cap * width)
            out = [torch.zeros(2 + cap * width, dtype=torch.float64, device=dev if dev is not None else "cpu")
                   for _ in range(world)]
            while True:
                fin = self._done.is_set()                  # read BEFORE the export: nothing is produced after it is set
                ip, res, inf = self.ep.export_records(cap)
                n = len(inf)
                host[0], host[1] = n, 1.0 if (fin and n < cap) else 0.0
                if n:
                    rec = host[2:2 + n * width].reshape(n, width)
                    rec[:, :k], rec[:, k:2 * k], rec[:, 2 * k] = ip, res, inf
                buf = torch.from_numpy(host[:2 + cap * width].copy())
                if dev is not None:
                    buf = buf.to(dev)
                self.dist.pg.all_gather(out, buf)
                got = torch.stack(out).cpu().numpy()
                self.rounds += 1
                self.sent += n
                for r in range(world):
                    m = int(got[r, 0])
                    if r == rank or m == 0:
                        continue
                    rec = got[r, 2:2 + m * width].reshape(m, width)
                    self.ep.import_records(rec[:, :k], rec[:, k:2 * k].astype(np.int32), rec[:, 2 * k].astype(np.int32))
                    self.received += m
                if all(got[r, 1] == 1.0 for r in range(world)):
                    return
                if not fin:
                    time.sleep(self.period_s)
        except Exception as e:                             # noqa: BLE001  (re-raised by __exit__ on the caller's thread)
            self.error = e


# ------------------------------------------------------------------------------------ backends
def default_workers():
    """Solver contexts (= host threads) per GPU: MOIP_WORKERS, else 24.  Every worker drives the B&B of its own IPs from its own
    host thread; with chained rounds (csrc/bbchain.h) that is a handful of launches and one wait per IP, so a worker needs a
    fraction of a core, and the pool lets its workers sleep on a blocking event whenever they outnumber the cores this rank
    may use (MOIP_SYNC=auto).  Measured, 3AP n=30 front on one B200 with 16 host cores: 16 spinning workers 6.6 s, 24 spinning
    6.2 s, 24 sleeping 5.5 - 5.7 s, 32 spinning 6.2 s (profiles/r02_fronts.md)."""
    if os.environ.get("MOIP_WORKERS"):
        return int(os.environ["MOIP_WORKERS"])
    return 24


class GpuBackend:
    """Product backend on this rank's B200: one solver context for the sequential generator, a pool of
    `workers` contexts (one host thread each, reference src/aira.cpp:1920-1933) for the EPP strips."""

    def __init__(self, path, device=0, stream=None, workers=None):
        import moip_aira_b200 as mb
        self.mb = mb
        self.device, self.stream = device, stream
        self.problem = mb.Problem(path)
        self.k = self.problem.objcnt
        self.sense = self.problem.objsen
        self.workers = max(1, default_workers() if workers is None else int(workers))
        self._ctx = None
        self._pool = None

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = self.mb.Context(self.problem, device=self.device, stream=self.stream)
        return self._ctx

    @property
    def pool(self):
        if self._pool is None:
            self._pool = self.mb.WorkerPool(self.problem, self.device, self.workers)
        return self._pool

    def get_limit(self, obj, rhs):
        st, res = self.pool.get_limit(obj, rhs)
        return res

    def run_strips(self, n_obj, strips, claim, share=0, windows=None):
        """The strips of one EPP level: this rank's workers draw strip indices from `claim` (shared by all ranks),
        solve them concurrently and share `here`/`infeasibles` like the reference's threads.  share > 0: at most that
        many strips in flight on this rank (its part of the level when there are fewer strips than workers in the job).
        windows: one per entry of `strips` -- the entries are boxes (strip x window on objective 1)."""
        self.pool.set_max_workers(share)
        return self.pool.run_strips(n_obj, strips, claim, windows=windows)

    def exchange_endpoint(self):
        """what RecordExchange talks to: the pool's export_records / import_records"""
        return self.pool

    def sequential_front(self):
        return self.ctx.pareto_front()

    def synergistic_local(self, n_workers):
        """-t N without --split on this rank's GPU alone: min(N, k) cooperative workers on a pool of their own."""
        w = max(1, min(int(n_workers), self.k))
        pool = self.mb.WorkerPool(self.problem, self.device, w)
        try:
            return pool.synergistic_front(w)
        finally:
            st = pool.stats()
            self._coop_stats = {f: self._coop_stats.get(f, 0) + v for f, v in st.items()} if getattr(self, "_coop_stats", None) else st
            pool.close()

    def coop_worker(self, perm, limits):
        """This rank's cooperative worker (owns perm[-1]) on this rank's GPU; returns the points it found."""
        mb = self.mb
        all_sols, inf = mb.Solutions(self.ctx), mb.Solutions(self.ctx)
        try:
            self.ctx.coop_optimise(mb.make_worker(self.k, perm=perm), limits, all_sols, inf)
            return [tuple(r) for r in all_sols.sort_unique()]
        finally:
            all_sols.close()
            inf.close()

    def split_strips(self, biggest, smallest, num_threads, split_normal):
        return self.mb.split_strips(self.sense, biggest, smallest, num_threads, split_normal)

    def stats(self):
        """counters summed over everything this backend has run (sequential context, strip pool, cooperative pools)"""
        tot = {}
        for st in ([self._ctx.stats()] if self._ctx is not None else []) + ([self._pool.stats()] if self._pool is not None else []) + \
                ([self._coop_stats] if getattr(self, "_coop_stats", None) else []):
            for f, v in st.items():
                tot[f] = tot.get(f, 0) + v
        return tot

    def ip_count(self):
        return self.stats().get("ip_solved", 0)


def windows_for(n_obj, num_threads, world):
    """How many windows on objective 1 the boxes of a level get: the largest power of two (at most 16) that leaves at least
    8 strips of the last objective.  MOIP_WINDOWS=<n> fixes it (1 = strips only, the reference's cut)."""
    if n_obj < 3:
        return 1
    env = os.environ.get("MOIP_WINDOWS")
    if env:
        return max(1, min(int(env), num_threads))
    nwin = 1
    while nwin < 16 and num_threads // (nwin * 2) >= 8:
        nwin *= 2
    return nwin


def window_edges(values, nwin, is_min):
    """(near edge, far edge) of `nwin` windows on an objective whose values over the level below are `values` (sorted,
    distinct): quantile cuts; the first window is open towards "free", the last has no far edge.  Bounds are upper limits
    for MIN models and lower limits for MAX models, and a window ends one unit before the next one starts."""
    big = 1e20
    if nwin <= 1 or len(values) < 2 * nwin:
        return [(big if is_min else -big, -big if is_min else big)]
    cuts = sorted({values[len(values) * i // nwin] for i in range(1, nwin)})
    if is_min:
        near = [big] + [float(c) for c in reversed(cuts)]
        return [(e, near[i + 1] + 1 if i + 1 < len(near) else -big) for i, e in enumerate(near)]
    near = [-big] + [float(c) for c in cuts]
    return [(e, near[i + 1] - 1 if i + 1 < len(near) else big) for i, e in enumerate(near)]


def epp_front(be, dist: Dist, num_threads: int, split_normal: bool, stats: list | None = None):
    """split_setup (src/aira.cpp:1945-1990) with each level's strips sharded over the ranks.  stats (a list) receives one
    dict per level: strips, exchange rounds, cache records sent / received by this rank."""
    k = be.k
    is_min = be.sense == 0
    free = [1e20 if is_min else -1e20] * k

    def level(n_obj):
        if n_obj == 1:
            r = be.get_limit(0, free)
            return [tuple(r)] if r is not None else []
        lower = level(n_obj - 1)
        if not lower:
            return []
        res = be.get_limit(n_obj - 1, free)
        if res is None:
            return []
        if is_min:
            smallest, biggest = res[n_obj - 1], max([INT_MIN] + [s[n_obj - 1] for s in lower])
            if biggest == smallest:
                biggest = INT_MAX
        else:
            biggest, smallest = res[n_obj - 1], min([INT_MAX] + [s[n_obj - 1] for s in lower])
            if biggest == smallest:
                smallest = INT_MIN
        t_level = time.monotonic()
        # The levels below the top one only supply the range of the next objective and are small fronts: with the top level's
        # strip count (one per worker of the whole job) most of their strips would hold nothing but their start-up.
        # MOIP_LOWER_STRIPS=<n> overrides (0 = as many as the top level, the reference's behaviour).
        if n_obj < k and (hasattr(be, "exchange_endpoint") or getattr(be, "boxes_ok", False)):
            cap_lower = int(os.environ.get("MOIP_LOWER_STRIPS", str(max(16, 4 * dist.world))))
            if cap_lower > 0:
                num_threads_here = min(num_threads, cap_lower)
            else:
                num_threads_here = num_threads
        else:
            num_threads_here = num_threads
        # Boxes (moip_worker::window; no counterpart in the reference): with many more workers than a level has heavy strips,
        # the level is cut both ways -- `nwin` windows on objective 1 (edges = quantiles of that objective over the level
        # below, the first window open towards "free", the last without a far edge) times num_threads_here / nwin strips of the
        # last objective.  A strip starts with an (n_obj-1)-objective front of its own; the boxes of one strip share that
        # start-up between them instead of each paying for it.
        nwin = windows_for(n_obj, num_threads_here, dist.world) if hasattr(be, "exchange_endpoint") or getattr(be, "boxes_ok", False) else 1
        windows = None
        if nwin > 1:
            edges = window_edges(sorted({s[1] for s in lower}), nwin, is_min)
            nwin = len(edges)
        if nwin > 1:
            base = be.split_strips(biggest, smallest, max(1, num_threads_here // nwin), split_normal)
            strips = [st for st in base for _ in range(nwin)]
            windows = [w for _ in base for w in edges]
        else:
            strips = be.split_strips(biggest, smallest, num_threads_here, split_normal)
        # Strip s belongs to rank s mod world: every rank gets an interleaved sample of the range -- light strips from its
        # ends and heavy ones from its middle (the points crowd there) -- so the ranks carry about the same load without
        # talking to each other; inside a rank, idle workers then cut busy strips in two (moip_pool_run_strips_claim).
        # MOIP_GLOBAL_STRIP_COUNTER=1: the ranks draw strips from one job-wide counter in the c10d store instead.
        if dist.world > 1 and not os.environ.get("MOIP_GLOBAL_STRIP_COUNTER"):
            owner = (lambda b: (b // nwin + b % nwin) % dist.world) if nwin > 1 else (lambda b: b % dist.world)   # boxes: a Latin square of strips and windows
            mine = iter([s_ for s_ in range(len(strips)) if owner(s_) == dist.rank] + [len(strips)] * 4096)
            lock = __import__("threading").Lock()

            def claim():
                with lock:
                    return next(mine, len(strips))
            share = 0
        else:
            claim = dist.counter("level%d" % n_obj)
            share = -(-len(strips) // dist.world) if dist.world > 1 else 0  # ceil: no rank claims more than its part at once
        endpoint = be.exchange_endpoint() if hasattr(be, "exchange_endpoint") else None
        with RecordExchange(dist, endpoint, k) as ex:
            rows = be.run_strips(n_obj, strips, claim, share, windows) if windows is not None else be.run_strips(n_obj, strips, claim, share)
        if stats is not None:
            stats.append({"n_obj": n_obj, "strips": len(strips), "windows": nwin, "exchange_rounds": ex.rounds, "records_sent": ex.sent,
                          "records_received": ex.received, "rows_here": len(rows),
                          "seconds": round(time.monotonic() - t_level, 3)})
        return dist.allgather_rows(rows, k)

    rows = level(k)
    return sorted(set(rows), key=lambda r: tuple(-v for v in r))     # sort_unique (src/aira.cpp:336)


def synergistic_front(be, dist: Dist, poll_s: float = 0.001):
    """`-t W` without --split across ranks: rank r is cooperative worker r (r-th rotation of the objective order, owns
    its last objective) on its own GPU.  The only exchange while solving is the W published limits -- one small value
    per owned objective in the job's c10d store, mirrored into the local limits handle by a host thread (publishing is
    fetch-min / fetch-max, so repeated or reordered updates are harmless) -- and one all-gather of the points at the end.
    Ranks beyond k (there is one owner per objective) contribute nothing and only take part in the gather."""
    import threading
    import moip_aira_b200 as mb
    k, world, rank = be.k, dist.world, dist.rank
    workers = min(world, k)
    perms = mb.coop_workers(k, workers)
    owned = [p[-1] for p in perms]
    limits = mb.CoopLimits(k, be.sense, owned if workers > 1 else [])
    store = None
    if world > 1:
        from torch.distributed import distributed_c10d
        store = distributed_c10d._get_default_store()
        dist._coop_seq = getattr(dist, "_coop_seq", 0) + 1
        key = lambda j: "moip_coop_%d_%d" % (dist._coop_seq, j)      # noqa: E731
        if rank < workers:
            store.set(key(owned[rank]), "free")
        dist.barrier()                                               # every owner's key exists before anybody reads
    stop = threading.Event()

    def mirror():
        last = None
        while True:
            final = stop.is_set()
            if rank < workers:                                       # my own limit -> store
                st, v = limits.read(owned[rank])
                cur = "done" if st == 2 else ("free" if st == 0 else str(v))
                if cur != last:
                    store.set(key(owned[rank]), cur)
                    last = cur
            for w in range(workers):                                 # the others' limits -> my handle
                if w == rank:
                    continue
                val = store.get(key(owned[w])).decode()
                if val == "done":
                    limits.publish(owned[w], done=True)
                elif val != "free":
                    limits.publish(owned[w], int(val))
            if final:
                return
            time.sleep(poll_s)

    th = None
    if store is not None:
        th = threading.Thread(target=mirror, daemon=True)
        th.start()
    try:
        rows = be.coop_worker(perms[rank], limits) if rank < workers else []
    finally:
        stop.set()
        if th is not None:
            th.join()
    rows = dist.allgather_rows(rows, k)
    limits.close()
    return sorted(set(rows), key=lambda r: tuple(-v for v in r))


def format_out(front, cpu_s, wall_s, ips, tag):
    lines = ["", f"Using improved algorithm at {tag}"]
    lines += ["".join(f"{v}\t" for v in row) for row in front]
    lines += ["", "---", f"{cpu_s:8.3f} CPU seconds", f"{wall_s:8.3f} elapsed seconds",
              f"{ips:8d} IPs solved", f"{len(front):8d} Solutions found"]
    return "\n".join(lines) + "\n"


def main(argv=None, backend_factory=None):
    ap = argparse.ArgumentParser(prog="aira", description="Options for aira")
    ap.add_argument("-p", "--lp", dest="lp", help="The LP file to solve. Required.")
    ap.add_argument("-o", "--output", dest="output", help="The output file. Optional.")
    ap.add_argument("--split", action="store_true", help="Split the range of the last objective into one strip per thread")
    ap.add_argument("--split-normal", action="store_true", dest="split_normal")
    ap.add_argument("-s", "--spread", action="store_true", default=True)
    ap.add_argument("-t", "--threads", type=int, default=1)
    ap.add_argument("-c", "--cplex_threads", type=int, default=1)
    args = ap.parse_args(argv)
    if args.split_normal and args.threads > 12:                       # src/aira.cpp:199-203
        print("Error: split_normal can only handle at most 12 threads.", file=sys.stderr)
        return 1
    if not args.lp:
        print("Error: You must pass in a problem file. All other parameters are optional.", file=sys.stderr)
        ap.print_help()
        return 1
    out_path = args.output or (args.lp[:args.lp.rfind(".")] + ".out" if "." in args.lp else args.lp + ".out")
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if backend_factory is None:
        import torch
        dev = None
        # one rank per GPU over NCCL; MOIP_DIST_BACKEND=gloo keeps the collectives on the host (several ranks sharing a
        # GPU, e.g. the 2-rank GPU test on a one-GPU box -- NCCL refuses two ranks on one device)
        gloo = os.environ.get("MOIP_DIST_BACKEND", "nccl") == "gloo"
        gpu = local_rank % max(1, torch.cuda.device_count()) if gloo else local_rank
        if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not gloo:
            torch.cuda.set_device(gpu)
            dev = torch.device("cuda", gpu)
        dist = Dist(dev)
        be = GpuBackend(args.lp, device=gpu)
    else:
        dist = Dist(None)
        be = backend_factory(args.lp)
    if be.k >= 5:
        print("Error: at most 4 objectives are supported.", file=sys.stderr)
        return 2
    t0, c0 = time.monotonic(), time.process_time()
    strips_for_workers = bool(os.environ.get("MOIP_THREADS_AS_STRIPS"))
    coop = not args.split and (args.threads > 1 or dist.world > 1) and not strips_for_workers
    if args.split_normal and max(args.threads, dist.world) > 12:
        print("Error: split_normal can only handle at most 12 threads (strips: max(-t, ranks)).", file=sys.stderr)
        return 1
    if coop:
        # -t N without --split = the reference's synergistic mode (src/aira.cpp:277-308), its default: the cooperative
        # workers of csrc/generator.cpp -- one per rank (one GPU each) under torchrun, else min(N, k) workers on this
        # GPU's pool.  There is one owner per objective, so at most k workers are active (the reference caps at k!).
        if dist.rank == 0 and max(args.threads, dist.world) > be.k:
            print("note: %d workers asked for, %d objectives: %d cooperative workers run (one owner per objective); "
                  "--split -t N uses every thread / GPU" % (max(args.threads, dist.world), be.k, be.k), file=sys.stderr)
        front = synergistic_front(be, dist) if dist.world > 1 else be.synergistic_local(args.threads)
    elif args.split or args.threads > 1 or dist.world > 1:
        front = epp_front(be, dist, max(1, args.threads, dist.world), args.split_normal)
    else:
        front = be.sequential_front()
    wall, cpu = time.monotonic() - t0, time.process_time() - c0
    if dist.rank == 0:
        with open(out_path, "w") as fh:
            fh.write(format_out(front, cpu, wall, int(be.ip_count()), "moip_b200"))
    dist.barrier()
    return 0


if __name__ == "__main__":
    sys.exit(main())
