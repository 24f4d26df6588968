"""TEST INFRASTRUCTURE / CPU BASELINE ONLY -- never imported by the product.

Time-to-front on the host cores: the restated generator of aira_oracle.py (reference src/aira.cpp:538-1884, EPP driver
:1886-1990) with HiGHS (`scipy.optimize.milp`) in place of CPLEX -- which cannot be installed here
-- and the strips of every EPP level dealt to a pool of worker processes, one strip per core, like the reference's
`--split -t <cores>` (one thread per strip, src/aira.cpp:1920-1933).  Used by bench.py (`cpu_baseline.front` and the
`--impl reference` arm); it is a stand-in, NOT the reference's own solver."""
from __future__ import annotations

import time

from . import aira_oracle as ao
from .lpformat import read_model

_MODEL = None
_PATH = None


def _init(path):
    """Pool initializer: parse the model and import scipy once per worker process (outside any timed region)."""
    global _MODEL, _PATH
    _MODEL, _PATH = read_model(path), path
    from scipy.optimize import linprog, milp  # noqa: F401


def _warm(_):
    ao.MilpOracle(_MODEL).get_limit(0, [ao.INF if _MODEL.sense == "MIN" else -ao.INF] * _MODEL.k)
    return 0


def _strip(args):
    n_obj, a, b = args
    k = _MODEL.k
    oracle = ao.MilpOracle(_MODEL)
    here, inf = ao.Solutions(k), ao.Solutions(k)
    t = time.perf_counter()
    ao.optimise(_MODEL, oracle, here, inf, ao.Worker(range(k), n_obj, a, b), split=True)
    return [list(r.result) for r in here.store if not r.infeasible], oracle.ip_calls, time.perf_counter() - t


def make_pool(path, procs):
    import multiprocessing as mp
    pool = mp.get_context("spawn").Pool(procs, initializer=_init, initargs=(path,))
    pool.map(_warm, range(procs * 2))
    return pool


def epp_front(path, strips, pool):
    """split_setup (reference src/aira.cpp:1945-1990) with each level's `strips` strips solved in parallel by `pool`.
    Returns (front rows sorted like the reference's output, IPs solved, seconds per level)."""
    model = read_model(path)
    k = model.k
    MIN = model.sense == "MIN"
    free = [ao.INF if MIN else -ao.INF] * k
    oracle = ao.MilpOracle(model)
    ips = [0]
    level_s = []

    def level(n_obj):
        if n_obj == 1:
            return [oracle.get_limit(0, free)]
        lower = level(n_obj - 1)
        t = time.perf_counter()
        res = oracle.get_limit(n_obj - 1, free)
        if MIN:
            smallest, biggest = res[n_obj - 1], max([ao.INT_MIN] + [s[n_obj - 1] for s in lower])
            if biggest == smallest:
                biggest = ao.INT_MAX
        else:
            biggest, smallest = res[n_obj - 1], min([ao.INT_MAX] + [s[n_obj - 1] for s in lower])
            if biggest == smallest:
                smallest = ao.INT_MIN
        out = []
        for rows, n_ip, _ in pool.map(_strip, [(n_obj, a, b) for a, b in ao.strip_bounds(MIN, biggest, smallest, strips, False)]):
            out += rows
            ips[0] += n_ip
        level_s.append(time.perf_counter() - t)
        return out

    rows = level(k)
    front = sorted({tuple(r) for r in rows}, key=lambda r: tuple(-v for v in r))
    return front, ips[0] + oracle.ip_calls, level_s
