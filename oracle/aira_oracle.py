"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

CPU restatement of moip_aira's hot path, used as the checker in tests/,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg:

* `FeasibleSet` / `MilpOracle` : exact lexicographic IP oracles standing in for the
  CPLEX calls of reference `src/aira.cpp:452-536` (`solve`) and `:367-450` (`get_limit`).
  The result vector of `solve()` is a lexicographic optimum, unique in objective space,
  hence solver independent.
* `Solutions`   : restatement of reference `src/solutions.cpp:11-101`, `src/result.cpp:9-46`.
* `optimise`    : restatement of the sequential / EPP paths of reference
  `src/aira.cpp:538-1884` (the 4-way "next rhs" table), including
  the INT_MIN-1 wrap that the reference relies on.
* `split_setup` : restatement of reference `src/aira.cpp:1886-1990` (EPP strips).
* `pareto_front`: whole-program result = what `main` prints (`src/aira.cpp:336-346`).

Pinned against the 7 committed Examples/*.out fronts (tests/test_oracle.py).
"""
from __future__ import annotations

import itertools

import numpy as np

from .lpformat import INF, Model

INT_MAX = 2147483647
INT_MIN = -2147483648
# CPLEX status codes the reference tests for (src/aira.cpp:489-492, 644, 840)
CPXMIP_OPTIMAL = 101
CPXMIP_INFEASIBLE = 103
CPXMIP_INForUNBD = 119
# normal-quantile strip table of reference src/aira.cpp:55-69 (values are data, not code)
NORMAL_VALUES = {
    1: [0, 1], 2: [0, 0.5, 1], 3: [0, 0.356, 0.644, 1], 4: [0, 0.275, 0.5, 0.725, 1],
    5: [0, 0.219, 0.416, 0.584, 0.781, 1], 6: [0, 0.178, 0.256, 0.5, 0.644, 0.822, 1],
    7: [0, 0.144, 0.311, 0.44, 0.56, 0.689, 0.856, 1],
    8: [0, 0.117, 0.275, 0.394, 0.5, 0.606, 0.725, 0.883, 1],
    9: [0, 0.093, 0.245, 0.356, 0.453, 0.547, 0.644, 0.755, 0.907, 1],
    10: [0, 0.073, 0.219, 0.325, 0.416, 0.5, 0.584, 0.675, 0.781, 0.927, 1],
    11: [0, 0.055, 0.197, 0.298, 0.384, 0.462, 0.538, 0.616, 0.702, 0.803, 0.945, 1],
    12: [0, 0.039, 0.178, 0.275, 0.356, 0.430, 0.5, 0.570, 0.644, 0.725, 0.822, 0.961, 1],
}


def wrap32(v: int) -> int:
    """Two's-complement int32 wrap (what `max[d]-1` does when max[d]==INT_MIN)."""
    return (int(v) + 2 ** 31) % 2 ** 32 - 2 ** 31


# --------------------------------------------------------------------------- IP oracles
class FeasibleSet:
    """All feasible integer points of a small model, enumerated once (solver free)."""

    def __init__(self, model: Model, limit: int = 5_000_000):
        self.model = model
        self.X = self._enumerate(model, limit)
        self.P = np.rint(self.X @ model.C.T).astype(np.int64)   # objective vectors
        self.ip_calls = 0

    @staticmethod
    def _enumerate(m: Model, limit):
        n = m.n
        lo = np.ceil(m.lb - 1e-9).astype(np.int64)
        ubf = m.ub.copy()
        # finite implied upper bounds from <= rows with non-negative data (knapsack rows)
        for i in range(m.ms):
            if m.row_sense[i] in ("L", "E") and np.all(m.A[i] >= 0) and np.all(lo >= 0):
                with np.errstate(divide="ignore"):
                    cap = np.where(m.A[i] > 0, np.floor(m.b[i] / np.where(m.A[i] > 0, m.A[i], 1) + 1e-9), INF)
                ubf = np.minimum(ubf, cap)
        if np.any(ubf >= INF):
            raise ValueError("unbounded column; cannot enumerate")
        hi = np.floor(ubf + 1e-9).astype(np.int64)
        A, b = m.A, m.b
        Apos, Aneg = np.maximum(A, 0), np.minimum(A, 0)
        out = []
        x = lo.copy()

        def feasible_possible(j):
            # activity bounds with x[:j] fixed, rest in [lo,hi]
            fixed = A[:, :j] @ x[:j]
            mn = fixed + Apos[:, j:] @ lo[j:] + Aneg[:, j:] @ hi[j:]
            mx = fixed + Apos[:, j:] @ hi[j:] + Aneg[:, j:] @ lo[j:]
            for i, s in enumerate(m.row_sense):
                if s in ("L", "E") and mn[i] > b[i] + 1e-9:
                    return False
                if s in ("G", "E") and mx[i] < b[i] - 1e-9:
                    return False
            return True

        def rec(j):
            if not feasible_possible(j):
                return
            if j == n:
                out.append(x.copy())
                if len(out) > limit:
                    raise ValueError("too many feasible points")
                return
            for v in range(lo[j], hi[j] + 1):
                x[j] = v
                rec(j + 1)
            x[j] = lo[j]

        rec(0)
        return np.array(out, dtype=np.int64).reshape(-1, n)

    def lex_solve(self, perm, n_obj, rhs):
        """Restates reference src/aira.cpp:452-536. Returns (status, result[k] or None)."""
        m = self.model
        k = m.k
        srhs = np.array(rhs, dtype=float)
        sgn = 1 if m.sense == "MIN" else -1
        mask = np.ones(len(self.P), dtype=bool)
        result = [0] * k
        for jp in range(n_obj):
            j = perm[jp]
            self.ip_calls += 1
            mask = np.all(sgn * self.P <= sgn * srhs + 0.0, axis=1)
            if not mask.any():
                return CPXMIP_INFEASIBLE, None
            best = (sgn * self.P[mask, j]).min() * sgn
            result[j] = int(best)
            srhs[j] = float(best)
        mask = np.all(sgn * self.P <= sgn * srhs, axis=1)
        i = int(np.flatnonzero(mask)[0])
        for jp in range(n_obj, k):
            result[perm[jp]] = int(self.P[i, perm[jp]])
        return CPXMIP_OPTIMAL, result

    def get_limit(self, obj, rhs):
        """Restates reference src/aira.cpp:367-450: returns result[k] or None if infeasible."""
        m = self.model
        sgn = 1 if m.sense == "MIN" else -1
        self.ip_calls += 1
        mask = np.all(sgn * self.P <= sgn * np.array(rhs, dtype=float), axis=1)
        if not mask.any():
            return None
        idx = np.flatnonzero(mask)
        i = idx[np.argmin(sgn * self.P[idx, obj])]
        return [int(v) for v in self.P[i]]


class MilpOracle:
    """HiGHS (scipy.optimize.milp) stand-in for larger instances. NOT CPLEX."""

    def __init__(self, model: Model):
        from scipy.optimize import Bounds, LinearConstraint  # noqa: F401
        self.model = model
        self.ip_calls = 0

    def _milp(self, c, srhs):
        from scipy.optimize import Bounds, LinearConstraint, milp
        m = self.model
        sgn = 1 if m.sense == "MIN" else -1
        cons = []
        lo = np.where(np.array(m.row_sense) == "L", -np.inf, m.b)
        hi = np.where(np.array(m.row_sense) == "G", np.inf, m.b)
        if m.ms:
            cons.append(LinearConstraint(m.A, lo, hi))
        act = np.abs(srhs) < 1e19
        if act.any():
            if sgn == 1:
                cons.append(LinearConstraint(m.C[act], -np.inf, srhs[act]))
            else:
                cons.append(LinearConstraint(m.C[act], srhs[act], np.inf))
        ub = np.where(m.ub >= INF, np.inf, m.ub)
        res = milp(sgn * c, constraints=cons, integrality=m.is_int.astype(int),
                   bounds=Bounds(m.lb, ub), options={"mip_rel_gap": 0.0})
        self.ip_calls += 1
        if res.status == 2 or res.x is None:
            return None
        return np.rint(res.x)

    def lex_solve(self, perm, n_obj, rhs):
        m = self.model
        srhs = np.array(rhs, dtype=float)
        result = [0] * m.k
        x = None
        for jp in range(n_obj):
            j = perm[jp]
            x = self._milp(m.C[j], srhs)
            if x is None:
                return CPXMIP_INFEASIBLE, None
            result[j] = int(round(float(m.C[j] @ x)))
            srhs[j] = result[j]
        for jp in range(n_obj, m.k):
            result[perm[jp]] = int(round(float(m.C[perm[jp]] @ x)))
        return CPXMIP_OPTIMAL, result

    def get_limit(self, obj, rhs):
        m = self.model
        x = self._milp(m.C[obj], np.array(rhs, dtype=float))
        if x is None:
            return None
        return [int(round(float(m.C[j] @ x))) for j in range(m.k)]


def make_oracle(model: Model):
    try:
        return FeasibleSet(model, limit=2_000_000)
    except (ValueError, RecursionError):
        return MilpOracle(model)


# --------------------------------------------------------------------------- Solutions
class Result:
    __slots__ = ("ip", "result", "infeasible")

    def __init__(self, ip, result, infeasible):
        self.ip = [float(v) for v in ip]
        self.infeasible = bool(infeasible)
        self.result = None if infeasible else [int(v) for v in result]


class Solutions:
    """Restates reference src/solutions.cpp:11-101 and src/solutions.h:41-57."""

    def __init__(self, k):
        self.k = k
        self.store = []
        self.compared = 0

    def find(self, ip, sense):
        k = self.k
        for idx, r in enumerate(self.store):
            self.compared += 1
            ok = True
            for i in range(k):
                if sense == "MIN":
                    if r.ip[i] < ip[i] or (not r.infeasible and r.result[i] > ip[i]):
                        ok = False
                        break
                else:
                    if r.ip[i] > ip[i] or (not r.infeasible and r.result[i] < ip[i]):
                        ok = False
                        break
            if ok:
                return idx, r
        return -1, None

    def insert(self, ip, result, infeasible):
        self.store.append(Result(ip, result, infeasible))

    def merge(self, other):
        self.store[0:0] = other.store       # splice at begin (solutions.h:41-44)
        other.store = []

    def sort_unique(self):
        """Descending lexicographic, infeasible first, then dedupe (result.cpp:9-46)."""
        inf = [r for r in self.store if r.infeasible]
        feas = sorted((r for r in self.store if not r.infeasible),
                      key=lambda r: tuple(-v for v in r.result))
        out = []
        for r in inf[:1] + feas:
            if out and out[-1].infeasible == r.infeasible and (r.infeasible or out[-1].result == r.result):
                continue
            out.append(r)
        self.store = out


# --------------------------------------------------------------------------- generator
class Worker:
    """The fields of reference `Thread` the sequential / EPP paths read (src/thread.h:37-43)."""

    def __init__(self, perm, n_obj, split_start=0.0, split_stop=0.0):
        self.perm = list(perm)
        self.n_obj = n_obj
        self.split_start = split_start
        self.split_stop = split_stop


def optimise(model: Model, oracle, all_sols: Solutions, infeasibles: Solutions, t: Worker,
             split: bool, trace=None):
    """One worker's subproblem stream (reference src/aira.cpp:538-1884 without bound sharing).

    `trace`, if a list, receives one tuple per generator iteration:
    (rhs tuple, hit(bool), infeasible(bool), result tuple|None).
    """
    k = model.k
    MIN = model.sense == "MIN"
    inf_rhs = INF if MIN else -INF
    perm = t.perm
    s = Solutions(k)
    rhs = [inf_rhs] * k
    if split:
        rhs[perm[t.n_obj - 1]] = t.split_start                    # :607
    status, result = oracle.lex_solve(perm, t.n_obj, rhs)          # :614
    if status == CPXMIP_INFEASIBLE:
        infeasibles.insert(rhs, None, True)                        # :644-645
    else:
        (all_sols if split else s).insert(rhs, result, False)      # :647-650
    if trace is not None:
        trace.append((tuple(rhs), False, status == CPXMIP_INFEASIBLE, None if result is None else tuple(result)))
    if status == CPXMIP_INFEASIBLE:
        # nothing lies inside these bounds; the reference's max[]/min[] trackers are uninitialised
        # from here on (:693-697), so the defined behaviour is to stop
        all_sols.merge(s)
        return
    if split:
        t.split_stop += -1 if MIN else 1                           # :653-657
    mx = list(result)
    mn = list(result)                                              # :693-697

    def step(d):
        """rhs[d] = max[d]-1 ; max[d] = INT_MIN (MIN) / mirrored (MAX), with int32 wrap."""
        if MIN:
            rhs[d] = float(wrap32(mx[d] - 1))
            mx[d] = INT_MIN
        else:
            rhs[d] = float(wrap32(mn[d] + 1))
            mn[d] = INT_MAX

    for oc in range(1, t.n_obj):                                   # :723
        objective = perm[oc]
        depth_level = 1
        depth = perm[depth_level]
        onwalk = False
        infcnt = 0
        inflast = False
        for jp in range(1, k):                                     # :733-756
            rhs[perm[jp]] = inf_rhs
        if split:
            rhs[perm[t.n_obj - 1]] = t.split_start                 # :757-759
        rhs[objective] = float(wrap32(mx[objective] - 1)) if MIN else float(wrap32(mn[objective] + 1))
        if split:                                                  # :778-801
            if MIN and rhs[t.n_obj - 1] < t.split_stop:
                break
            if (not MIN) and rhs[t.n_obj - 1] > t.split_stop:
                break
        mx[objective] = INT_MIN                                    # :802-803
        mn[objective] = INT_MAX
        while infcnt < oc:                                         # :804
            idx, rel = infeasibles.find(rhs, model.sense)          # :816
            if rel is None:
                idx, rel = (all_sols if split else s).find(rhs, model.sense)   # :818-822
            hit = rel is not None
            if hit:
                infeasible = rel.infeasible
                result = rel.result
            else:
                status, result = oracle.lex_solve(perm, t.n_obj, rhs)          # :835
                infeasible = status in (CPXMIP_INFEASIBLE, CPXMIP_INForUNBD)
                if infeasible:
                    infeasibles.insert(rhs, None, True)
                else:
                    (all_sols if split else s).insert(rhs, result, False)
            if trace is not None:
                trace.append((tuple(rhs), hit, infeasible, None if infeasible else tuple(result)))
            if split:                                              # :877-922
                if not infeasible:
                    if infcnt == t.n_obj - 2:
                        if MIN and rhs[t.n_obj - 1] < t.split_stop:
                            infeasible = True
                        if (not MIN) and rhs[t.n_obj - 1] > t.split_stop:
                            infeasible = True
                    for j in range(k):
                        mx[j] = max(mx[j], result[j])
                        mn[j] = min(mn[j], result[j])
                if infeasible:
                    infcnt += 1
                    inflast = True
                else:
                    infcnt = 0
                    inflast = False
            else:                                                  # :1087-1107
                if infeasible:
                    infcnt += 1
                    inflast = True
                else:
                    infcnt = 0
                    inflast = False
                    for j in range(k):
                        mx[j] = max(mx[j], result[j])
                        mn[j] = min(mn[j], result[j])
            # next rhs: the 4-way table (:1575-1832)
            if infeasible and infcnt == oc - 1:
                for j in range(k):
                    rhs[j] = inf_rhs
                if split:
                    rhs[t.n_obj - 1] = t.split_start               # :1649-1651
                step(objective)
                depth_level = 1
                depth = perm[depth_level]
                onwalk = False
            elif inflast and infcnt != oc:
                rhs[depth] = inf_rhs
                depth_level += 1
                depth = perm[depth_level]
                step(depth)
                onwalk = True
            elif (not onwalk) and infcnt != 1:
                step(depth)
            elif onwalk and infcnt != 1:
                depth_level = 1
                depth = perm[depth_level]
                step(depth)
                onwalk = False
    s.sort_unique()                                                # :1877-1879
    all_sols.merge(s)


def split_setup(model: Model, oracle, n_obj: int, num_threads: int, split_normal: bool = False,
                shared_cache: bool = True):
    """EPP recursion (reference src/aira.cpp:1945-1990); returns list of int result vectors."""
    k = model.k
    MIN = model.sense == "MIN"
    p_rhs = [INF if MIN else -INF] * k
    if n_obj == 1:
        r = oracle.get_limit(0, p_rhs)
        return [r]
    sols = split_setup(model, oracle, n_obj - 1, num_threads, split_normal, shared_cache)
    res = oracle.get_limit(n_obj - 1, p_rhs)
    if MIN:
        smallest = res[n_obj - 1]
        biggest = max([INT_MIN] + [s_[n_obj - 1] for s_ in sols])
        if biggest == smallest:
            biggest = INT_MAX
    else:
        biggest = res[n_obj - 1]
        smallest = min([INT_MAX] + [s_[n_obj - 1] for s_ in sols])
        if biggest == smallest:
            smallest = INT_MIN
    return split_optimise(model, oracle, n_obj, biggest, smallest, num_threads, split_normal, shared_cache)


def strip_bounds(MIN, mx, mn, num_threads, split_normal):
    """Strip edges of reference src/aira.cpp:1886-1917 -> list of (split_start, split_stop)."""
    start_point, stop_point = (float(mx), float(mn)) if MIN else (float(mn), float(mx))
    out = []
    split_start = start_point
    step_size = (stop_point - start_point) / num_threads
    for t in range(num_threads):
        if split_normal:
            nv = NORMAL_VALUES[num_threads]
            if MIN:
                gap = start_point - stop_point
                stop = nv[t] * gap + stop_point
                start = nv[t + 1] * gap + stop_point
            else:
                gap = stop_point - start_point
                start = nv[t] * gap + start_point
                stop = nv[t + 1] * gap + start_point
            out.append((start, stop))
        else:
            split_stop = split_start + step_size
            out.append((split_start, split_stop))
            split_start = split_stop
    return out


def split_optimise(model, oracle, n_obj, mx, mn, num_threads, split_normal, shared_cache=True):
    k = model.k
    MIN = model.sense == "MIN"
    here = Solutions(k)
    infeasibles = Solutions(k)
    for (a, b) in strip_bounds(MIN, mx, mn, num_threads, split_normal):
        w = Worker(range(k), n_obj, a, b)
        if shared_cache:
            optimise(model, oracle, here, infeasibles, w, split=True)
        else:
            mine, minf = Solutions(k), Solutions(k)
            optimise(model, oracle, mine, minf, w, split=True)
            here.merge(mine)
    return [list(r.result) for r in here.store if not r.infeasible]


def pareto_front(model: Model, oracle=None, split=False, num_threads=1, split_normal=False,
                 perm=None, trace=None, shared_cache=True):
    """What `main` prints (reference src/aira.cpp:264-346): sorted, deduplicated front rows."""
    oracle = oracle or make_oracle(model)
    k = model.k
    all_sols = Solutions(k)
    if split:
        for r in split_setup(model, oracle, k, num_threads, split_normal, shared_cache):
            all_sols.insert([0.0] * k, r, False)
    else:
        infeasibles = Solutions(k)
        optimise(model, oracle, all_sols, infeasibles, Worker(perm or range(k), k), split=False, trace=trace)
    all_sols.sort_unique()
    return [tuple(r.result) for r in all_sols.store if not r.infeasible]


def brute_force_front(model: Model, fs: FeasibleSet = None):
    """Solver-free non-dominated set, in the reference's output order."""
    fs = fs or FeasibleSet(model)
    sgn = 1 if model.sense == "MIN" else -1
    P = np.unique(fs.P, axis=0)
    keep = []
    for p in P:
        dom = np.all(sgn * P <= sgn * p, axis=1) & np.any(sgn * P < sgn * p, axis=1)
        if not dom.any():
            keep.append(tuple(int(v) for v in p))
    return sorted(keep, key=lambda r: tuple(-v for v in r))


def format_out(front, cpu_s=0.0, wall_s=0.0, ips=0, tag="b200"):
    """`.out` text in the reference's layout (src/aira.cpp:252, 336-358)."""
    lines = ["", f"Using improved algorithm at {tag}"]
    for row in front:
        lines.append("".join(f"{v}\t" for v in row))
    lines += ["", "---", f"{cpu_s:8.3f} CPU seconds", f"{wall_s:8.3f} elapsed seconds",
              f"{ips:8d} IPs solved", f"{len(front):8d} Solutions found", ""]
    return "\n".join(lines)
