"""TEST INFRASTRUCTURE ONLY -- ctypes access to oracle/_build/liboracle.so (pdhg_ref.c) and the
HiGHS stand-in for LP values (NOT CPLEX).  Never imported by the product."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .lpformat import INF, Model

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-s", "-C", _HERE, "oracle_c"])
        _LIB = C.CDLL(path)
        pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int)
        _LIB.pdhg_ref_batch.restype = C.c_int
        _LIB.pdhg_ref_batch.argtypes = [C.c_int, C.c_int, pd, C.c_int, pd, pd, pd, pd, pd, C.c_double, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, pd, pd, pd, pi, pi]
    return _LIB


def unpack_masks(model: Model, masks):
    """2-bit/var fixing masks -> (l, u) arrays [B][n] (0 free, 2 at lower bound, 3 at upper bound)."""
    masks = np.asarray(masks, dtype=np.uint32)
    B = masks.shape[0]
    l = np.tile(model.lb, (B, 1))
    u = np.tile(np.where(model.ub >= INF, np.inf, model.ub), (B, 1))
    for j in range(model.n):
        c = (masks[:, j >> 4] >> ((j & 15) * 2)) & 3
        u[c == 2, j] = l[c == 2, j]
        l[c == 3, j] = u[c == 3, j]
    return l, u


def node_lps(model: Model, cost_idx, rhs, masks=None):
    """Min-form dense LP data of a node batch: K (m x n), c[B][n], lo/hi[B][m], l/u[B][n]."""
    sgn = 1.0 if model.sense == "MIN" else -1.0
    B = len(cost_idx)
    K = np.vstack([model.A, sgn * model.C]).astype(np.float64)
    rs = np.array(model.row_sense)
    slo = np.where(rs == "L", -np.inf, model.b) if model.ms else np.zeros(0)
    shi = np.where(rs == "G", np.inf, model.b) if model.ms else np.zeros(0)
    rhs = np.asarray(rhs, dtype=np.float64).reshape(B, model.k)
    lo = np.hstack([np.tile(slo, (B, 1)), np.full((B, model.k), -np.inf)])
    hi = np.hstack([np.tile(shi, (B, 1)), np.where(np.abs(rhs) < 1e19, sgn * rhs, np.inf)])
    c = sgn * model.C[np.asarray(cost_idx)]
    if masks is None:
        masks = np.zeros((B, (model.n + 15) // 16), dtype=np.uint32)
    l, u = unpack_masks(model, masks)
    return K, c, lo, hi, l, u


def pdhg_ref(model: Model, cost_idx, rhs, masks=None, eps=1e-8, max_iter=100000, check_every=32, fixed_iters=0,
             norm_every=1, cutoff=np.inf, cutoff_slack=0.0, threads=0):
    """Runs the C restatement; returns objective values in the MODEL's sense like the C ABI does."""
    K, c, lo, hi, l, u = node_lps(model, cost_idx, rhs, masks)
    B = len(cost_idx)
    m, n = K.shape
    sgn = 1.0 if model.sense == "MIN" else -1.0
    x = np.zeros((B, n))
    pobj, lb = np.zeros(B), np.zeros(B)
    iters, status = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
    pd, pi = C.POINTER(C.c_double), C.POINTER(C.c_int)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (K, c, lo, hi, l, u)]
    _lib().pdhg_ref_batch(m, n, arrs[0].ctypes.data_as(pd), B, arrs[1].ctypes.data_as(pd), arrs[2].ctypes.data_as(pd),
                          arrs[3].ctypes.data_as(pd), arrs[4].ctypes.data_as(pd), arrs[5].ctypes.data_as(pd),
                          eps, max_iter, check_every, fixed_iters, norm_every,
                          float(cutoff) if np.isfinite(cutoff) else float("inf"), cutoff_slack, threads,
                          x.ctypes.data_as(pd), pobj.ctypes.data_as(pd), lb.ctypes.data_as(pd),
                          iters.ctypes.data_as(pi), status.ctypes.data_as(pi))
    return {"primal_obj": sgn * pobj, "dual_bound": sgn * lb, "status": status, "iters": iters, "x": x}


def highs_lp(model: Model, cost_idx, rhs, masks=None):
    """HiGHS stand-in: returns (status[B] 0 ok / 2 infeasible, objective[B] in the model's sense)."""
    from scipy.optimize import linprog
    K, c, lo, hi, l, u = node_lps(model, cost_idx, rhs, masks)
    sgn = 1.0 if model.sense == "MIN" else -1.0
    B = len(cost_idx)
    st, obj = np.zeros(B, dtype=np.int32), np.full(B, np.nan)
    for b in range(B):
        eq = lo[b] == hi[b]
        A_ub, b_ub = [], []
        fin_hi = np.isfinite(hi[b]) & ~eq
        fin_lo = np.isfinite(lo[b]) & ~eq
        if fin_hi.any():
            A_ub.append(K[fin_hi]); b_ub.append(hi[b][fin_hi])
        if fin_lo.any():
            A_ub.append(-K[fin_lo]); b_ub.append(-lo[b][fin_lo])
        r = linprog(c[b], A_ub=np.vstack(A_ub) if A_ub else None, b_ub=np.concatenate(b_ub) if b_ub else None,
                    A_eq=K[eq] if eq.any() else None, b_eq=lo[b][eq] if eq.any() else None,
                    bounds=list(zip(l[b], u[b])), method="highs")
        st[b] = r.status
        if r.status == 0:
            obj[b] = sgn * r.fun
    return st, obj


def sample_node_batch(model: Model, B, seed=7, maxdepth=20, loosen=0.3, feasible_bias=True):
    """Node batches of SURVEY.md section 8d item 6: root LP + random depth-d fixing + rhs drawn between
    ideal and nadir of each bounded objective + random cost index.  Returns (cost_idx, rhs, masks)."""
    rng = np.random.default_rng(seed)
    k, n = model.k, model.n
    sgn = 1.0 if model.sense == "MIN" else -1.0
    free = np.full((k, k), INF * sgn)
    _, _ = None, None
    from scipy.optimize import linprog
    vals = []
    for j in range(k):
        K, c, lo, hi, l, u = node_lps(model, [j], free[j:j + 1])
        eq = lo[0] == hi[0]
        fin = np.isfinite(hi[0]) & ~eq
        r = linprog(c[0], A_ub=K[fin] if fin.any() else None, b_ub=hi[0][fin] if fin.any() else None,
                    A_eq=K[eq] if eq.any() else None, b_eq=lo[0][eq] if eq.any() else None,
                    bounds=list(zip(l[0], u[0])), method="highs")
        vals.append(model.C @ r.x)
    vals = np.array(vals)
    words = (n + 15) // 16
    cost = rng.integers(0, k, size=B).astype(np.int32)
    rhs = np.zeros((B, k))
    masks = np.zeros((B, words), dtype=np.uint32)
    is_ap = model.ms > 1 and int(round(n ** 0.5)) ** 2 == n
    for b in range(B):
        d = int(rng.integers(0, maxdepth + 1))
        js = rng.choice(n, size=min(d, n), replace=False)
        if is_ap and feasible_bias:
            nn = int(round(n ** 0.5))
            pi = rng.permutation(nn)
            fixv = [int(pi[j // nn] == j % nn) for j in js]
        else:
            fixv = rng.integers(0, 2, size=len(js))
        for j, v in zip(js, fixv):
            masks[b, j >> 4] |= np.uint32((2 + int(v)) << ((j & 15) * 2))
        for o in range(k):
            a, c_ = vals[:, o].min(), vals[:, o].max()
            if sgn > 0:
                rhs[b, o] = np.floor(rng.uniform(a + loosen * (c_ - a), c_))
            else:
                rhs[b, o] = np.ceil(rng.uniform(a, c_ - loosen * (c_ - a)))
        rhs[b, cost[b]] = INF * sgn
    return cost, rhs, masks
