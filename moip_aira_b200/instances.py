"""Synthetic workloads of BASELINE.json / SURVEY.md section 8d (items 4-6), written in the reference's
extended LP dialect so the same file feeds every path.  numpy only; no solver code here."""
from __future__ import annotations

import math

import numpy as np


def _expr(row, names):
    parts = [f"{'+' if v >= 0 else '-'} {abs(int(v))} {names[j]}" for j, v in enumerate(row) if v != 0]
    lines, cur = [], ""
    for p in parts:
        if len(cur) + len(p) > 200:
            lines.append(cur)
            cur = ""
        cur += " " + p
    lines.append(cur)
    return "\n".join(lines)


def write_ap(path, n, k, seed):
    """k-objective assignment problem, n*n binaries `XiXj`, costs i.i.d. U{0..19} (section 8d item 4)."""
    rng = np.random.default_rng(seed)
    C = rng.integers(0, 20, size=(k, n * n))
    names = [f"X{i + 1}X{j + 1}" for i in range(n) for j in range(n)]
    with open(path, "w") as fh:
        fh.write("\\ synthetic assignment instance\nMinimize 0\nsubject to\n")
        for i in range(n):
            fh.write(" + ".join(names[i * n + j] for j in range(n)) + " = 1\n")
        for j in range(n):
            fh.write(" + ".join(names[i * n + j] for i in range(n)) + " = 1\n")
        for o in range(k):
            # every column appears in the assignment rows above, so zero costs can be dropped
            fh.write(f"{_expr(C[o], names)} < {o + 1}\n")
        fh.write("BINARY\n" + "\n".join(names) + "\nEND\n")
    return C


def write_kp(path, n, k, seed):
    """k-objective binary knapsack, w and values U{10..100}, capacity floor(sum w / 2) (item 5)."""
    rng = np.random.default_rng(seed)
    w = rng.integers(10, 101, size=n)
    V = rng.integers(10, 101, size=(k, n))
    names = [f"x{i}" for i in range(n)]
    with open(path, "w") as fh:
        fh.write("\\ synthetic knapsack instance\nMaximize 0\nsubject to\n")
        fh.write(f"{_expr(w, names)} <= {math.floor(w.sum() / 2)}\n")
        for o in range(k):
            fh.write(f"{_expr(V[o], names)} > {o + 1}\n")
        fh.write("BINARY\n" + "\n".join(names) + "\nEND\n")
    return w, V


def sample_node_batch(ctx, B, seed=7, maxdepth=20, loosen=0.3):
    """Node batch of section 8d item 6: random depth-d fixing (d ~ U{0..20}; for assignment models the
    fixings are consistent with a random permutation so that most nodes stay feasible), rhs drawn
    between the ideal and nadir LP values of every bounded objective, cost index ~ U{0..k-1}.
    The ideal/nadir values come from k root LPs solved on the GPU through the same C ABI."""
    pr = ctx.problem
    k, n = pr.objcnt, pr.n
    is_min = pr.objsen == 0
    inf = 1e20 if is_min else -1e20
    rng = np.random.default_rng(seed)
    root = ctx.lp_batch_solve(np.arange(k), np.full((k, k), inf), None, ctx.lp_params(eps=1e-6), want_x=True)
    vals = root["x"] @ pr.objcoef.T                # vals[j][o] = objective o at the optimiser of objective j
    words = pr.mask_words
    cost = rng.integers(0, k, size=B).astype(np.int32)
    rhs = np.zeros((B, k))
    masks = np.zeros((B, words), dtype=np.uint32)
    nn = int(round(n ** 0.5))
    is_ap = pr.ms == 2 * nn and nn * nn == n
    for b in range(B):
        d = int(rng.integers(0, maxdepth + 1))
        js = rng.choice(n, size=min(d, n), replace=False)
        if is_ap:
            pi = rng.permutation(nn)
            fixv = [int(pi[j // nn] == j % nn) for j in js]
        else:
            fixv = rng.integers(0, 2, size=len(js))
        for j, v in zip(js, fixv):
            masks[b, j >> 4] |= np.uint32((2 + int(v)) << ((j & 15) * 2))
        for o in range(k):
            a, c = vals[:, o].min(), vals[:, o].max()
            rhs[b, o] = math.floor(rng.uniform(a + loosen * (c - a), c)) if is_min else math.ceil(rng.uniform(a, c - loosen * (c - a)))
        rhs[b, cost[b]] = inf
    return cost, rhs, masks
