"""TEST INFRASTRUCTURE ONLY -- numpy prototype used to choose the PDHG variant for kernel K1.

Not imported by the product.  `python -m oracle.pdhg_proto` prints iteration counts of the
candidate variants on node LPs of the synthetic instances against HiGHS.
"""
from __future__ import annotations

import sys
import time

import numpy as np

from .lpformat import INF, Model, synthetic_ap, synthetic_kp


def node_lp(model: Model, cost_idx, rhs, fix=None):
    """min c x  s.t. lo <= K x <= hi, l <= x <= u   (MAX models are negated)."""
    sgn = 1.0 if model.sense == "MIN" else -1.0
    K = np.vstack([model.A, sgn * model.C])
    lo = np.concatenate([np.where(np.array(model.row_sense) == "L", -np.inf, model.b),
                         np.full(model.k, -np.inf)])
    hi = np.concatenate([np.where(np.array(model.row_sense) == "G", np.inf, model.b),
                         np.where(np.abs(rhs) < 1e19, sgn * np.asarray(rhs, float), np.inf)])
    c = sgn * model.C[cost_idx]
    l = model.lb.copy()
    u = np.where(model.ub >= INF, np.inf, model.ub)
    if fix is not None:
        for j, v in fix.items():
            l[j] = u[j] = v
    return K, lo, hi, c, l, u


def highs(K, lo, hi, c, l, u):
    from scipy.optimize import linprog
    A_ub, b_ub, A_eq, b_eq = [], [], [], []
    for i in range(K.shape[0]):
        if lo[i] == hi[i]:
            A_eq.append(K[i]); b_eq.append(lo[i])
        else:
            if np.isfinite(hi[i]):
                A_ub.append(K[i]); b_ub.append(hi[i])
            if np.isfinite(lo[i]):
                A_ub.append(-K[i]); b_ub.append(-lo[i])
    r = linprog(c, A_ub=np.array(A_ub) if A_ub else None, b_ub=b_ub or None,
                A_eq=np.array(A_eq) if A_eq else None, b_eq=b_eq or None,
                bounds=list(zip(l, u)), method="highs")
    return r.status, (r.fun if r.status == 0 else None), r.x


def ruiz(K, iters=10):
    m, n = K.shape
    dr, dc = np.ones(m), np.ones(n)
    S = K.copy()
    for _ in range(iters):
        r = np.sqrt(np.abs(S).max(axis=1)); r[r == 0] = 1
        cc = np.sqrt(np.abs(S).max(axis=0)); cc[cc == 0] = 1
        S = S / r[:, None] / cc[None, :]
        dr /= r; dc /= cc
    # Pock-Chambolle alpha=1
    r = np.sqrt(np.abs(S).sum(axis=1)); r[r == 0] = 1
    cc = np.sqrt(np.abs(S).sum(axis=0)); cc[cc == 0] = 1
    S = S / r[:, None] / cc[None, :]
    dr /= r; dc /= cc
    return S, dr, dc


def dual_bound(K, lo, hi, c, l, u, y):
    """Lagrangian bound (valid for any y whose sign pattern respects finite sides)."""
    r = c - K.T @ y
    yp, ym = np.maximum(y, 0), np.maximum(-y, 0)
    with np.errstate(invalid="ignore"):
        t = np.where(yp > 0, yp * lo, 0).sum() - np.where(ym > 0, ym * hi, 0).sum()
        t += np.where(r > 0, r * l, 0).sum() + np.where(r < 0, r * u, 0).sum()
    return t


def kkt(K, lo, hi, c, l, u, x, y):
    Kx = K @ x
    pr = np.linalg.norm(Kx - np.clip(Kx, lo, hi))
    r = c - K.T @ y
    # reduced costs that can be absorbed by finite bounds
    lam = np.where(r > 0, np.where(np.isfinite(l), r, 0), np.where(np.isfinite(u), r, 0))
    dr = np.linalg.norm(r - lam)
    pobj = c @ x
    dobj = dual_bound(K, lo, hi, c, l, u, y)
    bn = np.linalg.norm(np.where(np.isfinite(lo), lo, 0)) + np.linalg.norm(np.where(np.isfinite(hi), hi, 0))
    rel = max(pr / (1 + bn), dr / (1 + np.linalg.norm(c)), abs(pobj - dobj) / (1 + abs(pobj) + abs(dobj)))
    return rel, pobj, dobj


def T(S, lo, hi, c, l, u, x, y, tau, sigma):
    xn = np.clip(x - tau * (c - S.T @ y), l, u)
    xb = 2 * xn - x
    v = y / sigma - S @ xb
    yn = sigma * (v - np.clip(v, -hi, -lo))
    return xn, yn


def solve(K, lo, hi, c, l, u, variant="halpern", eps=1e-6, max_iter=100000, check=40, verbose=False,
          ref=None, x0=None, y0=None):
    S, dr, dc = ruiz(K)
    # scaled problem: x = dc * xs, y = dr * ys ; Ks = diag(dr) K diag(dc)
    cs, ls, us = c * dc, l / dc, u / dc
    los, his = lo * dr, hi * dr
    nrm = np.linalg.norm(S, 2)
    eta = 0.99 / nrm
    bn = np.linalg.norm(np.where(np.isfinite(his), his, 0)) + np.linalg.norm(np.where(np.isfinite(los), los, 0))
    cn = np.linalg.norm(cs)
    w = cn / bn if bn > 0 and cn > 0 else 1.0
    m, n = S.shape
    x = np.clip(np.zeros(n) if x0 is None else x0 / dc, ls, us)
    y = np.zeros(m) if y0 is None else y0 / dr
    hist = []

    def mnorm2(dx, dy, w):
        return (w / eta) * dx @ dx - 2 * dy @ (S @ dx) + dy @ dy / (eta * w)

    it = 0
    if variant == "avg":
        x0r, y0r = x.copy(), y.copy()
        xs_, ys_ = np.zeros(n), np.zeros(m)
        cnt = 0
        last_restart_err = None
        prev_err = None
        while it < max_iter:
            x, y = T(S, los, his, cs, ls, us, x, y, eta / w, eta * w)
            xs_ += x; ys_ += y; cnt += 1; it += 1
            if it % check == 0:
                xa, ya = xs_ / cnt, ys_ / cnt
                ea, pa, da = kkt(S, los, his, cs, ls, us, xa, ya)
                ec, pc, dcur = kkt(S, los, his, cs, ls, us, x, y)
                if ea < ec:
                    e, xc, yc, pp, dd = ea, xa, ya, pa, da
                else:
                    e, xc, yc, pp, dd = ec, x, y, pc, dcur
                hist.append((it, e, pp, dd))
                if e <= eps:
                    return it, xc * dc, yc * dr, pp, dd, hist
                if last_restart_err is None:
                    last_restart_err = kkt(S, los, his, cs, ls, us, x0r, y0r)[0]
                do = e <= 0.2 * last_restart_err or (e <= 0.8 * last_restart_err and prev_err is not None and e > prev_err) or cnt >= 0.36 * it
                prev_err = e
                if do:
                    dx, dy = np.linalg.norm(xc - x0r), np.linalg.norm(yc - y0r)
                    if dx > 1e-10 and dy > 1e-10:
                        w = np.exp(0.5 * np.log(dy / dx) + 0.5 * np.log(w))
                    x, y = xc.copy(), yc.copy()
                    x0r, y0r = x.copy(), y.copy()
                    xs_[:] = 0; ys_[:] = 0; cnt = 0
                    last_restart_err = e; prev_err = None
        return it, x * dc, y * dr, None, None, hist
    # reflected restarted Halpern PDHG
    xa, ya = x.copy(), y.copy()          # anchor
    kk = 0
    r0 = None
    rprev = None
    while it < max_iter:
        xt, yt = T(S, los, his, cs, ls, us, x, y, eta / w, eta * w)
        it += 1
        fp = np.sqrt(max(mnorm2(xt - x, yt - y, w), 0.0))
        if kk == 0:
            r0 = fp
        do_check = it % check == 0
        restart = False
        if kk > 0:
            if fp <= 0.2 * r0 or (fp <= 0.8 * r0 and rprev is not None and fp > rprev) or kk >= 0.36 * it:
                restart = True
        rprev = fp
        if do_check or restart:
            e, pp, dd = kkt(S, los, his, cs, ls, us, xt, yt)
            hist.append((it, e, pp, dd))
            if e <= eps:
                return it, xt * dc, yt * dr, pp, dd, hist
        if restart:
            dx, dy = np.linalg.norm(xt - xa), np.linalg.norm(yt - ya)
            if dx > 1e-10 and dy > 1e-10:
                w = np.exp(0.5 * np.log(dy / dx) + 0.5 * np.log(w))
            x, y = xt, yt
            xa, ya = x.copy(), y.copy()
            kk = 0
            rprev = None
            continue
        rho = 2.0 if variant == "halpern" else 1.0   # reflection
        a = (kk + 1.0) / (kk + 2.0)
        x = a * (x + rho * (xt - x)) + (1 - a) * xa
        y = a * (y + rho * (yt - y)) + (1 - a) * ya
        kk += 1
    return it, xt * dc, yt * dr, None, None, hist


def sample_nodes(model, count, seed=7, maxdepth=20, loosen=0.3):
    """Node batches of SURVEY.md section 8d item 6."""
    rng = np.random.default_rng(seed)
    # ideal/nadir-ish ranges via the LP relaxation of each objective alone is overkill; use data ranges
    sgn = 1.0 if model.sense == "MIN" else -1.0
    out = []
    k = model.k
    # per-objective ideal via LP
    ideal, nadir = [], []
    vals = []
    for j in range(k):
        K, lo, hi, c, l, u = node_lp(model, j, [INF * sgn] * k)
        st, f, x = highs(K, lo, hi, c, l, u)
        vals.append(model.C @ x)
    vals = np.array(vals)
    for _ in range(count):
        d = int(rng.integers(0, maxdepth + 1))
        js = rng.choice(model.n, size=d, replace=False)
        if len(model.row_sense) > 1:      # assignment: fixings consistent with a random permutation
            nn_ = int(round(model.n ** 0.5))
            pi = rng.permutation(nn_)
            fix = {int(j): float(pi[int(j) // nn_] == int(j) % nn_) for j in js}
        else:
            fix = {int(j): float(rng.integers(0, 2)) for j in js}
        ci = int(rng.integers(0, k))
        rhs = []
        for j in range(k):
            a, b = vals[:, j].min(), vals[:, j].max()
            rhs.append(float(np.floor(rng.uniform(a + loosen * (b - a), b))))
        rhs[ci] = INF * sgn          # the optimised objective is unbounded in stage 0
        out.append((ci, rhs, fix))
    return out


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "ap"
    nn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    model = synthetic_ap(nn, 3, 1) if which == "ap" else synthetic_kp(nn, 4, 1)
    nodes = sample_nodes(model, 6)
    for ci, rhs, fix in nodes:
        K, lo, hi, c, l, u = node_lp(model, ci, rhs, fix)
        st, f, xh = highs(K, lo, hi, c, l, u)
        print(f"node cost={ci} rhs={rhs} nfix={len(fix)} highs status={st} obj={f}")
        if st != 0:
            for variant in ("halpern",):
                it, x, y, pp, dd, hist = solve(K, lo, hi, c, l, u, variant=variant, max_iter=3000)
                print("   infeasible; bound trajectory:", [(h[0], round(h[3], 2)) for h in hist[::10]][:8])
            continue
        for variant in ("avg", "halpern"):
            t = time.time()
            it, x, y, pp, dd, hist = solve(K, lo, hi, c, l, u, variant=variant)
            db = dual_bound(K, lo, hi, c, l, u, y)
            # iterations until the Lagrangian bound is within 0.5 / 0.05 of the optimum
            def first(tol):
                for h in hist:
                    if f - h[3] <= tol:
                        return h[0]
                return None
            print(f"   {variant:8s} iters={it:6d} pobj={pp} dobj={dd} relerr={abs((pp or 0)-f)/max(1,abs(f)):.2e} "
                  f"bound(0.5)@{first(0.5)} bound(0.05)@{first(0.05)} t={time.time()-t:.1f}s")
