// K1 (small-model path) -- batched node-LP relaxations for models whose rows are all dense (knapsack-type:
// the k objective-bound rows plus a few long structural rows) and that have at most 64 columns.
// Same mathematics as the other K1 kernels (reflected restarted Halpern PDHG in fp64, k1_pdhg.cu); what changes
// is the mapping.  A 40-column LP gives a 32-lane warp about one lane-wide instruction of work per column phase
// between five warp-wide reductions, so one warp per node (k1_fast.cu) is bound by shuffle and dependent-issue
// latency: 1.57 us per iteration, 0.14 of the streaming HBM roofline (profiles/r01_sweep_batch.md).  Here
//   * 8 lanes own one node LP and a warp iterates 4 node LPs in lockstep; every group pulls its next node from the
//     batch counter on its own, so a finished group does not wait for its neighbours;
//   * lane g of a group owns columns g, g+8, ...: their dense-row values stay in registers for the life of the CTA,
//     the reflected point and the anchor stay in registers for the life of the node, {l,u} and the PDHG point sit
//     in a thread-private shared slot;
//   * the KD row sums and the two restart norms are reduced together by one transposed butterfly over the 8 lanes
//     (7 shuffles instead of 8 x 3): lane r ends up with the total of value r, and lane r is the owner of row r,
//     so the dual step needs no further exchange; the new duals come back to the columns by KD lane broadcasts;
//   * there is no block-level barrier and no shared-memory reduction anywhere in the iteration.
// Control flow is warp-uniform (conditions are agreed with __any_sync and applied per group through predicates),
// so every shuffle runs with the full mask.
#include <cfloat>
#include <cmath>
#include <cstdlib>

#include "device.h"

namespace moip {
namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kG = 8;                       // lanes per node LP

__device__ __forceinline__ double shx(double v, int o) { return __shfl_xor_sync(kFull, v, o); }
__device__ __forceinline__ double shi(double v, int src) { return __shfl_sync(kFull, v, src); }
__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double clampd(double v, double a, double b) { return dmin(dmax(v, a), b); }

// Sums 8 per-lane values over the 8 lanes of a group: lane L of the group returns the total of v[L].
__device__ __forceinline__ double reduce8(const double (&v)[8], int gl) {
  const bool b2 = gl & 4, b1 = gl & 2, b0 = gl & 1;
  double w[4], z[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = (b2 ? v[4 + i] : v[i]) + shx(b2 ? v[i] : v[4 + i], 4);
#pragma unroll
  for (int i = 0; i < 2; ++i) z[i] = (b1 ? w[2 + i] : w[i]) + shx(b1 ? w[i] : w[2 + i], 2);
  return (b0 ? z[1] : z[0]) + shx(b0 ? z[0] : z[1], 1);
}

constexpr int kHalpTab = 2048;
struct HalpTab {
  double v[kHalpTab];
  constexpr HalpTab() : v() {
    for (int i = 0; i < kHalpTab; ++i) v[i] = (double)(i + 1) / (double)(i + 2);
  }
};
__constant__ HalpTab c_halp_small = HalpTab();
__device__ __forceinline__ double halpern_weight(int kk) {
  return kk < kHalpTab ? c_halp_small.v[kk] : (double)(kk + 1) / (double)(kk + 2);
}

enum { COLD_BEST_LB = 0, COLD_POBJ, COLD_DOBJ, COLD_OBJ_UPPER, COLD_KKT_BINV, COLD_W, COLD_R0SQ, COLD_RPREV, COLD_N = 8 };

template <int KD, int CPT, int MINB>
__global__ void __launch_bounds__(128, MINB)
k1_small_kernel(const DevModel dm, const LpBatch b, const LpParams p) {
  constexpr int NT = 128;
  static_assert(KD + 2 <= 8, "row sums and norms share one 8-value reduction");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* lu = reinterpret_cast<double2*>(smem_raw);        // [CPT][NT] thread-private {l,u}
  double* xts = reinterpret_cast<double*>(lu + CPT * NT);    // [CPT][NT] thread-private PDHG point xt
  double* cold = xts + CPT * NT;                             // [NT/8][8] per-group scalars of rare use
  const int tid = threadIdx.x, lane = tid & 31, gl = lane & 7, gbase = lane & ~7;
  const int n = dm.n, k = dm.k;
  double2* lu_t = lu + tid;
  double* xts_t = xts + tid;
  double* cold_g = cold + (tid >> 3) * COLD_N;
  const int norm_mask = p.norm_every - 1;   // power of two (host guarantees)
  const bool row_lane = gl < KD;            // this lane owns kernel row gl (k objective rows, then the long rows)
  const double dr_mine = row_lane ? dm.dr_k[gl] : 1.0;

  // ---- the model, once per CTA: dense-row values of this lane's columns
  double rc[CPT][KD];
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int j = gl + kG * c;
#pragma unroll
    for (int d = 0; d < KD; ++d) rc[c][d] = j < n ? dm.D2[(size_t)d * n + j] : 0.0;
  }

  // ---- per-group state (identical in the 8 lanes of a group unless noted)
  bool alive = true, need_load = true, live = false;
  int node = 0, cost = 0, it = 0, kk = 0, status = MOIP_LP_ITERLIMIT, next_check = 0;
  size_t srow = 0;
  double inv_dr_cost = 1.0, tau = 0, sigma = 1, inv_sigma = 1, ah = 0;
  double xb[CPT], xa[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) { xb[c] = 0; xa[c] = 0; }
  // row state (per lane: the row this lane owns)
  double r_nlo = HUGE_VAL, r_nhi = -HUGE_VAL, r_y = 0, r_ya = 0, r_sx = 0, r_sxa = 0, r_yt = 0, r_sxt = 0;
  const int iter_cap = p.fixed_iters > 0 ? p.fixed_iters : p.max_iter;
  const int batchB = b.B_dev ? *b.B_dev : b.B;          // chained rounds: the batch size lives on the device

  for (;;) {
    // ================================================================ node load
    if (__any_sync(kFull, need_load)) {
      int nd = 0;
      if (need_load && gl == 0) {
        for (;;) {
          nd = atomicAdd(b.work_counter, 1);
          if (nd >= batchB || !(b.skip && b.skip[nd])) break;
          b.status[nd] = -1; b.iters[nd] = 0;            // already decided by K2
        }
      }
      nd = __shfl_sync(kFull, nd, gbase);
      const bool ld = need_load && nd < batchB;
      if (need_load && nd >= batchB) alive = false;
      need_load = false;
      unsigned act = 0;
      const double* nrhs = b.rhs;
      if (ld) {
        node = nd;
        srow = b.slot ? (size_t)b.slot[node] : (size_t)(b.slot_base + node);
        cost = b.cost_idx[(size_t)node * b.cost_stride];
        nrhs = b.rhs + (size_t)node * b.rhs_stride;
        inv_dr_cost = 1.0 / dm.dr_k[cost];
        for (int o = 0; o < k; ++o)
          if (fabs(nrhs[o]) < 1e19) act |= 1u << o;
        for (int t = k; t < KD; ++t) act |= 1u << t;
      }
      double v8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v8[i] = 0;
      if (ld) {
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          const int j = gl + kG * c;
          double lj = 0, uj = 0, xj = 0;
          if (j < n) {
            const double idc = 1.0 / dm.dc[j];
            lj = (double)b.lb[srow * n + j] * idc;
            uj = (double)b.ub[srow * n + j] * idc;
            xj = b.warm_x ? b.warm_x[srow * n + j] * idc : 0.0;
            xj = clampd(xj, lj, uj);
          }
          lu_t[c * NT] = make_double2(lj, uj);
          xts_t[c * NT] = xj;
          xb[c] = xj; xa[c] = xj;
          double cj = 0;
#pragma unroll
          for (int d = 0; d < KD; ++d) { if (d == cost) cj = rc[c][d]; v8[d] = fma(rc[c][d], xj, v8[d]); }
          cj *= inv_dr_cost;
          v8[KD] = fma(cj, cj, v8[KD]);
          v8[KD + 1] += dmax(cj * lj, cj * uj);
        }
      }
      const double q0 = reduce8(v8, gl);               // lane d < KD: (S x0)_d, lane KD: |c|^2, lane KD+1: max c x on the box
      const double cn2 = shi(q0, gbase + KD), cup = shi(q0, gbase + KD + 1);
      if (ld) {
        live = false;
        r_nlo = HUGE_VAL; r_nhi = -HUGE_VAL; r_y = 0; r_ya = 0; r_sx = 0; r_sxa = 0; r_yt = 0; r_sxt = 0;
        if (row_lane) {
          double lo, hi;
          if (gl < k) { lo = -HUGE_VAL; hi = ((act >> gl) & 1u) ? dm.sgn * nrhs[gl] * dr_mine : HUGE_VAL; }
          else { lo = dm.lo_k[gl]; hi = dm.hi_k[gl]; }
          double yi = b.warm_y ? b.warm_y[srow * dm.m + gl] / dr_mine : 0.0;
          if (lo == -HUGE_VAL) yi = dmin(yi, 0.0);
          if (hi == HUGE_VAL) yi = dmax(yi, 0.0);
          live = (act >> gl) & 1u;
          if (!live) yi = 0.0;
          r_nlo = -lo; r_nhi = -hi;
          r_y = yi; r_ya = yi; r_sx = q0; r_sxa = q0;
        }
        // primal weight w = |c| / |b| (scaled), unscaled |b| for the KKT denominator
        double bn2 = 0, bn2_unscaled = dm.norm_row_bounds2;
        for (int i = 0; i < KD; ++i) {
          double loi, hii;
          if (i < k) {
            loi = -HUGE_VAL;
            hii = ((act >> i) & 1u) ? dm.sgn * nrhs[i] * dm.dr_k[i] : HUGE_VAL;
            if ((act >> i) & 1u) bn2_unscaled += nrhs[i] * nrhs[i];
          } else { loi = dm.lo_k[i]; hii = dm.hi_k[i]; }
          const double t = (hii != HUGE_VAL) ? hii : ((loi != -HUGE_VAL) ? loi : 0.0);
          bn2 += t * t;
        }
        const double w = (cn2 > 0 && bn2 > 0) ? sqrt(cn2 / bn2) : 1.0;
        tau = dm.eta / w; sigma = dm.eta * w; inv_sigma = 1.0 / sigma;
        if (gl == 0) {
          cold_g[COLD_W] = w; cold_g[COLD_R0SQ] = 0.0; cold_g[COLD_RPREV] = -1.0;
          cold_g[COLD_BEST_LB] = -HUGE_VAL; cold_g[COLD_POBJ] = 0.0; cold_g[COLD_DOBJ] = -HUGE_VAL;
          cold_g[COLD_OBJ_UPPER] = cup; cold_g[COLD_KKT_BINV] = 1.0 / (1.0 + sqrt(bn2_unscaled));
        }
        it = 0; kk = 0; ah = 0.0; status = MOIP_LP_ITERLIMIT;
        next_check = p.fixed_iters > 0 ? 0x7fffffff : p.check_every;
      }
      __syncwarp();
    }
    if (!__any_sync(kFull, alive)) break;

    // ================================================================ one PDHG iteration of 4 node LPs
    ++it;
    const bool norm_it = alive && (kk & norm_mask) == 0;
    const bool check_it = alive && it == next_check;
    const bool last_it = alive && it >= iter_cap;
    const bool eval_it = check_it || last_it;
    const bool any_norm = __any_sync(kFull, norm_it);
    const bool any_eval = __any_sync(kFull, eval_it);
    // duals of the dense rows back to the columns; the objective rides on its own row: y_cost - 1/dr_cost
    double yd[KD];
    {
      const double ypub = row_lane ? r_y - (gl == cost ? inv_dr_cost : 0.0) : 0.0;
#pragma unroll
      for (int d = 0; d < KD; ++d) yd[d] = shi(ypub, gbase + d);
    }
    double v8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v8[i] = 0;
    {
      const double ah1 = 1.0 - ah;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const double2 bx = lu_t[c * NT];
        double g = 0;
#pragma unroll
        for (int d = 0; d < KD; ++d) g = fma(rc[c][d], yd[d], g);
        const double xaj = xa[c];
        const double xj = fma(ah, xb[c], ah1 * xaj);
        const double xtj = clampd(fma(tau, g, xj), bx.x, bx.y);
        const double xbn = fma(2.0, xtj, -xj);
        xb[c] = xbn;
        if (any_norm || any_eval) {                     // (warp-uniform) xt is only read on such iterations
          xts_t[c * NT] = xtj;
          const double d1 = xtj - xj, d2 = xtj - xaj;
          v8[KD] = fma(d1, d1, v8[KD]);
          v8[KD + 1] = fma(d2, d2, v8[KD + 1]);
        }
#pragma unroll
        for (int d = 0; d < KD; ++d) v8[d] = fma(rc[c][d], xbn, v8[d]);
      }
    }
    const double q = reduce8(v8, gl);                  // lane r < KD: (S xbar)_r; lanes KD, KD+1: the two norms
    // ---- dual step of the row this lane owns
    r_sxt = 0.5 * (q + r_sx);
    {
      const double v = fma(r_y, inv_sigma, -q);
      r_yt = live ? sigma * (v - clampd(v, r_nhi, r_nlo)) : 0.0;
    }
    bool restart = false;
    if (any_norm) {
      __syncwarp();                                     // cold scalars written by lane 0 earlier are visible
      double a8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a8[i] = 0;
      if (live && norm_it) {
        const double dy = r_yt - r_y, dya = r_yt - r_ya;
        a8[0] = dy * dy; a8[1] = dy * (r_sxt - r_sx); a8[2] = dya * dya;
      }
      const double s = reduce8(a8, gl);
      const double ab0 = shi(s, gbase), ab1 = shi(s, gbase + 1), ab2 = shi(s, gbase + 2);
      const double nx1 = shi(q, gbase + KD), nx2 = shi(q, gbase + KD + 1);
      if (norm_it) {
        const double wc = cold_g[COLD_W], r0sq = cold_g[COLD_R0SQ], rprev = cold_g[COLD_RPREV];
        const double fp2 = dmax(0.0, fma(wc / dm.eta, nx1, fma(-2.0, ab1, ab0 * inv_sigma)));
        if (kk != 0 && (fp2 <= 0.04 * r0sq || (fp2 <= 0.64 * r0sq && rprev >= 0.0 && fp2 > rprev) || 25 * kk >= 9 * it))
          restart = true;
        double w = wc;
        if (restart) {
          const double dxn = sqrt(nx2), dyn = sqrt(ab2);
          if (dxn > 1e-10 && dyn > 1e-10) w = exp(0.5 * log(dyn / dxn) + 0.5 * log(w));
          tau = dm.eta / w; sigma = dm.eta * w; inv_sigma = 1.0 / sigma;
        }
        __syncwarp();                                   // every lane of the group has read the old values
        if (gl == 0) {
          if (kk == 0) cold_g[COLD_R0SQ] = fp2;
          cold_g[COLD_RPREV] = restart ? -1.0 : fp2;
          if (restart) cold_g[COLD_W] = w;
        }
      } else {
        __syncwarp();
      }
    }
    // ---- termination tests at (xt, yt); on the last iteration they also produce the outputs
    bool stop = last_it;
    if (any_eval) {
      __syncwarp();
      double ytd[KD];
      {
        const double ytp = row_lane ? r_yt : 0.0;
#pragma unroll
        for (int d = 0; d < KD; ++d) ytd[d] = shi(ytp, gbase + d);
      }
      double c8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) c8[i] = 0;             // pobj, dual (columns), dual (rows), primal residual^2 (unscaled)
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const double2 bx = lu_t[c * NT];
        double g = 0, cj = 0;
#pragma unroll
        for (int d = 0; d < KD; ++d) { g = fma(rc[c][d], ytd[d], g); if (d == cost) cj = rc[c][d]; }
        cj *= inv_dr_cost;
        const double r = cj - g;
        c8[0] = fma(cj, xts_t[c * NT], c8[0]);
        c8[1] += (r > 0) ? r * bx.x : r * bx.y;
        const double fk = (g < 0) ? -g * bx.x : -g * bx.y;          // the same bound with the objective dropped (Farkas)
        c8[4] += fk; c8[5] += fabs(fk);
      }
      if (live) {
        if (r_yt > 0) c8[2] = -r_yt * r_nlo;
        else if (r_yt < 0) c8[2] = -r_yt * r_nhi;
        c8[5] += fabs(c8[2]);
        const double viol = dmax(0.0, dmax(r_sxt + r_nhi, -r_nlo - r_sxt)) / dr_mine;
        c8[3] = viol * viol;
      }
      const double s = reduce8(c8, gl);
      const double pobj = shi(s, gbase), dobj = shi(s, gbase + 1) + shi(s, gbase + 2), pres2 = shi(s, gbase + 3);
      // Farkas certificate (see k1_reg.cuh): F(y) > 0 proves the node LP infeasible
      const bool farkas = shi(s, gbase + 4) + shi(s, gbase + 2) > 1e-9 * shi(s, gbase + 5) + 1e-9;
      // one lane reads the cutoff for its group: with chained rounds other CTAs lower it while this one runs, and the 8
      // lanes of a node must take the same way out
      const double cutoff_g = shi(b.cutoff ? *((volatile const double*)b.cutoff) : HUGE_VAL, gbase);
      if (eval_it) {
        if (check_it) next_check += p.check_every;
        double best_lb = cold_g[COLD_BEST_LB];
        const double obj_upper = cold_g[COLD_OBJ_UPPER], kkt_binv = cold_g[COLD_KKT_BINV];
        if (p.fixed_iters > 0) best_lb = dobj;
        else {
          if (dobj > best_lb) best_lb = dobj;
          const double gap = fabs(pobj - dobj);
          const double rel = dmax(sqrt(pres2) * kkt_binv, gap / (1.0 + fabs(pobj) + fabs(dobj)));
          const double cutoff = cutoff_g;
          if (best_lb >= cutoff - p.cutoff_slack) { status = MOIP_LP_CUTOFF; stop = true; }
          else if (farkas || best_lb > obj_upper + 1e-6 * (1.0 + fabs(obj_upper))) { status = MOIP_LP_INFEASIBLE; stop = true; }
          else if (rel <= p.eps) { status = MOIP_LP_CONVERGED; stop = true; }
          else if (p.int_obj && sqrt(pres2) * kkt_binv <= 1e-5 && ceil(best_lb - 1e-6) >= ceil(pobj - 1e-3)) {
            status = MOIP_LP_CONVERGED; stop = true;      // the integer-rounded bound cannot improve any further
          }
        }
        __syncwarp();
        if (gl == 0) { cold_g[COLD_BEST_LB] = best_lb; cold_g[COLD_POBJ] = pobj; cold_g[COLD_DOBJ] = dobj; }
      } else {
        __syncwarp();
      }
      __syncwarp();
    }
    // ---- restart or Halpern step (groups that stop keep their PDHG point for the output)
    if (!stop) {
      if (restart) {
#pragma unroll
        for (int c = 0; c < CPT; ++c) xa[c] = xts_t[c * NT];
        r_y = r_yt; r_ya = r_yt; r_sx = r_sxt; r_sxa = r_sxt;
        kk = 0; ah = 0.0;
      } else {
        ah = halpern_weight(kk);
        const double c1 = 1.0 - ah;
        r_y = fma(ah, 2.0 * r_yt - r_y, c1 * r_ya);
        r_sx = fma(ah, 2.0 * r_sxt - r_sx, c1 * r_sxa);
        ++kk;
      }
    }

    // ================================================================ node store
    if (__any_sync(kFull, stop)) {
      __syncwarp();
      const double ytp = row_lane ? r_yt : 0.0;
      double ytd[KD];
#pragma unroll
      for (int d = 0; d < KD; ++d) ytd[d] = shi(ytp, gbase + d);
      double xo[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) xo[c] = 0.0;
      if (stop) {
        const double pobj = cold_g[COLD_POBJ], best_lb = cold_g[COLD_BEST_LB], dobj_last = cold_g[COLD_DOBJ];
        // reduced-cost tightening against the cutoff (valid for any dual iterate, see k1_fast.cu)
        if (b.rc_fix && b.cutoff && status != MOIP_LP_CUTOFF && status != MOIP_LP_INFEASIBLE) {
          const double cutoff = *((volatile const double*)b.cutoff);
          const double room = cutoff - p.cutoff_slack - dobj_last;
          if (cutoff < HUGE_VAL && room >= 0.0 && dobj_last > -HUGE_VAL) {
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
              const int j = gl + kG * c;
              if (j < n) {
                double g = 0, cj = 0;
#pragma unroll
                for (int d = 0; d < KD; ++d) { g = fma(rc[c][d], ytd[d], g); if (d == cost) cj = rc[c][d]; }
                const double r = (cj * inv_dr_cost - g) / dm.dc[j];        // unscaled reduced cost
                int lbj = b.lb[srow * n + j], ubj = b.ub[srow * n + j];
                if (lbj < ubj) {
                  if (r > 1e-9) {
                    const double t = floor(room / r + 1e-9);
                    if (t < (double)(ubj - lbj)) { ubj = lbj + (int)t; b.ub[srow * n + j] = ubj; }
                  } else if (r < -1e-9) {
                    const double t = floor(room / (-r) + 1e-9);
                    if (t < (double)(ubj - lbj)) { lbj = ubj - (int)t; b.lb[srow * n + j] = lbj; }
                  }
                }
              }
            }
          }
        }
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          const int j = gl + kG * c;
          xo[c] = (j < n) ? xts_t[c * NT] * dm.dc[j] : 0.0;
          if (b.out_x && j < n) b.out_x[srow * n + j] = xo[c];
        }
        if (b.out_y && row_lane) b.out_y[srow * dm.m + gl] = r_yt * dr_mine;
        if (gl == 0) {
          b.primal_obj[node] = pobj;
          b.dual_bound[node] = best_lb;
          b.status[node] = status;
          b.iters[node] = it;
        }
      }
      if (b.branch_var) {           // the three most fractional columns, best first (group-wide arg max)
        int c0 = -1, c1 = -1;
        for (int r = 0; r < 3; ++r) {
          double bestf = -1.0, bestv = 0.0; int bestj = -1;
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            const int j = gl + kG * c;
            const double f = fabs(xo[c] - rint(xo[c]));
            if (stop && j < n && j != c0 && j != c1 && f > bestf) { bestf = f; bestj = j; bestv = xo[c]; }
          }
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) {
            const double of = shx(bestf, o), ov = shx(bestv, o);
            const int oj = __shfl_xor_sync(kFull, bestj, o);
            if (of > bestf || (of == bestf && oj >= 0 && (bestj < 0 || oj < bestj))) { bestf = of; bestj = oj; bestv = ov; }
          }
          const int pick = (bestf > 1e-6) ? bestj : -1;
          if (stop && gl == 0) {
            b.branch_var[(size_t)node * 3 + r] = pick;
            if (b.branch_val) b.branch_val[(size_t)node * 3 + r] = (pick >= 0) ? bestv : 0.0;
          }
          if (r == 0) c0 = pick; else c1 = pick;
        }
      }
      if (stop) need_load = true;
    }
  }
}


template <int KD, int CPT, int MINB>
int launch_small(const DevModel& dm, const LpBatch& b, LpParams p, int num_sms, cudaStream_t st) {
  auto kern = k1_small_kernel<KD, CPT, MINB>;
  const size_t smem = (size_t)CPT * 128 * (sizeof(double2) + sizeof(double)) + (size_t)(128 / 8) * COLD_N * sizeof(double);
  static LaunchCfg cfg;
  std::unique_lock<std::mutex> cfg_lock(launch_cfg_mutex());
  const int dev = current_device();
  size_t& configured = cfg.configured[dev];
  int& occ = cfg.occ[dev];
  if (configured == 0) {
    MOIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MOIP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, smem));
    if (occ < 1) return MOIP_ERR_LIMIT;
    configured = smem;
  }
  const int occ_now = occ;
  cfg_lock.unlock();
  const long long groups = (b.B + 15) / 16;          // 16 node LPs in flight per CTA
  long long grid = (long long)num_sms * occ_now;
  if (grid > groups) grid = groups;
  if (grid < 1) grid = 1;
  int ne = 1;                                        // restart-test cadence: largest power of two <= norm_every
  while (ne * 2 <= p.norm_every) ne *= 2;
  p.norm_every = ne;
  kern<<<(unsigned)grid, 128, smem, st>>>(dm, b, p);
  MOIP_CUDA(cudaGetLastError());
  return MOIP_OK;
}

template <int KD>
int launch_small_kd(const DevModel& dm, const LpBatch& b, const LpParams& p, int num_sms, cudaStream_t st) {
  if (dm.n <= 16) return launch_small<KD, 2, 4>(dm, b, p, num_sms, st);
  if (dm.n <= 40) return launch_small<KD, 5, 3>(dm, b, p, num_sms, st);
  return launch_small<KD, 8, 2>(dm, b, p, num_sms, st);
}

}  // namespace

bool k1_small_applies(const DevModel& dm) {
  return dm.fast_ok && dm.msS == 0 && dm.ell2_w == 0 && dm.KD >= 3 && dm.KD <= 5 && dm.n <= 64 && dm.m == dm.KD &&
         !std::getenv("MOIP_K1_NOSMALL");
}

int launch_k1_small(const DevModel& dm, const LpBatch& b, const LpParams& p, int num_sms, cudaStream_t st) {
  if (b.B <= 0) return MOIP_OK;
  if (!b.B_dev) MOIP_CUDA(cudaMemsetAsync(b.work_counter, 0, sizeof(int), st));   // (chained rounds: K5 resets it)
  switch (dm.KD) {
    case 3: return launch_small_kd<3>(dm, b, p, num_sms, st);
    case 4: return launch_small_kd<4>(dm, b, p, num_sms, st);
    case 5: return launch_small_kd<5>(dm, b, p, num_sms, st);
  }
  return MOIP_ERR_UNSUPPORTED;
}

}  // namespace moip
