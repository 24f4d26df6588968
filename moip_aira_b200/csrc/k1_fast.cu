// K1 (fast path) -- batched node-LP relaxations, one CTA per B&B node, reflected restarted Halpern
// PDHG in fp64 (same mathematics as k1_pdhg.cu, which stays as the generic fallback).
//
// Design, driven by the ncu profiles under profiles/ (v1: 17.9k, v2: 7.2k warp-instructions per
// node-iteration, both issue-bound on address arithmetic, not on fp64 or memory):
//   * rows are reordered [short structural | k objective | long structural]; every row is owned by
//     one group of LPR (1/2/4) lanes whose leader keeps the row's dual state (y, anchor, S x, bounds)
//     in REGISTERS for the whole solve; only the published y lives in shared memory;
//   * the column state is an array of 5-double records {xbar, anchor, xt, l, u} in shared memory
//     (one address per column, conflict-free 64-bit accesses: stride 10 words);
//   * the model is read through packed records: one 16-byte-aligned record per column
//     (ELL values | dense-row values | ELL row ids) fetched with 128-bit loads, and {value, column}
//     pairs for the short rows (row-ELL, coalesced in the row index, thread-per-row: no 32-lane
//     shuffle reduction per row);
//   * dense rows (the k objective-bound rows and structural rows longer than max(32, n/8), e.g. a
//     knapsack capacity row) are accumulated inside the column pass and block-reduced once;
//   * the Halpern average of x is folded into the next column pass (x is never stored); an ordinary
//     iteration is: column pass, one block reduction, row phase, one barrier;
//   * restart tests run every `norm_every` iterations on squared norms (no sqrt / division).
// HBM is touched once per node (bounds / warm start in, iterate out); the model image is shared by
// all CTAs and served from L1/L2.
#include <cfloat>
#include <cmath>
#include <cstdlib>

#include "device.h"

namespace moip {
namespace {

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// generic block sum (all threads get the totals); one barrier inside
template <int NV, int NT>
__device__ __forceinline__ void bsum(double (&v)[NV], double* red, int tid) {
  constexpr int NW = NT / 32;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = wsum(v[i]);
  if (NW == 1) { __syncthreads(); return; }
  if ((tid & 31) == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[(tid >> 5) * NV + i] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += red[w * NV + i];
    v[i] = s;
  }
}

__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }   // no NaN semantics needed
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double clampd(double v, double a, double b) { return dmin(dmax(v, a), b); }

constexpr int CS = 5;   // doubles per column record in shared memory: xbar, xa, xt, l, u
enum { C_XBAR = 0, C_XA = 1, C_XT = 2, C_L = 3, C_U = 4 };

template <int ELLW, int KD>
struct ColRec {
  static constexpr int UC = ELLW + KD + (ELLW + 1) / 2;
  static constexpr int UP = (UC + 1) & ~1;
  double u[UP];
  __device__ __forceinline__ void load(const double* __restrict__ base, int j) {
    const double2* p = reinterpret_cast<const double2*>(base) + (size_t)j * (UP / 2);
#pragma unroll
    for (int q = 0; q < UP / 2; ++q) { const double2 t = __ldg(p + q); u[2 * q] = t.x; u[2 * q + 1] = t.y; }
  }
  __device__ __forceinline__ double ell(int e) const { return u[e]; }
  __device__ __forceinline__ double dense(int d) const { return u[ELLW + d]; }
  __device__ __forceinline__ int row(int e) const {
    const double w = u[ELLW + KD + e / 2];
    return (e & 1) ? __double2hiint(w) : __double2loint(w);
  }
};

template <int NT, int KD, int ELLW, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k1_fast_kernel(const DevModel dm, const LpBatch b, const LpParams p, const int lpr_log2) {
  constexpr int NW = NT / 32;
  extern __shared__ double smem[];
  const int n = dm.n, msS = dm.msS, m = dm.m, k = dm.k, RW = dm.RW;
  double* col = smem;               // [n][CS]
  double* ysh = col + (size_t)n * CS;   // [m] current dual iterate, published for the column pass
  double* ytsh = ysh + m;           // [m] dual PDHG point (termination tests only)
  double* redA = ytsh + m;          // NW * 8
  double* redB = redA + NW * 8;     // NW * 8
  double* redC = redB + NW * 8;     // NW * 8
  __shared__ int s_node;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double eta = dm.eta;
  const int LPR = 1 << lpr_log2;
  const int row = tid >> lpr_log2, sub = tid & (LPR - 1);
  const bool leader = sub == 0 && row < m;
  const bool ell_row = row < msS;           // this group owns a short structural row
  const int dd = row - msS;                 // dense index of this group's row (if >= 0)
  const double* __restrict__ crec = dm.colrec;
  const double2* __restrict__ rrec = reinterpret_cast<const double2*>(dm.rowrec);
  const int norm_mask = p.norm_every - 1;   // norm_every is a power of two (host guarantees)

  for (;;) {
    if (tid == 0) s_node = atomicAdd(b.work_counter, 1);
    __syncthreads();
    const int node = s_node;
    if (node >= b.B) break;
    if (b.skip && b.skip[node]) {
      if (tid == 0) { b.status[node] = -1; b.iters[node] = 0; }
      __syncthreads();
      continue;
    }
    // ---------------------------------------------------------------- node load
    const size_t srow = b.slot ? (size_t)b.slot[node] : (size_t)node;   // row of the node's arrays
    const int cost = b.cost_idx[(size_t)node * b.cost_stride];
    const double* nrhs = b.rhs + (size_t)node * b.rhs_stride;
    const double inv_dr_cost = 1.0 / dm.dr_k[msS + cost];
    unsigned act = 0;                 // active dense rows (finite bound)
    for (int o = 0; o < k; ++o)
      if (fabs(nrhs[o]) < 1e19) act |= 1u << o;
    for (int t = k; t < KD; ++t) act |= 1u << t;
    double a0[KD + 2];
#pragma unroll
    for (int d = 0; d < KD + 2; ++d) a0[d] = 0;
    for (int j = tid; j < n; j += NT) {
      ColRec<ELLW, KD> rc;
      rc.load(crec, j);
      const double idc = 1.0 / dm.dc[j];
      const double lj = (double)b.lb[srow * n + j] * idc;
      const double uj = (double)b.ub[srow * n + j] * idc;
      double xj = b.warm_x ? b.warm_x[srow * n + j] * idc : 0.0;
      xj = clampd(xj, lj, uj);
      double* cj_ = col + j * CS;
      cj_[C_XBAR] = xj; cj_[C_XA] = xj; cj_[C_XT] = xj; cj_[C_L] = lj; cj_[C_U] = uj;
      double cj = 0;
#pragma unroll
      for (int d = 0; d < KD; ++d) { if (d == cost) cj = rc.dense(d); a0[d] = fma(rc.dense(d), xj, a0[d]); }
      cj *= inv_dr_cost;
      a0[KD] = fma(cj, cj, a0[KD]);
      a0[KD + 1] += dmax(cj * lj, cj * uj);
    }
    bsum<KD + 2, NT>(a0, redA, tid);      // barrier: column records visible
    const double obj_upper = a0[KD + 1];
    // row state in registers (leaders)
    double r_lo = -HUGE_VAL, r_hi = HUGE_VAL, r_y = 0, r_ya = 0, r_sx = 0, r_sxa = 0, r_idr = 1.0, r_yt = 0, r_sxt = 0;
    bool live = false;                    // leader of a row that can carry a nonzero dual
    {
      double q = 0;                       // S x0 of this group's short row (all lanes take part in the shuffles)
      if (ell_row)
        for (int e = sub; e < RW; e += LPR) { const double2 t = __ldg(rrec + e * msS + row); q = fma(t.x, col[__double2loint(t.y) * CS + C_XA], q); }
      for (int o = LPR >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      if (row < m) {
        const double dri = dm.dr_k[row];
        r_idr = 1.0 / dri;
        if (dd >= 0 && dd < k) { r_lo = -HUGE_VAL; r_hi = ((act >> dd) & 1u) ? dm.sgn * nrhs[dd] * dri : HUGE_VAL; }
        else { r_lo = dm.lo_k[row]; r_hi = dm.hi_k[row]; }
        double yi = b.warm_y ? b.warm_y[srow * m + row] * r_idr : 0.0;
        if (r_lo == -HUGE_VAL) yi = dmin(yi, 0.0);
        if (r_hi == HUGE_VAL) yi = dmax(yi, 0.0);
        r_y = yi; r_ya = yi;
        if (!ell_row) {
#pragma unroll
          for (int d = 0; d < KD; ++d) if (d == dd) q = a0[d];
        }
        r_sx = q; r_sxa = q;
        live = leader && (dd < 0 || ((act >> dd) & 1u));
        if (sub == 0) ysh[row] = yi;
      }
    }
    // primal weight w = |c| / |b| (scaled), unscaled |b| for the KKT denominator (uniform, m small)
    double bn2 = 0, bn2_unscaled = dm.norm_row_bounds2;
    for (int i = 0; i < m; ++i) {
      const int di = i - msS;
      double loi, hii;
      if (di >= 0 && di < k) {
        loi = -HUGE_VAL;
        hii = ((act >> di) & 1u) ? dm.sgn * nrhs[di] * dm.dr_k[i] : HUGE_VAL;
        if ((act >> di) & 1u) bn2_unscaled += nrhs[di] * nrhs[di];
      } else { loi = dm.lo_k[i]; hii = dm.hi_k[i]; }
      const double t = (hii != HUGE_VAL) ? hii : ((loi != -HUGE_VAL) ? loi : 0.0);
      bn2 += t * t;
    }
    double w = (a0[KD] > 0 && bn2 > 0) ? sqrt(a0[KD] / bn2) : 1.0;
    double tau = eta / w, sigma = eta * w, inv_sigma = 1.0 / sigma, w_over_eta = w / eta, inv_eta_w = 1.0 / (eta * w);
    const double kkt_binv = 1.0 / (1.0 + sqrt(bn2_unscaled));
    __syncthreads();                       // ysh visible

    int kk = 0, it = 0, status = MOIP_LP_ITERLIMIT;
    double r0sq = 0, rprev = -1.0, best_lb = -HUGE_VAL, pobj = 0, dobj_last = -HUGE_VAL;
    const int iter_cap = p.fixed_iters > 0 ? p.fixed_iters : p.max_iter;
    int next_check = p.fixed_iters > 0 ? 0x7fffffff : p.check_every;

    // ---------------------------------------------------------------- PDHG iterations
    for (;;) {
      ++it;
      const bool norm_it = (kk & norm_mask) == 0;
      const bool check_it = it == next_check;
      const bool last_it = it >= iter_cap;
      const double ah = (double)kk * __drcp_rn((double)(kk + 1)), ah1 = 1.0 - ah;   // x = ah*xbar + (1-ah)*xa
      // ---- column pass
      double yd[KD];
#pragma unroll
      for (int d = 0; d < KD; ++d) yd[d] = ((act >> d) & 1u) ? ysh[msS + d] : 0.0;
      double aA[KD + 2];
#pragma unroll
      for (int d = 0; d < KD + 2; ++d) aA[d] = 0;
#pragma unroll 2
      for (int j = tid; j < n; j += NT) {
        ColRec<ELLW, KD> rc;
        rc.load(crec, j);
        double* cj_ = col + j * CS;
        const double xbo = cj_[C_XBAR], xaj = cj_[C_XA], lj = cj_[C_L], uj = cj_[C_U];
        double g = 0, cj = 0;
#pragma unroll
        for (int e = 0; e < ELLW; ++e) g = fma(rc.ell(e), ysh[rc.row(e)], g);
#pragma unroll
        for (int d = 0; d < KD; ++d) { g = fma(rc.dense(d), yd[d], g); if (d == cost) cj = rc.dense(d); }
        cj *= inv_dr_cost;
        const double xj = fma(ah, xbo, ah1 * xaj);
        const double xtj = clampd(fma(-tau, cj - g, xj), lj, uj);
        const double xb = 2.0 * xtj - xj;
        cj_[C_XBAR] = xb; cj_[C_XT] = xtj;
        if (norm_it) {
          const double d1 = xtj - xj, d2 = xtj - xaj;
          aA[KD] = fma(d1, d1, aA[KD]);
          aA[KD + 1] = fma(d2, d2, aA[KD + 1]);
        }
#pragma unroll
        for (int d = 0; d < KD; ++d) aA[d] = fma(rc.dense(d), xb, aA[d]);
      }
      // block reduction of the dense-row products (+ the two norms on norm iterations)
#pragma unroll
      for (int d = 0; d < KD; ++d) if ((act >> d) & 1u) aA[d] = wsum(aA[d]);
      if (norm_it) { aA[KD] = wsum(aA[KD]); aA[KD + 1] = wsum(aA[KD + 1]); }
      if (NW > 1 && lane == 0) {
#pragma unroll
        for (int d = 0; d < KD + 2; ++d) redA[warp * 8 + d] = aA[d];
      }
      __syncthreads();                     // xbar visible, warp sums visible
      if (NW > 1 && norm_it) {
        double s0 = 0, s1 = 0;
#pragma unroll
        for (int wq = 0; wq < NW; ++wq) { s0 += redA[wq * 8 + KD]; s1 += redA[wq * 8 + KD + 1]; }
        aA[KD] = s0; aA[KD + 1] = s1;
      }
      // ---- row phase
      double aB[3] = {0, 0, 0};
      {
        double q = 0, q2 = 0;
        if (ell_row) {
          int e = sub;
          for (; e + LPR < RW; e += 2 * LPR) {
            const double2 t0 = __ldg(rrec + e * msS + row), t1 = __ldg(rrec + (e + LPR) * msS + row);
            q = fma(t0.x, col[__double2loint(t0.y) * CS + C_XBAR], q);
            q2 = fma(t1.x, col[__double2loint(t1.y) * CS + C_XBAR], q2);
          }
          if (e < RW) { const double2 t0 = __ldg(rrec + e * msS + row); q = fma(t0.x, col[__double2loint(t0.y) * CS + C_XBAR], q); }
          q += q2;
        }
        for (int o = LPR >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        if (live) {
          if (dd >= 0) {
            if (NW > 1) { q = 0;
#pragma unroll
              for (int wq = 0; wq < NW; ++wq) q += redA[wq * 8 + dd]; }
            else {
#pragma unroll
              for (int d = 0; d < KD; ++d) if (d == dd) q = aA[d];
            }
          }
          r_sxt = 0.5 * (q + r_sx);
          const double v = fma(r_y, inv_sigma, -q);
          r_yt = sigma * (v - clampd(v, -r_hi, -r_lo));
          if (norm_it) {
            const double dy = r_yt - r_y, dya = r_yt - r_ya;
            aB[0] = dy * dy; aB[1] = dy * (r_sxt - r_sx); aB[2] = dya * dya;
          }
        }
      }
      bool restart = false;
      if (norm_it) {
        bsum<3, NT>(aB, redB, tid);
        const double fp2 = dmax(0.0, fma(w_over_eta, aA[KD], fma(-2.0, aB[1], aB[0] * inv_eta_w)));
        if (kk == 0) r0sq = fp2;
        else if (fp2 <= 0.04 * r0sq || (fp2 <= 0.64 * r0sq && rprev >= 0.0 && fp2 > rprev) || 25 * kk >= 9 * it)
          restart = true;
        rprev = fp2;
      }
      const bool need_stop_eval = check_it || last_it;
      if (live && need_stop_eval) ytsh[row] = r_yt;
      // ---- termination tests at (xt, yt); on the last iteration they also produce the outputs
      bool stop = last_it;
      if (need_stop_eval) {
        if (check_it) next_check += p.check_every;
        __syncthreads();                   // ytsh visible
        double aC[4] = {0, 0, 0, 0};       // pobj, dual (columns), dual (rows), primal residual^2 (unscaled)
        double ytd[KD];
#pragma unroll
        for (int d = 0; d < KD; ++d) ytd[d] = ((act >> d) & 1u) ? ytsh[msS + d] : 0.0;
        for (int j = tid; j < n; j += NT) {
          ColRec<ELLW, KD> rc;
          rc.load(crec, j);
          const double* cj_ = col + j * CS;
          double g = 0, cj = 0;
#pragma unroll
          for (int e = 0; e < ELLW; ++e) g = fma(rc.ell(e), ytsh[rc.row(e)], g);
#pragma unroll
          for (int d = 0; d < KD; ++d) { g = fma(rc.dense(d), ytd[d], g); if (d == cost) cj = rc.dense(d); }
          cj *= inv_dr_cost;
          const double r = cj - g;
          aC[0] = fma(cj, cj_[C_XT], aC[0]);
          aC[1] += (r > 0) ? r * cj_[C_L] : r * cj_[C_U];
        }
        if (live) {
          if (r_yt > 0) aC[2] = r_yt * r_lo;
          else if (r_yt < 0) aC[2] = r_yt * r_hi;
          const double viol = dmax(0.0, dmax(r_sxt - r_hi, r_lo - r_sxt)) * r_idr;
          aC[3] = viol * viol;
        }
        bsum<4, NT>(aC, redC, tid);
        pobj = aC[0];
        const double dobj = aC[1] + aC[2];
        dobj_last = dobj;
        if (p.fixed_iters > 0) best_lb = dobj;
        else {
          if (dobj > best_lb) best_lb = dobj;
          const double gap = fabs(pobj - dobj);
          const double rel = dmax(sqrt(aC[3]) * kkt_binv, gap / (1.0 + fabs(pobj) + fabs(dobj)));
          const double cutoff = b.cutoff ? *((volatile const double*)b.cutoff) : HUGE_VAL;
          if (best_lb >= cutoff - p.cutoff_slack) { status = MOIP_LP_CUTOFF; stop = true; }
          else if (best_lb > obj_upper + 1e-6 * (1.0 + fabs(obj_upper))) { status = MOIP_LP_INFEASIBLE; stop = true; }
          else if (rel <= p.eps) { status = MOIP_LP_CONVERGED; stop = true; }
          else if (p.int_obj && sqrt(aC[3]) * kkt_binv <= 1e-5 && ceil(best_lb - 1e-6) >= ceil(pobj - 1e-3)) {
            status = MOIP_LP_CONVERGED; stop = true;      // the integer-rounded bound cannot improve any further
          }
        }
      }
      if (stop) break;
      // ---- restart or Halpern step of the row state; x follows in the next column pass
      if (restart) {
        const double dxn = sqrt(aA[KD + 1]), dyn = sqrt(aB[2]);
        if (dxn > 1e-10 && dyn > 1e-10) w = exp(0.5 * log(dyn / dxn) + 0.5 * log(w));
        tau = eta / w; sigma = eta * w; inv_sigma = 1.0 / sigma; w_over_eta = w / eta; inv_eta_w = 1.0 / (eta * w);
        for (int j = tid; j < n; j += NT) col[j * CS + C_XA] = col[j * CS + C_XT];
        if (live) { r_y = r_yt; r_ya = r_yt; r_sx = r_sxt; r_sxa = r_sxt; ysh[row] = r_y; }
        kk = 0; rprev = -1.0;
      } else {
        if (live) {
          const double a = (double)(kk + 1) * __drcp_rn((double)(kk + 2)), c1 = 1.0 - a;
          r_y = fma(a, 2.0 * r_yt - r_y, c1 * r_ya);
          r_sx = fma(a, 2.0 * r_sxt - r_sx, c1 * r_sxa);
          ysh[row] = r_y;
        }
        ++kk;
      }
      __syncthreads();                     // ysh / anchors visible; column records free for the next pass
    }

    // ---------------------------------------------------------------- node store
    // Reduced-cost tightening: with the Lagrangian bound L(yt) = dobj_last and reduced costs r, any
    // point with x_j >= l_j + t has objective >= L + r_j t (r_j > 0), so columns can be tightened against
    // the cutoff (strictly better integer solutions have objective <= cutoff - slack).  Valid for any yt.
    if (b.rc_fix && b.cutoff && status != MOIP_LP_CUTOFF && status != MOIP_LP_INFEASIBLE) {
      const double cutoff = *((volatile const double*)b.cutoff);
      const double room = cutoff - p.cutoff_slack - dobj_last;
      if (cutoff < HUGE_VAL && room >= 0.0 && dobj_last > -HUGE_VAL) {
        double ytd[KD];
#pragma unroll
        for (int d = 0; d < KD; ++d) ytd[d] = ((act >> d) & 1u) ? ytsh[msS + d] : 0.0;
        for (int j = tid; j < n; j += NT) {
          ColRec<ELLW, KD> rc;
          rc.load(crec, j);
          double g = 0, cj = 0;
#pragma unroll
          for (int e = 0; e < ELLW; ++e) g = fma(rc.ell(e), ytsh[rc.row(e)], g);
#pragma unroll
          for (int d = 0; d < KD; ++d) { g = fma(rc.dense(d), ytd[d], g); if (d == cost) cj = rc.dense(d); }
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
    if (b.out_x)
      for (int j = tid; j < n; j += NT) b.out_x[srow * n + j] = col[j * CS + C_XT] * dm.dc[j];
    if (b.out_y && leader) b.out_y[srow * m + row] = (live ? r_yt : 0.0) * dm.dr_k[row];
    if (b.branch_var) {           // the three most fractional columns, best first
      int c0 = -1, c1 = -1;
      for (int r = 0; r < 3; ++r) {
        double bestf = -1.0; int bestj = -1;
        for (int j = tid; j < n; j += NT) {
          const double v = col[j * CS + C_XT] * dm.dc[j];
          const double f = fabs(v - rint(v));
          if (j != c0 && j != c1 && f > bestf) { bestf = f; bestj = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double of = __shfl_xor_sync(0xffffffffu, bestf, o);
          const int oj = __shfl_xor_sync(0xffffffffu, bestj, o);
          if (of > bestf || (of == bestf && oj >= 0 && (bestj < 0 || oj < bestj))) { bestf = of; bestj = oj; }
        }
        __syncthreads();
        if (lane == 0) { redA[warp * 2] = bestf; redA[warp * 2 + 1] = (double)bestj; }
        __syncthreads();
        for (int wq = 0; wq < NW; ++wq) {     // every thread folds the warp results: uniform outcome
          const double of = redA[wq * 2]; const int oj = (int)redA[wq * 2 + 1];
          if (of > bestf || (of == bestf && oj >= 0 && (bestj < 0 || oj < bestj))) { bestf = of; bestj = oj; }
        }
        const int pick = (bestf > 1e-6) ? bestj : -1;
        if (tid == 0) {
          b.branch_var[(size_t)node * 3 + r] = pick;
          if (b.branch_val) b.branch_val[(size_t)node * 3 + r] = (pick >= 0) ? col[pick * CS + C_XT] * dm.dc[pick] : 0.0;
        }
        if (r == 0) c0 = pick; else c1 = pick;
      }
    }
    if (tid == 0) {
      b.primal_obj[node] = pobj;
      b.dual_bound[node] = best_lb;
      b.status[node] = status;
      b.iters[node] = it;
    }
    __syncthreads();
  }
}

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

template <int NT, int KD, int ELLW, int MINB>
int launch_fast(const DevModel& dm, const LpBatch& b, LpParams p, int num_sms, cudaStream_t st) {
  auto kern = k1_fast_kernel<NT, KD, ELLW, MINB>;
  const size_t smem = sizeof(double) * ((size_t)CS * dm.n + (size_t)2 * dm.m + (size_t)24 * (NT / 32));
  static LaunchCfg cfg;
  std::unique_lock<std::mutex> cfg_lock(launch_cfg_mutex());
  const int dev = current_device();
  size_t& configured = cfg.configured[dev];
  int& occ = cfg.occ[dev];
  if (smem > configured) {
    MOIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // leave the rest of the 256 KB array to L1: the packed model image (tens of KB) is re-read by every
    // CTA every iteration and must hit there
    MOIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    MOIP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    if (occ < 1) { std::fprintf(stderr, "moip_b200: node LP does not fit in shared memory (n=%d)\n", dm.n); return MOIP_ERR_LIMIT; }
    const int cap = env_int("MOIP_K1_OCC", 0);
    if (cap > 0 && cap < occ) occ = cap;
    int pct = (int)(((smem + 1024) * occ * 100 + 228 * 1024 - 1) / (228 * 1024));
    if (pct > 100) pct = 100;
    pct = env_int("MOIP_K1_CARVEOUT_PCT", pct);
    MOIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    configured = smem;
  }
  const int occ_now = occ;
  cfg_lock.unlock();
  long long grid = (long long)num_sms * occ_now;
  if (grid > b.B) grid = b.B;
  if (grid < 1) grid = 1;
  int lpr_log2 = 0;
  while (lpr_log2 < 2 && (dm.m << (lpr_log2 + 1)) <= NT) ++lpr_log2;
  int ne = 1;                              // restart-test cadence: largest power of two <= norm_every
  while (ne * 2 <= p.norm_every) ne *= 2;
  p.norm_every = ne;
  kern<<<(unsigned)grid, NT, smem, st>>>(dm, b, p, lpr_log2);
  MOIP_CUDA(cudaGetLastError());
  return MOIP_OK;
}

#define MOIP_KD_CASES(NT, E, MB) case 1: return launch_fast<NT, 1, E, MB>(dm, b, p, num_sms, st); case 2: return launch_fast<NT, 2, E, MB>(dm, b, p, num_sms, st); \
  case 3: return launch_fast<NT, 3, E, MB>(dm, b, p, num_sms, st); case 4: return launch_fast<NT, 4, E, MB>(dm, b, p, num_sms, st); \
  case 5: return launch_fast<NT, 5, E, MB>(dm, b, p, num_sms, st); case 6: return launch_fast<NT, 6, E, MB>(dm, b, p, num_sms, st);

template <int NT, int ELLW, int MINB>
int launch_kd(const DevModel& dm, const LpBatch& b, const LpParams& p, int num_sms, cudaStream_t st) {
  switch (dm.KD) { MOIP_KD_CASES(NT, ELLW, MINB) }
  return MOIP_ERR_UNSUPPORTED;
}

template <int NT, int MINB>
int launch_ell(const DevModel& dm, const LpBatch& b, const LpParams& p, int num_sms, cudaStream_t st) {
  switch (dm.ell2_w) {
    case 0: return launch_kd<NT, 0, MINB>(dm, b, p, num_sms, st);
    case 1: return launch_kd<NT, 1, MINB>(dm, b, p, num_sms, st);
    case 2: return launch_kd<NT, 2, MINB>(dm, b, p, num_sms, st);
  }
  return MOIP_ERR_UNSUPPORTED;
}

}  // namespace

int launch_k1_fast(const DevModel& dm, const LpBatch& b, const LpParams& p, int num_sms, cudaStream_t st) {
  if (b.B <= 0) return MOIP_OK;
  MOIP_CUDA(cudaMemsetAsync(b.work_counter, 0, sizeof(int), st));
  if (dm.n <= 64 && dm.m <= 32) return launch_ell<32, 16>(dm, b, p, num_sms, st);
  return launch_ell<128, 4>(dm, b, p, num_sms, st);     // m <= 128 guaranteed by Model::fast_ok
}

}  // namespace moip
