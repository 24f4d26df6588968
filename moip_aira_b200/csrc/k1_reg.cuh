// K1 (register-resident path) -- batched node-LP relaxations, one CTA per B&B node, reflected
// restarted Halpern PDHG in fp64.  Same mathematics as k1_fast.cu / k1_pdhg.cu; what changes is
// where the data lives.  The ncu profile of k1_fast (profiles/r01_k1_v3_summary.md) showed the SM's
// L1/shared data pipe at 76 % of peak: ~1 410 wavefronts per node-iteration, of which 337 re-read the
// packed column records, 337 moved the column state through shared memory and ~600 belonged to the
// row-ELL pass (record loads + gathers of xbar).  Here
//   * every thread owns CPT columns (j = tid + c*NT) for the whole life of the CTA: their packed model
//     records are loaded ONCE per CTA into registers, and the column state (reflected point, anchor,
//     PDHG point) never leaves registers; only the box {l,u} sits in a thread-private shared slot;
//   * S*xbar of the short rows is formed from per-entry products that the column pass scatters into
//     prod[row*RWP + pos] (position chosen on the host so that a half-warp's stores and the row
//     owners' loads fall in distinct banks); the row owners (LPR lanes per row) just add them up --
//     no row records, no gathers;
//   * the dense rows (objective-bound rows, long structural rows) and the two restart norms are NOT
//     reduced with warp shuffles in the column pass (that was ~100 of 570 instructions per thread and
//     iteration): every thread stores its partial sums to part[v][tid] and one half-warp per dense row
//     adds them up in the row phase, in parallel with the short-row owners;
//   * the objective is folded into the dual of its own dense row (y_cost - 1/dr_cost), so the reduced
//     cost needs no separate cost term;
//   * duals and product slots are addressed through 16-bit absolute shared-memory addresses packed in the
//     column record (ld.shared / st.shared on a register address, no pointer arithmetic).
// Per node-iteration the shared-memory traffic drops to: 2 y loads + 2 product stores + {l,u}, anchor load
// and xt store per column, one product load per nonzero, KD(+2) partial sums per thread.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdlib>

#include "device.h"
#include "k2_propagate.cuh"

namespace moip {
namespace k1reg {

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

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

__device__ __forceinline__ double dmax(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double dmin(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double clampd(double v, double a, double b) { return dmin(dmax(v, a), b); }

// One column's share of the model: ELLW short-row values | KD dense-row values | ELLW packed ids,
// id = (byte offset of the product slot << 16) | byte offset of the row's dual in ysh.
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
  __device__ __forceinline__ void clear(unsigned dummy_id) {
#pragma unroll
    for (int q = 0; q < UP; ++q) u[q] = 0.0;
    if (ELLW > 0) {
#pragma unroll
      for (int q = 0; q < (ELLW + 1) / 2; ++q) u[ELLW + KD + q] = __hiloint2double((int)dummy_id, (int)dummy_id);
    }
  }
  __device__ __forceinline__ double ell(int e) const { return u[e]; }
  __device__ __forceinline__ double dense(int d) const { return u[ELLW + d]; }
  __device__ __forceinline__ unsigned id(int e) const {
    const double w = u[ELLW + KD + e / 2];
    return (unsigned)((e & 1) ? __double2hiint(w) : __double2loint(w));
  }
  // both offsets become absolute shared-window addresses once per CTA (sbase < 64 KB - offsets, checked by the host)
  __device__ __forceinline__ void rebase(unsigned sbase) {
    if (ELLW > 0) {
#pragma unroll
      for (int q = 0; q < (ELLW + 1) / 2; ++q) {
        const double w = u[ELLW + KD + q];
        const unsigned add = sbase * 0x10001u;
        u[ELLW + KD + q] = __hiloint2double((int)((unsigned)__double2hiint(w) + add), (int)((unsigned)__double2loint(w) + add));
      }
    }
  }
  // volatile: keeps the two extractions inside the iteration loop.  Hoisted, the 4*CPT addresses would be
  // loop-invariant registers that ptxas then spills and reloads from local memory every iteration.
  __device__ __forceinline__ unsigned yaddr(int e) const {
    unsigned r; asm volatile("and.b32 %0, %1, 0xffff;" : "=r"(r) : "r"(id(e))); return r;
  }
  __device__ __forceinline__ unsigned paddr(int e) const {
    unsigned r; asm volatile("shr.u32 %0, %1, 16;" : "=r"(r) : "r"(id(e))); return r;
  }
};

// (kk+1)/(kk+2): Halpern weights, one constant-bank load instead of an fp64 division per iteration
constexpr int kHalpTab = 2048;
struct HalpTab {
  double v[kHalpTab];
  constexpr HalpTab() : v() {
    for (int i = 0; i < kHalpTab; ++i) v[i] = (double)(i + 1) / (double)(i + 2);
  }
};
__constant__ HalpTab c_halp = HalpTab();
__device__ __forceinline__ double halpern_weight(int kk) {
  return kk < kHalpTab ? c_halp.v[kk] : (double)(kk + 1) / (double)(kk + 2);
}

// volatile (not hoisted, not merged, ordered against the barriers) but no memory clobber: the compiler stays
// free to schedule the thread-private C++ loads and stores around them
__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); return v;
}
__device__ __forceinline__ void sts_f64(unsigned addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" :: "r"(addr), "d"(v));
}
template <int OFF>
__device__ __forceinline__ void lds_f64x2(unsigned addr, double& a, double& b) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(a), "=d"(b) : "r"(addr), "n"(OFF));
}
__device__ __forceinline__ double shx(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }
struct TrueT { static constexpr bool value = true; };
struct FalseT { static constexpr bool value = false; };
// role word of a thread in the row phase (kept opaque so that it is not re-derived from tid every iteration)
enum : unsigned { ROLE_ELL = 1u, ROLE_DENSE = 2u, ROLE_NORM = 4u, ROLE_LEADER = 8u, ROLE_WARP_DENSE = 16u, ROLE_WARP_ELL2 = 32u,
                  ROLE_LIVE = 64u };

// shared-memory map (bytes from the start of dynamic shared memory); the first two regions are fixed so that
// the host can pack byte offsets into the column records
constexpr int kYshBytes = 2048;            // ysh[256]   current dual iterate, published for the column pass
constexpr int kProdBase = kYshBytes;       // prod[msS*RWP + 2] products S_ij * xbar_j (+ dummy slot)
constexpr int kDL = 16;                    // lanes that add up one dense row / one norm
enum { COLD_BEST_LB = 0, COLD_POBJ, COLD_DOBJ, COLD_OBJ_UPPER, COLD_KKT_BINV, COLD_W, COLD_R0SQ, COLD_RPREV, COLD_CUTOFF, COLD_N = 10 };

__device__ __forceinline__ int cost_of(const LpBatch& b, int node) { return b.cost_idx[(size_t)node * b.cost_stride]; }

// FUSED: the B&B instantiation (LpBatch::fused): K2 propagation in front of the LP, K4 rounding/verification behind it.
// The plain instantiation (batch API, bench `value`) carries none of that code, so its iteration loop keeps its registers.
// FARKAS: the termination test also evaluates the Farkas certificate (always in the B&B instantiation).  In the plain batch
// instantiation it is a launch-time choice (LpBatch::farkas): its two extra sums change the register allocation of the
// whole iteration loop (ptxas: 28 -> 40-56 bytes of spill loads), which costs a batch of feasible node LPs 5 %.
template <int NT, int CPT, int KD, int ELLW, int MINB, bool FUSED, bool FARKAS>
__global__ void __launch_bounds__(NT, MINB)
k1_reg_kernel(const DevModel dm, const LpBatch b, const LpParams p, const int lpr_log2) {
  constexpr int NW = NT / 32;
  constexpr int NV = KD + 2;               // partial sums per thread: KD dense rows + 2 restart norms
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = dm.n, msS = dm.msS, m = dm.m, k = dm.k;
  double* ysh = reinterpret_cast<double*>(smem_raw);
  double* prod = reinterpret_cast<double*>(smem_raw + kProdBase);
  const int prod_len = (msS * dm.RWP + 2 + 1) & ~1;
  double* ytsh = prod + prod_len;        // [256] dual PDHG point (termination tests only)
  double* ydsh = ytsh + 256;             // [8]   duals of the dense rows with the objective folded in
  double* cold = ydsh + 8;               // [8]   per-node scalars used on norm / check iterations only
  double* nrm = cold + COLD_N;           // [2]   the two restart norms of a norm iteration
  double* redB = nrm + 2;                // NW * 8
  double* redC = redB + NW * 8;          // NW * 8
  int* s_node = reinterpret_cast<int*>(redC + NW * 8);       // [2]
  double2* lu = reinterpret_cast<double2*>(redC + NW * 8 + 2);   // [CPT][NT] thread-private {l,u}
  double* xts = reinterpret_cast<double*>(lu + NT * CPT);    // [CPT][NT] thread-private PDHG point xt
  double* xas = xts + NT * CPT;                              // [CPT][NT] thread-private Halpern anchor
  double* part = xas + NT * CPT;                             // [NV][NT]  per-thread partial sums of the dense rows / norms
  __shared__ int s_prop[FUSED ? 4 : 1];                      // flags of the fused propagation
  // chained rounds: the batch size comes from the device; with dynamic scheduling the first batchB CTAs take every node,
  // so the others leave before they load the model
  const int batchB = (FUSED && b.B_dev) ? *b.B_dev : b.B;
  if (FUSED && (int)blockIdx.x >= batchB) return;
  const int tid = threadIdx.x, lane = tid & 31;
  const int LPR = 1 << lpr_log2;
  // ---- roles in the row phase: LPR lanes per short structural row from thread 0 up, one half-warp per dense
  // row and per restart norm from the top thread down
  int row;                                  // kernel row owned by this thread's group, -1 = none
  unsigned role, psum;                      // flags; shared address of the first 16 bytes this thread adds up
  {
    const bool is_ell = tid < (msS << lpr_log2);
    const int vrow = (NT - 1 - tid) >> 4;
    const bool is_dense = vrow < NV;
    const int sub = is_ell ? (tid & (LPR - 1)) : (tid & (kDL - 1));
    row = is_ell ? (tid >> lpr_log2) : (is_dense && vrow < KD ? msS + vrow : -1);
    role = (is_ell ? ROLE_ELL : 0u) | (is_dense ? ROLE_DENSE : 0u) | (is_dense && vrow >= KD ? ROLE_NORM : 0u) |
           (row >= 0 && sub == 0 ? ROLE_LEADER : 0u);
    if (is_dense && vrow >= KD && sub == 0) row = KD - 2 - vrow;           // leader of a norm group: -2 / -3 = slot 0 / 1 of nrm[]
    if (__any_sync(0xffffffffu, is_dense)) role |= ROLE_WARP_DENSE;
    if (LPR == 2 && __any_sync(0xffffffffu, is_ell)) role |= ROLE_WARP_ELL2;
    const double* src = is_ell ? prod + (tid >> lpr_log2) * dm.RWP + 8 * sub : (is_dense ? part + vrow * NT + 2 * sub : prod);
    psum = (unsigned)__cvta_generic_to_shared(src);
    asm volatile("" : "+r"(role), "+r"(psum), "+r"(row));
  }
  const int dd = row - msS;                 // dense index of this thread's row (if >= 0)
  const int trips = dm.reg_trips;           // 64-byte steps per lane over its share of a short row
  const int norm_mask = p.norm_every - 1;   // norm_every is a power of two (host guarantees)
  double2* lu_t = lu + tid;
  double* xts_t = xts + tid;
  double* xas_t = xas + tid;
  double* part_t = part + tid;

  // ---- the model, once per CTA
  ColRec<ELLW, KD> rc[CPT];
  {
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned dummy = (unsigned)kProdBase + (unsigned)(msS * dm.RWP) * 8u;
    if (sbase + dummy + 16u >= 65536u) __trap();             // 16-bit shared addresses (host keeps dummy < 60000)
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int j = tid + c * NT;
      if (j < n) rc[c].load(dm.colrec2, j);
      else rc[c].clear(dummy << 16);
      rc[c].rebase(sbase);
    }
    for (int i = tid; i < prod_len; i += NT) prod[i] = 0.0;   // padding slots stay zero for ever
  }

  // S*v of this thread's row from the products / partial sums of the last column pass (after a barrier)
  auto row_sum = [&](const bool with_norms) -> double {
    double q = 0, q2 = 0;
    if (role & ROLE_ELL) {
      unsigned a = psum;
      for (int t = 0; t < trips; ++t) {
        double x0, y0, x1, y1, x2, y2, x3, y3;
        lds_f64x2<0>(a, x0, y0); lds_f64x2<16>(a, x1, y1); lds_f64x2<32>(a, x2, y2); lds_f64x2<48>(a, x3, y3);
        q += x0; q2 += y0; q += x1; q2 += y1; q += x2; q2 += y2; q += x3; q2 += y3;
        a += 64u << lpr_log2;
      }
    } else if ((role & ROLE_DENSE) && (with_norms || !(role & ROLE_NORM))) {
      double x[NT / 32], y[NT / 32];
#pragma unroll
      for (int t = 0; t < NT / 32; ++t) {
        if (t == 0) lds_f64x2<0>(psum, x[t], y[t]); else if (t == 1) lds_f64x2<256>(psum, x[t], y[t]);
        else if (t == 2) lds_f64x2<512>(psum, x[t], y[t]); else if (t == 3) lds_f64x2<768>(psum, x[t], y[t]);
        else if (t == 4) lds_f64x2<1024>(psum, x[t], y[t]); else if (t == 5) lds_f64x2<1280>(psum, x[t], y[t]);
        else if (t == 6) lds_f64x2<1536>(psum, x[t], y[t]); else if (t == 7) lds_f64x2<1792>(psum, x[t], y[t]);
        else if (t == 8) lds_f64x2<2048>(psum, x[t], y[t]); else if (t == 9) lds_f64x2<2304>(psum, x[t], y[t]);
        else if (t == 10) lds_f64x2<2560>(psum, x[t], y[t]); else if (t == 11) lds_f64x2<2816>(psum, x[t], y[t]);
        else if (t == 12) lds_f64x2<3072>(psum, x[t], y[t]); else if (t == 13) lds_f64x2<3328>(psum, x[t], y[t]);
        else if (t == 14) lds_f64x2<3584>(psum, x[t], y[t]); else lds_f64x2<3840>(psum, x[t], y[t]);
      }
#pragma unroll
      for (int t = 0; t < NT / 32; ++t) { q += x[t]; q2 += y[t]; }
    }
    q += q2;
    if (role & ROLE_WARP_DENSE) {
      double qd = (role & ROLE_DENSE) ? q : 0.0;
      qd += shx(qd, 8); qd += shx(qd, 4); qd += shx(qd, 2); qd += shx(qd, 1);
      const double qe = LPR == 2 ? q + shx(q, 1) : q;        // (uniform condition: every lane shuffles)
      q = (role & ROLE_DENSE) ? qd : qe;
    } else if (role & ROLE_WARP_ELL2) q += shx(q, 1);
    return q;
  };

  for (;;) {
    if (tid == 0) s_node[0] = atomicAdd(b.work_counter, 1);
    __syncthreads();
    const int node = s_node[0];
    if (node >= batchB) break;
    const size_t srow = b.slot ? (size_t)b.slot[node] : (size_t)(b.slot_base + node);   // row of the node's arrays
    if (FUSED) {
      // ---- fused round, part 1: K2 propagation of this node (xts|xas hold the 4n ints, part the per-warp partial sums)
      const int f = k2::propagate_node<NT>(dm, b.lb + srow * n, b.ub + srow * n, b.f_obj_lo, b.f_obj_hi, b.f_max_rounds,
                                           reinterpret_cast<int*>(xts), reinterpret_cast<long long*>(part), s_prop,
                                           b.f_leaf_obj + (size_t)node * k);
      if (tid == 0) b.f_flag[node] = f;
      if (f != 0) {                                           // infeasible or leaf: nothing to solve
        if (tid == 0) {
          b.status[node] = -1; b.iters[node] = 0;
          if (f == 2 && b.f_inc) {                            // a leaf is a verified point: it may be the new incumbent
            const long long v = (long long)dm.sgn * b.f_leaf_obj[(size_t)node * k + cost_of(b, node)];
            if (v < atomicMin(b.f_inc, v)) *b.f_cutoff_rw = (double)v;
          }
        }
        __syncthreads();
        continue;
      }
    } else if (b.skip && b.skip[node]) {
      if (tid == 0) { b.status[node] = -1; b.iters[node] = 0; }
      __syncthreads();
      continue;
    }
    // ---------------------------------------------------------------- node load
    const int cost = b.cost_idx[(size_t)node * b.cost_stride];
    const double* nrhs = b.rhs + (size_t)node * b.rhs_stride;
    const double inv_dr_cost = 1.0 / dm.dr_k[msS + cost];
    unsigned act = 0;                 // active dense rows (finite bound)
    for (int o = 0; o < k; ++o)
      if (fabs(nrhs[o]) < 1e19) act |= 1u << o;
    for (int t = k; t < KD; ++t) act |= 1u << t;
    double xb[CPT];
    double a0[NV];
#pragma unroll
    for (int d = 0; d < NV; ++d) a0[d] = 0;
    double cn2 = 0, cup = 0;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int j = tid + c * NT;
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
      xas_t[c * NT] = xj;
      xb[c] = xj;
      double cj = 0;
#pragma unroll
      for (int d = 0; d < KD; ++d) { if (d == cost) cj = rc[c].dense(d); a0[d] = fma(rc[c].dense(d), xj, a0[d]); }
      cj *= inv_dr_cost;
      cn2 = fma(cj, cj, cn2);
      cup += dmax(cj * lj, cj * uj);
#pragma unroll
      for (int e = 0; e < ELLW; ++e) sts_f64(rc[c].paddr(e), rc[c].ell(e) * xj);
    }
#pragma unroll
    for (int d = 0; d < KD; ++d) part_t[d * NT] = a0[d];
    double cs[2] = {cn2, cup};
    bsum<2, NT>(cs, redB, tid);           // barrier: products and partial sums visible
    // row state in registers (leaders); nlo/nhi are the negated row bounds (the dual step clamps to [-hi, -lo])
    double r_nlo = HUGE_VAL, r_nhi = -HUGE_VAL, r_y = 0, r_ya = 0, r_sx = 0, r_sxa = 0, r_yt = 0, r_sxt = 0;
    role &= ~ROLE_LIVE;                   // LIVE: leader of a row that can carry a nonzero dual
    {
      const double q = row_sum(false);    // S x0 of this thread's row
      if (row >= 0) {
        const double dri = dm.dr_k[row];
        double lo, hi;
        if (dd >= 0 && dd < k) { lo = -HUGE_VAL; hi = ((act >> dd) & 1u) ? dm.sgn * nrhs[dd] * dri : HUGE_VAL; }
        else { lo = dm.lo_k[row]; hi = dm.hi_k[row]; }
        double yi = b.warm_y ? b.warm_y[srow * m + row] / dri : 0.0;
        if (lo == -HUGE_VAL) yi = dmin(yi, 0.0);
        if (hi == HUGE_VAL) yi = dmax(yi, 0.0);
        const bool live = (role & ROLE_LEADER) && (dd < 0 || ((act >> dd) & 1u));
        if (live) role |= ROLE_LIVE; else yi = 0.0;
        r_nlo = -lo; r_nhi = -hi;
        r_y = yi; r_ya = yi;
        r_sx = q; r_sxa = q;
        if (role & ROLE_LEADER) {
          if (dd < 0) ysh[row] = yi;
          else ydsh[dd] = yi - (dd == cost ? inv_dr_cost : 0.0);
        }
      }
      asm volatile("" : "+r"(role));
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
    double tau, sigma, inv_sigma;
    {
      const double w = (cs[0] > 0 && bn2 > 0) ? sqrt(cs[0] / bn2) : 1.0;
      tau = dm.eta / w; sigma = dm.eta * w; inv_sigma = 1.0 / sigma;
      if (tid == 0) {
        cold[COLD_W] = w; cold[COLD_R0SQ] = 0.0; cold[COLD_RPREV] = -1.0;
        cold[COLD_BEST_LB] = -HUGE_VAL; cold[COLD_POBJ] = 0.0; cold[COLD_DOBJ] = -HUGE_VAL;
        cold[COLD_OBJ_UPPER] = cs[1]; cold[COLD_KKT_BINV] = 1.0 / (1.0 + sqrt(bn2_unscaled));
      }
    }
    __syncthreads();                       // ysh / ydsh visible; products consumed

    int kk = 0, it = 0, status = MOIP_LP_ITERLIMIT;
    double ah = 0.0;                       // Halpern weight of this iteration's column pass: x = ah*xbar + (1-ah)*xa
    const int iter_cap = p.fixed_iters > 0 ? p.fixed_iters : p.max_iter;
    int next_check = p.fixed_iters > 0 ? 0x7fffffff : p.check_every;

    // ---- column pass.  ydsh carries the objective (y_cost - 1/dr_cost), so  -(c - S^T y) = sum_i S_ij yd_i.
    // NORM: also accumulate the two restart norms; xt is only written out when somebody will read it
    // (restart / termination tests happen on norm, check and last iterations).
    auto column_pass = [&](auto norm_tag, const double ah, const double tau, const bool keep_xt) {
      constexpr bool NORM = decltype(norm_tag)::value;
      const double ah1 = 1.0 - ah;
      double yd[KD];
#pragma unroll
      for (int d = 0; d < KD; ++d) yd[d] = ydsh[d];
      double ye[CPT][ELLW > 0 ? ELLW : 1];
#pragma unroll
      for (int c = 0; c < CPT; ++c)
#pragma unroll
        for (int e = 0; e < ELLW; ++e) ye[c][e] = lds_f64(rc[c].yaddr(e));
      double aA[NV];
#pragma unroll
      for (int d = 0; d < NV; ++d) aA[d] = 0;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const double2 bx = lu_t[c * NT];
        const double xaj = xas_t[c * NT];
        double g = 0, gd = 0;
#pragma unroll
        for (int e = 0; e < ELLW; ++e) g = fma(rc[c].ell(e), ye[c][e], g);
#pragma unroll
        for (int d = 0; d < KD; ++d) gd = fma(rc[c].dense(d), yd[d], gd);
        g += gd;
        const double xj = fma(ah, xb[c], ah1 * xaj);
        const double xtj = clampd(fma(tau, g, xj), bx.x, bx.y);
        const double xbn = fma(2.0, xtj, -xj);
        xb[c] = xbn;
        if (NORM || keep_xt) xts_t[c * NT] = xtj;
        if (NORM) {
          const double d1 = xtj - xj, d2 = xtj - xaj;
          aA[KD] = fma(d1, d1, aA[KD]);
          aA[KD + 1] = fma(d2, d2, aA[KD + 1]);
        }
#pragma unroll
        for (int d = 0; d < KD; ++d) aA[d] = fma(rc[c].dense(d), xbn, aA[d]);
      }
#pragma unroll
      for (int c = 0; c < CPT; ++c)
#pragma unroll
        for (int e = 0; e < ELLW; ++e) sts_f64(rc[c].paddr(e), rc[c].ell(e) * xb[c]);
#pragma unroll
      for (int d = 0; d < (NORM ? NV : KD); ++d) part_t[d * NT] = aA[d];
    };

    // ---------------------------------------------------------------- PDHG iterations
    for (;;) {
      ++it;
      const bool norm_it = (kk & norm_mask) == 0;
      const bool check_it = it == next_check;
      const bool last_it = it >= iter_cap;
      if (norm_it) column_pass(TrueT{}, ah, tau, true);
      else column_pass(FalseT{}, ah, tau, check_it || last_it);
      __syncthreads();                     // products and partial sums visible
      // ---- row phase
      double aB[3] = {0, 0, 0};
      {
        const double q = row_sum(norm_it);
        if (norm_it && row < -1) nrm[-2 - row] = q;
        r_sxt = 0.5 * (q + r_sx);
        const double v = fma(r_y, inv_sigma, -q);
        r_yt = (role & ROLE_LIVE) ? sigma * (v - clampd(v, r_nhi, r_nlo)) : 0.0;
        if (norm_it && (role & ROLE_LIVE)) {
          const double dy = r_yt - r_y, dya = r_yt - r_ya;
          aB[0] = dy * dy; aB[1] = dy * (r_sxt - r_sx); aB[2] = dya * dya;
        }
      }
      bool restart = false;
      double fp2 = 0.0, w_new = 0.0;
      const bool first_norm = kk == 0;
      if (norm_it) {
        bsum<3, NT>(aB, redB, tid);        // barrier: nrm visible
        const double nx1 = nrm[0], nx2 = nrm[1];
        const double wc = cold[COLD_W], r0sq = cold[COLD_R0SQ], rprev = cold[COLD_RPREV];
        fp2 = dmax(0.0, fma(wc / dm.eta, nx1, fma(-2.0, aB[1], aB[0] * inv_sigma)));
        if (kk != 0 && (fp2 <= 0.04 * r0sq || (fp2 <= 0.64 * r0sq && rprev >= 0.0 && fp2 > rprev) || 25 * kk >= 9 * it))
          restart = true;
        if (restart) {
          const double dxn = sqrt(nx2), dyn = sqrt(aB[2]);
          double w = wc;
          if (dxn > 1e-10 && dyn > 1e-10) w = exp(0.5 * log(dyn / dxn) + 0.5 * log(w));
          tau = dm.eta / w; sigma = dm.eta * w; inv_sigma = 1.0 / sigma;
          w_new = w;
        }
      }
      const bool need_stop_eval = check_it || last_it;
      // ---- termination tests at (xt, yt); on the last iteration they also produce the outputs
      bool stop = last_it;
      if (need_stop_eval) {
        if (role & ROLE_LEADER) ytsh[row] = r_yt;
        if (check_it) next_check += p.check_every;
        __syncthreads();                   // ytsh visible
        // pobj, dual (columns), dual (rows), primal residual^2 (unscaled), Farkas value (columns), sum of its |terms|
        double aC[4] = {0, 0, 0, 0};
        double aF[2] = {0, 0};             // (reduced on their own: six values in one reduction cost the iteration loop spills)
        double ytd[KD];
#pragma unroll
        for (int d = 0; d < KD; ++d) ytd[d] = ytsh[msS + d];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          const double2 bx = lu_t[c * NT];
          double g = 0, cj = 0;
#pragma unroll
          for (int e = 0; e < ELLW; ++e) g = fma(rc[c].ell(e), lds_f64(rc[c].yaddr(e) + (unsigned)((const char*)ytsh - (const char*)ysh)), g);
#pragma unroll
          for (int d = 0; d < KD; ++d) { g = fma(rc[c].dense(d), ytd[d], g); if (d == cost) cj = rc[c].dense(d); }
          cj *= inv_dr_cost;
          const double r = cj - g;
          aC[0] = fma(cj, xts_t[c * NT], aC[0]);
          aC[1] += (r > 0) ? r * bx.x : r * bx.y;
          if (FARKAS) {
            const double fk = (g < 0) ? -g * bx.x : -g * bx.y;      // the same bound with the objective dropped (Farkas)
            aF[0] += fk; aF[1] += fabs(fk);
          }
        }
        if (role & ROLE_LIVE) {
          if (r_yt > 0) aC[2] = -r_yt * r_nlo;
          else if (r_yt < 0) aC[2] = -r_yt * r_nhi;
          if (FARKAS) aF[1] += fabs(aC[2]);
          const double viol = dmax(0.0, dmax(r_sxt + r_nhi, -r_nlo - r_sxt)) / dm.dr_k[row];
          aC[3] = viol * viol;
        }
        // chained rounds: other CTAs lower the cutoff while this one runs -- one thread reads it, so that every thread of
        // the CTA takes the same way out of the loop
        if (FUSED && tid == 0) cold[COLD_CUTOFF] = b.cutoff ? *((volatile const double*)b.cutoff) : HUGE_VAL;
        if (FUSED) {                       // (B&B instantiation: one reduction over all six)
          double a6[6] = {aC[0], aC[1], aC[2], aC[3], aF[0], aF[1]};
          bsum<6, NT>(a6, redC, tid);
          aC[0] = a6[0]; aC[1] = a6[1]; aC[2] = a6[2]; aC[3] = a6[3]; aF[0] = a6[4]; aF[1] = a6[5];
        } else {
          bsum<4, NT>(aC, redC, tid);
          if (FARKAS) {
            __syncthreads();               // redC is read by every thread before it is written again
            bsum<2, NT>(aF, redC, tid);
          }
        }
        const double pobj = aC[0], dobj = aC[1] + aC[2];
        // Farkas certificate: F(y) = min over the box of (-S^T y) x + (y+ lo - y- hi) <= 0 for every point that satisfies
        // the rows, so F(y) > 0 proves the node LP infeasible -- long before the Lagrangian bound (the same sum plus the
        // objective) climbs past the objective's maximum on the box.  The margin is 1000x the rounding error of the sum.
        const bool farkas = FARKAS && aF[0] + aC[2] > 1e-9 * aF[1] + 1e-9;
        double best_lb = cold[COLD_BEST_LB];
        const double obj_upper = cold[COLD_OBJ_UPPER], kkt_binv = cold[COLD_KKT_BINV];
        if (p.fixed_iters > 0) best_lb = dobj;
        else {
          if (dobj > best_lb) best_lb = dobj;
          const double gap = fabs(pobj - dobj);
          const double rel = dmax(sqrt(aC[3]) * kkt_binv, gap / (1.0 + fabs(pobj) + fabs(dobj)));
          const double cutoff = FUSED ? cold[COLD_CUTOFF] : (b.cutoff ? *((volatile const double*)b.cutoff) : HUGE_VAL);
          if (best_lb >= cutoff - p.cutoff_slack) { status = MOIP_LP_CUTOFF; stop = true; }
          else if (farkas || best_lb > obj_upper + 1e-6 * (1.0 + fabs(obj_upper))) { status = MOIP_LP_INFEASIBLE; stop = true; }
          else if (rel <= p.eps) { status = MOIP_LP_CONVERGED; stop = true; }
          else if (p.int_obj && sqrt(aC[3]) * kkt_binv <= 1e-5 && ceil(best_lb - 1e-6) >= ceil(pobj - 1e-3)) {
            status = MOIP_LP_CONVERGED; stop = true;      // the integer-rounded bound cannot improve any further
          }
        }
        __syncthreads();                   // every thread has read the cold scalars
        if (tid == 0) { cold[COLD_BEST_LB] = best_lb; cold[COLD_POBJ] = pobj; cold[COLD_DOBJ] = dobj; }
      }
      if (stop) break;
      // ---- restart or Halpern step of the row state; x follows in the next column pass
      if (restart) {
#pragma unroll
        for (int c = 0; c < CPT; ++c) xas_t[c * NT] = xts_t[c * NT];
        r_y = r_yt; r_ya = r_yt; r_sx = r_sxt; r_sxa = r_sxt;
        kk = 0; ah = 0.0;
      } else {
        ah = halpern_weight(kk);           // (kk+1)/(kk+2)
        const double c1 = 1.0 - ah;
        r_y = fma(ah, 2.0 * r_yt - r_y, c1 * r_ya);
        r_sx = fma(ah, 2.0 * r_sxt - r_sx, c1 * r_sxa);
        ++kk;
      }
      if (role & ROLE_LEADER) {
        if (dd < 0) ysh[row] = r_y;
        else ydsh[dd] = r_y - (dd == cost ? inv_dr_cost : 0.0);
      }
      __syncthreads();                     // duals visible; products consumed before the next column pass
      if (norm_it && tid == 0) {           // deferred: every thread read the old values before the barrier
        if (first_norm) cold[COLD_R0SQ] = fp2;
        cold[COLD_RPREV] = restart ? -1.0 : fp2;
        if (restart) cold[COLD_W] = w_new;
      }
    }
    __syncthreads();                       // cold scalars of the last evaluation visible
    const double pobj = cold[COLD_POBJ], best_lb = cold[COLD_BEST_LB], dobj_last = cold[COLD_DOBJ];

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
        for (int d = 0; d < KD; ++d) {
          ytd[d] = ytsh[msS + d];
          if (d == cost) ytd[d] -= inv_dr_cost;
        }
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          const int j = tid + c * NT;
          if (j < n) {
            double g = 0;
#pragma unroll
            for (int e = 0; e < ELLW; ++e) g = fma(rc[c].ell(e), lds_f64(rc[c].yaddr(e) + (unsigned)((const char*)ytsh - (const char*)ysh)), g);
#pragma unroll
            for (int d = 0; d < KD; ++d) g = fma(rc[c].dense(d), ytd[d], g);
            const double r = -g / dm.dc[j];        // unscaled reduced cost
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
    // unscaled PDHG point
    double xo[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int j = tid + c * NT;
      xo[c] = (j < n) ? xts_t[c * NT] * dm.dc[j] : 0.0;
      if (b.out_x && j < n) b.out_x[srow * n + j] = xo[c];
    }
    if (b.out_y && (role & ROLE_LEADER)) b.out_y[srow * m + row] = r_yt * dm.dr_k[row];
    if (b.branch_var) {           // the three most fractional columns, best first
      int c0 = -1, c1 = -1;
      for (int r = 0; r < 3; ++r) {
        double bestf = -1.0, bestv = 0.0; int bestj = -1;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          const int j = tid + c * NT;
          const double f = fabs(xo[c] - rint(xo[c]));
          if (j < n && j != c0 && j != c1 && f > bestf) { bestf = f; bestj = j; bestv = xo[c]; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double of = __shfl_xor_sync(0xffffffffu, bestf, o);
          const double ov = __shfl_xor_sync(0xffffffffu, bestv, o);
          const int oj = __shfl_xor_sync(0xffffffffu, bestj, o);
          if (of > bestf || (of == bestf && oj >= 0 && (bestj < 0 || oj < bestj))) { bestf = of; bestj = oj; bestv = ov; }
        }
        __syncthreads();
        if (lane == 0) { redB[(tid >> 5) * 3] = bestf; redB[(tid >> 5) * 3 + 1] = (double)bestj; redB[(tid >> 5) * 3 + 2] = bestv; }
        __syncthreads();
        for (int wq = 0; wq < NW; ++wq) {     // every thread folds the warp results: uniform outcome
          const double of = redB[wq * 3]; const int oj = (int)redB[wq * 3 + 1];
          if (of > bestf || (of == bestf && oj >= 0 && (bestj < 0 || oj < bestj))) { bestf = of; bestj = oj; bestv = redB[wq * 3 + 2]; }
        }
        const int pick = (bestf > 1e-6) ? bestj : -1;
        if (tid == 0) {
          b.branch_var[(size_t)node * 3 + r] = pick;
          if (b.branch_val) b.branch_val[(size_t)node * 3 + r] = (pick >= 0) ? bestv : 0.0;
        }
        if (r == 0) c0 = pick; else c1 = pick;
      }
    }
    if (FUSED) {
      // ---- fused round, part 2: round the LP point three ways (nearest / down / up, clipped to the node's box) and
      // verify each candidate exactly in int64 -- objective values by a block reduction, structural rows a thread per row
      // on the candidates staged in shared memory (xts|xas are free: xo holds the point)
      __syncthreads();
      int* xr_s = reinterpret_cast<int*>(xts);                // [3][n]
      long long co[3][MOIP_MAX_OBJ];
#pragma unroll
      for (int md = 0; md < 3; ++md)
#pragma unroll
        for (int o = 0; o < MOIP_MAX_OBJ; ++o) co[md][o] = 0;
      int ff = INT_MAX;                                       // first column that is not fixed yet (fallback branching column)
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const int j = tid + c * NT;
        if (j < n) {
          const int lj = b.lb[srow * n + j], uj = b.ub[srow * n + j];
          const double v = xo[c];
          int r[3] = {(int)llrint(v), (int)floor(v + 1e-6), (int)ceil(v - 1e-6)};
          if (lj < uj && j < ff) ff = j;
#pragma unroll
          for (int md = 0; md < 3; ++md) {
            r[md] = max(lj, min(uj, r[md]));
            xr_s[md * n + j] = r[md];
            b.f_xr[((size_t)node * 3 + md) * n + j] = r[md];
          }
#pragma unroll
          for (int o = 0; o < MOIP_MAX_OBJ; ++o)
            if (o < k) {
              const long long cj = dm.ci[(size_t)o * n + j];
#pragma unroll
              for (int md = 0; md < 3; ++md) co[md][o] += cj * (long long)r[md];
            }
        }
      }
      constexpr int RS = 3 * MOIP_MAX_OBJ + 1;                // per-warp partial results: 3 x 4 objective sums + first free column
      long long* red = reinterpret_cast<long long*>(part);
#pragma unroll
      for (int md = 0; md < 3; ++md)
#pragma unroll
        for (int o = 0; o < MOIP_MAX_OBJ; ++o) {
          const long long t = k2::wsum_ll(co[md][o]);
          if (lane == 0) red[(tid >> 5) * RS + md * MOIP_MAX_OBJ + o] = t;
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ff = min(ff, __shfl_xor_sync(0xffffffffu, ff, o));
      if (lane == 0) red[(tid >> 5) * RS + 3 * MOIP_MAX_OBJ] = ff;
      __syncthreads();                                        // candidates and partial results visible
      int bad0 = 0, bad1 = 0, bad2 = 0;
      for (int i = tid; i < dm.ms; i += NT) {
        long long a0 = 0, a1 = 0, a2 = 0;
        for (int e = dm.s_ptr[i]; e < dm.s_ptr[i + 1]; ++e) {
          const int col = dm.s_col[e];
          const long long av = dm.ai_val[e];
          a0 += av * (long long)xr_s[col]; a1 += av * (long long)xr_s[n + col]; a2 += av * (long long)xr_s[2 * n + col];
        }
        const long long lo = dm.ri_lo[i], hi = dm.ri_hi[i];
        bad0 |= (a0 < lo || a0 > hi); bad1 |= (a1 < lo || a1 > hi); bad2 |= (a2 < lo || a2 > hi);
      }
      bad0 = __syncthreads_or(bad0); bad1 = __syncthreads_or(bad1); bad2 = __syncthreads_or(bad2);
      if (tid < 3 * k) {
        const int md = tid / k, o = tid - md * k;
        long long t = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) t += red[w * RS + md * MOIP_MAX_OBJ + o];
        b.f_cand_obj[((size_t)node * 3 + md) * k + o] = t;
      }
      if (tid == 0) {
        b.f_cand_feas[(size_t)node * 3] = bad0 ? 0 : 1;
        b.f_cand_feas[(size_t)node * 3 + 1] = bad1 ? 0 : 1;
        b.f_cand_feas[(size_t)node * 3 + 2] = bad2 ? 0 : 1;
        long long fm = INT_MAX;
#pragma unroll
        for (int w = 0; w < NW; ++w) fm = min(fm, red[w * RS + 3 * MOIP_MAX_OBJ]);
        const int f1 = (int)fm;
        b.f_first_free[(size_t)node * 3] = f1 == INT_MAX ? -1 : f1;
        b.f_first_free[(size_t)node * 3 + 1] = f1 == INT_MAX ? 0 : b.lb[srow * n + f1];
        b.f_first_free[(size_t)node * 3 + 2] = f1 == INT_MAX ? 0 : b.ub[srow * n + f1];
      }
      if (b.f_inc) {                                          // chained rounds: the incumbent value is kept on the device
        __syncthreads();                                      // the candidates' objective values are in global memory
        if (tid == 0) {
          long long best = LLONG_MAX;
          for (int md = 0; md < 3; ++md) {
            if (!b.f_cand_feas[(size_t)node * 3 + md]) continue;
            const long long* ov = b.f_cand_obj + ((size_t)node * 3 + md) * k;
            bool ok = true;
            for (int o = 0; o < k; ++o) ok = ok && ov[o] >= b.f_lim_lo[o] && ov[o] <= b.f_lim_hi[o];
            const long long v = (long long)dm.sgn * ov[cost];
            if (ok && v < best) best = v;
          }
          // (a stale, larger cutoff left by a racing writer is still valid: it is only ever compared against bounds)
          if (best != LLONG_MAX && best < atomicMin(b.f_inc, best)) *b.f_cutoff_rw = (double)best;
        }
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

inline int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

template <int NT, int CPT, int KD, int ELLW, int MINB, bool FUSED, bool FARKAS>
int launch_reg_f(const DevModel& dm, const LpBatch& b, LpParams p, int num_sms, cudaStream_t st) {
  auto kern = k1_reg_kernel<NT, CPT, KD, ELLW, MINB, FUSED, FARKAS>;
  const size_t prod_len = ((size_t)dm.msS * dm.RWP + 2 + 1) & ~(size_t)1;
  const size_t smem = kYshBytes + sizeof(double) * (prod_len + 256 + 8 + COLD_N + 2 + (size_t)16 * (NT / 32) + 2 +
                                                    (size_t)4 * NT * CPT + (size_t)(KD + 2) * NT);
  static LaunchCfg cfg;
  std::unique_lock<std::mutex> cfg_lock(launch_cfg_mutex());
  const int dev = current_device();
  size_t& configured = cfg.configured[dev];
  int& occ = cfg.occ[dev];
  if (smem > configured) {
    MOIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MOIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    MOIP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    if (occ < 1) { std::fprintf(stderr, "moip_b200: node LP does not fit on an SM (n=%d)\n", dm.n); return MOIP_ERR_LIMIT; }
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
  int ne = 1;                              // restart-test cadence: largest power of two <= norm_every
  while (ne * 2 <= p.norm_every) ne *= 2;
  p.norm_every = ne;
  kern<<<(unsigned)grid, NT, smem, st>>>(dm, b, p, dm.reg_lpr_log2);
  MOIP_CUDA(cudaGetLastError());
  return MOIP_OK;
}

template <int NT, int CPT, int KD, int ELLW, int MINB>
int launch_reg(const DevModel& dm, const LpBatch& b, const LpParams& p, int num_sms, cudaStream_t st) {
  if (b.fused) {
    if (!b.slot && !b.B_dev) return MOIP_ERR_ARG;
    return launch_reg_f<NT, CPT, KD, ELLW, MINB, true, true>(dm, b, p, num_sms, st);
  }
  if (b.farkas) return launch_reg_f<NT, CPT, KD, ELLW, MINB, false, true>(dm, b, p, num_sms, st);
  return launch_reg_f<NT, CPT, KD, ELLW, MINB, false, false>(dm, b, p, num_sms, st);
}

// shape dispatch for one KD (one translation unit per KD keeps the build parallel)
template <int KD>
int launch_reg_kd(const DevModel& dm, const LpBatch& b, const LpParams& p, int num_sms, cudaStream_t st) {
  // fewer, fatter threads win: with more warps per node the barriers and the partial-sum phase cost more than
  // the shorter column pass saves (measured on B200, 3AP30: 512x2 and 320x3 are 15-25 % slower than 256x4)
  if (dm.ell2_w == 2) {
    if (dm.n <= 256) return launch_reg<128, 2, KD, 2, 4>(dm, b, p, num_sms, st);
    if (dm.n <= 512) return launch_reg<256, 2, KD, 2, 2>(dm, b, p, num_sms, st);
    return launch_reg<256, 4, KD, 2, 2>(dm, b, p, num_sms, st);
  }
  if (dm.ell2_w == 0) {
    if (dm.n <= 256) return launch_reg<128, 2, KD, 0, 4>(dm, b, p, num_sms, st);
    if (dm.n <= 512) return launch_reg<256, 2, KD, 0, 2>(dm, b, p, num_sms, st);
    return launch_reg<256, 4, KD, 0, 2>(dm, b, p, num_sms, st);
  }
  return MOIP_ERR_UNSUPPORTED;
}

}  // namespace k1reg
}  // namespace moip
