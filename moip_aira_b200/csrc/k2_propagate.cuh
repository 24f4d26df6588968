// K2 propagation of ONE node by one CTA: exact int64 activity-based bound tightening over the structural rows and the k
// objective-bound rows (incl. the incumbent cut-off on the optimised objective) to a fixpoint, classification
// open / infeasible / leaf, exact evaluation of leaves.  Shared by the stand-alone kernel (k2_nodepool.cu) and the fused
// round of the register-resident K1 (k1_reg.cuh: propagate -> LP -> round/verify in one CTA).
#pragma once
#include <climits>

#include "device.h"

namespace moip {
namespace k2 {

__device__ __forceinline__ long long wsum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ long long floor_div(long long a, long long b) {   // b != 0
  long long q = a / b, r = a % b;
  return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}
__device__ __forceinline__ long long ceil_div(long long a, long long b) {
  long long q = a / b, r = a % b;
  return (r != 0 && ((r < 0) == (b < 0))) ? q + 1 : q;
}

// One tightening step of column j against a row  rlo <= sum a x <= rhi  with activity range [mn, mx] over the node's
// current box.  The new bound floor(room / a) (or its mirror images) only matters when it cuts into the box, and that is
// decided by one multiplication -- room < a * u  <=>  floor(room / a) < u  (a > 0, u integer) -- so the 64-bit division,
// ~100 instructions on this machine, is paid only for the few columns that actually move.
__device__ __forceinline__ void tighten_column(long long a, int j, int lj, int uj, long long mn, long long mx, long long rlo,
                                               long long rhi, bool use_hi, bool use_lo, int* nlb, int* nub, int* changed) {
  if (a == 0 || lj == uj) return;
  if (use_hi) {
    const long long room = rhi - (mn - (a > 0 ? a * lj : a * uj));
    if (a > 0) {
      if (room < a * (long long)uj) { const long long t = floor_div(room, a); atomicMin(&nub[j], (int)max(t, (long long)INT_MIN / 2)); *changed = 1; }
    } else if (room < a * (long long)lj) { const long long t = ceil_div(room, a); atomicMax(&nlb[j], (int)min(t, (long long)INT_MAX / 2)); *changed = 1; }
  }
  if (use_lo) {
    const long long need = rlo - (mx - (a > 0 ? a * uj : a * lj));
    if (a > 0) {
      if (need > a * (long long)lj) { const long long t = ceil_div(need, a); atomicMax(&nlb[j], (int)min(t, (long long)INT_MAX / 2)); *changed = 1; }
    } else if (need > a * (long long)uj) { const long long t = floor_div(need, a); atomicMin(&nub[j], (int)max(t, (long long)INT_MIN / 2)); *changed = 1; }
  }
}

// All NT threads of the CTA call this.  plb/pub: the node's bounds in the pool (tightened in place when the node stays open);
// sm_i: 4n ints of shared memory; s_act: (NT/32) * 2 * MOIP_MAX_OBJ long longs; s_flags: 3 ints; leaf_obj: k values (leaves).
// Returns 0 open / 1 infeasible / 2 leaf, the same value in every thread; the shared buffers are free again on return.
template <int NT>
__device__ int propagate_node(const DevModel& dm, int* plb, int* pub, const long long* obj_lo, const long long* obj_hi,
                              int max_rounds, int* sm_i, long long* s_act, int* s_flags, long long* leaf_obj) {
  const int n = dm.n, ms = dm.ms, k = dm.k;
  int* lb = sm_i;
  int* ub = lb + n;
  int* nlb = ub + n;
  int* nub = nlb + n;
  constexpr int NW = NT / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __syncthreads();                           // whatever used the shared buffers before is done with them
  if (tid == 0) s_flags[1] = 0;
  __syncthreads();                         // the reset must not overtake another warp's "crossed bounds" store below
  for (int j = tid; j < n; j += NT) {
    const int a = plb[j], b2 = pub[j];
    lb[j] = a; ub[j] = b2; nlb[j] = a; nub[j] = b2;
    if (a > b2) s_flags[1] = 1;
  }
  // objective rows that carry a bound (the others cannot tighten anything)
  unsigned act_rows = 0;
  for (int o = 0; o < k; ++o)
    if (obj_lo[o] != LLONG_MIN || obj_hi[o] != LLONG_MAX) act_rows |= 1u << o;
  for (int round = 0; round < max_rounds; ++round) {
    __syncthreads();                       // (A) bounds of this round visible
    if (s_flags[1]) break;               // uniform: nobody writes it between (A) and (B)
    if (tid == 0) s_flags[0] = 0;
    // ---- the k objective rows are dense (every column): the whole CTA adds up their activity ranges, a thread per
    // column, instead of one warp walking 900 entries per row
    long long dmn[MOIP_MAX_OBJ] = {0, 0, 0, 0}, dmx[MOIP_MAX_OBJ] = {0, 0, 0, 0};
    if (act_rows) {
      for (int j = tid; j < n; j += NT) {
        const long long lj = lb[j], uj = ub[j];
#pragma unroll
        for (int o = 0; o < MOIP_MAX_OBJ; ++o) {
          if (!((act_rows >> o) & 1u)) continue;
          const long long a = dm.ci[(size_t)o * n + j];
          if (a > 0) { dmn[o] += a * lj; dmx[o] += a * uj; }
          else if (a < 0) { dmn[o] += a * uj; dmx[o] += a * lj; }
        }
      }
#pragma unroll
      for (int o = 0; o < MOIP_MAX_OBJ; ++o) {
        if (!((act_rows >> o) & 1u)) continue;
        dmn[o] = wsum_ll(dmn[o]); dmx[o] = wsum_ll(dmx[o]);
        if (lane == 0) { s_act[warp * 2 * MOIP_MAX_OBJ + 2 * o] = dmn[o]; s_act[warp * 2 * MOIP_MAX_OBJ + 2 * o + 1] = dmx[o]; }
      }
    }
    __syncthreads();                       // (B) partial sums and the s_flags[0] reset visible
    if (act_rows) {
      bool use_hi[MOIP_MAX_OBJ], use_lo[MOIP_MAX_OBJ];
      bool any = false;
#pragma unroll
      for (int o = 0; o < MOIP_MAX_OBJ; ++o) {
        use_hi[o] = use_lo[o] = false;
        if (!((act_rows >> o) & 1u)) continue;
        long long mn = 0, mx = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) { mn += s_act[w * 2 * MOIP_MAX_OBJ + 2 * o]; mx += s_act[w * 2 * MOIP_MAX_OBJ + 2 * o + 1]; }
        dmn[o] = mn; dmx[o] = mx;
        const long long rlo = obj_lo[o], rhi = obj_hi[o];
        if ((rhi != LLONG_MAX && mn > rhi) || (rlo != LLONG_MIN && mx < rlo)) { if (tid == 0) s_flags[1] = 1; continue; }
        use_hi[o] = rhi != LLONG_MAX && mx > rhi;   // row can still be violated from above
        use_lo[o] = rlo != LLONG_MIN && mn < rlo;
        any = any || use_hi[o] || use_lo[o];
      }
      if (any) {
        int ch = 0;
        for (int j = tid; j < n; j += NT) {
          const int lj = lb[j], uj = ub[j];
          if (lj == uj) continue;
#pragma unroll
          for (int o = 0; o < MOIP_MAX_OBJ; ++o)
            if (use_hi[o] || use_lo[o])
              tighten_column(dm.ci[(size_t)o * n + j], j, lj, uj, dmn[o], dmx[o], obj_lo[o], obj_hi[o], use_hi[o], use_lo[o], nlb, nub, &ch);
        }
        if (ch) s_flags[0] = 1;
      }
    }
    // ---- structural rows: a warp per row
    for (int i = warp; i < ms; i += NW) {
      const long long rlo = dm.ri_lo[i], rhi = dm.ri_hi[i];
      if (rlo == LLONG_MIN && rhi == LLONG_MAX) continue;
      const int e0 = dm.s_ptr[i], e1 = dm.s_ptr[i + 1];
      long long mn = 0, mx = 0;
      for (int e = e0 + lane; e < e1; e += 32) {
        const int j = dm.s_col[e];
        const long long a = dm.ai_val[e];
        if (a > 0) { mn += a * lb[j]; mx += a * ub[j]; }
        else if (a < 0) { mn += a * ub[j]; mx += a * lb[j]; }
      }
      mn = wsum_ll(mn); mx = wsum_ll(mx);
      if ((rhi != LLONG_MAX && mn > rhi) || (rlo != LLONG_MIN && mx < rlo)) { if (lane == 0) s_flags[1] = 1; continue; }
      const bool use_hi = rhi != LLONG_MAX && mx > rhi;
      const bool use_lo = rlo != LLONG_MIN && mn < rlo;
      if (!use_hi && !use_lo) continue;
      int ch = 0;
      for (int e = e0 + lane; e < e1; e += 32) {
        const int j = dm.s_col[e];
        tighten_column(dm.ai_val[e], j, lb[j], ub[j], mn, mx, rlo, rhi, use_hi, use_lo, nlb, nub, &ch);
      }
      if (ch) s_flags[0] = 1;
    }
    __syncthreads();                       // (C)
    if (!s_flags[0] || s_flags[1]) break; // uniform: written before (C), rewritten only after the next (A)
    for (int j = tid; j < n; j += NT) {
      const int a = nlb[j], b2 = nub[j];
      lb[j] = a; ub[j] = b2;
      if (a > b2) s_flags[1] = 1;
    }
  }
  __syncthreads();
  if (tid == 0) s_flags[2] = 0;
  __syncthreads();
  if (!s_flags[1]) {
    int unf = 0;
    for (int j = tid; j < n; j += NT) {
      plb[j] = lb[j]; pub[j] = ub[j];
      if (lb[j] != ub[j]) unf = 1;
    }
    if (unf) s_flags[2] = 1;
  }
  __syncthreads();
  int f = s_flags[1] ? 1 : (s_flags[2] ? 0 : 2);
  __syncthreads();                         // every thread has its f before a leaf evaluation may raise the flag again
  if (f == 2) {
    // leaf: evaluate everything exactly at x = lb
    for (int i = warp; i < ms + k; i += NW) {
      const bool dense = i >= ms;
      long long a = 0;
      if (dense) { const long long* cv = dm.ci + (size_t)(i - ms) * n; for (int j = lane; j < n; j += 32) a += cv[j] * lb[j]; }
      else for (int e = dm.s_ptr[i] + lane; e < dm.s_ptr[i + 1]; e += 32) a += dm.ai_val[e] * lb[dm.s_col[e]];
      a = wsum_ll(a);
      if (lane == 0) {
        if (dense) {
          if (leaf_obj) leaf_obj[i - ms] = a;
          if (a < obj_lo[i - ms] || a > obj_hi[i - ms]) s_flags[1] = 1;
        } else if (a < dm.ri_lo[i] || a > dm.ri_hi[i]) s_flags[1] = 1;
      }
    }
    __syncthreads();
    if (s_flags[1]) f = 1;
  }
  __syncthreads();
  return f;
}

}  // namespace k2
}  // namespace moip
