// K2 -- GPU node pool of the branch-and-bound that replaces CPXmipopt's tree
// (reference src/aira.cpp:480; no reference source).
//   * k2_propagate : exact int64 activity-based bound tightening of every node of a batch over
//                    the structural rows and the k objective-bound rows (incl. the incumbent
//                    cut-off on the optimised objective); classifies nodes as open / infeasible /
//                    leaf (all columns fixed) and evaluates leaves exactly.
//   * k2_branch    : creates children in pool slots (copy bounds + warm start, tighten one column).
// Everything here is integer arithmetic on the exact image of the model; fp64 only enters through
// K1's bounds and branching guidance.
#include <climits>

#include "device.h"
#include "nodepool.h"

namespace moip {
namespace {

__device__ __forceinline__ long long wsum(long long v) {
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

constexpr int kPropThreads = 128;

__global__ void __launch_bounds__(kPropThreads) k2_propagate_kernel(const DevModel dm, const PoolView pool, int B,
                                                                    const int* ids, const long long* obj_lo,
                                                                    const long long* obj_hi, int max_rounds, int* flag,
                                                                    long long* leaf_obj) {
  extern __shared__ int sm_i[];
  const int n = dm.n, ms = dm.ms, k = dm.k;
  int* lb = sm_i;
  int* ub = lb + n;
  int* nlb = ub + n;
  int* nub = nlb + n;
  __shared__ int s_changed, s_infeasible, s_unfixed;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kPropThreads / 32;
  for (int bi = blockIdx.x; bi < B; bi += gridDim.x) {
    const int slot = ids[bi];
    int* plb = pool.lb + (size_t)slot * n;
    int* pub = pool.ub + (size_t)slot * n;
    if (tid == 0) s_infeasible = 0;
    __syncthreads();                         // the reset must not overtake another warp's "crossed bounds" store below
    for (int j = tid; j < n; j += kPropThreads) {
      const int a = plb[j], b2 = pub[j];
      lb[j] = a; ub[j] = b2; nlb[j] = a; nub[j] = b2;
      if (a > b2) s_infeasible = 1;
    }
    for (int round = 0; round < max_rounds; ++round) {
      __syncthreads();                       // (A) bounds of this round visible
      if (s_infeasible) break;               // uniform: nobody writes it between (A) and (B)
      if (tid == 0) s_changed = 0;
      __syncthreads();                       // (B)
      for (int i = warp; i < ms + k; i += NW) {
        long long rlo, rhi;
        const bool dense = i >= ms;
        const long long* cv = dense ? dm.ci + (size_t)(i - ms) * n : nullptr;
        int e0, e1;
        if (dense) { rlo = obj_lo[i - ms]; rhi = obj_hi[i - ms]; e0 = 0; e1 = n; }
        else { rlo = dm.ri_lo[i]; rhi = dm.ri_hi[i]; e0 = dm.s_ptr[i]; e1 = dm.s_ptr[i + 1]; }
        if (rlo == LLONG_MIN && rhi == LLONG_MAX) continue;
        long long mn = 0, mx = 0;
        for (int e = e0 + lane; e < e1; e += 32) {
          const int j = dense ? e : dm.s_col[e];
          const long long a = dense ? cv[e] : dm.ai_val[e];
          if (a > 0) { mn += a * lb[j]; mx += a * ub[j]; }
          else if (a < 0) { mn += a * ub[j]; mx += a * lb[j]; }
        }
        mn = wsum(mn); mx = wsum(mx);
        if ((rhi != LLONG_MAX && mn > rhi) || (rlo != LLONG_MIN && mx < rlo)) { if (lane == 0) s_infeasible = 1; continue; }
        const bool use_hi = rhi != LLONG_MAX && mx > rhi;   // row can still be violated from above
        const bool use_lo = rlo != LLONG_MIN && mn < rlo;
        if (!use_hi && !use_lo) continue;
        for (int e = e0 + lane; e < e1; e += 32) {
          const int j = dense ? e : dm.s_col[e];
          const long long a = dense ? cv[e] : dm.ai_val[e];
          if (a == 0) continue;
          const int lj = lb[j], uj = ub[j];
          if (lj == uj) continue;
          if (use_hi) {
            const long long rest = mn - (a > 0 ? a * lj : a * uj);
            const long long room = rhi - rest;
            if (a > 0) { const long long t = floor_div(room, a); if (t < uj) { atomicMin(&nub[j], (int)max(t, (long long)INT_MIN / 2)); s_changed = 1; } }
            else { const long long t = ceil_div(room, a); if (t > lj) { atomicMax(&nlb[j], (int)min(t, (long long)INT_MAX / 2)); s_changed = 1; } }
          }
          if (use_lo) {
            const long long rest = mx - (a > 0 ? a * uj : a * lj);
            const long long need = rlo - rest;
            if (a > 0) { const long long t = ceil_div(need, a); if (t > lj) { atomicMax(&nlb[j], (int)min(t, (long long)INT_MAX / 2)); s_changed = 1; } }
            else { const long long t = floor_div(need, a); if (t < uj) { atomicMin(&nub[j], (int)max(t, (long long)INT_MIN / 2)); s_changed = 1; } }
          }
        }
      }
      __syncthreads();                       // (C)
      if (!s_changed || s_infeasible) break; // uniform: written before (C), rewritten only after the next (A)
      for (int j = tid; j < n; j += kPropThreads) {
        const int a = nlb[j], b2 = nub[j];
        lb[j] = a; ub[j] = b2;
        if (a > b2) s_infeasible = 1;
      }
    }
    __syncthreads();
    if (tid == 0) s_unfixed = 0;
    __syncthreads();
    if (!s_infeasible) {
      int unf = 0;
      for (int j = tid; j < n; j += kPropThreads) {
        plb[j] = lb[j]; pub[j] = ub[j];
        if (lb[j] != ub[j]) unf = 1;
      }
      if (unf) s_unfixed = 1;
    }
    __syncthreads();
    int f = s_infeasible ? 1 : (s_unfixed ? 0 : 2);
    if (f == 2) {
      // leaf: evaluate everything exactly at x = lb
      for (int i = warp; i < ms + k; i += NW) {
        const bool dense = i >= ms;
        long long a = 0;
        if (dense) { const long long* cv = dm.ci + (size_t)(i - ms) * n; for (int j = lane; j < n; j += 32) a += cv[j] * lb[j]; }
        else for (int e = dm.s_ptr[i] + lane; e < dm.s_ptr[i + 1]; e += 32) a += dm.ai_val[e] * lb[dm.s_col[e]];
        a = wsum(a);
        if (lane == 0) {
          if (dense) {
            leaf_obj[(size_t)bi * k + (i - ms)] = a;
            if (a < obj_lo[i - ms] || a > obj_hi[i - ms]) s_infeasible = 1;
          } else if (a < dm.ri_lo[i] || a > dm.ri_hi[i]) s_infeasible = 1;
        }
      }
      __syncthreads();
      if (s_infeasible) f = 1;
    }
    if (tid == 0) flag[bi] = f;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(128) k2_branch_kernel(const DevModel dm, const PoolView pool, int C, const BranchOp* ops) {
  const int n = dm.n, m = dm.m;
  for (int ci = blockIdx.x; ci < C; ci += gridDim.x) {
    const BranchOp op = ops[ci];
    const int* slb = pool.lb + (size_t)op.parent * n;
    const int* sub = pool.ub + (size_t)op.parent * n;
    int* dlb = pool.lb + (size_t)op.child * n;
    int* dub = pool.ub + (size_t)op.child * n;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      int a = slb[j], b2 = sub[j];
#pragma unroll
      for (int q = 0; q < 3; ++q)
        if (q < op.nv && j == op.var[q]) { a = max(a, op.new_lb[q]); b2 = min(b2, op.new_ub[q]); }
      dlb[j] = a; dub[j] = b2;
    }
    const double* sx = pool.wx + (size_t)op.parent * n;
    double* dx = pool.wx + (size_t)op.child * n;
    for (int j = threadIdx.x; j < n; j += blockDim.x) dx[j] = sx[j];
    const double* sy = pool.wy + (size_t)op.parent * m;
    double* dy = pool.wy + (size_t)op.child * m;
    for (int i = threadIdx.x; i < m; i += blockDim.x) dy[i] = sy[i];
  }
}


}  // namespace

int launch_k2_propagate(const DevModel& dm, const PoolView& pool, int B, const int* ids, const long long* obj_lo,
                        const long long* obj_hi, int max_rounds, int* flag, long long* leaf_obj, cudaStream_t st) {
  if (B <= 0) return MOIP_OK;
  const size_t smem = sizeof(int) * 4 * (size_t)dm.n;
  static LaunchCfg cfg;
  std::unique_lock<std::mutex> cfg_lock(launch_cfg_mutex());
  size_t& configured = cfg.configured[current_device()];
  if (smem > 48 * 1024 && smem > configured) {
    MOIP_CUDA(cudaFuncSetAttribute(k2_propagate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  cfg_lock.unlock();
  static LaunchCfg carve;
  if (set_aux_carveout(k2_propagate_kernel, carve)) return MOIP_ERR_CUDA;
  const int gcap = aux_grid_cap();
  int grid = B < gcap ? B : gcap;
  k2_propagate_kernel<<<grid, kPropThreads, smem, st>>>(dm, pool, B, ids, obj_lo, obj_hi, max_rounds, flag, leaf_obj);
  MOIP_CUDA(cudaGetLastError());
  return MOIP_OK;
}

int launch_k2_branch(const DevModel& dm, const PoolView& pool, int C, const BranchOp* ops, cudaStream_t st) {
  if (C <= 0) return MOIP_OK;
  static LaunchCfg carve;
  if (set_aux_carveout(k2_branch_kernel, carve)) return MOIP_ERR_CUDA;
  int grid = C < 148 * 8 ? C : 148 * 8;
  k2_branch_kernel<<<grid, 128, 0, st>>>(dm, pool, C, ops);
  MOIP_CUDA(cudaGetLastError());
  return MOIP_OK;
}

}  // namespace moip
