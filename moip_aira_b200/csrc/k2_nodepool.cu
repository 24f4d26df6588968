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
#include "k2_propagate.cuh"
#include "nodepool.h"

namespace moip {
namespace {

constexpr int kPropThreads = 256;

__global__ void __launch_bounds__(kPropThreads) k2_propagate_kernel(const DevModel dm, const PoolView pool, int B,
                                                                    const int* ids, const long long* obj_lo,
                                                                    const long long* obj_hi, int max_rounds, int* flag,
                                                                    long long* leaf_obj, const ChainRef ch) {
  extern __shared__ int sm_i[];
  __shared__ int s_flags[4];
  __shared__ long long s_act[(kPropThreads / 32) * 2 * MOIP_MAX_OBJ];
  if (ch.B_dev) B = *ch.B_dev;
  for (int bi = blockIdx.x; bi < B; bi += gridDim.x) {
    const int slot = ch.B_dev ? ch.slot_base + bi : ids[bi];
    const int f = k2::propagate_node<kPropThreads>(dm, pool.lb + (size_t)slot * dm.n, pool.ub + (size_t)slot * dm.n, obj_lo, obj_hi,
                                                   max_rounds, sm_i, s_act, s_flags, leaf_obj + (size_t)bi * dm.k);
    if (threadIdx.x == 0) {
      flag[bi] = f;
      if (f == 2 && ch.inc) {              // a leaf is a verified point: it may be the new incumbent (chained rounds)
        const long long v = (long long)dm.sgn * leaf_obj[(size_t)bi * dm.k + *ch.cost];
        if (v < atomicMin(ch.inc, v)) *ch.cutoff = (double)v;
      }
    }
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
                        const long long* obj_hi, int max_rounds, int* flag, long long* leaf_obj, cudaStream_t st,
                        const ChainRef& ch) {
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
  k2_propagate_kernel<<<grid, kPropThreads, smem, st>>>(dm, pool, B, ids, obj_lo, obj_hi, max_rounds, flag, leaf_obj, ch);
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
