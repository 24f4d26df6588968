// Node pool shared between the host B&B driver and the K2 kernels.
#pragma once
#include "device.h"

namespace moip {

// Pool slot s: integer column bounds + warm start of one open B&B node (HBM resident).
struct PoolView {
  int* lb;      // [slots][n]
  int* ub;      // [slots][n]
  double* wx;   // [slots][n]  unscaled LP primal iterate of the parent / of the node itself
  double* wy;   // [slots][m]
};

struct BranchOp {           // child = copy of parent with up to 3 columns tightened
  int parent, child, nv;
  int var[3], new_lb[3], new_ub[3];
};

int launch_k2_propagate(const DevModel& dm, const PoolView& pool, int B, const int* ids, const long long* obj_lo,
                        const long long* obj_hi, int max_rounds, int* flag, long long* leaf_obj, cudaStream_t st,
                        const ChainRef& ch = ChainRef());
int launch_k2_branch(const DevModel& dm, const PoolView& pool, int C, const BranchOp* ops, cudaStream_t st);

}  // namespace moip
