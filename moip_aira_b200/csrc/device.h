// Device-side views shared by the kernels (sm_100a) and the host driver.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <mutex>

#include "../../include/moip_b200.h"

namespace moip {

#define MOIP_CUDA(call)                                                                      \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      std::fprintf(stderr, "moip_b200: CUDA error %s at %s:%d (%s)\n", cudaGetErrorString(e_), \
                   __FILE__, __LINE__, #call);                                               \
      return MOIP_ERR_CUDA;                                                                  \
    }                                                                                        \
  } while (0)

// Model image in HBM (read-only for all kernels; ~nnz*12 + k*n*16 bytes, L1/L2 resident).
struct DevModel {
  int n, ms, k, m, ell_w, nnz;
  double sgn;                 // +1 MIN, -1 MAX: kernels work on the min-form  min sgn*c x
  double eta;                 // 0.98 / ||S||_2
  double norm_row_bounds2;    // sum of squares of finite *unscaled* structural rhs
  // scaled LP image
  const double* ellT_val;     // [ell_w][n]  structural part of S^T, column-ELL
  const int* ellT_row;        // [ell_w][n]
  const int* s_ptr;           // [ms+1]      structural part of S, CSR
  const int* s_col;           // [nnz]
  const double* s_val;        // [nnz]
  const double* D;            // [k][n]      dense scaled objective block (rows ms..m-1 of S)
  const double* s_lo;         // [ms] scaled structural row bounds (+-inf)
  const double* s_hi;
  const double* dr;           // [m]
  const double* dc;           // [n]
  // fast-kernel image (rows in kernel order: short structural | k objective | long structural)
  int fast_ok, msS, nL, KD, RW, ell2_w;
  const double* rowell_val;   // [RW][msS]
  const int* rowell_col;
  const double* ellT2_val;    // [ell2_w][n]
  const int* ellT2_row;
  const double* D2;           // [KD][n]
  int col_units;
  const double* colrec;       // [n][col_units] packed column records (see model.h)
  const double* rowrec;       // [RW][msS][2]   packed row-ELL entries
  int reg_ok, RWP, reg_lpr_log2, reg_trips;   // register-resident kernel image (k1_reg.cuh)
  const double* colrec2;      // [n][col_units] ids packed as (product byte offset << 16 | dual byte offset)
  const double* dr_k;         // [m]
  const double* lo_k;         // [m] scaled structural bounds in kernel order
  const double* hi_k;
  // exact integer image
  const long long* ai_val;    // [nnz] (pattern s_ptr/s_col)
  const long long* ri_lo;     // [ms]  LLONG_MIN = free
  const long long* ri_hi;     // [ms]  LLONG_MAX = free
  const long long* ci;        // [k][n]
  const int* lbI;             // [n]
  const int* ubI;             // [n]
};

// One batch of node LPs (K1).  All arrays are device pointers.
struct LpBatch {
  int B;
  const int* cost_idx;        // [B]
  const double* rhs;          // [B][k] objective-bound rows in the model's own sign (+-1e20 free)
  int* lb;                    // [B][n] (or [slots][n] with `slot`) integer column bounds of the node;
  int* ub;                    //        tightened in place when rc_fix is set
  const int* slot;            // optional indirection: node b lives in row slot[b] of lb/ub/warm_*/out_* (node pool)
  int rc_fix;                 // reduced-cost bound tightening against the cutoff at node end (fast kernel)
  const double* warm_x;       // [B][n] unscaled warm start or nullptr
  const double* warm_y;       // [B][m]
  double* out_x;              // [B][n] or nullptr
  double* out_y;              // [B][m] or nullptr
  double* primal_obj;         // [B] min-form objective of the last primal iterate
  double* dual_bound;         // [B] best valid Lagrangian bound seen (min-form)
  int* status;                // [B] MOIP_LP_*
  int* iters;                 // [B]
  int* branch_var;            // [B][3] the three most fractional columns (-1 = none), best first
  double* branch_val;         // [B][3] their (fractional) values
  const int* skip;            // [B] nonzero: node already decided by K2, do not solve (or nullptr)
  int cost_stride, rhs_stride;// 0 = all nodes share cost_idx[0] / rhs[0..k)
  const double* cutoff;       // device scalar (min-form); node stops once bound >= *cutoff - cutoff_slack
  int* work_counter;          // device int, zeroed before launch (dynamic node scheduling)
  double* scratch;            // generic kernel, large models: per-CTA iterate storage in HBM (nullptr = shared memory)
  size_t scratch_stride;      // doubles per CTA (k1_scratch_stride)
  int scratch_slots;          // CTAs the scratch has room for
  // Fused B&B round (register-resident K1 only, `slot` set): the CTA that pulls a node first propagates it (K2,
  // k2_propagate.cuh) and, once its LP is solved, rounds the LP point three ways and verifies the candidates exactly
  // (what k4_round_verify_kernel does) -- one launch per round instead of three, the node's bounds read once.
  int farkas;                 // plain batch on the register-resident K1: also test the Farkas certificate (k1_reg.cuh FARKAS)
  int fused;                  // 0 = plain LP batch (the f_* fields are ignored)
  const long long* f_obj_lo;  // [k] integer limits of the objective rows for the propagation (incl. the incumbent cut-off)
  const long long* f_obj_hi;
  int f_max_rounds;
  int* f_flag;                // [B] out: 0 open / 1 infeasible / 2 leaf
  long long* f_leaf_obj;      // [B][k] exact objective values of leaves
  int* f_xr;                  // [B][3][n] rounded candidates (nearest / down / up, clipped to the node's box)
  long long* f_cand_obj;      // [B][3][k]
  unsigned char* f_cand_feas; // [B][3] structural rows satisfied
  int* f_first_free;          // [B][3] first unfixed column, its lb, ub
  // Chained rounds (bbchain.h): the batch size and the incumbent live on the device, so that the rounds of one IP can be
  // enqueued back to back without a host round trip in between.
  const int* B_dev;           // non-null: the batch size is *B_dev (B is then only the grid hint of the launch)
  int slot_base;              // no `slot` array: node b lives in row slot_base + b
  long long* f_inc;           // non-null: min-form incumbent value; every verified candidate / leaf that beats it lowers it (atomicMin)
  const long long* f_lim_lo;  // [k] the IP's own objective limits (f_obj_* without the incumbent cut-off): candidates must respect them
  const long long* f_lim_hi;
  double* f_cutoff_rw;        // == cutoff, writable: lowered together with f_inc
};

// Chained rounds (bbchain.h) for the helper kernels K2 / K4: where the batch size, the rows of the nodes and the incumbent
// live on the device.  B_dev == nullptr: plain launch (everything in the arguments).
struct ChainRef {
  const int* B_dev = nullptr;        // batch size
  int slot_base = 0;                 // node b lives in pool row slot_base + b (no ids array)
  long long* inc = nullptr;          // min-form incumbent value, lowered with atomicMin
  double* cutoff = nullptr;          // (double)*inc for K1
  const int* cost = nullptr;         // optimised objective
  const long long* lim_lo = nullptr; // [k] the IP's own objective limits
  const long long* lim_hi = nullptr;
};

struct LpParams {
  double eps;
  int max_iter;
  int check_every;
  int fixed_iters;
  int norm_every;             // restart criteria evaluated every this many iterations
  double cutoff_slack;        // stop when bound >= cutoff - slack (integer objectives: 1 - 1e-6)
  int int_obj;                // objective is integer valued: stop once ceil(bound) cannot rise any more
};

// serialises the one-time cudaFuncSetAttribute blocks of the launchers (worker threads share the kernels)
std::mutex& launch_cfg_mutex();
// cudaFuncSetAttribute / the shared-memory carve-out apply to the CURRENT device only: one process may hold contexts on
// several GPUs (Seam 1 MOIP_B200_DEVICES, one worker per GPU), so every launcher keeps its one-time state per device
constexpr int kMaxDevices = 64;
struct LaunchCfg {
  size_t configured[kMaxDevices] = {};
  int occ[kMaxDevices] = {};
};
inline int current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
  return d;
}

int launch_k1_any(const DevModel& dm, const LpBatch& b, const LpParams& p, int num_sms, cudaStream_t st);   // best path the model allows
int launch_k1_reg(const DevModel& dm, const LpBatch& b, const LpParams& p, int num_sms, cudaStream_t st);
bool k1_small_applies(const DevModel& dm);   // all rows dense, n <= 64: 8 lanes per node (k1_small.cu)
int launch_k1_small(const DevModel& dm, const LpBatch& b, const LpParams& p, int num_sms, cudaStream_t st);
int launch_k1_fast(const DevModel& dm, const LpBatch& b, const LpParams& p, int num_sms, cudaStream_t st);
int launch_k1(const DevModel& dm, const LpBatch& b, const LpParams& p, int num_sms, cudaStream_t st);
size_t k1_scratch_stride(const DevModel& dm);   // > 0: the generic kernel needs LpBatch::scratch for this model
int launch_expand_masks(const DevModel& dm, int B, const uint32_t* masks, int mask_words, int* lb, int* ub,
                        cudaStream_t st);

// K3: one cached subproblem = one 64-byte record (ip: the k bounds it was solved under,
// result: its lexicographic optimum, reference src/result.h:10-20)
struct alignas(16) CacheRecord {
  double ip[MOIP_MAX_OBJ];
  int result[MOIP_MAX_OBJ];
  int infeasible;
  int pad[3];
};
static_assert(sizeof(CacheRecord) == 64, "cache record must be 64 bytes");
struct DevCache {
  int k, size;
  const CacheRecord* rec;     // [size], insertion order
};
// searches up to two stores per query (infeasibles first, then solutions: reference
// src/aira.cpp:816-823); first_match[q] = index in store `which[q]` (0/1) or -1
int launch_k3(const DevCache& c0, const DevCache& c1, int Q, const double* queries, int sense, int* first_match,
              int* which, cudaStream_t st);

// one query, answered into mapped pinned host memory (the generator's scan; see k3_scan1_kernel)
struct K3Query { double ip[MOIP_MAX_OBJ]; };
struct alignas(16) K3Answer {
  CacheRecord rec;            // the matching record (valid when first_match >= 0)
  int first_match, which;     // index in store `which` (0/1), or -1
  int seq, pad;               // written last: the host polls it
};
int launch_k3_one(const DevCache& c0, const DevCache& c1, const K3Query& q, int sense, K3Answer* answer_dev, int seq,
                  cudaStream_t st);

// K4
int launch_k4_round(const DevModel& dm, int B, const int* slot, const double* wx, const int* lb, const int* ub,
                    int* xr /*[B][3][n]*/, long long* obj_out /*[B][3][k]*/, unsigned char* feasible_out /*[B][3]*/,
                    int* first_free /*[B][3]: first unfixed column, its lb, ub (or nullptr)*/,
                    const int* skip /*[B] nonzero: node decided by K2, not rounded (or nullptr)*/, cudaStream_t st,
                    const ChainRef& ch = ChainRef());
// Largest grid of the short per-round helper kernels (K2 propagate / K4 round).  Their CTAs are latency-bound and each
// one that lands on an SM holds registers a K1 CTA of another worker could use (K1: 2 CTAs x 32 K registers fill an SM),
// so they are kept on few SMs and loop over the nodes instead of spreading one CTA per node (MOIP_AUX_GRID overrides).
int aux_grid_cap();
// Shared-memory carve-out of the helper kernels (K2 / K3 / K4), percent, or -1 = the driver's default (MOIP_AUX_CARVEOUT).
// Kernels with different carve-outs cannot share an SM, and switching an SM over needs it drained.
int aux_carveout_pct();
template <class K>
inline int set_aux_carveout(K kern, LaunchCfg& cfg) {
  const int pct = aux_carveout_pct();
  if (pct < 0) return MOIP_OK;
  std::lock_guard<std::mutex> lk(launch_cfg_mutex());
  const int dev = current_device();
  if (cfg.occ[dev] == pct + 1) return MOIP_OK;
  MOIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
  cfg.occ[dev] = pct + 1;
  return MOIP_OK;
}
int launch_k4(const DevModel& dm, int B, const int* x, const double* rhs, long long* obj_out,
              unsigned char* feasible_out, cudaStream_t st);

}  // namespace moip
