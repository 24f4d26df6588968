// Host driver state behind the opaque handles of include/moip_b200.h.
#pragma once
#include <chrono>
#include <mutex>
#include <string>
#include <vector>

#include "device.h"
#include "model.h"
#include "nodepool.h"

struct moip_model {
  moip::Model M;
};

namespace moip {

template <class T>
struct DBuf {   // grow-only device buffer
  T* p = nullptr;
  size_t cap = 0;
  int ensure(size_t n, bool keep = false, cudaStream_t st = 0) {
    if (n <= cap) return MOIP_OK;
    size_t nc = cap ? cap : 16;
    while (nc < n) nc *= 2;
    T* q = nullptr;
    MOIP_CUDA(cudaMalloc(&q, nc * sizeof(T)));
    if (keep && p && cap) {
      MOIP_CUDA(cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st));
      MOIP_CUDA(cudaStreamSynchronize(st));
    }
    if (p) cudaFree(p);
    p = q; cap = nc;
    return MOIP_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <class T>
struct HBuf {   // grow-only pinned host buffer
  T* p = nullptr;
  size_t cap = 0;
  int ensure(size_t n) {
    if (n <= cap) return MOIP_OK;
    size_t nc = cap ? cap : 16;
    while (nc < n) nc *= 2;
    if (p) cudaFreeHost(p);
    p = nullptr;
    MOIP_CUDA(cudaMallocHost(&p, nc * sizeof(T)));
    cap = nc;
    return MOIP_OK;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct IpResult {
  int status = MOIP_MIP_INFEASIBLE;
  long long obj = 0;             // model-sense objective value of the optimised objective
  std::vector<int> x;
};

}  // namespace moip

struct moip_cache {
  moip_ctx* ctx = nullptr;
  int device = 0;                        // copy: the cache may outlive its context
  int k = 0;
  std::vector<moip::CacheRecord> host;   // insertion order
  moip::DBuf<moip::CacheRecord> dev;
  size_t synced = 0;                     // records [0, synced) are on the device
  std::mutex mu;                         // the stores are shared by the worker threads of a pool (src/solutions.h:41-44)
  int sync_to_device(cudaStream_t st);
  moip::DevCache view() const;
};

struct moip_ctx {
  moip_model* model = nullptr;
  int device = 0;
  cudaStream_t stream = 0;
  bool owns_stream = false;   // moip_ctx_create_own_stream: destroyed with the context
  int num_sms = 148;
  moip::DevModel dm{};
  std::vector<void*> model_allocs;
  moip_stats stats{};
  // ---- K1 batch API (resident batch)
  int batch_B = 0;
  moip::DBuf<int> b_cost, b_lb, b_ub, b_status, b_iters, b_branch, b_counter;
  moip::DBuf<double> b_rhs, b_pobj, b_dbound, b_x, b_cutoff;
  moip::DBuf<uint32_t> b_masks;
  // ---- K3 / K4 staging
  moip::DBuf<double> q_ip;
  moip::DBuf<int> q_out, q_which;
  moip::HBuf<int> h_q;
  moip::DBuf<int> v_x;
  moip::DBuf<double> v_rhs;
  moip::DBuf<long long> v_obj;
  moip::DBuf<unsigned char> v_feas;
  // ---- branch and bound
  int pool_slots = 0;
  moip::DBuf<int> p_lb, p_ub;
  moip::DBuf<double> p_wx, p_wy;
  std::vector<int> free_slots;
  moip::DBuf<int> r_xr, r_counter;      // rounded candidates [B][3][n]; K1 work counter
  moip::DBuf<double> r_pobj;
  moip::DBuf<moip::BranchOp> r_ops;
  moip::HBuf<unsigned char> h_round;    // packed D2H results of one round
  moip::DBuf<unsigned char> r_in, r_out; // one H2D block in / one D2H block out per round (solve_ip)
  moip::HBuf<unsigned char> h_in;
  moip::HBuf<moip::BranchOp> h_ops;
  moip::DBuf<double> d_root_x, d_root_y;            // [k][n] / [k][m] root iterate per objective (warm start of the next IP)
  std::vector<char> root_valid;
  moip::DBuf<int> d_inc;                            // incumbent of the IP being solved
  // tunables (env MOIP_*)
  int bb_batch = 0;            // 0 = SMs * occupancy
  int bb_max_iter = 0;            // fixed node-LP iteration cap; 0 = controlled (see solve_ip)
  double bb_cap_lo = 0.38, bb_cap_hi = 0.50;   // share of a round's LPs allowed to hit the cap
  int lp_cap_dyn = 0;             // current cap
  long long cap_seen = 0, cap_hit = 0;
  double bb_eps = 1e-5;
  int bb_check = 32;
  int norm_every = 1;
  int bb_levels = 3;           // max tree levels expanded per round while the device is under-filled

  double prof_t[7] = {0, 0, 0, 0, 0, 0, 0};   // MOIP_PROFILE_ROUNDS: enqueue / device wait / host seconds, rounds, nodes

  moip::DBuf<double> k1_scratch;   // streaming mode of the generic K1 kernel (models too large for shared memory)
  int attach_k1_scratch(moip::LpBatch& b);
  int ensure_pool(int slots);
  int alloc_slot();
  int solve_ip(int cost, const double* srhs, const std::vector<int>* inc_x, moip::IpResult& out);
  int lex_solve(const int* perm, int n_obj, const double* rhs, int* result, int* mip_status);
  int get_limit(int obj, int sense, const double* rhs, int* result, int* mip_status);
};
