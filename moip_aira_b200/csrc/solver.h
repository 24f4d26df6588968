// Host driver state behind the opaque handles of include/moip_b200.h.
#pragma once
#include <chrono>
#include <array>
#include <atomic>
#include <mutex>
#include <set>
#include <shared_mutex>
#include <string>
#include <vector>

#include "device.h"
#include "model.h"
#include "nodepool.h"
#include "bbchain.h"

namespace moip {

// Feasible integer points met while solving -- the optimiser of every IP any context of this model has solved -- kept as
// MIP starts for later IPs.  CPLEX keeps the last solution as a start when only right-hand sides change (what the
// reference gets for free between the CPXmipopt calls of solve(), src/aira.cpp:467-517, and from one subproblem to the
// next); here a new IP starts from the best stored point that satisfies its objective bounds.  Points are exact facts
// (verified in int64 before they became incumbents), so starting from one never changes a result, only the node count.
struct PointStore {
  static constexpr size_t kMaxBytes = (size_t)256 << 20;
  std::shared_mutex mu;
  int n = 0, k = 0;
  bool narrow = false;                 // every column fits int8 (binaries): 1 byte per column
  std::vector<long long> obj;          // [P][k]
  std::vector<int8_t> x8;              // [P][n]
  std::vector<int> x32;                // [P][n]
  std::set<std::array<long long, MOIP_MAX_OBJ>> seen;   // one point per objective vector is enough
  long long hits = 0, queries = 0;
  void init(int n_, int k_, bool narrow_) { n = n_; k = k_; narrow = narrow_; }
  size_t size() const { return k ? obj.size() / (size_t)k : 0; }
  void add(const int* x, const long long* ov) {
    std::array<long long, MOIP_MAX_OBJ> key{};
    for (int o = 0; o < k; ++o) key[o] = ov[o];
    std::unique_lock<std::shared_mutex> lk(mu);
    if ((size() + 1) * (size_t)n * (narrow ? 1 : 4) > kMaxBytes) return;
    if (!seen.insert(key).second) return;
    obj.insert(obj.end(), ov, ov + k);
    if (narrow) for (int j = 0; j < n; ++j) x8.push_back((int8_t)x[j]);
    else x32.insert(x32.end(), x, x + n);
  }
  // best stored point (smallest sgn * obj[cost]) with olo <= obj <= ohi and value < below (min-form); false = none
  bool best(int cost, long long sgn, const long long* olo, const long long* ohi, long long below, std::vector<int>& x_out,
            long long& val_out) {
    std::shared_lock<std::shared_mutex> lk(mu);
    const size_t P = size();
    size_t arg = P;
    long long bv = below;
    for (size_t p = 0; p < P; ++p) {
      const long long* ov = obj.data() + p * k;
      const long long v = sgn * ov[cost];
      if (v >= bv) continue;
      bool ok = true;
      for (int o = 0; o < k && ok; ++o) ok = ov[o] >= olo[o] && ov[o] <= ohi[o];
      if (ok) { bv = v; arg = p; }
    }
    if (arg == P) return false;
    x_out.resize(n);
    if (narrow) for (int j = 0; j < n; ++j) x_out[j] = x8[arg * n + j];
    else for (int j = 0; j < n; ++j) x_out[j] = x32[arg * n + j];
    val_out = bv;
    return true;
  }
};

}  // namespace moip

struct moip_model {
  moip::Model M;
  moip::PointStore points;
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
  std::vector<moip::CacheRecord*> retired;   // device arrays outgrown while scans of other workers may still read them: freed with the store
  size_t synced = 0;                     // records [0, synced) are on the device
  std::mutex mu;                         // the stores are shared by the worker threads of a pool (src/solutions.h:41-44)
  int sync_to_device(cudaStream_t st);   // caller holds mu
  moip::DevCache view() const;           // caller holds mu; the snapshot stays valid after the lock is released
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
  moip::K3Answer* k3_answer = nullptr;       // mapped pinned: the single-query scan writes its answer here
  moip::K3Answer* k3_answer_dev = nullptr;   // the same memory as the device sees it
  int k3_seq = 0;
  bool k3_poll = true;                       // MOIP_K3_POLL=0: always the copy + synchronise path
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
  // how a worker waits for its round: spinning in cudaStreamSynchronize (lowest latency, one core per worker) or sleeping on
  // a blocking event (frees the core; ~20 us later).  MOIP_SYNC=spin|block|auto; auto blocks when the workers of all ranks
  // on this host outnumber its cores.
  bool block_sync = false;
  cudaEvent_t sync_ev = nullptr;
  int wait_stream();
  bool use_fused = true;       // register-resident K1: one fused propagate -> LP -> round/verify launch per round (MOIP_FUSED_ROUND=0: three kernels)
  bool use_points = true;      // start every IP from the best stored feasible point (PointStore); MOIP_POINT_STORE=0 disables
  long long start_hits = 0;    // IPs that began with a stored point as incumbent

  // CUDA-event split per kernel class (moip_ctx_set_kernel_timing)
  bool ktiming = false, kev_branch_pending = false;
  cudaEvent_t kev[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  moip_kernel_times ktimes{};
  void kmark(int i) { if (ktiming) cudaEventRecord(kev[i], stream); }
  double kspan(int a, int b) { float ms = 0.f; return cudaEventElapsedTime(&ms, kev[a], kev[b]) == cudaSuccess ? (double)ms : 0.0; }

  long long stage_ips[MOIP_MAX_OBJ + 1] = {0, 0, 0, 0, 0}, stage_nodes[MOIP_MAX_OBJ + 1] = {0, 0, 0, 0, 0},
            stage_lps[MOIP_MAX_OBJ + 1] = {0, 0, 0, 0, 0}, stage_rounds[MOIP_MAX_OBJ + 1] = {0, 0, 0, 0, 0},
            stage_root_solved[MOIP_MAX_OBJ + 1] = {0, 0, 0, 0, 0};   // per stage of the lexicographic chain (MOIP_PROFILE_ROUNDS)
  int cur_stage = MOIP_MAX_OBJ;       // stage of the IP being solved (MOIP_MAX_OBJ = outside lex_solve)
  // what the worker is doing right now, for the pool's watchdog (MOIP_WATCHDOG=<seconds>): 0 idle, 1 cache scan, 2 B&B round
  std::atomic<int> dbg_where{0};
  std::atomic<long long> dbg_rounds{0}, dbg_open{0}, dbg_strip{-1};
  double prof_t[7] = {0, 0, 0, 0, 0, 0, 0};   // MOIP_PROFILE_ROUNDS: enqueue / device wait / host seconds, rounds, nodes

  // ---- chained rounds (bbchain.h): the tree of an IP advances on the device, the host looks in once per chunk of rounds
  bool batch_farkas = false;      // moip_lp_batch_* on the register-resident K1: Farkas certificate in the termination test (MOIP_K1_BATCH_FARKAS=1);
                                  // off, infeasible node LPs are recognised by the Lagrangian bound alone, as in round 1
  long long ip_node_budget = 0;   // > 0: solve_ip gives up (MOIP_ERR_BUDGET) beyond this many nodes (moip_ctx_set_ip_node_budget)
  bool use_chain = true;       // MOIP_CHAIN=0: every round through the host (solve_ip's own loop)
  bool chain_debug = false;    // MOIP_CHAIN_DEBUG=1: synchronise after every launch and say which one failed
  int chain_q = 4096;          // pool rows per parity = widest tree level the device handles (MOIP_CHAIN_Q)
  moip::DBuf<moip::BbCtl> d_ctl;
  moip::HBuf<moip::BbCtl> h_ctl;
  moip::HBuf<int> h_inc;
  moip::DBuf<double> c_bound;
  moip::DBuf<int> c_depth;
  double chain_rounds_avg[MOIP_MAX_OBJ + 1] = {6, 6, 6, 6, 6};   // rounds per IP of each stage (sizes the first chunk)
  long long chain_ips = 0, chain_fallbacks = 0, chain_chunks = 0, chain_idle_rounds = 0;
  // returns MOIP_OK with `handled` set, or leaves `handled` false (level wider than chain_q ...): the caller's loop takes over,
  // starting from the incumbent found so far (inc_val / inc_x updated)
  int solve_ip_chained(int cost, const double* srhs, const long long* olo, const long long* ohi, bool& have_inc,
                       long long& inc_val, std::vector<int>& inc_x, bool& handled);

  moip::DBuf<double> k1_scratch;   // streaming mode of the generic K1 kernel (models too large for shared memory)
  int attach_k1_scratch(moip::LpBatch& b);
  int ensure_pool(int slots);
  int alloc_slot();
  int solve_ip(int cost, const double* srhs, const std::vector<int>* inc_x, moip::IpResult& out);
  int lex_solve(const int* perm, int n_obj, const double* rhs, int* result, int* mip_status);
  int get_limit(int obj, int sense, const double* rhs, int* result, int* mip_status);
};
