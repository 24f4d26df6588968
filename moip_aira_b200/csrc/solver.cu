// Host driver: contexts, the batched node-LP API, the solution cache, int64 verification and the
// branch-and-bound + lexicographic chain that replace the bodies of solve() / get_limit()
// (reference src/aira.cpp:452-536, :367-450).  Host C++ keeps the control; all arithmetic on
// model data runs in the kernels of k1_pdhg.cu / k2_nodepool.cu / k3_k4.cu.
#include "solver.h"

#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>

using namespace moip;

std::mutex& moip::launch_cfg_mutex() {
  static std::mutex m;
  return m;
}

int moip::aux_carveout_pct() {
  static const int pct = [] {
    const char* s = std::getenv("MOIP_AUX_CARVEOUT");
    return s ? std::atoi(s) : -1;
  }();
  return pct;
}

int moip::aux_grid_cap() {
  static const int cap = [] {
    const char* s = std::getenv("MOIP_AUX_GRID");
    const int v = s ? std::atoi(s) : 0;
    return v > 0 ? v : 148 * 8;
  }();
  return cap;
}

namespace {

// Worker contexts put one stream each on the device; with the default of 8 hardware work queues, streams share a queue
// and a small copy of one worker waits behind another worker's K1 launch (measured: 109 us of copy time per B&B round
// with 12 workers, 28 us with 32 queues).  The driver reads the variable when it creates the device context, so it is
// set when the library is loaded -- an explicit setting by the user wins.
struct WorkQueueEnv {
  WorkQueueEnv() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }
} g_work_queue_env;

template <class T>
int upload(moip_ctx* c, const std::vector<T>& v, const T** out) {
  T* p = nullptr;
  size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  MOIP_CUDA(cudaMalloc(&p, bytes));
  if (!v.empty()) MOIP_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  c->model_allocs.push_back(p);
  *out = p;
  return MOIP_OK;
}

int env_int(const char* name, int dflt) {
  const char* s = std::getenv(name);
  return s ? std::atoi(s) : dflt;
}
double env_double(const char* name, double dflt) {
  const char* s = std::getenv(name);
  return s ? std::atof(s) : dflt;
}

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// ------------------------------------------------------------------------------------ context
extern "C" int moip_ctx_create(moip_model* m, int device, void* stream, moip_ctx** out) {
  if (!m || !out) return MOIP_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    std::fprintf(stderr, "moip_b200: no CUDA device available (this library has no CPU fallback)\n");
    return MOIP_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) return MOIP_ERR_ARG;
  MOIP_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  MOIP_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    std::fprintf(stderr, "moip_b200: device %d is sm_%d%d; kernels are built for sm_100a only\n", device, prop.major, prop.minor);
    return MOIP_ERR_CUDA;
  }
  moip_ctx* c = new moip_ctx();
  c->model = m;
  c->device = device;
  c->stream = (cudaStream_t)stream;
  c->num_sms = prop.multiProcessorCount;
  const Model& M = m->M;
  DevModel& d = c->dm;
  d.n = M.n; d.ms = M.ms; d.k = M.k; d.m = M.m; d.ell_w = M.ell_w; d.nnz = (int)M.a_val.size();
  d.sgn = M.sense == 0 ? 1.0 : -1.0;
  d.eta = M.eta;
  double nb2 = 0;
  for (int i = 0; i < M.ms; ++i) nb2 += M.rhs[i] * M.rhs[i];
  d.norm_row_bounds2 = nb2;
  int rc = 0;
  std::vector<long long> ai(M.ai_val.begin(), M.ai_val.end()), rlo(M.ri_lo.begin(), M.ri_lo.end()),
      rhi(M.ri_hi.begin(), M.ri_hi.end()), ci(M.ci.begin(), M.ci.end());
  for (auto& v : rlo) if (v == INT64_MIN) v = LLONG_MIN;
  for (auto& v : rhi) if (v == INT64_MAX) v = LLONG_MAX;
  rc |= upload(c, M.ellT_val, &d.ellT_val);
  rc |= upload(c, M.ellT_row, &d.ellT_row);
  rc |= upload(c, M.a_ptr, &d.s_ptr);
  rc |= upload(c, M.a_col, &d.s_col);
  rc |= upload(c, M.s_val, &d.s_val);
  rc |= upload(c, M.D, &d.D);
  rc |= upload(c, M.s_lo, &d.s_lo);
  rc |= upload(c, M.s_hi, &d.s_hi);
  rc |= upload(c, M.dr, &d.dr);
  rc |= upload(c, M.dc, &d.dc);
  d.fast_ok = M.fast_ok ? 1 : 0; d.msS = M.msS; d.nL = M.nL; d.KD = M.KD; d.RW = M.RW; d.ell2_w = M.ell2_w;
  if (std::getenv("MOIP_K1_GENERIC")) { d.fast_ok = 0; d.reg_ok = 0; }
  rc |= upload(c, M.rowell_val, &d.rowell_val);
  rc |= upload(c, M.rowell_col, &d.rowell_col);
  rc |= upload(c, M.ellT2_val, &d.ellT2_val);
  rc |= upload(c, M.ellT2_row, &d.ellT2_row);
  rc |= upload(c, M.D2, &d.D2);
  d.col_units = M.col_units;
  rc |= upload(c, M.colrec, &d.colrec);
  rc |= upload(c, M.rowrec, &d.rowrec);
  d.reg_ok = M.reg_ok ? 1 : 0; d.RWP = M.RWP; d.reg_lpr_log2 = M.reg_lpr_log2; d.reg_trips = M.reg_trips;
  if (std::getenv("MOIP_K1_NOREG") || std::getenv("MOIP_K1_GENERIC")) d.reg_ok = 0;
  rc |= upload(c, M.colrec2, &d.colrec2);
  rc |= upload(c, M.dr_k, &d.dr_k);
  rc |= upload(c, M.lo_k, &d.lo_k);
  rc |= upload(c, M.hi_k, &d.hi_k);
  rc |= upload(c, ai, &d.ai_val);
  rc |= upload(c, rlo, &d.ri_lo);
  rc |= upload(c, rhi, &d.ri_hi);
  rc |= upload(c, ci, &d.ci);
  rc |= upload(c, M.lbI, &d.lbI);
  rc |= upload(c, M.ubI, &d.ubI);
  if (rc) { moip_ctx_destroy(c); return MOIP_ERR_CUDA; }
  c->root_valid.assign(M.k, 0);
  c->bb_batch = env_int("MOIP_BB_BATCH", 0);
  c->bb_max_iter = env_int("MOIP_BB_MAX_ITER", 0);
  c->bb_cap_lo = env_double("MOIP_BB_CAP_LO", 0.38);
  c->bb_cap_hi = env_double("MOIP_BB_CAP_HI", 0.50);
  c->bb_eps = env_double("MOIP_BB_EPS", 1e-5);
  c->bb_check = env_int("MOIP_BB_CHECK", 32);
  c->norm_every = env_int("MOIP_NORM_EVERY", 16);
  c->bb_levels = env_int("MOIP_BB_LEVELS", 3);
  c->use_points = env_int("MOIP_POINT_STORE", 1) != 0;
  c->use_fused = env_int("MOIP_FUSED_ROUND", 1) != 0;
  c->use_chain = env_int("MOIP_CHAIN", 1) != 0;
  c->batch_farkas = env_int("MOIP_K1_BATCH_FARKAS", 0) != 0;
  c->chain_q = std::max(64, env_int("MOIP_CHAIN_Q", 4096));
  c->chain_debug = env_int("MOIP_CHAIN_DEBUG", 0) != 0;
  c->k3_poll = env_int("MOIP_K3_POLL", 1) != 0;
  if (const char* sm = std::getenv("MOIP_SYNC")) c->block_sync = std::strcmp(sm, "block") == 0;
  {
    std::unique_lock<std::shared_mutex> lk(m->points.mu);
    if (m->points.n == 0) {
      bool narrow = true;
      for (int j = 0; j < M.n; ++j) narrow = narrow && M.lbI[j] >= -127 && M.ubI[j] <= 127;
      m->points.init(M.n, M.k, narrow);
    }
  }
  if (env_int("MOIP_KERNEL_TIMING", 0) && moip_ctx_set_kernel_timing(c, 1)) { moip_ctx_destroy(c); return MOIP_ERR_CUDA; }
  *out = c;
  return MOIP_OK;
}

int moip_ctx::wait_stream() {
  if (!block_sync) { MOIP_CUDA(cudaStreamSynchronize(stream)); return MOIP_OK; }
  if (!sync_ev) MOIP_CUDA(cudaEventCreateWithFlags(&sync_ev, cudaEventBlockingSync | cudaEventDisableTiming));
  MOIP_CUDA(cudaEventRecord(sync_ev, stream));
  MOIP_CUDA(cudaEventSynchronize(sync_ev));
  return MOIP_OK;
}

extern "C" int moip_ctx_set_sync_mode(moip_ctx* c, int blocking) {
  if (!c) return MOIP_ERR_ARG;
  c->block_sync = blocking != 0;
  return MOIP_OK;
}

extern "C" int moip_ctx_set_ip_node_budget(moip_ctx* c, long long nodes) {
  if (!c) return MOIP_ERR_ARG;
  c->ip_node_budget = nodes > 0 ? nodes : 0;
  return MOIP_OK;
}

extern "C" int moip_ctx_set_kernel_timing(moip_ctx* c, int on) {
  if (!c) return MOIP_ERR_ARG;
  MOIP_CUDA(cudaSetDevice(c->device));
  if (on && !c->kev[0])
    for (auto& e : c->kev) MOIP_CUDA(cudaEventCreate(&e));
  c->ktiming = on != 0;
  c->kev_branch_pending = false;
  return MOIP_OK;
}
extern "C" int moip_ctx_kernel_times(const moip_ctx* c, moip_kernel_times* out) {
  if (!c || !out) return MOIP_ERR_ARG;
  *out = c->ktimes;
  return MOIP_OK;
}

extern "C" void moip_ctx_destroy(moip_ctx* c) {
  if (!c) return;
  if (c->prof_t[3] > 0)
    std::fprintf(stderr, "moip_b200: %.0f B&B rounds (%.1f nodes each), node-LP cap %d (%.0f %% of the LPs hit it): enqueue %.1f us, wait for the device %.1f us, host %.1f us per round\n",
                 c->prof_t[3], c->prof_t[4] / c->prof_t[3], c->bb_max_iter > 0 ? c->bb_max_iter : c->lp_cap_dyn, 100.0 * c->prof_t[6] / std::max(1.0, c->prof_t[5]), 1e6 * c->prof_t[0] / c->prof_t[3], 1e6 * c->prof_t[1] / c->prof_t[3],
                 1e6 * c->prof_t[2] / c->prof_t[3]);
  if (c->prof_t[3] > 0)
    for (int sg = 0; sg <= MOIP_MAX_OBJ; ++sg)
      if (c->stage_ips[sg])
        std::fprintf(stderr, "moip_b200:   stage %d%s: %lld IPs, %.1f nodes, %.1f node LPs, %.2f rounds per IP; %.0f %% of them done after the root round; %lld started from a stored point\n",
                     sg, sg == MOIP_MAX_OBJ ? " (get_limit / mip_solve)" : "", c->stage_ips[sg], (double)c->stage_nodes[sg] / c->stage_ips[sg],
                     (double)c->stage_lps[sg] / c->stage_ips[sg], (double)c->stage_rounds[sg] / c->stage_ips[sg],
                     100.0 * c->stage_root_solved[sg] / c->stage_ips[sg], sg == 0 ? c->start_hits : 0LL);
  if (c->chain_ips > 0 && std::getenv("MOIP_CHAIN_STATS")) {
    std::fprintf(stderr, "moip_b200: chained rounds: %lld IPs, %.2f host waits and %.2f idle rounds per IP, %lld handed back to the host loop; node-LP cap %d\n",
                 c->chain_ips, (double)c->chain_chunks / c->chain_ips, (double)c->chain_idle_rounds / c->chain_ips, c->chain_fallbacks,
                 c->bb_max_iter > 0 ? c->bb_max_iter : c->lp_cap_dyn);
    for (int sg = 0; sg <= MOIP_MAX_OBJ; ++sg)
      if (c->stage_ips[sg])
        std::fprintf(stderr, "moip_b200:   stage %d: %lld IPs, %.1f nodes, %.1f node LPs, %.2f rounds per IP; %.0f %% done after the root round\n",
                     sg, c->stage_ips[sg], (double)c->stage_nodes[sg] / c->stage_ips[sg], (double)c->stage_lps[sg] / c->stage_ips[sg],
                     (double)c->stage_rounds[sg] / c->stage_ips[sg], 100.0 * c->stage_root_solved[sg] / c->stage_ips[sg]);
  }
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto& e : c->kev) if (e) cudaEventDestroy(e);
  if (c->k3_answer) cudaFreeHost(c->k3_answer);
  if (c->sync_ev) cudaEventDestroy(c->sync_ev);
  for (void* p : c->model_allocs) cudaFree(p);
  c->b_cost.release(); c->b_lb.release(); c->b_ub.release(); c->b_status.release(); c->b_iters.release();
  c->b_branch.release(); c->b_counter.release(); c->b_rhs.release(); c->b_pobj.release(); c->b_dbound.release();
  c->b_x.release(); c->b_cutoff.release(); c->b_masks.release();
  c->q_ip.release(); c->q_out.release(); c->q_which.release(); c->h_q.release();
  c->v_x.release(); c->v_rhs.release(); c->v_obj.release(); c->v_feas.release();
  c->p_lb.release(); c->p_ub.release(); c->p_wx.release(); c->p_wy.release();
  c->r_xr.release(); c->r_counter.release(); c->r_pobj.release();
  c->r_ops.release(); c->h_round.release(); c->r_in.release(); c->r_out.release(); c->h_in.release(); c->h_ops.release(); c->d_inc.release(); c->d_root_x.release(); c->d_root_y.release(); c->k1_scratch.release();
  c->d_ctl.release(); c->h_ctl.release(); c->h_inc.release(); c->c_bound.release(); c->c_depth.release();
  if (c->owns_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

// Context on a non-blocking stream of its own: what a worker thread that holds no CUDA handles needs (the link-level
// CPLEX seam: one context per CPXLPptr, reference src/aira.cpp:561-585).
extern "C" int moip_ctx_create_own_stream(moip_model* m, int device, moip_ctx** out) {
  if (!m || !out) return MOIP_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    std::fprintf(stderr, "moip_b200: no CUDA device available (this library has no CPU fallback)\n");
    return MOIP_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) return MOIP_ERR_ARG;
  cudaStream_t st = nullptr;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
    std::fprintf(stderr, "moip_b200: cannot create a stream on device %d\n", device);
    return MOIP_ERR_CUDA;
  }
  int rc = moip_ctx_create(m, device, st, out);
  if (rc) { cudaStreamDestroy(st); return rc; }
  (*out)->owns_stream = true;
  return MOIP_OK;
}

extern "C" int moip_device_count(void) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) return 0;
  return ndev;
}

extern "C" int moip_ctx_stats(const moip_ctx* c, moip_stats* out) {
  if (!c || !out) return MOIP_ERR_ARG;
  *out = c->stats;
  return MOIP_OK;
}
extern "C" int moip_ctx_reset_stats(moip_ctx* c) {
  if (!c) return MOIP_ERR_ARG;
  c->stats = moip_stats{};
  return MOIP_OK;
}

// ------------------------------------------------------------------------------------ K1 batch API
extern "C" void moip_lp_default_params(moip_lp_params* p) {
  if (!p) return;
  p->eps = 1e-8;
  p->max_iter = 100000;
  p->check_every = 32;
  p->fixed_iters = 0;
  p->cutoff = MOIP_INFBOUND;
}

extern "C" int moip_lp_batch_upload(moip_ctx* c, int B, const int* cost_idx, const double* rhs, const uint32_t* fix_masks) {
  if (!c || B < 0 || (B > 0 && (!cost_idx || !rhs))) return MOIP_ERR_ARG;
  MOIP_CUDA(cudaSetDevice(c->device));
  const DevModel& d = c->dm;
  const int words = (d.n + 15) / 16;
  for (int b = 0; b < B; ++b)
    if (cost_idx[b] < 0 || cost_idx[b] >= d.k) return MOIP_ERR_ARG;
  int rc = 0;
  rc |= c->b_cost.ensure(B); rc |= c->b_rhs.ensure((size_t)B * d.k); rc |= c->b_masks.ensure((size_t)B * words);
  rc |= c->b_lb.ensure((size_t)B * d.n); rc |= c->b_ub.ensure((size_t)B * d.n);
  rc |= c->b_status.ensure(B); rc |= c->b_iters.ensure(B); rc |= c->b_branch.ensure((size_t)3 * B);
  rc |= c->b_pobj.ensure(B); rc |= c->b_dbound.ensure(B); rc |= c->b_x.ensure((size_t)B * d.n);
  rc |= c->b_counter.ensure(1); rc |= c->b_cutoff.ensure(1);
  if (rc) return MOIP_ERR_CUDA;
  c->batch_B = B;
  if (B == 0) return MOIP_OK;
  MOIP_CUDA(cudaMemcpyAsync(c->b_cost.p, cost_idx, sizeof(int) * B, cudaMemcpyHostToDevice, c->stream));
  MOIP_CUDA(cudaMemcpyAsync(c->b_rhs.p, rhs, sizeof(double) * B * d.k, cudaMemcpyHostToDevice, c->stream));
  if (fix_masks)
    MOIP_CUDA(cudaMemcpyAsync(c->b_masks.p, fix_masks, sizeof(uint32_t) * B * words, cudaMemcpyHostToDevice, c->stream));
  rc = launch_expand_masks(d, B, fix_masks ? c->b_masks.p : nullptr, words, c->b_lb.p, c->b_ub.p, c->stream);
  c->stats.kernel_launches += 1;
  return rc;
}

extern "C" int moip_lp_batch_run(moip_ctx* c, const moip_lp_params* params) {
  if (!c) return MOIP_ERR_ARG;
  MOIP_CUDA(cudaSetDevice(c->device));
  if (c->batch_B == 0) return MOIP_OK;
  moip_lp_params dp;
  moip_lp_default_params(&dp);
  if (params) dp = *params;
  const DevModel& d = c->dm;
  // public cutoff is in the model's sense; kernels use the min-form
  double cut = std::fabs(dp.cutoff) >= 1e19 ? HUGE_VAL : d.sgn * dp.cutoff;
  MOIP_CUDA(cudaMemcpyAsync(c->b_cutoff.p, &cut, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  LpBatch b{};
  b.B = c->batch_B;
  b.cost_idx = c->b_cost.p; b.rhs = c->b_rhs.p; b.lb = c->b_lb.p; b.ub = c->b_ub.p;
  b.warm_x = nullptr; b.warm_y = nullptr; b.out_x = c->b_x.p; b.out_y = nullptr;
  b.primal_obj = c->b_pobj.p; b.dual_bound = c->b_dbound.p; b.status = c->b_status.p; b.iters = c->b_iters.p;
  b.branch_var = c->b_branch.p; b.branch_val = nullptr; b.skip = nullptr; b.slot = nullptr; b.rc_fix = 0;
  b.farkas = c->batch_farkas ? 1 : 0;
  b.cost_stride = 1; b.rhs_stride = d.k;
  b.cutoff = c->b_cutoff.p; b.work_counter = c->b_counter.p;
  LpParams p{};
  p.eps = dp.eps; p.max_iter = dp.max_iter; p.check_every = dp.check_every > 0 ? dp.check_every : 32;
  p.fixed_iters = dp.fixed_iters; p.norm_every = c->norm_every > 0 ? c->norm_every : 1;
  p.cutoff_slack = 0.0; p.int_obj = 0;
  c->stats.kernel_launches += 1;
  c->stats.node_lps += b.B;
  if (c->attach_k1_scratch(b)) return MOIP_ERR_CUDA;
  return launch_k1_any(d, b, p, c->num_sms, c->stream);
}

extern "C" int moip_lp_batch_download(moip_ctx* c, double* primal_obj, double* dual_bound, int* status, int* iters,
                                      double* x_out) {
  if (!c) return MOIP_ERR_ARG;
  MOIP_CUDA(cudaSetDevice(c->device));
  const int B = c->batch_B;
  const DevModel& d = c->dm;
  if (B > 0) {
    if (primal_obj) MOIP_CUDA(cudaMemcpyAsync(primal_obj, c->b_pobj.p, sizeof(double) * B, cudaMemcpyDeviceToHost, c->stream));
    if (dual_bound) MOIP_CUDA(cudaMemcpyAsync(dual_bound, c->b_dbound.p, sizeof(double) * B, cudaMemcpyDeviceToHost, c->stream));
    if (status) MOIP_CUDA(cudaMemcpyAsync(status, c->b_status.p, sizeof(int) * B, cudaMemcpyDeviceToHost, c->stream));
    if (iters) MOIP_CUDA(cudaMemcpyAsync(iters, c->b_iters.p, sizeof(int) * B, cudaMemcpyDeviceToHost, c->stream));
    if (x_out) MOIP_CUDA(cudaMemcpyAsync(x_out, c->b_x.p, sizeof(double) * B * d.n, cudaMemcpyDeviceToHost, c->stream));
  }
  MOIP_CUDA(cudaStreamSynchronize(c->stream));
  for (int b = 0; b < B; ++b) {
    if (primal_obj) primal_obj[b] *= d.sgn;
    if (dual_bound) dual_bound[b] *= d.sgn;
    if (iters) c->stats.lp_iterations += iters[b];
  }
  return MOIP_OK;
}

extern "C" int moip_lp_batch_solve(moip_ctx* c, int B, const int* cost_idx, const double* rhs, const uint32_t* fix_masks,
                                   const moip_lp_params* params, double* primal_obj, double* dual_bound, int* status,
                                   int* iters, double* x_out) {
  int rc = moip_lp_batch_upload(c, B, cost_idx, rhs, fix_masks);
  if (rc) return rc;
  rc = moip_lp_batch_run(c, params);
  if (rc) return rc;
  return moip_lp_batch_download(c, primal_obj, dual_bound, status, iters, x_out);
}

// ------------------------------------------------------------------------------------ K3 cache
int moip_cache::sync_to_device(cudaStream_t st) {
  if (synced == host.size()) return MOIP_OK;
  MOIP_CUDA(cudaSetDevice(device));
  if (host.size() > dev.cap) {
    // grow: a new array, refilled from the host mirror.  The old one is only retired -- a scan another worker launched
    // from its snapshot may still be reading it -- and freed with the store.
    size_t nc = dev.cap ? dev.cap * 2 : 1024;
    while (nc < host.size()) nc *= 2;
    CacheRecord* q = nullptr;
    MOIP_CUDA(cudaMalloc(&q, nc * sizeof(CacheRecord)));
    if (dev.p) retired.push_back(dev.p);
    dev.p = q; dev.cap = nc;
    synced = 0;
  }
  MOIP_CUDA(cudaMemcpyAsync(dev.p + synced, host.data() + synced, sizeof(CacheRecord) * (host.size() - synced),
                            cudaMemcpyHostToDevice, st));
  // the records must have landed before `synced` is published: scans of OTHER workers run on other streams
  MOIP_CUDA(cudaStreamSynchronize(st));
  synced = host.size();
  return MOIP_OK;
}
DevCache moip_cache::view() const {
  DevCache v;
  v.k = k; v.size = (int)synced; v.rec = dev.p;
  return v;
}

extern "C" int moip_cache_create(moip_ctx* c, moip_cache** out) {
  if (!c || !out) return MOIP_ERR_ARG;
  moip_cache* s = new moip_cache();
  s->ctx = c; s->k = c->dm.k; s->device = c->device;
  *out = s;
  return MOIP_OK;
}
extern "C" void moip_cache_destroy(moip_cache* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  s->dev.release();
  for (CacheRecord* q : s->retired) cudaFree(q);
  delete s;
}
extern "C" int moip_cache_insert(moip_cache* s, const double* ip, const int* result, int infeasible) {
  if (!s || !ip || (!infeasible && !result)) return MOIP_ERR_ARG;
  CacheRecord r{};
  for (int i = 0; i < s->k; ++i) { r.ip[i] = ip[i]; r.result[i] = (infeasible || !result) ? 0 : result[i]; }
  r.infeasible = infeasible ? 1 : 0;
  std::lock_guard<std::mutex> lk(s->mu);
  s->host.push_back(r);
  return MOIP_OK;
}
extern "C" int moip_cache_size(const moip_cache* s) {
  if (!s) return -1;
  std::lock_guard<std::mutex> lk(const_cast<moip_cache*>(s)->mu);
  return (int)s->host.size();
}

// shared by the public batch call and the generator (two stores, one launch, one sync).  A store's lock is held only
// while its new records go to the device and the {array, size} snapshot is taken -- not across the scan: the stores of
// an EPP level are shared by every worker of the pool (src/aira.cpp:1918-1933), and a 4-objective knapsack front asks
// them 10x more often than it solves.  Records are append-only while workers run, so a snapshot stays valid.
int cache_find2(moip_ctx* c, moip_cache* s0, moip_cache* s1, int Q, const double* ip, int sense, int* first_match, int* which,
                CacheRecord* rec_out) {
  MOIP_CUDA(cudaSetDevice(c->device));
  const int k = c->dm.k;
  c->dbg_where.store(1, std::memory_order_relaxed);
  struct Leave { moip_ctx* c; ~Leave() { c->dbg_where.store(0, std::memory_order_relaxed); } } leave{c};
  DevCache e{}; e.k = k; e.size = 0; e.rec = nullptr;
  DevCache v0 = e, v1 = e;
  if (s0) {
    std::lock_guard<std::mutex> lk(s0->mu);
    if (s0->sync_to_device(c->stream)) return MOIP_ERR_CUDA;
    v0 = s0->view();
  }
  if (s1 && s1 != s0) {
    std::lock_guard<std::mutex> lk(s1->mu);
    if (s1->sync_to_device(c->stream)) return MOIP_ERR_CUDA;
    v1 = s1->view();
  }
  if (Q == 1 && rec_out && c->k3_poll) {
    // the generator's scan: query by kernel parameter, answer into mapped pinned memory, host polls the sequence number
    if (!c->k3_answer) {
      MOIP_CUDA(cudaHostAlloc((void**)&c->k3_answer, sizeof(K3Answer), cudaHostAllocMapped));
      MOIP_CUDA(cudaHostGetDevicePointer((void**)&c->k3_answer_dev, c->k3_answer, 0));
      std::memset(c->k3_answer, 0, sizeof(K3Answer));
    }
    K3Query qq{};
    for (int i = 0; i < k; ++i) qq.ip[i] = ip[i];
    const int seq = ++c->k3_seq;
    c->kmark(8);
    int rc = launch_k3_one(v0, v1, qq, sense, c->k3_answer_dev, seq, c->stream);
    if (rc) return rc;
    c->kmark(9);
    c->stats.kernel_launches += 1;
    c->stats.cache_queries += 1;
    volatile int* vseq = &c->k3_answer->seq;
    for (long spins = 0; *vseq != seq; ++spins) {
      if (c->block_sync && (spins & 0x3f) == 0x3f) std::this_thread::yield();   // more workers than cores: do not hog one
      if ((spins & 0xfff) == 0xfff) {                    // a failed launch would never publish: ask the stream now and then
        const cudaError_t e = cudaStreamQuery(c->stream);
        if (e != cudaSuccess && e != cudaErrorNotReady) { MOIP_CUDA(e); }
        if (e == cudaSuccess && *vseq != seq) { MOIP_CUDA(cudaStreamSynchronize(c->stream)); break; }
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (c->ktiming) { cudaEventSynchronize(c->kev[9]); c->ktimes.k3_ms += c->kspan(8, 9); c->ktimes.scans += 1; }
    first_match[0] = c->k3_answer->first_match;
    if (which) which[0] = c->k3_answer->which;
    if (first_match[0] >= 0) rec_out[0] = c->k3_answer->rec;
    return MOIP_OK;
  }
  if (c->q_ip.ensure((size_t)Q * k) || c->q_out.ensure(Q) || c->q_which.ensure(Q) || c->h_q.ensure((size_t)2 * Q)) return MOIP_ERR_CUDA;
  c->kmark(8);
  MOIP_CUDA(cudaMemcpyAsync(c->q_ip.p, ip, sizeof(double) * Q * k, cudaMemcpyHostToDevice, c->stream));
  int rc = launch_k3(v0, v1, Q, c->q_ip.p, sense, c->q_out.p, c->q_which.p, c->stream);
  if (rc) return rc;
  c->stats.kernel_launches += 1;
  c->stats.cache_queries += Q;
  MOIP_CUDA(cudaMemcpyAsync(c->h_q.p, c->q_out.p, sizeof(int) * Q, cudaMemcpyDeviceToHost, c->stream));
  MOIP_CUDA(cudaMemcpyAsync(c->h_q.p + Q, c->q_which.p, sizeof(int) * Q, cudaMemcpyDeviceToHost, c->stream));
  c->kmark(9);
  MOIP_CUDA(cudaStreamSynchronize(c->stream));
  if (c->ktiming) { c->ktimes.k3_ms += c->kspan(8, 9); c->ktimes.scans += 1; }
  std::memcpy(first_match, c->h_q.p, sizeof(int) * Q);
  if (which) std::memcpy(which, c->h_q.p + Q, sizeof(int) * Q);
  if (rec_out)
    for (int q = 0; q < Q; ++q)
      if (first_match[q] >= 0) {
        moip_cache* s = c->h_q.p[Q + q] == 0 ? s0 : s1;
        std::lock_guard<std::mutex> lk(s->mu);
        rec_out[q] = s->host[first_match[q]];
      }
  return MOIP_OK;
}

extern "C" int moip_cache_find_batch(moip_cache* s, int Q, const double* ip, int sense, int* first_match) {
  if (!s || Q < 0 || (Q > 0 && (!ip || !first_match))) return MOIP_ERR_ARG;
  if (Q == 0) return MOIP_OK;
  return cache_find2(s->ctx, s, nullptr, Q, ip, sense, first_match, nullptr, nullptr);
}
extern "C" int moip_cache_get(const moip_cache* s, int i, double* ip, int* result, int* infeasible) {
  if (!s) return MOIP_ERR_ARG;
  std::lock_guard<std::mutex> lk(const_cast<moip_cache*>(s)->mu);
  if (i < 0 || i >= (int)s->host.size()) return MOIP_ERR_ARG;
  const CacheRecord& r = s->host[i];
  for (int j = 0; j < s->k; ++j) { if (ip) ip[j] = r.ip[j]; if (result) result[j] = r.result[j]; }
  if (infeasible) *infeasible = r.infeasible;
  return MOIP_OK;
}
extern "C" int moip_cache_merge(moip_cache* s, moip_cache* other) {
  if (!s || !other || s == other || s->k != other->k) return MOIP_ERR_ARG;
  std::scoped_lock lk(s->mu, other->mu);
  std::vector<CacheRecord> merged;
  merged.reserve(s->host.size() + other->host.size());
  merged.insert(merged.end(), other->host.begin(), other->host.end());   // splice at begin()
  merged.insert(merged.end(), s->host.begin(), s->host.end());
  s->host.swap(merged);
  s->synced = 0;
  other->host.clear();
  other->synced = 0;
  return MOIP_OK;
}
extern "C" int moip_cache_sort_unique(moip_cache* s, int* rows, int cap) {
  if (!s) return -1;
  std::lock_guard<std::mutex> lk(s->mu);
  const int k = s->k;
  // Result::operator< : infeasible first, then descending lexicographic (reference src/result.cpp:9-29)
  std::stable_sort(s->host.begin(), s->host.end(), [k](const CacheRecord& a, const CacheRecord& b) {
    if (a.infeasible != b.infeasible) return a.infeasible > b.infeasible;
    if (a.infeasible) return false;
    for (int i = 0; i < k; ++i) if (a.result[i] != b.result[i]) return a.result[i] > b.result[i];
    return false;
  });
  auto same = [k](const CacheRecord& a, const CacheRecord& b) {
    if (a.infeasible != b.infeasible) return false;
    if (a.infeasible) return true;
    for (int i = 0; i < k; ++i) if (a.result[i] != b.result[i]) return false;
    return true;
  };
  s->host.erase(std::unique(s->host.begin(), s->host.end(), same), s->host.end());
  s->synced = 0;
  int nrows = 0;
  for (auto& r : s->host) {
    if (r.infeasible) continue;
    if (rows && nrows < cap) for (int i = 0; i < k; ++i) rows[(size_t)nrows * k + i] = r.result[i];
    ++nrows;
  }
  return nrows;
}

// ------------------------------------------------------------------------------------ K4
extern "C" int moip_verify_int64(moip_ctx* c, int B, const int32_t* x, const double* rhs, int64_t* obj_out,
                                 uint8_t* feasible_out) {
  if (!c || B < 0 || (B > 0 && (!x || !obj_out || !feasible_out))) return MOIP_ERR_ARG;
  if (B == 0) return MOIP_OK;
  MOIP_CUDA(cudaSetDevice(c->device));
  const DevModel& d = c->dm;
  if (c->v_x.ensure((size_t)B * d.n) || c->v_obj.ensure((size_t)B * d.k) || c->v_feas.ensure(B) ||
      c->v_rhs.ensure((size_t)B * d.k)) return MOIP_ERR_CUDA;
  MOIP_CUDA(cudaMemcpyAsync(c->v_x.p, x, sizeof(int) * (size_t)B * d.n, cudaMemcpyHostToDevice, c->stream));
  if (rhs) MOIP_CUDA(cudaMemcpyAsync(c->v_rhs.p, rhs, sizeof(double) * (size_t)B * d.k, cudaMemcpyHostToDevice, c->stream));
  int rc = launch_k4(d, B, c->v_x.p, rhs ? c->v_rhs.p : nullptr, c->v_obj.p, c->v_feas.p, c->stream);
  if (rc) return rc;
  c->stats.kernel_launches += 1;
  static_assert(sizeof(long long) == sizeof(int64_t), "int64");
  MOIP_CUDA(cudaMemcpyAsync(obj_out, c->v_obj.p, sizeof(int64_t) * (size_t)B * d.k, cudaMemcpyDeviceToHost, c->stream));
  MOIP_CUDA(cudaMemcpyAsync(feasible_out, c->v_feas.p, (size_t)B, cudaMemcpyDeviceToHost, c->stream));
  MOIP_CUDA(cudaStreamSynchronize(c->stream));
  return MOIP_OK;
}

// ------------------------------------------------------------------------------------ B&B
// large models that run on the generic kernel keep the iterate of every resident CTA in HBM (streaming mode)
int moip_ctx::attach_k1_scratch(moip::LpBatch& b) {
  b.scratch = nullptr; b.scratch_stride = 0; b.scratch_slots = 0;
  if (dm.reg_ok || dm.fast_ok) return MOIP_OK;
  size_t stride = k1_scratch_stride(dm);
  if (stride == 0 && std::getenv("MOIP_K1_FORCE_STREAMING")) stride = ((size_t)5 * dm.n + (size_t)8 * dm.m + 15) & ~(size_t)15;   // tests
  if (stride == 0) return MOIP_OK;
  const int slots = num_sms * 4;
  if (k1_scratch.ensure(stride * (size_t)slots)) return MOIP_ERR_CUDA;
  b.scratch = k1_scratch.p; b.scratch_stride = stride; b.scratch_slots = slots;
  return MOIP_OK;
}

int moip_ctx::ensure_pool(int slots) {
  if (slots <= pool_slots) return MOIP_OK;
  int ns = pool_slots ? pool_slots : 256;
  while (ns < slots) ns *= 2;
  if (p_lb.ensure((size_t)ns * dm.n, true, stream) || p_ub.ensure((size_t)ns * dm.n, true, stream) ||
      p_wx.ensure((size_t)ns * dm.n, true, stream) || p_wy.ensure((size_t)ns * dm.m, true, stream)) return MOIP_ERR_CUDA;
  for (int s = ns - 1; s >= pool_slots; --s) free_slots.push_back(s);
  pool_slots = ns;
  return MOIP_OK;
}

int moip_ctx::alloc_slot() {
  if (free_slots.empty()) {
    if (ensure_pool(pool_slots ? pool_slots * 2 : 256)) return -1;
  }
  int s = free_slots.back();
  free_slots.pop_back();
  return s;
}

// (The thread-sanitizer build of the HOST code, tests/tsan, compiles this file with MOIP_HOST_DOUBLE and supplies an
// enumeration solve_ip of its own: everything else in this file -- contexts, caches, lexicographic chain -- is the code under test.)
#ifndef MOIP_HOST_DOUBLE
namespace {
struct OpenNode {
  int slot;
  double bound;   // min-form valid lower bound inherited from the parent
  int depth;
  int pref;       // how many branched columns went to the side nearer to the parent's LP value
};
}  // namespace

// One IP with its rounds chained on the device (bbchain.h).  The host enqueues a chunk of rounds -- node LPs with the fused
// propagation / rounding (k1_reg.cuh) followed by K5 (k5_chain.cu) -- copies the control block back and waits once per
// chunk; the first chunk is as long as this stage's IPs usually need.  Same exactness model as the host loop below: int64
// propagation and verification, pruning on valid Lagrangian bounds only.
int moip_ctx::solve_ip_chained(int cost, const double* srhs, const long long* olo, const long long* ohi, bool& have_inc,
                               long long& inc_val, std::vector<int>& inc_x, bool& handled) {
  handled = false;
  const Model& M = model->M;
  const int n = dm.n, k = dm.k, m = dm.m;
  const int Q = chain_q;
  int rc = 0;
  if (ensure_pool(2 * Q)) return MOIP_ERR_CUDA;
  auto up8 = [](size_t v) { return (v + 7) & ~(size_t)7; };
  struct OutLayout { size_t flag, status, iters, branch, ff, dbound, bval, leaf, cobj, cfeas, end; } LO;
  LO.flag = 0; LO.status = up8(LO.flag + sizeof(int) * Q); LO.iters = up8(LO.status + sizeof(int) * Q);
  LO.branch = up8(LO.iters + sizeof(int) * Q); LO.ff = up8(LO.branch + sizeof(int) * 3 * Q);
  LO.dbound = up8(LO.ff + sizeof(int) * 3 * Q);
  LO.bval = LO.dbound + sizeof(double) * Q; LO.leaf = LO.bval + sizeof(double) * 3 * Q;
  LO.cobj = LO.leaf + sizeof(long long) * Q * k; LO.cfeas = LO.cobj + sizeof(long long) * Q * 3 * k;
  LO.end = up8(LO.cfeas + (size_t)Q * 3);
  rc |= r_out.ensure(LO.end); rc |= r_xr.ensure((size_t)Q * 3 * n); rc |= r_pobj.ensure(Q);
  rc |= d_inc.ensure(n); rc |= d_root_x.ensure((size_t)k * n); rc |= d_root_y.ensure((size_t)k * m);
  rc |= d_ctl.ensure(1); rc |= h_ctl.ensure(1); rc |= h_inc.ensure(n); rc |= c_bound.ensure((size_t)2 * Q); rc |= c_depth.ensure((size_t)2 * Q);
  if (rc) return MOIP_ERR_CUDA;
  const PoolView pool{p_lb.p, p_ub.p, p_wx.p, p_wy.p};
  BbCtl* ctl = d_ctl.p;
  const int Bmax = bb_batch > 0 ? bb_batch : num_sms * (dm.n <= 64 ? 16 : 4);
  const bool fused = dm.reg_ok;          // register-resident K1: K2 and K4 run inside its CTAs; else (8-lanes-per-node K1) as kernels of their own

  const bool cold_root = !root_valid[cost];
  if (lp_cap_dyn <= 0) lp_cap_dyn = std::min(2000, std::max(200, (int)(20.0 * std::sqrt((double)n))));
  const int lp_cap = bb_max_iter > 0 ? bb_max_iter : lp_cap_dyn;
  LpParams lp{};
  lp.eps = bb_eps; lp.check_every = bb_check; lp.fixed_iters = 0;
  lp.norm_every = norm_every > 0 ? norm_every : 1; lp.cutoff_slack = 1.0 - 1e-6; lp.int_obj = 1;

  BbInit init{};
  for (int o = 0; o < MOIP_MAX_OBJ; ++o) {
    init.olo[o] = o < k ? olo[o] : LLONG_MIN; init.ohi[o] = o < k ? ohi[o] : LLONG_MAX; init.rhs[o] = o < k ? srhs[o] : 0.0;
  }
  const long long inc0 = have_inc ? inc_val : LLONG_MAX;
  init.inc_val = inc0; init.cost = cost; init.sense = M.sense; init.qcap = Q; init.bmax = Bmax;
  init.levels_max = std::max(1, bb_levels); init.warm = cold_root ? 0 : 1;
  init.root_x = d_root_x.p + (size_t)cost * n; init.root_y = d_root_y.p + (size_t)cost * m;
  if (launch_bb_init(dm, pool, ctl, init, c_bound.p, c_depth.p, stream)) return MOIP_ERR_CUDA;

  BbRound R{};
  R.cold_root = cold_root ? 1 : 0;
  R.flag = reinterpret_cast<const int*>(r_out.p + LO.flag); R.status = reinterpret_cast<const int*>(r_out.p + LO.status);
  R.iters = reinterpret_cast<const int*>(r_out.p + LO.iters); R.branch = reinterpret_cast<const int*>(r_out.p + LO.branch);
  R.bval = reinterpret_cast<const double*>(r_out.p + LO.bval); R.ff = reinterpret_cast<const int*>(r_out.p + LO.ff);
  R.dbound = reinterpret_cast<const double*>(r_out.p + LO.dbound); R.leaf = reinterpret_cast<const long long*>(r_out.p + LO.leaf);
  R.cobj = reinterpret_cast<const long long*>(r_out.p + LO.cobj); R.cfeas = r_out.p + LO.cfeas; R.xr = r_xr.p;
  R.node_bound = c_bound.p; R.node_depth = c_depth.p; R.inc_x = d_inc.p;
  R.root_x = d_root_x.p + (size_t)cost * n; R.root_y = d_root_y.p + (size_t)cost * m;

  LpBatch b{};
  b.rhs = ctl->rhs; b.lb = pool.lb; b.ub = pool.ub; b.slot = nullptr; b.rc_fix = 1;
  b.fused = fused ? 1 : 0; b.f_obj_lo = ctl->plo; b.f_obj_hi = ctl->phi; b.f_max_rounds = 16;
  b.f_flag = reinterpret_cast<int*>(r_out.p + LO.flag); b.f_leaf_obj = reinterpret_cast<long long*>(r_out.p + LO.leaf);
  b.f_xr = r_xr.p; b.f_cand_obj = reinterpret_cast<long long*>(r_out.p + LO.cobj); b.f_cand_feas = r_out.p + LO.cfeas;
  b.f_first_free = reinterpret_cast<int*>(r_out.p + LO.ff);
  b.warm_x = pool.wx; b.warm_y = pool.wy; b.out_x = pool.wx; b.out_y = pool.wy;
  b.primal_obj = r_pobj.p; b.dual_bound = reinterpret_cast<double*>(r_out.p + LO.dbound);
  b.status = reinterpret_cast<int*>(r_out.p + LO.status); b.iters = reinterpret_cast<int*>(r_out.p + LO.iters);
  b.branch_var = reinterpret_cast<int*>(r_out.p + LO.branch); b.branch_val = reinterpret_cast<double*>(r_out.p + LO.bval);
  b.skip = fused ? nullptr : b.f_flag; b.cost_stride = 0; b.rhs_stride = 0; b.cost_idx = &ctl->cost;
  b.cutoff = &ctl->cutoff; b.f_cutoff_rw = &ctl->cutoff; b.work_counter = &ctl->work_counter;
  b.f_inc = &ctl->inc_val; b.f_lim_lo = ctl->olo; b.f_lim_hi = ctl->ohi;
  ChainRef ch;
  ch.inc = &ctl->inc_val; ch.cutoff = &ctl->cutoff; ch.cost = &ctl->cost; ch.lim_lo = ctl->olo; ch.lim_hi = ctl->ohi;

  dbg_where.store(2, std::memory_order_relaxed);
  const int stage = cur_stage;
  int chunk = std::min(16, std::max(2, (int)std::ceil(chain_rounds_avg[stage] + 0.5)));
  int r = 0;
  bool over_budget = false;
  const BbCtl* H = h_ctl.p;
  for (;;) {
    for (int c = 0; c < chunk; ++c, ++r) {
      const int p = r & 1;
      const int hint = r < 4 ? std::min(Q, 1 << (3 * r)) : Q;          // a level is at most 8 times the one before
      b.B = hint; b.B_dev = &ctl->count[p]; b.slot_base = p * Q;
      lp.max_iter = (r == 0 && cold_root) ? std::max(lp_cap, 20000) : lp_cap;
      ch.B_dev = b.B_dev; ch.slot_base = b.slot_base;
      if (!fused && launch_k2_propagate(dm, pool, hint, nullptr, ctl->plo, ctl->phi, 16, b.f_flag, b.f_leaf_obj, stream, ch)) return MOIP_ERR_CUDA;
      if (launch_k1_any(dm, b, lp, num_sms, stream)) return MOIP_ERR_CUDA;
      if (!fused && launch_k4_round(dm, hint, nullptr, pool.wx, pool.lb, pool.ub, r_xr.p, b.f_cand_obj, b.f_cand_feas, b.f_first_free,
                                    b.f_flag, stream, ch)) return MOIP_ERR_CUDA;
      if (chain_debug) {
        const cudaError_t e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) { std::fprintf(stderr, "moip_b200: chained round %d: K1 failed: %s (n=%d cost=%d cold=%d cap=%d)\n", r, cudaGetErrorString(e), n, cost, (int)cold_root, lp.max_iter); return MOIP_ERR_CUDA; }
      }
      R.parity = p; R.round = r;
      if (launch_bb_advance(dm, pool, ctl, R, hint, stream)) return MOIP_ERR_CUDA;
      if (chain_debug) {
        const cudaError_t e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) { std::fprintf(stderr, "moip_b200: chained round %d: K5 failed: %s (n=%d)\n", r, cudaGetErrorString(e), n); return MOIP_ERR_CUDA; }
      }
    }
    stats.kernel_launches += (fused ? 2 : 4) * chunk;
    chain_chunks += 1;
    MOIP_CUDA(cudaMemcpyAsync(h_ctl.p, ctl, sizeof(BbCtl), cudaMemcpyDeviceToHost, stream));
    MOIP_CUDA(cudaMemcpyAsync(h_inc.p, d_inc.p, sizeof(int) * n, cudaMemcpyDeviceToHost, stream));
    if (wait_stream()) return MOIP_ERR_CUDA;
    dbg_rounds.store(r, std::memory_order_relaxed); dbg_open.store(H->count[r & 1], std::memory_order_relaxed);
    if (H->overflow || H->count[r & 1] == 0) break;
    if (ip_node_budget > 0 && (long long)H->n_nodes > ip_node_budget) { over_budget = true; break; }
    chunk = 3;
  }
  // ---- what the device did
  stats.bb_nodes += (long long)H->n_nodes; stage_nodes[stage] += (long long)H->n_nodes;
  stats.node_lps += (long long)H->n_lps; stage_lps[stage] += (long long)H->n_lps;
  stats.lp_iterations += (long long)H->n_iters;
  stage_rounds[stage] += H->rounds_live;
  if (H->rounds_live <= 1) stage_root_solved[stage] += 1;
  chain_ips += 1; chain_idle_rounds += r - H->rounds_live;
  if (H->root_solved) root_valid[cost] = 1;
  cap_seen += (long long)H->n_solved; cap_hit += (long long)H->n_capped;
  prof_t[5] += (double)H->n_solved; prof_t[6] += (double)H->n_capped;
  if (bb_max_iter <= 0 && cap_seen >= 64) {       // cap controller (see the host loop): once per IP here
    const double share = (double)cap_hit / (double)cap_seen;
    if (share > bb_cap_hi) lp_cap_dyn = std::min(6000, lp_cap_dyn + lp_cap_dyn / 4 + 1);
    else if (share < bb_cap_lo) lp_cap_dyn = std::max(100, lp_cap_dyn - lp_cap_dyn / 8);
    cap_seen = 0; cap_hit = 0;
  }
  if (H->inc_val < inc0) {
    if (H->inc_seen != H->inc_val) {
      std::fprintf(stderr, "moip_b200: chained rounds lost the incumbent's point (value %lld, point of %lld)\n", H->inc_val, H->inc_seen);
      return MOIP_ERR_CUDA;
    }
    have_inc = true; inc_val = H->inc_val;
    inc_x.assign(h_inc.p, h_inc.p + n);
  }
  if (over_budget) {                     // drain the rounds still queued: the pool rows are reused by the next IP
    MOIP_CUDA(cudaMemsetAsync(ctl->count, 0, sizeof(int) * 2, stream));
    return MOIP_ERR_BUDGET;
  }
  if (H->overflow) { chain_fallbacks += 1; return MOIP_OK; }      // the caller's loop solves the IP from the root, with this incumbent
  chain_rounds_avg[stage] += 0.125 * ((double)H->rounds_live - chain_rounds_avg[stage]);
  handled = true;
  return MOIP_OK;
}

// One single-objective IP:  optimise objective `cost` s.t. the structural rows and C x (<=|>=) srhs.
// Exactness: pruning uses only valid Lagrangian bounds from K1 and int64 propagation from K2;
// incumbents are accepted only after exact int64 evaluation (K2 leaves / K4 rounded LP points).
int moip_ctx::solve_ip(int cost, const double* srhs, const std::vector<int>* inc_x_in, IpResult& out) {
  MOIP_CUDA(cudaSetDevice(device));
  const Model& M = model->M;
  const int n = dm.n, k = dm.k, m = dm.m;
  const double sgn = dm.sgn;
  stats.ip_solved += 1;
  stage_ips[cur_stage] += 1;
  long long ip_rounds = 0, ip_nodes = 0;
  static const long long explode_log = std::getenv("MOIP_EXPLODE_LOG") ? std::atoll(std::getenv("MOIP_EXPLODE_LOG")) : 0;
  bool explode_said = false;
  out.status = MOIP_MIP_INFEASIBLE;
  out.x.clear();
  if (M.int_infeasible) return MOIP_OK;
  // integer limits of the k objective rows (model space)
  std::vector<long long> olo(k, LLONG_MIN), ohi(k, LLONG_MAX);
  for (int o = 0; o < k; ++o) {
    if (std::fabs(srhs[o]) >= 1e19) continue;
    if (M.sense == 0) ohi[o] = (long long)std::floor(srhs[o] + 1e-9);
    else olo[o] = (long long)std::ceil(srhs[o] - 1e-9);
  }
  // incumbent (min-form value)
  bool have_inc = false;
  long long inc_val = LLONG_MAX;
  std::vector<int> inc_x;
  bool inc_on_device = false;
  if (inc_x_in && (int)inc_x_in->size() == n) {
    long long v = 0;
    for (int j = 0; j < n; ++j) v += M.ci[(size_t)cost * n + j] * (long long)(*inc_x_in)[j];
    have_inc = true; inc_val = (long long)sgn * v; inc_x = *inc_x_in;
  }
  if (use_points) {     // MIP start: the best point on record that satisfies the objective bounds (and beats the given start)
    long long v = 0;
    if (model->points.best(cost, (long long)sgn, olo.data(), ohi.data(), have_inc ? inc_val : LLONG_MAX, inc_x, v)) {
      have_inc = true; inc_val = v; start_hits += 1;
    }
  }
  // what every way out of this function does with the incumbent
  auto finish = [&](bool inc_on_device) -> int {
    if (!have_inc) return MOIP_OK;
    if (inc_on_device) {
      inc_x.resize(n);
      MOIP_CUDA(cudaMemcpyAsync(inc_x.data(), d_inc.p, sizeof(int) * n, cudaMemcpyDeviceToHost, stream));
      MOIP_CUDA(cudaStreamSynchronize(stream));
    }
    out.status = MOIP_MIP_OPTIMAL;
    out.obj = (long long)sgn * inc_val;
    out.x = inc_x;
    if (use_points) {
      long long ov[MOIP_MAX_OBJ] = {0, 0, 0, 0};
      for (int o = 0; o < k; ++o) {
        long long v = 0;
        const int64_t* co = M.ci.data() + (size_t)o * n;
        for (int j = 0; j < n; ++j) v += co[j] * (long long)inc_x[j];
        ov[o] = v;
      }
      model->points.add(inc_x.data(), ov);
    }
    return MOIP_OK;
  };
  // chained rounds: the tree advances on the device; the host loop below only takes over when a level outgrows the pool
  if (use_chain && ((use_fused && dm.reg_ok) || (!dm.reg_ok && k1_small_applies(dm))) && !ktiming && !std::getenv("MOIP_PROFILE_ROUNDS")) {
    bool handled = false;
    if (int rcc = solve_ip_chained(cost, srhs, olo.data(), ohi.data(), have_inc, inc_val, inc_x, handled)) return rcc;
    if (handled) { dbg_where.store(0, std::memory_order_relaxed); return finish(false); }
  }
  // batch geometry
  int Bmax = bb_batch > 0 ? bb_batch : num_sms * (dm.n <= 64 ? 16 : 4);
  if (ensure_pool(std::max(256, 4 * Bmax))) return MOIP_ERR_CUDA;
  int rc = 0;
  // One round = one H2D block in, one D2H block out (both pinned on the host).  The arrays of a round are laid
  // out back to back for the round's own batch size B, so that a single copy moves exactly what is needed.
  auto up8 = [](size_t v) { return (v + 7) & ~(size_t)7; };
  struct InLayout { size_t ids, cost, olo, ohi, cutoff, rhs, end; };
  struct OutLayout { size_t flag, status, iters, branch, ff, dbound, bval, leaf, cobj, cfeas, end; };
  auto in_layout = [&](int B) {
    InLayout L;
    L.ids = 0; L.cost = up8(sizeof(int) * B); L.olo = L.cost + 8; L.ohi = L.olo + sizeof(long long) * k;
    L.cutoff = L.ohi + sizeof(long long) * k; L.rhs = L.cutoff + 8; L.end = L.rhs + sizeof(double) * k;
    return L;
  };
  auto out_layout = [&](int B) {
    OutLayout L;
    L.flag = 0; L.status = up8(L.flag + sizeof(int) * B); L.iters = up8(L.status + sizeof(int) * B);
    L.branch = up8(L.iters + sizeof(int) * B); L.ff = up8(L.branch + sizeof(int) * 3 * B);
    L.dbound = up8(L.ff + sizeof(int) * 3 * B);
    L.bval = L.dbound + sizeof(double) * B; L.leaf = L.bval + sizeof(double) * 3 * B;
    L.cobj = L.leaf + sizeof(long long) * B * k; L.cfeas = L.cobj + sizeof(long long) * B * 3 * k;
    L.end = up8(L.cfeas + (size_t)B * 3);
    return L;
  };
  rc |= r_in.ensure(in_layout(Bmax).end); rc |= r_out.ensure(out_layout(Bmax).end);
  rc |= h_in.ensure(in_layout(Bmax).end); rc |= h_round.ensure(out_layout(Bmax).end);
  rc |= r_xr.ensure((size_t)Bmax * 3 * n); rc |= r_counter.ensure(1); rc |= r_pobj.ensure(Bmax);
  rc |= r_ops.ensure((size_t)8 * Bmax); rc |= h_ops.ensure((size_t)8 * Bmax);
  rc |= d_inc.ensure(n); rc |= d_root_x.ensure((size_t)k * n); rc |= d_root_y.ensure((size_t)k * m);
  if (rc) return MOIP_ERR_CUDA;
  PoolView pool{p_lb.p, p_ub.p, p_wx.p, p_wy.p};

  // root node
  std::vector<OpenNode> open;
  {
    int s = alloc_slot();
    if (s < 0) return MOIP_ERR_CUDA;
    pool = PoolView{p_lb.p, p_ub.p, p_wx.p, p_wy.p};
    MOIP_CUDA(cudaMemcpyAsync(pool.lb + (size_t)s * n, dm.lbI, sizeof(int) * n, cudaMemcpyDeviceToDevice, stream));
    MOIP_CUDA(cudaMemcpyAsync(pool.ub + (size_t)s * n, dm.ubI, sizeof(int) * n, cudaMemcpyDeviceToDevice, stream));
    if (root_valid[cost]) {     // the root iterate of the last IP on this objective (kept on the device)
      MOIP_CUDA(cudaMemcpyAsync(pool.wx + (size_t)s * n, d_root_x.p + (size_t)cost * n, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
      MOIP_CUDA(cudaMemcpyAsync(pool.wy + (size_t)s * m, d_root_y.p + (size_t)cost * m, sizeof(double) * m, cudaMemcpyDeviceToDevice, stream));
    } else {
      MOIP_CUDA(cudaMemsetAsync(pool.wx + (size_t)s * n, 0, sizeof(double) * n, stream));
      MOIP_CUDA(cudaMemsetAsync(pool.wy + (size_t)s * m, 0, sizeof(double) * m, stream));
    }
    open.push_back({s, -HUGE_VAL, 0, 0});
  }
  std::vector<int> ids;
  std::vector<OpenNode> batch;
  std::vector<BranchOp> ops;
  std::vector<int> to_free;
  bool first_round = true;
  // Node LPs are cut short: a round is as slow as its slowest LP, and a weaker (still valid) bound only costs
  // nodes -- until the cap falls below what the LPs of this model need, where the node count explodes (2KP n=100:
  // 8.0e6 nodes at 200 iterations, 2.3e5 at 400, 1.1e5 at 800).  The cap therefore follows the share of node LPs
  // that hit it (kept between bb_cap_lo and bb_cap_hi of a round's LPs), starting from 20 sqrt(n); it is a property
  // of the model and carries over from one IP to the next.  Only the cold root LP of an objective runs long.
  const bool cold_root = !root_valid[cost];
  if (lp_cap_dyn <= 0) lp_cap_dyn = std::min(2000, std::max(200, (int)(20.0 * std::sqrt((double)n))));
  int lp_cap = bb_max_iter > 0 ? bb_max_iter : lp_cap_dyn;
  LpParams lp{};
  lp.eps = bb_eps; lp.max_iter = cold_root ? std::max(lp_cap, 20000) : lp_cap; lp.check_every = bb_check; lp.fixed_iters = 0;
  lp.norm_every = norm_every > 0 ? norm_every : 1; lp.cutoff_slack = 1.0 - 1e-6; lp.int_obj = 1;

  auto prunable = [&](double bound) {
    return have_inc && bound > -HUGE_VAL && std::ceil(bound - 1e-6) >= (double)inc_val;
  };

  const bool prof = std::getenv("MOIP_PROFILE_ROUNDS") != nullptr;
  while (!open.empty()) {
    // ---- select the batch: dive (deepest first) until an incumbent exists, then best bound first
    if (have_inc)
      std::sort(open.begin(), open.end(), [](const OpenNode& a, const OpenNode& b) {
        if (a.bound != b.bound) return a.bound > b.bound;      // best (smallest) bound at the back
        return a.depth < b.depth;
      });
    else
      std::sort(open.begin(), open.end(), [](const OpenNode& a, const OpenNode& b) {
        if (a.depth != b.depth) return a.depth < b.depth;      // deepest at the back
        if (a.pref != b.pref) return a.pref < b.pref;
        return a.bound > b.bound;
      });
    batch.clear(); ids.clear();
    while (!open.empty() && (int)batch.size() < Bmax) {
      OpenNode nd = open.back();
      open.pop_back();
      if (prunable(nd.bound)) { free_slots.push_back(nd.slot); continue; }
      batch.push_back(nd);
      ids.push_back(nd.slot);
    }
    const int B = (int)batch.size();
    if (B == 0) break;
    ip_nodes += B;
    if (explode_log > 0 && ip_nodes > explode_log && !explode_said) {
      explode_said = true;
      std::fprintf(stderr, "moip_b200: IP past %lld nodes: stage %d cost %d rhs [%g %g %g %g] incumbent %s%lld, open %zu, round %lld, node-LP cap %d\n", explode_log, cur_stage, cost,
                   srhs[0], k > 1 ? srhs[1] : 0.0, k > 2 ? srhs[2] : 0.0, k > 3 ? srhs[3] : 0.0, have_inc ? "" : "none ", have_inc ? inc_val : 0LL, open.size(), ip_rounds, lp_cap);
    }
    if (ip_node_budget > 0 && ip_nodes > ip_node_budget) {     // given up: nothing is recorded, the caller decides what to do
      for (auto& nd : batch) free_slots.push_back(nd.slot);
      for (auto& nd : open) free_slots.push_back(nd.slot);
      dbg_where.store(0, std::memory_order_relaxed);
      return MOIP_ERR_BUDGET;
    }
    stats.bb_nodes += B;
    stage_nodes[cur_stage] += B; stage_rounds[cur_stage] += 1; ++ip_rounds;
    dbg_where.store(2, std::memory_order_relaxed); dbg_rounds.store(ip_rounds, std::memory_order_relaxed); dbg_open.store((long long)open.size() + B, std::memory_order_relaxed);
    // ---- device round: propagate -> gather -> LP -> scatter/round -> verify
    std::vector<long long> plo = olo, phi = ohi;
    if (have_inc) {   // incumbent cut-off row on the optimised objective: must be strictly better
      if (M.sense == 0) phi[cost] = std::min(phi[cost], inc_val - 1);
      else plo[cost] = std::max(plo[cost], -inc_val + 1);
    }
    const InLayout LI = in_layout(B);
    const OutLayout LO = out_layout(B);
    const double tp0 = prof ? now_s() : 0.0;
    {
      unsigned char* hi = h_in.p;          // free again: the previous round synchronised after its last use
      std::memcpy(hi + LI.ids, ids.data(), sizeof(int) * B);
      std::memcpy(hi + LI.cost, &cost, sizeof(int));
      std::memcpy(hi + LI.olo, plo.data(), sizeof(long long) * k);
      std::memcpy(hi + LI.ohi, phi.data(), sizeof(long long) * k);
      const double cut = have_inc ? (double)inc_val : HUGE_VAL;
      std::memcpy(hi + LI.cutoff, &cut, sizeof(double));
      std::memcpy(hi + LI.rhs, srhs, sizeof(double) * k);
      kmark(0);
      MOIP_CUDA(cudaMemcpyAsync(r_in.p, hi, LI.end, cudaMemcpyHostToDevice, stream));
      kmark(1);
    }
    const int* d_ids = reinterpret_cast<const int*>(r_in.p + LI.ids);
    const int* d_cost = reinterpret_cast<const int*>(r_in.p + LI.cost);
    const long long* d_olo = reinterpret_cast<const long long*>(r_in.p + LI.olo);
    const long long* d_ohi = reinterpret_cast<const long long*>(r_in.p + LI.ohi);
    const double* d_cutoff = reinterpret_cast<const double*>(r_in.p + LI.cutoff);
    const double* d_rhs = reinterpret_cast<const double*>(r_in.p + LI.rhs);
    int* d_flag = reinterpret_cast<int*>(r_out.p + LO.flag);
    int* d_status = reinterpret_cast<int*>(r_out.p + LO.status);
    int* d_iters = reinterpret_cast<int*>(r_out.p + LO.iters);
    int* d_branch = reinterpret_cast<int*>(r_out.p + LO.branch);
    int* d_ff = reinterpret_cast<int*>(r_out.p + LO.ff);
    double* d_dbound = reinterpret_cast<double*>(r_out.p + LO.dbound);
    double* d_bval = reinterpret_cast<double*>(r_out.p + LO.bval);
    long long* d_leaf = reinterpret_cast<long long*>(r_out.p + LO.leaf);
    long long* d_cobj = reinterpret_cast<long long*>(r_out.p + LO.cobj);
    unsigned char* d_cfeas = r_out.p + LO.cfeas;
    // register-resident K1: propagate -> LP -> round/verify of a node happen in ONE CTA of one launch (fused round)
    const bool fused = dm.reg_ok && use_fused;
    if (!fused && launch_k2_propagate(dm, pool, B, d_ids, d_olo, d_ohi, 16, d_flag, d_leaf, stream)) return MOIP_ERR_CUDA;
    kmark(2);
    LpBatch b{};
    b.B = B; b.rhs = d_rhs; b.lb = pool.lb; b.ub = pool.ub; b.slot = d_ids; b.rc_fix = have_inc ? 1 : 0;
    b.farkas = 1;
    if (fused) {
      b.fused = 1; b.f_obj_lo = d_olo; b.f_obj_hi = d_ohi; b.f_max_rounds = 16; b.f_flag = d_flag; b.f_leaf_obj = d_leaf;
      b.f_xr = r_xr.p; b.f_cand_obj = d_cobj; b.f_cand_feas = d_cfeas; b.f_first_free = d_ff;
    }
    b.warm_x = pool.wx; b.warm_y = pool.wy; b.out_x = pool.wx; b.out_y = pool.wy;   // iterate returns to the node's slot
    b.primal_obj = r_pobj.p; b.dual_bound = d_dbound; b.status = d_status; b.iters = d_iters;
    b.branch_var = d_branch; b.branch_val = d_bval; b.skip = fused ? nullptr : d_flag;
    b.cost_stride = 0; b.rhs_stride = 0; b.cutoff = d_cutoff; b.work_counter = r_counter.p;
    b.cost_idx = d_cost;           // one shared cost index (cost_stride = 0)
    if (attach_k1_scratch(b)) return MOIP_ERR_CUDA;
    if (launch_k1_any(dm, b, lp, num_sms, stream)) return MOIP_ERR_CUDA;
    kmark(3);
    if (!fused && launch_k4_round(dm, B, d_ids, pool.wx, pool.lb, pool.ub, r_xr.p, d_cobj, d_cfeas, d_ff, d_flag, stream)) return MOIP_ERR_CUDA;
    stats.kernel_launches += fused ? 1 : 3;
    kmark(4);
    unsigned char* H = h_round.p;
    MOIP_CUDA(cudaMemcpyAsync(H, r_out.p, LO.end, cudaMemcpyDeviceToHost, stream));
    kmark(5);
    const double tp1 = prof ? now_s() : 0.0;
    if (wait_stream()) return MOIP_ERR_CUDA;
    const double tp2 = prof ? now_s() : 0.0;
    if (ktiming) {
      ktimes.copy_ms += kspan(0, 1) + kspan(4, 5);
      ktimes.k2_ms += kspan(1, 2) + (kev_branch_pending ? kspan(6, 7) : 0.0);
      ktimes.k1_ms += kspan(2, 3);
      ktimes.k4_ms += kspan(3, 4);
      ktimes.rounds += 1;
      kev_branch_pending = false;
    }
    const int* flag = (const int*)(H + LO.flag);
    const int* status = (const int*)(H + LO.status);
    const int* iters = (const int*)(H + LO.iters);
    const int* branch = (const int*)(H + LO.branch);
    const int* ffree = (const int*)(H + LO.ff);
    const double* dbound = (const double*)(H + LO.dbound);
    const double* bval = (const double*)(H + LO.bval);
    const long long* leaf = (const long long*)(H + LO.leaf);
    const long long* cobj = (const long long*)(H + LO.cobj);
    const unsigned char* cfeas = H + LO.cfeas;
    // ---- incumbent candidates of this round (exactly evaluated on the device)
    int best_src = -1; bool best_is_leaf = false;
    long long best_val = inc_val;
    auto within = [&](const long long* ov) {
      for (int o = 0; o < k; ++o) if (ov[o] < olo[o] || ov[o] > ohi[o]) return false;
      return true;
    };
    std::vector<long long> node_cand(B, LLONG_MAX);   // best verified rounding of each node (min-form), if any
    for (int i = 0; i < B; ++i) {
      if (flag[i] == 2) {
        const long long v = (long long)sgn * leaf[(size_t)i * k + cost];
        if (v < best_val) { best_val = v; best_src = i * 3; best_is_leaf = true; }
      } else if (flag[i] == 0) {
        stats.node_lps += 1;
        stage_lps[cur_stage] += 1;
        stats.lp_iterations += iters[i];
        for (int c3 = 0; c3 < 3; ++c3) {
          const size_t w3 = (size_t)i * 3 + c3;
          if (cfeas[w3] && within(cobj + w3 * k)) {
            const long long v = (long long)sgn * cobj[w3 * k + cost];
            if (v < node_cand[i]) node_cand[i] = v;
            if (v < best_val) { best_val = v; best_src = (int)w3; best_is_leaf = false; }
          }
        }
      }
    }
    if (best_src >= 0) {         // the incumbent stays on the device until the IP is solved (no extra synchronise per round)
      const int* src = best_is_leaf ? pool.lb + (size_t)ids[best_src / 3] * n : r_xr.p + (size_t)best_src * n;
      MOIP_CUDA(cudaMemcpyAsync(d_inc.p, src, sizeof(int) * n, cudaMemcpyDeviceToDevice, stream));
      inc_val = best_val; have_inc = true; inc_on_device = true;
    }
    if (!(first_round && cold_root)) {   // cap controller
      int solved = 0, capped = 0;
      for (int i = 0; i < B; ++i)
        if (flag[i] == 0) { ++solved; capped += status[i] == MOIP_LP_ITERLIMIT; }
      cap_seen += solved; cap_hit += capped;
      prof_t[5] += solved; prof_t[6] += capped;
      if (bb_max_iter <= 0 && cap_seen >= 64) {
        const double share = (double)cap_hit / (double)cap_seen;
        if (share > bb_cap_hi) lp_cap_dyn = std::min(6000, lp_cap_dyn + lp_cap_dyn / 4 + 1);
        else if (share < bb_cap_lo) lp_cap_dyn = std::max(100, lp_cap_dyn - lp_cap_dyn / 8);
        cap_seen = 0; cap_hit = 0;
      }
      if (bb_max_iter <= 0) lp_cap = lp_cap_dyn;
    }
    lp.max_iter = lp_cap;
    if (first_round && flag[0] == 0) {   // remember the root iterate as warm start for the next IP on this objective
      MOIP_CUDA(cudaMemcpyAsync(d_root_x.p + (size_t)cost * n, pool.wx + (size_t)ids[0] * n, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
      MOIP_CUDA(cudaMemcpyAsync(d_root_y.p + (size_t)cost * m, pool.wy + (size_t)ids[0] * m, sizeof(double) * m, cudaMemcpyDeviceToDevice, stream));
      root_valid[cost] = 1;
    }
    first_round = false;
    // ---- branch.  While the device is under-filled the tree is expanded several levels per round
    // (children over the 2 or 3 most fractional columns at once): rounds are latency-bound, idle SMs are free.
    ops.clear(); to_free.clear();
    std::vector<int> todo;
    for (int i = 0; i < B; ++i) {
      const OpenNode& nd = batch[i];
      to_free.push_back(nd.slot);
      if (flag[i] != 0) continue;                                  // infeasible or leaf: done
      if (status[i] == MOIP_LP_CUTOFF || status[i] == MOIP_LP_INFEASIBLE) continue;
      const double lbnd = std::max(nd.bound, dbound[i]);
      if (prunable(lbnd)) continue;
      if (branch[3 * i] < 0 && node_cand[i] != LLONG_MAX && (double)node_cand[i] <= std::ceil(lbnd - 1e-6)) continue;   // solved by its rounding
      todo.push_back(i);
    }
    int levels = 1;
    if (bb_levels > 1) {
      while (levels < bb_levels && (long long)todo.size() * (2LL << levels) + (long long)open.size() <= (long long)Bmax) ++levels;
    }
    for (int i : todo) {
      const OpenNode& nd = batch[i];
      const double lbnd = std::max(nd.bound, dbound[i]);
      int vars[3]; double vals[3]; int nv = 0;
      for (int q = 0; q < levels; ++q) if (branch[3 * i + q] >= 0) { vars[nv] = branch[3 * i + q]; vals[nv] = bval[3 * i + q]; ++nv; }
      if (nv == 0) {
        // LP point integral but not accepted: branch on the first unfixed column (found by K4) at its midpoint
        if (ffree[3 * i] >= 0) { vars[0] = ffree[3 * i]; vals[0] = 0.5 * ((double)ffree[3 * i + 1] + (double)ffree[3 * i + 2]); nv = 1; }
        if (nv == 0) continue;   // fully fixed: K2 will have classified it as a leaf
      }
      const int nchild = 1 << nv;
      for (int cmb = 0; cmb < nchild; ++cmb) {
        const int cs = alloc_slot();
        if (cs < 0) return MOIP_ERR_CUDA;
        pool = PoolView{p_lb.p, p_ub.p, p_wx.p, p_wy.p};
        BranchOp op{};
        op.parent = nd.slot; op.child = cs; op.nv = nv;
        int closeness = 0;      // how many columns go to the side nearer to the LP value (explored first when diving)
        for (int q = 0; q < nv; ++q) {
          const int fl = (int)std::floor(vals[q]);
          const bool up = (cmb >> q) & 1;
          op.var[q] = vars[q];
          op.new_lb[q] = up ? fl + 1 : INT_MIN;
          op.new_ub[q] = up ? INT_MAX : fl;
          if (up == ((vals[q] - fl) >= 0.5)) ++closeness;
        }
        ops.push_back(op);
        open.push_back({cs, lbnd, nd.depth + nv, closeness});
      }
    }
    if (!ops.empty()) {
      if (r_ops.ensure(ops.size()) || h_ops.ensure(ops.size())) return MOIP_ERR_CUDA;
      std::memcpy(h_ops.p, ops.data(), sizeof(BranchOp) * ops.size());     // pinned; rewritten only after the next round's sync
      MOIP_CUDA(cudaMemcpyAsync(r_ops.p, h_ops.p, sizeof(BranchOp) * ops.size(), cudaMemcpyHostToDevice, stream));
      kmark(6);
      if (launch_k2_branch(dm, pool, (int)ops.size(), r_ops.p, stream)) return MOIP_ERR_CUDA;
      kmark(7);
      kev_branch_pending = ktiming;
      stats.kernel_launches += 1;
    }
    for (int s : to_free) free_slots.push_back(s);
    if (prof) { const double tp3 = now_s(); prof_t[0] += tp1 - tp0; prof_t[1] += tp2 - tp1; prof_t[2] += tp3 - tp2; prof_t[3] += 1; prof_t[4] += B; }
  }
  for (auto& nd : open) free_slots.push_back(nd.slot);
  dbg_where.store(0, std::memory_order_relaxed);
  if (ip_rounds <= 1) stage_root_solved[cur_stage] += 1;
  return finish(inc_on_device);
}

#endif  // MOIP_HOST_DOUBLE

// int solve(Env&, Problem&, int* result, double* rhs, Thread* t)  -- reference src/aira.cpp:452-536
int moip_ctx::lex_solve(const int* perm, int n_obj, const double* rhs, int* result, int* mip_status) {
  const Model& M = model->M;
  const int k = dm.k, n = dm.n;
  const double t0 = now_s();
  std::vector<double> srhs(rhs, rhs + k);            // :457-462
  std::vector<int> x;
  int st = MOIP_MIP_INFEASIBLE;
  for (int jp = 0; jp < n_obj; ++jp) {               // :467
    const int j = perm[jp];
    IpResult r;
    cur_stage = jp;
    int rc = solve_ip(j, srhs.data(), x.empty() ? nullptr : &x, r);
    cur_stage = MOIP_MAX_OBJ;
    if (rc) return rc;
    st = r.status;
    if (st == MOIP_MIP_INFEASIBLE) break;            // :489-492
    x = r.x;
    result[j] = (int)r.obj;                          // :517 (exact, no rounding needed)
    srhs[j] = (double)r.obj;
  }
  if (st != MOIP_MIP_INFEASIBLE) {
    for (int jp = n_obj; jp < k; ++jp) {             // :520-530
      const int j = perm[jp];
      long long v = 0;
      for (int q = 0; q < n; ++q) v += M.ci[(size_t)j * n + q] * (long long)x[q];
      result[j] = (int)v;
    }
  }
  if (mip_status) *mip_status = st;
  stats.solver_seconds += now_s() - t0;
  return MOIP_OK;
}

// void get_limit(Env&, Problem&, int obj, double* rhs, int* result, Sense) -- reference src/aira.cpp:367-450
int moip_ctx::get_limit(int obj, int sense, const double* rhs, int* result, int* mip_status) {
  const Model& M = model->M;
  if (sense != M.sense) {
    std::fprintf(stderr, "moip_b200: get_limit with a sense different from the model's is not supported\n");
    return MOIP_ERR_UNSUPPORTED;
  }
  const double t0 = now_s();
  IpResult r;
  int rc = solve_ip(obj, rhs, nullptr, r);
  if (rc) return rc;
  if (mip_status) *mip_status = r.status;
  if (r.status != MOIP_MIP_INFEASIBLE) {             // :437-447; untouched when infeasible (:410-412)
    for (int j = 0; j < dm.k; ++j) {
      long long v = 0;
      for (int q = 0; q < dm.n; ++q) v += M.ci[(size_t)j * dm.n + q] * (long long)r.x[q];
      result[j] = (int)v;
    }
  }
  stats.solver_seconds += now_s() - t0;
  return MOIP_OK;
}

// One CPXmipopt (+ CPXgetstat / CPXgetobjval / CPXgetx): optimise objective `obj` in the model's sense subject to the
// structural rows and the k objective-bound rows at `rhs`.  x_start, when given, must satisfy all of them.
extern "C" int moip_mip_solve(moip_ctx* c, int obj, const double* rhs, const int32_t* x_start, int32_t* x_out,
                              int64_t* objval, int* mip_status) {
  if (!c || !rhs || obj < 0 || obj >= c->dm.k) return MOIP_ERR_ARG;
  const double t0 = now_s();
  std::vector<int> start;
  if (x_start) start.assign(x_start, x_start + c->dm.n);
  IpResult r;
  int rc = c->solve_ip(obj, rhs, x_start ? &start : nullptr, r);
  if (rc) return rc;
  if (mip_status) *mip_status = r.status;
  if (r.status != MOIP_MIP_INFEASIBLE) {
    if (x_out) for (int j = 0; j < c->dm.n; ++j) x_out[j] = r.x[j];
    if (objval) *objval = r.obj;
  }
  c->stats.solver_seconds += now_s() - t0;
  return MOIP_OK;
}

extern "C" int moip_lex_solve(moip_ctx* c, const int* perm, int n_obj, const double* rhs, int* result, int* mip_status) {
  if (!c || !perm || !rhs || !result || n_obj < 1 || n_obj > c->dm.k) return MOIP_ERR_ARG;
  for (int i = 0; i < c->dm.k; ++i) if (perm[i] < 0 || perm[i] >= c->dm.k) return MOIP_ERR_ARG;
  return c->lex_solve(perm, n_obj, rhs, result, mip_status);
}
extern "C" int moip_get_limit(moip_ctx* c, int obj, int sense, const double* rhs, int* result, int* mip_status) {
  if (!c || !rhs || !result || obj < 0 || obj >= c->dm.k) return MOIP_ERR_ARG;
  return c->get_limit(obj, sense, rhs, result, mip_status);
}
