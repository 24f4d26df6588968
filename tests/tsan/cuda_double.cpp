// TEST INFRASTRUCTURE ONLY -- never linked into or loaded by the product.
//
// CPU stand-ins that let the HOST code of the library -- csrc/solver.cu (contexts, caches, lexicographic chain),
// csrc/generator.cpp (generator state machine, worker pool, work stealing, record exchange, cooperative workers),
// csrc/capi.cpp, csrc/model.cpp and seam1/cpx_shim.cpp -- be compiled with g++ -fsanitize=thread and run in a container
// without a GPU (oracle/Makefile target `tsan`, tests/test_tsan_host.py).  Three groups:
//   * the few CUDA runtime entry points that host code calls: "device" memory is host memory, streams and events are
//     tokens, copies are memcpy;
//   * the kernel launchers: K3 (the cache scan) runs on the CPU with the semantics of k3_scan_kernel / k3_scan1_kernel
//     (reference src/solutions.cpp:11-81); K1/K2/K4 are not available here and abort;
//   * moip_ctx::solve_ip: one exact IP by depth-first enumeration with activity-bound pruning on the int64 image of the
//     model (small models only) -- solver.cu leaves its GPU branch-and-bound out under MOIP_HOST_DOUBLE.
// What is under test is the locking / atomics discipline of the real host code, not these stand-ins.
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../moip_aira_b200/csrc/solver.h"

// ------------------------------------------------------------------------------------ CUDA runtime
extern "C" {
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp* p, int) { std::memset(p, 0, sizeof(*p)); p->major = 10; p->multiProcessorCount = 148; return cudaSuccess; }
cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void** p, size_t n) { *p = std::malloc(n ? n : 1); return cudaSuccess; }
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = std::malloc(n ? n : 1); return cudaSuccess; }
cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return cudaSuccess; }
cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)std::malloc(8); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)std::malloc(8); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)std::malloc(8); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { std::free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "host double"; }
}

// ------------------------------------------------------------------------------------ kernel launchers
namespace moip {

static int unavailable(const char* what) {
  std::fprintf(stderr, "host double: %s needs the GPU\n", what);
  std::abort();
  return MOIP_ERR_CUDA;
}
int launch_k1_any(const DevModel&, const LpBatch&, const LpParams&, int, cudaStream_t) { return unavailable("K1"); }
int launch_expand_masks(const DevModel&, int, const uint32_t*, int, int*, int*, cudaStream_t) { return unavailable("expand_masks"); }
size_t k1_scratch_stride(const DevModel&) { return 0; }
int launch_k2_propagate(const DevModel&, const PoolView&, int, const int*, const long long*, const long long*, int, int*, long long*, cudaStream_t) { return unavailable("K2"); }
int launch_k2_branch(const DevModel&, const PoolView&, int, const BranchOp*, cudaStream_t) { return unavailable("K2"); }
int launch_k4_round(const DevModel&, int, const int*, const double*, const int*, const int*, int*, long long*, unsigned char*, int*, const int*, cudaStream_t) { return unavailable("K4"); }
int launch_k4(const DevModel&, int, const int*, const double*, long long*, unsigned char*, cudaStream_t) { return unavailable("K4"); }

// Solutions::find over one store (reference src/solutions.cpp:11-81): first record that is a relaxation of ip whose
// answer stays valid
static int scan(const DevCache& c, const double* ip, int sense) {
  for (int r = 0; r < c.size; ++r) {
    const CacheRecord& R = c.rec[r];
    bool ok = true;
    for (int i = 0; i < c.k && ok; ++i) {
      if (sense == MOIP_SENSE_MIN) ok = !(R.ip[i] < ip[i]) && !(!R.infeasible && (double)R.result[i] > ip[i]);
      else ok = !(R.ip[i] > ip[i]) && !(!R.infeasible && (double)R.result[i] < ip[i]);
    }
    if (ok) return r;
  }
  return -1;
}
int launch_k3(const DevCache& c0, const DevCache& c1, int Q, const double* queries, int sense, int* first_match, int* which, cudaStream_t) {
  for (int q = 0; q < Q; ++q) {
    int f = scan(c0, queries + (size_t)q * c0.k, sense), w = 0;
    if (f < 0) { f = scan(c1, queries + (size_t)q * c0.k, sense); w = 1; }
    first_match[q] = f;
    if (which) which[q] = f >= 0 ? w : -1;
  }
  return MOIP_OK;
}
int launch_k3_one(const DevCache& c0, const DevCache& c1, const K3Query& q, int sense, K3Answer* a, int seq, cudaStream_t) {
  int f = scan(c0, q.ip, sense), w = 0;
  if (f < 0) { f = scan(c1, q.ip, sense); w = 1; }
  if (f >= 0) a->rec = (w == 0 ? c0 : c1).rec[f];
  a->first_match = f; a->which = f >= 0 ? w : -1;
  std::atomic_thread_fence(std::memory_order_release);
  *reinterpret_cast<volatile int*>(&a->seq) = seq;
  return MOIP_OK;
}


}  // namespace moip

// ------------------------------------------------------------------------------------ solve_ip by enumeration
namespace {
struct Enum {
  const moip::Model& M;
  int n, k, cost;
  long long sgn;
  std::vector<long long> olo, ohi;
  std::vector<std::vector<long long>> rows;   // dense structural rows + k objective rows
  std::vector<long long> rlo, rhi, act, sufmin_c;
  std::vector<std::vector<long long>> sufmin, sufmax;
  std::vector<int> x, best_x;
  long long best = LLONG_MAX;
  bool have = false;
  long long nodes = 0;
  Enum(const moip::Model& M_, int cost_, const std::vector<long long>& lo, const std::vector<long long>& hi)
      : M(M_), n(M_.n), k(M_.k), cost(cost_), sgn(M_.sense == 0 ? 1 : -1), olo(lo), ohi(hi) {
    for (int i = 0; i < M.ms; ++i) {
      std::vector<long long> r(n, 0);
      for (int e = M.a_ptr[i]; e < M.a_ptr[i + 1]; ++e) r[M.a_col[e]] += M.ai_val[e];
      rows.push_back(r);
      rlo.push_back(M.ri_lo[i] == INT64_MIN ? LLONG_MIN / 4 : M.ri_lo[i]);
      rhi.push_back(M.ri_hi[i] == INT64_MAX ? LLONG_MAX / 4 : M.ri_hi[i]);
    }
    for (int o = 0; o < k; ++o) {
      rows.emplace_back(M.ci.begin() + (size_t)o * n, M.ci.begin() + (size_t)(o + 1) * n);
      rlo.push_back(olo[o] == LLONG_MIN ? LLONG_MIN / 4 : olo[o]);
      rhi.push_back(ohi[o] == LLONG_MAX ? LLONG_MAX / 4 : ohi[o]);
    }
    const size_t R = rows.size();
    sufmin.assign(R, std::vector<long long>(n + 1, 0));
    sufmax.assign(R, std::vector<long long>(n + 1, 0));
    for (size_t r = 0; r < R; ++r)
      for (int j = n - 1; j >= 0; --j) {
        const long long a = rows[r][j], l = M.lbI[j], u = M.ubI[j];
        sufmin[r][j] = sufmin[r][j + 1] + std::min(a * l, a * u);
        sufmax[r][j] = sufmax[r][j + 1] + std::max(a * l, a * u);
      }
    sufmin_c.assign(n + 1, 0);
    for (int j = n - 1; j >= 0; --j) {
      const long long c = sgn * M.ci[(size_t)cost * n + j];
      sufmin_c[j] = sufmin_c[j + 1] + std::min(c * M.lbI[j], c * M.ubI[j]);
    }
    act.assign(R, 0);
    x.assign(n, 0);
  }
  void dfs(int j, long long cur) {
    if (++nodes > 50000000LL) { std::fprintf(stderr, "host double: model too large for enumeration\n"); std::abort(); }
    for (size_t r = 0; r < rows.size(); ++r)
      if (act[r] + sufmin[r][j] > rhi[r] || act[r] + sufmax[r][j] < rlo[r]) return;
    if (cur + sufmin_c[j] >= best) return;
    if (j == n) { best = cur; best_x = x; have = true; return; }
    const long long c = sgn * M.ci[(size_t)cost * n + j];
    const int lo = M.lbI[j], hi = M.ubI[j];
    for (int t = 0; t <= hi - lo; ++t) {
      const int v = c < 0 ? hi - t : lo + t;
      x[j] = v;
      for (size_t r = 0; r < rows.size(); ++r) act[r] += rows[r][j] * v;
      dfs(j + 1, cur + c * v);
      for (size_t r = 0; r < rows.size(); ++r) act[r] -= rows[r][j] * v;
    }
  }
};
}  // namespace

int moip_ctx::solve_ip(int cost, const double* srhs, const std::vector<int>* inc_x_in, moip::IpResult& out) {
  const moip::Model& M = model->M;
  const int k = dm.k;
  stats.ip_solved += 1;
  stats.bb_nodes += 1;
  out.status = MOIP_MIP_INFEASIBLE;
  out.x.clear();
  if (M.int_infeasible) return MOIP_OK;
  // node budget (moip_ctx_set_ip_node_budget; the box scheduler's postponement): a tiny budget makes every other budgeted
  // IP give up, so that the put-back / retry path of run_boxes is exercised
  if (ip_node_budget > 0 && ip_node_budget <= 8) {
    static std::atomic<unsigned> flip{0};
    if (flip.fetch_add(1) % 2 == 0) return MOIP_ERR_BUDGET;
  }
  std::vector<long long> olo(k, LLONG_MIN), ohi(k, LLONG_MAX);
  for (int o = 0; o < k; ++o) {
    if (std::fabs(srhs[o]) >= 1e19) continue;
    if (M.sense == 0) ohi[o] = (long long)std::floor(srhs[o] + 1e-9);
    else olo[o] = (long long)std::ceil(srhs[o] - 1e-9);
  }
  Enum e(M, cost, olo, ohi);
  // MIP starts: the caller's point and the model's point store are real host code under test
  if (inc_x_in && (int)inc_x_in->size() == dm.n) {
    long long v = 0;
    for (int j = 0; j < dm.n; ++j) v += M.ci[(size_t)cost * dm.n + j] * (long long)(*inc_x_in)[j];
    e.best = e.sgn * v; e.best_x = *inc_x_in; e.have = true;
  }
  if (use_points) {
    long long v = 0;
    std::vector<int> px;
    if (model->points.best(cost, e.sgn, olo.data(), ohi.data(), e.have ? e.best : LLONG_MAX, px, v)) { e.best = v; e.best_x = px; e.have = true; start_hits += 1; }
  }
  e.dfs(0, 0);
  if (!e.have) return MOIP_OK;
  out.status = MOIP_MIP_OPTIMAL;
  out.obj = e.sgn * e.best;
  out.x = e.best_x;
  if (use_points) {
    long long ov[MOIP_MAX_OBJ] = {0, 0, 0, 0};
    for (int o = 0; o < k; ++o)
      for (int j = 0; j < dm.n; ++j) ov[o] += M.ci[(size_t)o * dm.n + j] * (long long)out.x[j];
    model->points.add(out.x.data(), ov);
  }
  return MOIP_OK;
}
