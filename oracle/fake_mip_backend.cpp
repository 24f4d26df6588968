// TEST INFRASTRUCTURE ONLY -- never linked into or loaded by the product.
//
// LD_PRELOAD test double for the solver entry points that the link-level CPLEX seam
// (moip_aira_b200/seam1/cpx_shim.cpp) calls: moip_ctx_create_own_stream, moip_mip_solve, moip_ctx_destroy,
// moip_ctx_stats, moip_device_count.  It answers one single-objective IP by exhaustive depth-first enumeration with activity-bound
// pruning on the CPU (binary models of a few dozen columns only), using nothing but the public model getters of
// include/moip_b200.h.  Purpose: the host logic of the seam and of the UNMODIFIED reference driver built on it
// (oracle/_ref/aira_seam1: option parsing, Problem::read_lp_problem through CPXgetrows, the solve()/get_limit()
// call sequences, -t N threads, --split) can be exercised by `pytest -m "not gpu"` in a container without a GPU.
// The same tests run on the GPU box WITHOUT this preload, i.e. on the real kernels (tests/test_seam1.py).
//
// What it restates: the contract of CPXmipopt as the reference uses it (src/aira.cpp:480-521) -- minimise /
// maximise objective `obj` subject to the structural rows and the k objective-bound rows `C x <= rhs` (MIN) /
// `>= rhs` (MAX), +-1e20 = free; status 101 optimal / 103 infeasible.
#include <cmath>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "moip_b200.h"

struct moip_ctx {
  moip_model* model;
  moip_model_info info;
  std::vector<double> a, rhs, lb, ub, c;   // dense structural rows, objectives k*n
  std::vector<char> sense;
  long long ip_solved = 0;
};

namespace {
struct Row {
  std::vector<long long> a, sufmin, sufmax;   // coefficients, suffix sums of min(0,a) / max(0,a)
  long long lo, hi;
};

struct Search {
  int n;
  std::vector<Row> rows;
  std::vector<long long> c, csufmin;   // min-form objective
  std::vector<long long> act;
  std::vector<int> x, best_x;
  long long best = LLONG_MAX;
  long long nodes = 0;
  bool have = false;
  void dfs(int j, long long cur) {
    if (++nodes > 200000000LL) { std::fprintf(stderr, "fake_mip_backend: model too large for enumeration\n"); std::abort(); }
    for (size_t r = 0; r < rows.size(); ++r) {
      const Row& R = rows[r];
      if (act[r] + R.sufmin[j] > R.hi || act[r] + R.sufmax[j] < R.lo) return;
    }
    if (cur + csufmin[j] >= best) return;
    if (j == n) { best = cur; best_x = x; have = true; return; }
    const int first = c[j] < 0 ? 1 : 0;
    for (int t = 0; t < 2; ++t) {
      const int v = t == 0 ? first : 1 - first;
      x[j] = v;
      if (v) for (size_t r = 0; r < rows.size(); ++r) act[r] += rows[r].a[j];
      dfs(j + 1, cur + (v ? c[j] : 0));
      if (v) for (size_t r = 0; r < rows.size(); ++r) act[r] -= rows[r].a[j];
    }
    x[j] = 0;
  }
};

void finish(Row& R, int n) {
  R.sufmin.assign(n + 1, 0);
  R.sufmax.assign(n + 1, 0);
  for (int j = n - 1; j >= 0; --j) {
    R.sufmin[j] = R.sufmin[j + 1] + (R.a[j] < 0 ? R.a[j] : 0);
    R.sufmax[j] = R.sufmax[j + 1] + (R.a[j] > 0 ? R.a[j] : 0);
  }
}
}  // namespace

// FAKE_MIP_DEVICES=G pretends G devices are visible and FAKE_MIP_LOG_DEVICES=1 prints the device of every context, so
// that the seam's placement of worker contexts over GPUs (MOIP_B200_DEVICE / MOIP_B200_DEVICES) is testable on the CPU.
extern "C" int moip_device_count(void) {
  const char* s = std::getenv("FAKE_MIP_DEVICES");
  return (s && *s) ? std::atoi(s) : 1;
}

extern "C" int moip_ctx_create_own_stream(moip_model* m, int device, moip_ctx** out) {
  if (!m || !out) return MOIP_ERR_ARG;
  if (device < 0 || device >= moip_device_count()) return MOIP_ERR_ARG;
  if (std::getenv("FAKE_MIP_LOG_DEVICES")) std::fprintf(stderr, "fake_mip_backend: context on device %d\n", device);
  moip_ctx* c = new moip_ctx();
  c->model = m;
  if (moip_model_get_info(m, &c->info)) return MOIP_ERR_ARG;
  const int n = c->info.n, ms = c->info.ms, k = c->info.k;
  c->a.resize((size_t)ms * n); c->rhs.resize(ms); c->sense.resize(ms); c->lb.resize(n); c->ub.resize(n);
  std::vector<uint8_t> isint(n);
  if (moip_model_dense(m, c->a.data(), c->sense.data(), c->rhs.data(), c->lb.data(), c->ub.data(), isint.data())) return MOIP_ERR_ARG;
  for (int j = 0; j < n; ++j)
    if (!isint[j] || c->lb[j] != 0.0 || c->ub[j] != 1.0) {
      std::fprintf(stderr, "fake_mip_backend: binary models only\n");
      return MOIP_ERR_UNSUPPORTED;
    }
  c->c.resize((size_t)k * n);
  for (int o = 0; o < k; ++o) moip_model_objcoef(m, o, c->c.data() + (size_t)o * n);
  *out = c;
  return MOIP_OK;
}

extern "C" void moip_ctx_destroy(moip_ctx* c) { delete c; }

extern "C" int moip_ctx_stats(const moip_ctx* c, moip_stats* out) {
  if (!c || !out) return MOIP_ERR_ARG;
  *out = moip_stats{};
  out->ip_solved = c->ip_solved;   // no kernels, no node LPs: this is the CPU test double
  return MOIP_OK;
}

extern "C" int moip_mip_solve(moip_ctx* c, int obj, const double* rhs, const int32_t* /*x_start*/, int32_t* x_out,
                              int64_t* objval, int* mip_status) {
  if (!c || !rhs || obj < 0 || obj >= c->info.k) return MOIP_ERR_ARG;
  const int n = c->info.n, ms = c->info.ms, k = c->info.k;
  c->ip_solved += 1;
  const long long sgn = c->info.sense == MOIP_SENSE_MIN ? 1 : -1;
  Search S;
  S.n = n;
  for (int r = 0; r < ms; ++r) {
    Row R;
    R.a.resize(n);
    for (int j = 0; j < n; ++j) R.a[j] = (long long)std::llround(c->a[(size_t)r * n + j]);
    R.lo = LLONG_MIN / 4; R.hi = LLONG_MAX / 4;
    if (c->sense[r] == 'L' || c->sense[r] == 'E') R.hi = (long long)std::floor(c->rhs[r] + 1e-9);
    if (c->sense[r] == 'G' || c->sense[r] == 'E') R.lo = (long long)std::ceil(c->rhs[r] - 1e-9);
    finish(R, n);
    S.rows.push_back(R);
  }
  for (int o = 0; o < k; ++o) {
    if (std::fabs(rhs[o]) >= 1e19) continue;
    Row R;
    R.a.resize(n);
    for (int j = 0; j < n; ++j) R.a[j] = (long long)std::llround(c->c[(size_t)o * n + j]);
    R.lo = LLONG_MIN / 4; R.hi = LLONG_MAX / 4;
    if (sgn > 0) R.hi = (long long)std::floor(rhs[o] + 1e-9);
    else R.lo = (long long)std::ceil(rhs[o] - 1e-9);
    finish(R, n);
    S.rows.push_back(R);
  }
  S.c.resize(n);
  for (int j = 0; j < n; ++j) S.c[j] = sgn * (long long)std::llround(c->c[(size_t)obj * n + j]);
  S.csufmin.assign(n + 1, 0);
  for (int j = n - 1; j >= 0; --j) S.csufmin[j] = S.csufmin[j + 1] + (S.c[j] < 0 ? S.c[j] : 0);
  S.act.assign(S.rows.size(), 0);
  S.x.assign(n, 0);
  S.dfs(0, 0);
  if (mip_status) *mip_status = S.have ? MOIP_MIP_OPTIMAL : MOIP_MIP_INFEASIBLE;
  if (S.have) {
    if (x_out) for (int j = 0; j < n; ++j) x_out[j] = S.best_x[j];
    if (objval) *objval = sgn * S.best;
  }
  return MOIP_OK;
}
