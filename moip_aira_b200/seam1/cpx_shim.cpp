// cpx_shim.cpp -- the 23 CPLEX Callable Library entry points moip_aira binds ("Seam 1"),
// implemented on the C ABI of the B200 solver core (include/moip_b200.h).  Built as libcplex_moip_b200.so; the
// UNMODIFIED reference sources (src/aira.cpp, src/problem.cpp ...) compile against seam1/include/ilcplex/cplex.h and
// link against this library, so `aira -p X.lp -t N [--split]` runs with every CPXmipopt served by the GPU
// branch and bound (K1 node LPs, K2 node pool, K4 int64 verification).  Nothing here computes: a call either
// answers from the loaded model (sizes, rows, names) or forwards to libmoip_b200; without a B200 CPXmipopt fails
// loudly (status != 0 and a diagnostic), there is no CPU solve.
//
// What the reference does with a problem object is narrow (src/problem.cpp:28-152, :157-340, src/aira.cpp:367-536):
// read the file, ask sizes / the last k rows / column names, turn the k objective rows into bound rows, then
// repeat { CPXchgobj(one of the k objectives), CPXchgrhs(the k bound rows), CPXmipopt, CPXgetstat, CPXgetobjval,
// CPXgetx }.  The shim supports exactly that family and refuses anything outside it (an objective vector that is
// none of the model's k objectives, a change to a structural row, a solve against the model's sense) with a
// nonzero status and a message on stderr, never with a wrong answer.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "../../include/moip_b200.h"
#include "include/ilcplex/cplex.h"

namespace {

// CPLEX error codes the callable library documents for these situations (values from the public manual)
constexpr int kErrNoEnv = 1002;        // CPXERR_NO_ENVIRONMENT
constexpr int kErrNullPointer = 1004;  // CPXERR_NULL_POINTER
constexpr int kErrNoProblem = 1009;    // CPXERR_NO_PROBLEM
constexpr int kErrBadArg = 1003;       // CPXERR_BAD_ARGUMENT
constexpr int kErrIndexRange = 1200;   // CPXERR_INDEX_RANGE
constexpr int kErrNegSurplus = 1207;   // CPXERR_NEGATIVE_SURPLUS
constexpr int kErrNoSoln = 1217;       // CPXERR_NO_SOLN
constexpr int kErrFailOpenRead = 1423; // CPXERR_FAIL_OPEN_READ
constexpr int kErrUnsupported = 1811;  // CPXERR_UNSUPPORTED_OPERATION
constexpr int kErrSolver = 3003;       // solver backend (GPU) failure -- see the MOIP_ERR_* diagnostic printed with it

std::atomic<int> g_lp_serial{0};

// MOIP_B200_SEAM_STATS=1: one line on stderr at process exit with what the GPU did for this process (the reference's
// worker threads never free their problem objects, src/aira.cpp:1877-1884, so the contexts are summed up here).
struct Registry {
  std::mutex mu;
  std::set<moip_cpxlp*> live;
  moip_stats done{};
  long long mipopt_calls = 0, mipopt_memo = 0;
  void add(const moip_stats& s) {
    done.ip_solved += s.ip_solved; done.bb_nodes += s.bb_nodes; done.node_lps += s.node_lps;
    done.lp_iterations += s.lp_iterations; done.kernel_launches += s.kernel_launches;
    done.cache_queries += s.cache_queries; done.solver_seconds += s.solver_seconds;
  }
  ~Registry();
};
Registry g_registry;

int env_int(const char* name, int dflt) {
  const char* s = std::getenv(name);
  return (s && *s) ? std::atoi(s) : dflt;
}

}  // namespace

struct moip_cpxenv {
  int threads = 1;
  int parallel = 0;
  int scrind = 0;
  double mipgap = 1e-4;
};

struct moip_cpxlp {
  std::string name;
  moip_model* model = nullptr;
  moip_ctx* ctx = nullptr;
  moip_model_info info{};
  bool loaded = false;
  bool is_mop = false;       // objective rows are not rows of the file; CPXaddrows appends them (src/problem.cpp:321)
  bool rows_added = false;
  int serial = 0;            // creation order, used to spread contexts over the visible GPUs
  std::vector<double> objcoef;      // k*n, the model's objectives
  std::vector<int> obj_nnz;         // nonzeros per objective row
  std::vector<double> objrow_rhs;   // current right-hand sides of the k objective-bound rows
  std::vector<double> cur_obj;      // objective vector as set through CPXchgobj
  int cur_obj_index = -1;           // which of the k objectives cur_obj equals, -1 = none
  int cur_sense = CPX_MIN;
  // structural part on demand (CPXgetrows / CPXgetrhs on structural rows)
  bool have_dense = false;
  std::vector<double> a, rhs_s;
  std::vector<char> sense_s;
  // last solve
  bool solved = false;
  int stat = 0;
  double objval = 0;
  std::vector<int32_t> x;
  bool have_x = false;
  int solved_obj = -1;
  std::vector<double> solved_rhs;

  int k() const { return info.k; }
  int n() const { return info.n; }
  int ms() const { return info.ms; }
  int numrows() const { return ms() + ((is_mop && !rows_added) ? 0 : k()); }
  void match_objective() {
    cur_obj_index = -1;
    for (int j = 0; j < k(); ++j)
      if (std::memcmp(cur_obj.data(), objcoef.data() + (size_t)j * n(), sizeof(double) * n()) == 0) {
        cur_obj_index = j;
        return;
      }
    // -0.0 vs 0.0 and the like: fall back to a value compare
    for (int j = 0; j < k(); ++j) {
      bool same = true;
      for (int q = 0; q < n() && same; ++q) same = cur_obj[q] == objcoef[(size_t)j * n() + q];
      if (same) {
        cur_obj_index = j;
        return;
      }
    }
  }
  int ensure_dense() {
    if (have_dense) return 0;
    a.assign((size_t)ms() * n(), 0.0);
    rhs_s.assign(ms(), 0.0);
    sense_s.assign(ms(), 'L');
    std::vector<double> lb(n()), ub(n());
    std::vector<uint8_t> isint(n());
    if (moip_model_dense(model, a.data(), sense_s.data(), rhs_s.data(), lb.data(), ub.data(), isint.data())) return kErrBadArg;
    have_dense = true;
    return 0;
  }
  // does the kept solution satisfy the k objective-bound rows at their current right-hand sides?
  bool x_within_bounds() const {
    if (!have_x) return false;
    for (int j = 0; j < k(); ++j) {
      const double r = objrow_rhs[j];
      if (std::fabs(r) >= CPX_INFBOUND) continue;
      double v = 0;   // integer data: exact in fp64 far beyond the int range the reference reports
      for (int q = 0; q < n(); ++q) v += objcoef[(size_t)j * n() + q] * (double)x[q];
      if (info.sense == MOIP_SENSE_MIN ? v > r : v < r) return false;
    }
    return true;
  }
};

namespace {
Registry::~Registry() {
  const char* want = std::getenv("MOIP_B200_SEAM_STATS");
  if (!want || !*want || *want == '0') return;
  for (moip_cpxlp* lp : live) {
    moip_stats st{};
    if (lp->ctx && moip_ctx_stats(lp->ctx, &st) == MOIP_OK) add(st);
  }
  std::fprintf(stderr, "cplex shim: %lld CPXmipopt calls (%lld answered from the previous identical solve), %lld IPs, "
               "%lld B&B nodes, %lld node LPs, %lld LP iterations, %lld kernel launches, %.3f s in the solver\n",
               mipopt_calls, mipopt_memo, (long long)done.ip_solved, (long long)done.bb_nodes, (long long)done.node_lps,
               (long long)done.lp_iterations, (long long)done.kernel_launches, done.solver_seconds);
}

bool has_suffix(const std::string& s, const char* suf) {
  const size_t l = std::strlen(suf);
  return s.size() >= l && s.compare(s.size() - l, l, suf) == 0;
}
}  // namespace

extern "C" {

CPXENVptr CPXopenCPLEX(int* status_p) {
  if (status_p) *status_p = 0;
  return new moip_cpxenv();
}

int CPXcloseCPLEX(CPXENVptr* env_p) {
  if (!env_p || !*env_p) return kErrNoEnv;
  delete *env_p;
  *env_p = nullptr;
  return 0;
}

CPXLPptr CPXcreateprob(CPXCENVptr env, int* status_p, const char* probname) {
  if (!env) {
    if (status_p) *status_p = kErrNoEnv;
    return nullptr;
  }
  moip_cpxlp* lp = new moip_cpxlp();
  lp->name = probname ? probname : "";
  lp->serial = g_lp_serial.fetch_add(1);
  {
    std::lock_guard<std::mutex> g(g_registry.mu);
    g_registry.live.insert(lp);
  }
  if (status_p) *status_p = 0;
  return lp;
}

int CPXfreeprob(CPXCENVptr env, CPXLPptr* lp_p) {
  if (!env) return kErrNoEnv;
  if (!lp_p || !*lp_p) return kErrNoProblem;
  moip_cpxlp* lp = *lp_p;
  {
    std::lock_guard<std::mutex> g(g_registry.mu);
    g_registry.live.erase(lp);
    moip_stats st{};
    if (lp->ctx && moip_ctx_stats(lp->ctx, &st) == MOIP_OK) g_registry.add(st);
  }
  if (lp->ctx) moip_ctx_destroy(lp->ctx);
  if (lp->model) moip_model_free(lp->model);
  delete lp;
  *lp_p = nullptr;
  return 0;
}

int CPXreadcopyprob(CPXCENVptr env, CPXLPptr lp, const char* filename, const char* /*filetype*/) {
  if (!env) return kErrNoEnv;
  if (!lp) return kErrNoProblem;
  if (!filename) return kErrNullPointer;
  if (lp->loaded) {
    std::fprintf(stderr, "cplex shim: CPXreadcopyprob on a problem that already holds a model is not supported\n");
    return kErrUnsupported;
  }
  if (moip_model_load(filename, &lp->model) != MOIP_OK || !lp->model) return kErrFailOpenRead;
  if (moip_model_get_info(lp->model, &lp->info) != MOIP_OK) return kErrFailOpenRead;
  lp->is_mop = has_suffix(filename, ".mop");
  const int k = lp->k(), n = lp->n();
  lp->objcoef.assign((size_t)k * n, 0.0);
  lp->obj_nnz.assign(k, 0);
  for (int j = 0; j < k; ++j) {
    if (moip_model_objcoef(lp->model, j, lp->objcoef.data() + (size_t)j * n) != MOIP_OK) return kErrFailOpenRead;
    for (int q = 0; q < n; ++q) lp->obj_nnz[j] += lp->objcoef[(size_t)j * n + q] != 0.0;
  }
  // the extended .lp numbers its objective rows 1..k on the right-hand side, the last one carrying the count
  // (src/problem.cpp:54-61); a .mop gets its rows from CPXaddrows
  lp->objrow_rhs.resize(k);
  for (int j = 0; j < k; ++j) lp->objrow_rhs[j] = j + 1;
  // objective function of the file: "maximize 0" in the .lp dialect, the first N row in a .mop
  lp->cur_obj.assign(n, 0.0);
  if (lp->is_mop && k > 0) {
    std::memcpy(lp->cur_obj.data(), lp->objcoef.data(), sizeof(double) * n);
    lp->cur_obj_index = 0;
  }
  lp->cur_sense = lp->info.sense == MOIP_SENSE_MIN ? CPX_MIN : CPX_MAX;
  lp->loaded = true;
  return 0;
}

int CPXgetnumcols(CPXCENVptr env, CPXCLPptr lp) { return (env && lp && lp->loaded) ? lp->n() : 0; }
int CPXgetnumrows(CPXCENVptr env, CPXCLPptr lp) { return (env && lp && lp->loaded) ? lp->numrows() : 0; }
int CPXgetnumnz(CPXCENVptr env, CPXCLPptr lp) {
  if (!env || !lp || !lp->loaded) return 0;
  int nz = lp->info.nnz;
  if (!lp->is_mop) for (int j = 0; j < lp->k(); ++j) nz += lp->obj_nnz[j];
  else if (lp->rows_added) nz += lp->k() * lp->n();   // CPXaddrows received dense rows (src/problem.cpp:222-235)
  return nz;
}

int CPXgetrhs(CPXCENVptr env, CPXCLPptr lp_c, double* rhs, int begin, int end) {
  moip_cpxlp* lp = const_cast<moip_cpxlp*>(lp_c);
  if (!env) return kErrNoEnv;
  if (!lp || !lp->loaded) return kErrNoProblem;
  if (!rhs) return kErrNullPointer;
  if (begin < 0 || end >= lp->numrows() || begin > end) return kErrIndexRange;
  for (int r = begin; r <= end; ++r) {
    if (r >= lp->ms()) rhs[r - begin] = lp->objrow_rhs[r - lp->ms()];
    else {
      if (int rc = lp->ensure_dense()) return rc;
      rhs[r - begin] = lp->rhs_s[r];
    }
  }
  return 0;
}

int CPXgetrows(CPXCENVptr env, CPXCLPptr lp_c, int* nzcnt_p, int* rmatbeg, int* rmatind, double* rmatval,
               int rmatspace, int* surplus_p, int begin, int end) {
  moip_cpxlp* lp = const_cast<moip_cpxlp*>(lp_c);
  if (!env) return kErrNoEnv;
  if (!lp || !lp->loaded) return kErrNoProblem;
  if (!nzcnt_p || !rmatbeg || !surplus_p || (rmatspace > 0 && (!rmatind || !rmatval))) return kErrNullPointer;
  if (begin < 0 || end >= lp->numrows() || begin > end) return kErrIndexRange;
  const int n = lp->n();
  int nz = 0;
  for (int r = begin; r <= end; ++r) {
    const double* row;
    if (r >= lp->ms()) row = lp->objcoef.data() + (size_t)(r - lp->ms()) * n;
    else {
      if (int rc = lp->ensure_dense()) return rc;
      row = lp->a.data() + (size_t)r * n;
    }
    rmatbeg[r - begin] = nz;
    for (int q = 0; q < n; ++q) {
      if (row[q] == 0.0) continue;
      if (nz < rmatspace) {
        rmatind[nz] = q;
        rmatval[nz] = row[q];
      }
      ++nz;
    }
  }
  *surplus_p = rmatspace - nz;
  *nzcnt_p = nz < rmatspace ? nz : rmatspace;
  return nz > rmatspace ? kErrNegSurplus : 0;
}

int CPXgetobjsen(CPXCENVptr env, CPXCLPptr lp) { return (env && lp && lp->loaded) ? lp->cur_sense : 0; }

int CPXchgsense(CPXCENVptr env, CPXLPptr lp, int cnt, const int* indices, const char* sense) {
  if (!env) return kErrNoEnv;
  if (!lp || !lp->loaded) return kErrNoProblem;
  if (cnt > 0 && (!indices || !sense)) return kErrNullPointer;
  // the only change the reference makes: the k objective rows become '<=' rows (MIN) / '>=' rows (MAX)
  // (src/problem.cpp:122-141, :299-330) -- which is what the solver core assumes for them
  const char want = lp->info.sense == MOIP_SENSE_MIN ? 'L' : 'G';
  for (int i = 0; i < cnt; ++i) {
    if (indices[i] < 0 || indices[i] >= lp->numrows()) return kErrIndexRange;
    if (indices[i] < lp->ms() || sense[i] != want) {
      std::fprintf(stderr, "cplex shim: CPXchgsense(row %d, '%c') is outside the supported family (objective rows to '%c')\n",
                   indices[i], sense[i], want);
      return kErrUnsupported;
    }
  }
  return 0;
}

int CPXchgrhs(CPXCENVptr env, CPXLPptr lp, int cnt, const int* indices, const double* values) {
  if (!env) return kErrNoEnv;
  if (!lp || !lp->loaded) return kErrNoProblem;
  if (cnt > 0 && (!indices || !values)) return kErrNullPointer;
  for (int i = 0; i < cnt; ++i) {
    if (indices[i] < 0 || indices[i] >= lp->numrows()) return kErrIndexRange;
    if (indices[i] < lp->ms()) {
      std::fprintf(stderr, "cplex shim: CPXchgrhs on structural row %d is not supported\n", indices[i]);
      return kErrUnsupported;
    }
  }
  for (int i = 0; i < cnt; ++i) lp->objrow_rhs[indices[i] - lp->ms()] = values[i];
  return 0;
}

int CPXgetcolname(CPXCENVptr env, CPXCLPptr lp, char** name, char* namestore, int storespace, int* surplus_p,
                  int begin, int end) {
  if (!env) return kErrNoEnv;
  if (!lp || !lp->loaded) return kErrNoProblem;
  if (!surplus_p) return kErrNullPointer;
  if (begin < 0 || end >= lp->n() || begin > end) return kErrIndexRange;
  int used = 0;
  bool fits = name && namestore;
  char buf[1024];
  for (int j = begin; j <= end; ++j) {
    if (moip_model_colname(lp->model, j, buf, (int)sizeof buf) != MOIP_OK) return kErrBadArg;
    const int len = (int)std::strlen(buf) + 1;
    if (fits && used + len <= storespace) {
      std::memcpy(namestore + used, buf, len);
      name[j - begin] = namestore + used;
    } else {
      fits = false;
    }
    used += len;
  }
  *surplus_p = storespace - used;
  return used > storespace ? kErrNegSurplus : 0;
}

int CPXaddrows(CPXCENVptr env, CPXLPptr lp, int ccnt, int rcnt, int nzcnt, const double* rhs, const char* sense,
               const int* rmatbeg, const int* rmatind, const double* rmatval, char** /*colname*/, char** /*rowname*/) {
  if (!env) return kErrNoEnv;
  if (!lp || !lp->loaded) return kErrNoProblem;
  // supported: the k objective rows of a .mop appended once, as dense rows in objective order (src/problem.cpp:222-321)
  if (!lp->is_mop || lp->rows_added || ccnt != 0 || rcnt != lp->k() || !rmatbeg || !rmatind || !rmatval) {
    std::fprintf(stderr, "cplex shim: CPXaddrows outside the supported family (the k objective rows of a .mop, once)\n");
    return kErrUnsupported;
  }
  const int n = lp->n();
  std::vector<double> row(n);
  for (int j = 0; j < rcnt; ++j) {
    std::fill(row.begin(), row.end(), 0.0);
    const int from = rmatbeg[j], to = (j + 1 < rcnt) ? rmatbeg[j + 1] : nzcnt;
    for (int e = from; e < to; ++e) {
      if (rmatind[e] < 0 || rmatind[e] >= n) return kErrIndexRange;
      row[rmatind[e]] += rmatval[e];
    }
    for (int q = 0; q < n; ++q)
      if (row[q] != lp->objcoef[(size_t)j * n + q]) {
        std::fprintf(stderr, "cplex shim: CPXaddrows row %d differs from objective %d of the model at column %d (%g vs %g)\n",
                     j, j, q, row[q], lp->objcoef[(size_t)j * n + q]);
        return kErrUnsupported;
      }
    if (rhs) lp->objrow_rhs[j] = rhs[j];
    const char want = lp->info.sense == MOIP_SENSE_MIN ? 'L' : 'G';
    if (sense && sense[j] != want) {
      std::fprintf(stderr, "cplex shim: CPXaddrows row %d has sense '%c', expected '%c'\n", j, sense[j], want);
      return kErrUnsupported;
    }
  }
  lp->rows_added = true;
  return 0;
}

int CPXchgobj(CPXCENVptr env, CPXLPptr lp, int cnt, const int* indices, const double* values) {
  if (!env) return kErrNoEnv;
  if (!lp || !lp->loaded) return kErrNoProblem;
  if (cnt > 0 && (!indices || !values)) return kErrNullPointer;
  for (int i = 0; i < cnt; ++i)
    if (indices[i] < 0 || indices[i] >= lp->n()) return kErrIndexRange;
  for (int i = 0; i < cnt; ++i) lp->cur_obj[indices[i]] = values[i];
  lp->match_objective();
  if (lp->cur_obj_index < 0) {
    std::fprintf(stderr, "cplex shim: CPXchgobj set an objective that is none of the model's %d objectives; the GPU path "
                         "optimises those only\n", lp->k());
    return kErrUnsupported;
  }
  return 0;
}

int CPXchgobjsen(CPXCENVptr env, CPXLPptr lp, int maxormin) {
  if (!env) return kErrNoEnv;
  if (!lp || !lp->loaded) return kErrNoProblem;
  if (maxormin != CPX_MIN && maxormin != CPX_MAX) return kErrBadArg;
  lp->cur_sense = maxormin;
  return 0;
}

int CPXsetintparam(CPXENVptr env, int whichparam, int newvalue) {
  if (!env) return kErrNoEnv;
  switch (whichparam) {
    case CPXPARAM_Threads: env->threads = newvalue; return 0;
    case CPXPARAM_Parallel: env->parallel = newvalue; return 0;
    case CPX_PARAM_SCRIND: env->scrind = newvalue; return 0;
    default: return 1013;  // CPXERR_BAD_PARAM_NUM
  }
}

int CPXsetdblparam(CPXENVptr env, int whichparam, double newvalue) {
  if (!env) return kErrNoEnv;
  if (whichparam != CPXPARAM_MIP_Tolerances_MIPGap) return 1013;
  env->mipgap = newvalue;   // recorded only: the solve is exact (gap 0), so a tighter gap changes nothing
  return 0;
}

int CPXmipopt(CPXCENVptr env, CPXLPptr lp) {
  if (!env) return kErrNoEnv;
  if (!lp || !lp->loaded) return kErrNoProblem;
  if (lp->is_mop && !lp->rows_added) {
    std::fprintf(stderr, "cplex shim: CPXmipopt before the objective rows of the .mop were added\n");
    return kErrUnsupported;
  }
  if (lp->cur_obj_index < 0) {
    std::fprintf(stderr, "cplex shim: CPXmipopt without one of the model's objectives set (CPXchgobj)\n");
    lp->solved = false;
    return kErrUnsupported;
  }
  if (lp->cur_sense != (lp->info.sense == MOIP_SENSE_MIN ? CPX_MIN : CPX_MAX)) {
    std::fprintf(stderr, "cplex shim: CPXmipopt against the model's own objective sense is not supported\n");
    lp->solved = false;
    return kErrUnsupported;
  }
  // the reference re-solves the same IP after tightening the MIP gap (src/aira.cpp:417-423, :497-503): same answer
  {
    std::lock_guard<std::mutex> g(g_registry.mu);
    ++g_registry.mipopt_calls;
    if (lp->solved && lp->solved_obj == lp->cur_obj_index && lp->solved_rhs == lp->objrow_rhs) {
      ++g_registry.mipopt_memo;
      return 0;
    }
  }
  if (!lp->ctx) {
    // one context per problem object = per worker thread (src/aira.cpp:561-585); MOIP_B200_DEVICES=G spreads the
    // problem objects of one process over G GPUs in creation order (one host thread per worker, each
    // bound to its own GPU), MOIP_B200_DEVICE picks the first one
    const int first = env_int("MOIP_B200_DEVICE", 0);
    int spread = env_int("MOIP_B200_DEVICES", 1);
    const int have = moip_device_count();
    if (spread < 1) spread = 1;
    if (have > 0 && first + spread > have) spread = have - first > 0 ? have - first : 1;
    const int dev = first + (spread > 1 ? lp->serial % spread : 0);
    const int rc = moip_ctx_create_own_stream(lp->model, dev, &lp->ctx);
    if (rc != MOIP_OK || !lp->ctx) {
      std::fprintf(stderr, "cplex shim: cannot create a B200 solver context on device %d (MOIP error %d); there is no CPU solve\n", dev, rc);
      lp->ctx = nullptr;
      lp->solved = false;
      return kErrSolver;
    }
  }
  const bool start_ok = lp->x_within_bounds();   // CPLEX keeps the incumbent of the previous solve as a MIP start
  std::vector<int32_t> xnew(lp->n());
  int64_t obj = 0;
  int st = 0;
  const int rc = moip_mip_solve(lp->ctx, lp->cur_obj_index, lp->objrow_rhs.data(), start_ok ? lp->x.data() : nullptr,
                                xnew.data(), &obj, &st);
  if (rc != MOIP_OK) {
    std::fprintf(stderr, "cplex shim: the GPU solve failed (MOIP error %d)\n", rc);
    lp->solved = false;
    return kErrSolver;
  }
  lp->solved = true;
  lp->stat = st;
  lp->solved_obj = lp->cur_obj_index;
  lp->solved_rhs = lp->objrow_rhs;
  if (st == MOIP_MIP_INFEASIBLE || st == MOIP_MIP_INFORUNBD) {
    lp->have_x = false;
  } else {
    lp->x.swap(xnew);
    lp->have_x = true;
    lp->objval = (double)obj;
  }
  return 0;
}

int CPXgetstat(CPXCENVptr env, CPXCLPptr lp) { return (env && lp && lp->solved) ? lp->stat : 0; }

int CPXgetobjval(CPXCENVptr env, CPXCLPptr lp, double* objval_p) {
  if (!env) return kErrNoEnv;
  if (!lp) return kErrNoProblem;
  if (!objval_p) return kErrNullPointer;
  if (!lp->solved || !lp->have_x) return kErrNoSoln;
  *objval_p = lp->objval;
  return 0;
}

int CPXgetx(CPXCENVptr env, CPXCLPptr lp, double* x, int begin, int end) {
  if (!env) return kErrNoEnv;
  if (!lp || !lp->loaded) return kErrNoProblem;
  if (!x) return kErrNullPointer;
  if (begin < 0 || end >= lp->n() || begin > end) return kErrIndexRange;
  if (!lp->solved || !lp->have_x) {
    // the reference reads x even after an infeasible solve and ignores this status (src/aira.cpp:489-521); hand
    // it zeros rather than whatever its buffer held
    for (int q = begin; q <= end; ++q) x[q - begin] = 0.0;
    return kErrNoSoln;
  }
  for (int q = begin; q <= end; ++q) x[q - begin] = (double)lp->x[q];
  return 0;
}

}  // extern "C"
