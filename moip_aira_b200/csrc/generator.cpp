// Subproblem generator re-hosted on the C ABI: the sequential ("-t 1") and EPP ("--split") paths of
// the reference's optimise<sense>() (reference src/aira.cpp:538-1884) and the EPP driver
// (reference src/aira.cpp:1886-1990), written as an explicit state machine over the bound vector.
// The inter-thread bound cells of the synergistic mode (src/aira.cpp:923-1086, :1111-1552) are out
// of scope here.
//
// State per worker: rhs[k] (current bounds), hi_seen[]/lo_seen[] (the reference's max[]/min[]
// trackers), misses (its infcnt), last_missed (inflast), level (depth_level), walking (onwalk).
#include <algorithm>
#include <sched.h>

#include <atomic>
#include <chrono>
#include <mutex>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

#include "solver.h"

int cache_find2(moip_ctx* c, moip_cache* s0, moip_cache* s1, int Q, const double* ip, int sense, int* first_match, int* which,
                moip::CacheRecord* rec_out);

namespace moip {

struct GenBackend {
  virtual ~GenBackend() {}
  virtual int solve(const int* perm, int n_obj, const double* rhs, int* result, int* status) = 0;
  virtual int find(const double* rhs, int* hit, int* infeasible, int* result) = 0;
  virtual int insert(const double* rhs, const int* result, int infeasible) = 0;
  // The top-level bound of the LAST stage (objective perm[n_obj-1]) has moved to new_rhs: every non-dominated point
  // beyond it has been recorded.  Only the cooperative workers (CoopBackend below) listen.
  virtual void outer_bound_moved(int /*objective*/, double /*new_rhs*/) {}
  // node budget of the IPs behind the next solve() calls (0 = none): solve() then may return MOIP_ERR_BUDGET
  virtual void set_budget(long long /*nodes*/) {}
};

// A strip whose END can move while it is being solved: an idle worker of the pool takes over the far half of the range
// a busy strip has not reached yet (moip_pool_run_strips_claim).  A strip is only a range of the last objective
// (src/aira.cpp:1895-1916), its generator state is the bound vector and the trackers, so handing off [mid, stop) is just
// "this strip now stops at mid" + a new strip that starts at mid; whatever the owner had already covered beyond mid when
// it notices is solved twice, never lost.
struct StripDyn {
  std::atomic<double> stop{0.0};   // current end of the range (as in moip_worker::split_stop)
  std::atomic<double> pos{0.0};    // bound of the last objective the owner is working under (its progress)
  std::atomic<int> state{0};       // 0 free, 1 being solved, 2 done
};

namespace {

// `max[d]-1` / `min[d]+1` in the reference are int expressions whose trackers may sit at
// (int)-CPX_INFBOUND = INT_MIN / (int)CPX_INFBOUND = INT_MAX; the front only comes out right when
// the two's-complement wrap the reference gets in practice is reproduced.
inline int wrap32(int64_t v) { return (int32_t)(uint32_t)(uint64_t)v; }

}  // namespace

// first_budget > 0 (boxes only): the cold first subproblem of a box is looked up in the stores first and, on a miss, solved
// under this node budget; MOIP_ERR_BUDGET then means "nothing done, try this box again later".
int run_worker(GenBackend& be, int k, int sense, const moip_worker& w, int64_t* n_iter, int64_t* n_hit, StripDyn* dyn = nullptr,
               long long first_budget = 0) {
  const bool is_min = sense == MOIP_SENSE_MIN;
  const double free_rhs = is_min ? kInf : -kInf;
  const int* perm = w.perm;
  const int n_obj = w.n_obj;
  const bool split = w.split != 0;
  const double split_start = w.split_start;
  double stop_adj = 0.0;                             // the reference moves the end by one once the first point is known (:653-657)
  std::vector<double> rhs(k, free_rhs);
  std::vector<int> res(k, 0), hi_seen(k, 0), lo_seen(k, 0);
  int status = 0, rc;
  const int last = perm[n_obj - 1];
  // window on the objective of the innermost sweeps (moip_worker::window): its "free" bound is the window's near edge,
  // and subproblems bounded beyond the far edge count as infeasible without being solved (or recorded)
  const bool windowed = w.window != 0 && n_obj >= 3;
  const int wobj = perm[1];
  auto free_of = [&](int j) { return (windowed && j == wobj) ? w.win_start : free_rhs; };
  auto beyond_window = [&]() { return windowed && (is_min ? rhs[wobj] < w.win_stop : rhs[wobj] > w.win_stop); };
  rhs[wobj] = free_of(wobj);
  if (split) rhs[last] = split_start;                                       // :607
  bool root_known = false, root_infeasible = false;
  if (windowed) {       // a box starts inside the front: its first subproblem may already be answered by a neighbour's records
    int hit = 0, inf0 = 0;
    if ((rc = be.find(rhs.data(), &hit, &inf0, res.data()))) return rc;
    if (hit) { root_known = true; root_infeasible = inf0 != 0; }
  }
  if (!root_known) {
    if (windowed && first_budget > 0) be.set_budget(first_budget);
    rc = be.solve(perm, n_obj, rhs.data(), res.data(), &status);              // :614
    if (windowed && first_budget > 0) be.set_budget(0);
    if (rc) return rc;
    root_infeasible = status == MOIP_MIP_INFEASIBLE;
    if ((rc = be.insert(rhs.data(), res.data(), root_infeasible ? 1 : 0))) return rc;   // :644-651
  }
  if (root_infeasible) return MOIP_OK;   // nothing lies inside these bounds (reference: trackers undefined)
  if (split) stop_adj = is_min ? -1.0 : 1.0;                                // :653-657
  hi_seen = res; lo_seen = res;                                             // :693-697

  auto tighten = [&](int d) {   // move the bound of objective d just past everything seen, reset its tracker
    if (is_min) { rhs[d] = (double)wrap32((int64_t)hi_seen[d] - 1); hi_seen[d] = INT_MIN; }
    else { rhs[d] = (double)wrap32((int64_t)lo_seen[d] + 1); lo_seen[d] = INT_MAX; }
  };
  // NB: the strip tests index rhs by position n_obj-1, not perm (src/aira.cpp:781, :882); EPP workers
  // use the identity permutation so both agree.
  auto crossed_stop = [&]() {
    const double split_stop = (dyn ? dyn->stop.load(std::memory_order_acquire) : w.split_stop) + stop_adj;
    return is_min ? rhs[n_obj - 1] < split_stop : rhs[n_obj - 1] > split_stop;
  };
  auto publish_progress = [&](int active) {          // the strip's sweep proper: the last stage moves the bound of the last objective
    if (dyn && split && active == n_obj - 1) dyn->pos.store(rhs[n_obj - 1], std::memory_order_release);
  };

  for (int active = 1; active < n_obj; ++active) {                          // :723 objective_counter
    const int objective = perm[active];
    int level = 1, depth = perm[level];
    bool walking = false, last_missed = false;
    int misses = 0;
    for (int jp = 1; jp < k; ++jp) rhs[perm[jp]] = free_of(perm[jp]);       // :733-756
    if (split) rhs[last] = split_start;                                     // :757-759
    rhs[objective] = is_min ? (double)wrap32((int64_t)hi_seen[objective] - 1)
                            : (double)wrap32((int64_t)lo_seen[objective] + 1);   // :761-777
    if (split && crossed_stop()) break;                                     // :778-801
    publish_progress(active);
    if (!split && active == n_obj - 1) be.outer_bound_moved(objective, rhs[objective]);
    hi_seen[objective] = INT_MIN;                                           // :802-803
    lo_seen[objective] = INT_MAX;
    long long spins = 0;
    while (misses < active) {                                               // :804
      if (++spins % 200000 == 0 && std::getenv("MOIP_DEBUG_GEN"))
        std::fprintf(stderr, "moip_b200: generator worker %d (n_obj %d, strip [%g, %g)) stage %d: %lld iterations, misses %d level %d walking %d rhs [%g %g %g %g] last result [%d %d %d %d] hi_seen [%d %d %d %d] lo_seen [%d %d %d %d]\n",
                     w.id, n_obj, split_start, w.split_stop, active, spins, misses, level, (int)walking, rhs[0], k > 1 ? rhs[1] : 0.0, k > 2 ? rhs[2] : 0.0,
                     k > 3 ? rhs[3] : 0.0, res[0], k > 1 ? res[1] : 0, k > 2 ? res[2] : 0, k > 3 ? res[3] : 0, hi_seen[0], k > 1 ? hi_seen[1] : 0,
                     k > 2 ? hi_seen[2] : 0, k > 3 ? hi_seen[3] : 0, lo_seen[0], k > 1 ? lo_seen[1] : 0, k > 2 ? lo_seen[2] : 0, k > 3 ? lo_seen[3] : 0);
      int hit = 0, infeasible = 0;
      if (beyond_window()) { hit = 1; infeasible = 1; }
      else if ((rc = be.find(rhs.data(), &hit, &infeasible, res.data()))) return rc;   // :816-827
      if (n_iter) ++*n_iter;
      if (hit) { if (n_hit) ++*n_hit; }
      else {
        if ((rc = be.solve(perm, n_obj, rhs.data(), res.data(), &status))) return rc;   // :835
        infeasible = (status == MOIP_MIP_INFEASIBLE || status == MOIP_MIP_INFORUNBD) ? 1 : 0;
        if ((rc = be.insert(rhs.data(), res.data(), infeasible))) return rc;            // :842-850
      }
      if (split) {                                                          // :877-922
        if (!infeasible) {
          if (misses == n_obj - 2 && crossed_stop()) infeasible = 1;
          for (int j = 0; j < k; ++j) { if (res[j] > hi_seen[j]) hi_seen[j] = res[j]; if (res[j] < lo_seen[j]) lo_seen[j] = res[j]; }
        }
      } else if (!infeasible) {                                             // :1087-1107
        for (int j = 0; j < k; ++j) { if (res[j] > hi_seen[j]) hi_seen[j] = res[j]; if (res[j] < lo_seen[j]) lo_seen[j] = res[j]; }
      }
      if (infeasible) { ++misses; last_missed = true; } else { misses = 0; last_missed = false; }
      // next bound vector (:1575-1832)
      if (infeasible && misses == active - 1) {
        // Boxes: a sweep that saw no point at all leaves the tracker of the stage's objective at its reset value, and the
        // reference's tighten() would wrap the bound to INT_MAX / INT_MIN -- "free" -- and walk the whole front beyond the
        // strip again (harmless duplicates there).  Inside a window that walk starts with the window's tightest subproblem
        // under no other bound, the one case plain B&B explodes on; nothing is left in the box, so the stage ends here.
        if (windowed && (is_min ? hi_seen[objective] == INT_MIN : lo_seen[objective] == INT_MAX)) break;
        for (int j = 0; j < k; ++j) rhs[j] = free_of(j);                    // :1586-1599
        if (split) rhs[n_obj - 1] = split_start;                            // :1649-1651
        tighten(objective);                                                 // :1655-1673
        publish_progress(active);
        if (!split && active == n_obj - 1) be.outer_bound_moved(objective, rhs[objective]);
        level = 1; depth = perm[level]; walking = false;
      } else if (last_missed && misses != active) {
        rhs[depth] = free_of(depth);                                        // :1722-1728
        depth = perm[++level];                                              // :1730-1731
        tighten(depth);                                                     // :1758-1760 / :1778-1780
        walking = true;
      } else if (!walking && misses != 1) {
        tighten(depth);                                                     // :1797-1798 / :1806-1807
      } else if (walking && misses != 1) {
        level = 1; depth = perm[level];
        tighten(depth);                                                     // :1810-1831
        walking = false;
      }
    }
  }
  return MOIP_OK;
}

namespace {

// GPU backend: solve = lexicographic B&B on the device, find = K3 over (infeasibles, solutions)
struct GpuBackend : GenBackend {
  moip_ctx* c;
  moip_cache* infeasibles;
  moip_cache* sols;
  int sense;
  int solve(const int* perm, int n_obj, const double* rhs, int* result, int* status) override {
    return c->lex_solve(perm, n_obj, rhs, result, status);
  }
  void set_budget(long long nodes) override { c->ip_node_budget = nodes; }
  int find(const double* rhs, int* hit, int* infeasible, int* result) override {
    int idx = -1, which = -1;
    CacheRecord r{};
    int rc = cache_find2(c, infeasibles, sols, 1, rhs, sense, &idx, &which, &r);
    if (rc) return rc;
    *hit = idx >= 0;
    if (idx >= 0) {
      *infeasible = r.infeasible;
      for (int j = 0; j < c->dm.k; ++j) result[j] = r.result[j];
    }
    return MOIP_OK;
  }
  int insert(const double* rhs, const int* result, int infeasible) override {
    return moip_cache_insert(infeasible ? infeasibles : sols, rhs, result, infeasible);
  }
};

struct CallbackBackend : GenBackend {
  moip_solve_fn solve_fn;
  moip_find_cb find_fn;
  moip_insert_cb insert_fn;
  void* user;
  int solve(const int* perm, int n_obj, const double* rhs, int* result, int* status) override {
    return solve_fn(user, perm, n_obj, rhs, result, status);
  }
  int find(const double* rhs, int* hit, int* infeasible, int* result) override {
    *hit = find_fn(user, rhs, infeasible, result);
    return *hit < 0 ? MOIP_ERR_ARG : MOIP_OK;
  }
  int insert(const double* rhs, const int* result, int infeasible) override {
    return insert_fn(user, rhs, result, infeasible);
  }
};

}  // namespace
}  // namespace moip

using namespace moip;

static int optimise_strip(moip_ctx* c, const moip_worker* w, moip_cache* all, moip_cache* infeasibles, moip::StripDyn* dyn,
                          long long first_budget = 0);

extern "C" int moip_optimise(moip_ctx* c, const moip_worker* w, moip_cache* all, moip_cache* infeasibles) {
  return optimise_strip(c, w, all, infeasibles, nullptr);
}

static int optimise_strip(moip_ctx* c, const moip_worker* w, moip_cache* all, moip_cache* infeasibles, moip::StripDyn* dyn,
                          long long first_budget) {
  if (!c || !w || !all || !infeasibles || w->n_obj < 1 || w->n_obj > c->dm.k) return MOIP_ERR_ARG;
  const int sense = c->model->M.sense;
  moip_cache* local = nullptr;                       // `Solutions s(p.objcnt)` (src/aira.cpp:587)
  int rc = moip_cache_create(c, &local);
  if (rc) return rc;
  GpuBackend be;
  be.c = c; be.infeasibles = infeasibles; be.sols = w->split ? all : local; be.sense = sense;
  rc = run_worker(be, c->dm.k, sense, *w, nullptr, nullptr, dyn, first_budget);
  if (!rc && moip_cache_size(local) > 0) {           // (EPP strips write straight into `all`: nothing to splice)
    moip_cache_sort_unique(local, nullptr, 0);       // :1877
    rc = moip_cache_merge(all, local);               // :1879
  }
  moip_cache_destroy(local);
  return rc;
}

extern "C" int moip_optimise_with(int k, int sense, const moip_worker* w, moip_solve_fn solve, moip_find_cb find,
                                  moip_insert_cb insert, void* user, int64_t* n_iterations, int64_t* n_hits) {
  if (!w || !solve || !find || !insert || k < 1 || k > MOIP_MAX_OBJ) return MOIP_ERR_ARG;
  CallbackBackend be;
  be.solve_fn = solve; be.find_fn = find; be.insert_fn = insert; be.user = user;
  if (n_iterations) *n_iterations = 0;
  if (n_hits) *n_hits = 0;
  return run_worker(be, k, sense, *w, n_iterations, n_hits);
}

// ---------------------------------------------------------------------------------------------------------------
// Cooperative ("synergistic") workers: a race-free re-hosting of the idea behind the reference's bound-sharing
// threads (src/aira.cpp:923-1086, :1111-1552).  Worker w runs the sequential generator above with a
// permutation whose LAST objective o_w is its own; the W workers own W different objectives.  Whenever the top-level
// bound of w's last stage moves to b, every non-dominated point with f_{o_w} beyond b is on record (the inner
// (k-1)-objective front under the previous bound is complete and b = (extreme f_{o_w} over it) -+ 1), so w
// publishes b as its LIMIT -- one monotone 64-bit atomic per objective.  Every other worker intersects each of its
// subproblems with f_{o_w} <= b (MIN; >= for MAX) before the cache scan / solve / insert: it stops looking where w
// has already looked.  Why this is sound without any waiting or locking:
//   * a limit only ever cuts off a region whose non-dominated points are recorded, and the limits are bounds on
//     objective VALUES, so the non-dominated points of a restricted set are exactly the non-dominated points of the
//     full set that lie in it; a worker run against shrinking limits enumerates a superset of what the final
//     limits require;
//   * reading a stale (looser) limit only means redundant work, never a lost point; limits are written with
//     fetch-min / fetch-max, so they are monotone whatever the interleaving;
//   * cache records are stored under the restricted bound vector actually solved, so Solutions::find's relaxation
//     test stays exact; infeasible records are valid for every permutation and may be shared;
//   * a worker that runs to completion has, together with the limits it honoured, covered everything: it publishes
//     "done" and the others answer their remaining subproblems as infeasible without solving them.
// The result is the union of the workers' points (all of them lexicographic optima, hence non-dominated).
namespace moip {
namespace {

struct CoopShared {
  int k = 0;
  bool is_min = true;
  bool owned[MOIP_MAX_OBJ] = {false, false, false, false};
  std::atomic<long long> limit[MOIP_MAX_OBJ];
  static constexpr long long kFree = LLONG_MAX, kDone = LLONG_MIN;   // MIN form; MAX workers publish negated values
  void init(int k_, bool is_min_) {
    k = k_; is_min = is_min_;
    for (int j = 0; j < MOIP_MAX_OBJ; ++j) { limit[j].store(kFree); owned[j] = false; }
  }
  // limits are kept in "min form" (value for MIN models, -value for MAX models) so that tighter = smaller
  void publish(int obj, long long v_minform) {
    long long cur = limit[obj].load(std::memory_order_relaxed);
    while (v_minform < cur && !limit[obj].compare_exchange_weak(cur, v_minform, std::memory_order_release,
                                                               std::memory_order_relaxed)) {}
  }
};

struct CoopBackend : GenBackend {
  GenBackend* inner = nullptr;
  CoopShared* sh = nullptr;
  int own = -1;
  int64_t solves = 0, skipped = 0;
  double raw[MOIP_MAX_OBJ], clamped[MOIP_MAX_OBJ];
  bool have = false, dead = false;

  // intersect the generator's bound vector with the partners' published limits
  void restrict_to_limits(const double* rhs) {
    const int k = sh->k;
    if (have && std::memcmp(raw, rhs, sizeof(double) * k) == 0) return;   // find -> solve -> insert of one subproblem
    dead = false;
    for (int j = 0; j < k; ++j) {
      raw[j] = clamped[j] = rhs[j];
      if (j == own || !sh->owned[j]) continue;
      const long long L = sh->limit[j].load(std::memory_order_acquire);
      if (L == CoopShared::kDone) { dead = true; continue; }
      if (L == CoopShared::kFree) continue;
      const double lim = sh->is_min ? (double)L : -(double)L;
      if (sh->is_min ? lim < clamped[j] : lim > clamped[j]) clamped[j] = lim;
    }
    have = true;
  }
  int solve(const int* perm, int n_obj, const double* rhs, int* result, int* status) override {
    restrict_to_limits(rhs);
    if (dead) { ++skipped; *status = MOIP_MIP_INFEASIBLE; return MOIP_OK; }
    ++solves;
    return inner->solve(perm, n_obj, clamped, result, status);
  }
  int find(const double* rhs, int* hit, int* infeasible, int* result) override {
    have = false;                                  // a new subproblem: read the limits afresh
    restrict_to_limits(rhs);
    if (dead) { *hit = 1; *infeasible = 1; ++skipped; return MOIP_OK; }
    return inner->find(clamped, hit, infeasible, result);
  }
  int insert(const double* rhs, const int* result, int infeasible) override {
    restrict_to_limits(rhs);
    if (dead) return MOIP_OK;                      // nothing was solved
    return inner->insert(clamped, result, infeasible);
  }
  void outer_bound_moved(int objective, double new_rhs) override {
    if (objective != own || std::fabs(new_rhs) >= 2147483647.0) return;   // +-1e20 and the INT_MIN-1 wrap: no claim
    sh->publish(own, sh->is_min ? (long long)new_rhs : -(long long)new_rhs);
  }
  void finished() { sh->limit[own].store(CoopShared::kDone, std::memory_order_release); }
};

// objective owned by a worker = last entry of its permutation; owners must be distinct
int coop_check(int k, int n_workers, const moip_worker* ws) {
  if (n_workers < 1 || n_workers > k) return MOIP_ERR_ARG;
  bool seen[MOIP_MAX_OBJ] = {false, false, false, false};
  for (int i = 0; i < n_workers; ++i) {
    if (ws[i].n_obj != k || ws[i].split) return MOIP_ERR_ARG;
    bool in_perm[MOIP_MAX_OBJ] = {false, false, false, false};
    for (int j = 0; j < k; ++j) {
      if (ws[i].perm[j] < 0 || ws[i].perm[j] >= k || in_perm[ws[i].perm[j]]) return MOIP_ERR_ARG;
      in_perm[ws[i].perm[j]] = true;
    }
    const int o = ws[i].perm[k - 1];
    if (seen[o]) return MOIP_ERR_ARG;
    seen[o] = true;
  }
  return MOIP_OK;
}

}  // namespace
}  // namespace moip

// Limits as a handle of their own: one worker per PROCESS (one rank per GPU) runs against a local handle that a host
// thread keeps in step with the other ranks (a few int64 values through the job's store / one small collective);
// publishing is fetch-min, so merging remote values in any order and any number of times is safe.
struct moip_coop {
  moip::CoopShared sh;
};

extern "C" int moip_coop_create(int k, int sense, int owned_mask, moip_coop** out) {
  if (!out || k < 1 || k > MOIP_MAX_OBJ || (sense != MOIP_SENSE_MIN && sense != MOIP_SENSE_MAX)) return MOIP_ERR_ARG;
  moip_coop* h = new moip_coop();
  h->sh.init(k, sense == MOIP_SENSE_MIN);
  for (int j = 0; j < k; ++j) h->sh.owned[j] = (owned_mask >> j) & 1;
  *out = h;
  return MOIP_OK;
}
extern "C" void moip_coop_destroy(moip_coop* h) { delete h; }

// value is in the model's own terms (the right-hand side of f_obj <= value for MIN, >= for MAX); done != 0: the owner is through
extern "C" int moip_coop_publish(moip_coop* h, int obj, long long value, int done) {
  if (!h || obj < 0 || obj >= h->sh.k) return MOIP_ERR_ARG;
  if (done) h->sh.limit[obj].store(CoopShared::kDone, std::memory_order_release);
  else h->sh.publish(obj, h->sh.is_min ? value : -value);
  return MOIP_OK;
}
// state: 0 = no limit yet, 1 = *value holds the limit, 2 = the owner is through
extern "C" int moip_coop_read(const moip_coop* h, int obj, long long* value, int* state) {
  if (!h || obj < 0 || obj >= h->sh.k || !state) return MOIP_ERR_ARG;
  const long long L = h->sh.limit[obj].load(std::memory_order_acquire);
  *state = L == CoopShared::kDone ? 2 : (L == CoopShared::kFree ? 0 : 1);
  if (value) *value = *state == 1 ? (h->sh.is_min ? L : -L) : 0;
  return MOIP_OK;
}

// one cooperative worker against a limits handle: the product path (solver context) ...
extern "C" int moip_coop_optimise(moip_ctx* c, const moip_worker* w, moip_coop* shared, moip_cache* all, moip_cache* infeasibles) {
  if (!c || !w || !shared || !all || !infeasibles || shared->sh.k != c->dm.k) return MOIP_ERR_ARG;
  if (int rc0 = coop_check(c->dm.k, 1, w)) return rc0;
  const int sense = c->model->M.sense, k = c->dm.k;
  moip_cache* local = nullptr;
  int rc = moip_cache_create(c, &local);
  if (rc) return rc;
  GpuBackend be;
  be.c = c; be.infeasibles = infeasibles; be.sols = local; be.sense = sense;
  CoopBackend co;
  co.inner = &be; co.sh = &shared->sh; co.own = w->perm[k - 1];
  rc = run_worker(co, k, sense, *w, nullptr, nullptr);
  if (!rc) {
    co.finished();
    moip_cache_sort_unique(local, nullptr, 0);
    rc = moip_cache_merge(all, local);
  }
  moip_cache_destroy(local);
  return rc;
}
// ... and the host-logic hook
extern "C" int moip_coop_optimise_one_with(int k, int sense, const moip_worker* w, moip_coop* shared, moip_solve_fn solve,
                                           moip_find_cb find, moip_insert_cb insert, void* user, int64_t* n_solves,
                                           int64_t* n_skipped) {
  if (!w || !shared || !solve || !find || !insert || k < 1 || k > MOIP_MAX_OBJ || shared->sh.k != k) return MOIP_ERR_ARG;
  if (int rc0 = coop_check(k, 1, w)) return rc0;
  CallbackBackend cb;
  cb.solve_fn = solve; cb.find_fn = find; cb.insert_fn = insert; cb.user = user;
  CoopBackend co;
  co.inner = &cb; co.sh = &shared->sh; co.own = w->perm[k - 1];
  int rc = run_worker(co, k, sense, *w, nullptr, nullptr);
  if (!rc) co.finished();
  if (n_solves) *n_solves = co.solves;
  if (n_skipped) *n_skipped = co.skipped;
  return rc;
}

extern "C" int moip_coop_workers(int k, int n_workers, moip_worker* out) {
  if (!out || k < 1 || k > MOIP_MAX_OBJ || n_workers < 1 || n_workers > k) return MOIP_ERR_ARG;
  for (int i = 0; i < n_workers; ++i) {            // rotations of the identity: worker i owns objective (k-1-i) mod k
    out[i] = moip_worker{};
    out[i].id = i; out[i].n_obj = k; out[i].split = 0;
    for (int j = 0; j < k; ++j) out[i].perm[j] = ((j - i) % k + k) % k;
  }
  return MOIP_OK;
}

extern "C" int moip_coop_optimise_with(int k, int sense, int n_workers, const moip_worker* workers, moip_solve_fn solve,
                                       moip_find_cb find, moip_insert_cb insert, void* const* users, int64_t* n_solves,
                                       int64_t* n_skipped) {
  if (!workers || !solve || !find || !insert || k < 1 || k > MOIP_MAX_OBJ) return MOIP_ERR_ARG;
  if (int rc = coop_check(k, n_workers, workers)) return rc;
  CoopShared sh;
  sh.init(k, sense == MOIP_SENSE_MIN);
  for (int i = 0; i < n_workers; ++i) sh.owned[workers[i].perm[k - 1]] = n_workers > 1;
  std::vector<CallbackBackend> cb(n_workers);
  std::vector<CoopBackend> co(n_workers);
  std::vector<int> rcs(n_workers, MOIP_OK);
  auto work = [&](int i) {
    cb[i].solve_fn = solve; cb[i].find_fn = find; cb[i].insert_fn = insert; cb[i].user = users ? users[i] : nullptr;
    co[i].inner = &cb[i]; co[i].sh = &sh; co[i].own = workers[i].perm[k - 1];
    rcs[i] = run_worker(co[i], k, sense, workers[i], nullptr, nullptr);
    if (!rcs[i]) co[i].finished();
  };
  std::vector<std::thread> th;
  for (int i = 1; i < n_workers; ++i) th.emplace_back(work, i);
  work(0);
  for (auto& t : th) t.join();
  for (int i = 0; i < n_workers; ++i) {
    if (n_solves) n_solves[i] = co[i].solves;
    if (n_skipped) n_skipped[i] = co[i].skipped;
    if (rcs[i]) return rcs[i];
  }
  return MOIP_OK;
}

extern "C" int moip_split_strips(int sense, int biggest, int smallest, int num_threads, int split_normal,
                                 double* start_stop) {
  // the quantile table is data of the reference (src/aira.cpp:55-69)
  static const double nv[13][13] = {
      {0}, {0, 1}, {0, 0.5, 1}, {0, 0.356, 0.644, 1}, {0, 0.275, 0.5, 0.725, 1},
      {0, 0.219, 0.416, 0.584, 0.781, 1}, {0, 0.178, 0.256, 0.5, 0.644, 0.822, 1},
      {0, 0.144, 0.311, 0.44, 0.56, 0.689, 0.856, 1}, {0, 0.117, 0.275, 0.394, 0.5, 0.606, 0.725, 0.883, 1},
      {0, 0.093, 0.245, 0.356, 0.453, 0.547, 0.644, 0.755, 0.907, 1},
      {0, 0.073, 0.219, 0.325, 0.416, 0.5, 0.584, 0.675, 0.781, 0.927, 1},
      {0, 0.055, 0.197, 0.298, 0.384, 0.462, 0.538, 0.616, 0.702, 0.803, 0.945, 1},
      {0, 0.039, 0.178, 0.275, 0.356, 0.430, 0.5, 0.570, 0.644, 0.725, 0.822, 0.961, 1}};
  if (!start_stop || num_threads < 1) return MOIP_ERR_ARG;
  if (split_normal && num_threads > 12) return MOIP_ERR_ARG;               // :199-203
  const bool is_min = sense == MOIP_SENSE_MIN;
  const double start_point = is_min ? (double)biggest : (double)smallest;  // :1888-1894
  const double stop_point = is_min ? (double)smallest : (double)biggest;
  double cur = start_point;
  const double step = (stop_point - start_point) / num_threads;           // :1897
  for (int t = 0; t < num_threads; ++t) {
    if (split_normal) {                                                    // :1900-1911
      double a, b;
      if (is_min) { const double gap = start_point - stop_point; b = nv[num_threads][t] * gap + stop_point; a = nv[num_threads][t + 1] * gap + stop_point; }
      else { const double gap = stop_point - start_point; a = nv[num_threads][t] * gap + start_point; b = nv[num_threads][t + 1] * gap + start_point; }
      start_stop[2 * t] = a; start_stop[2 * t + 1] = b;
    } else {                                                               // :1913-1915
      start_stop[2 * t] = cur; start_stop[2 * t + 1] = cur + step;
      cur += step;
    }
  }
  return MOIP_OK;
}

namespace {

// split_setup / split_optimise (src/aira.cpp:1886-1990) with the strips of one level solved one
// after another on this context, sharing `here` and `infeasibles` like the reference's threads do.
int epp_level(moip_ctx* c, int n_obj, int num_threads, int split_normal, std::vector<std::vector<int>>& sols) {
  const int k = c->dm.k, sense = c->model->M.sense;
  const bool is_min = sense == MOIP_SENSE_MIN;
  std::vector<double> free_rhs(k, is_min ? kInf : -kInf);
  std::vector<int> res(k, 0);
  int st = 0, rc;
  if (n_obj == 1) {                                                        // :1946-1949
    if ((rc = c->get_limit(0, sense, free_rhs.data(), res.data(), &st))) return rc;
    if (st != MOIP_MIP_INFEASIBLE) sols.push_back(res);
    return MOIP_OK;
  }
  std::vector<std::vector<int>> lower;
  if ((rc = epp_level(c, n_obj - 1, num_threads, split_normal, lower))) return rc;
  if (lower.empty()) return MOIP_OK;                                       // infeasible model
  if ((rc = c->get_limit(n_obj - 1, sense, free_rhs.data(), res.data(), &st))) return rc;   // :1959 / :1971
  if (st == MOIP_MIP_INFEASIBLE) return MOIP_OK;
  int biggest, smallest;
  if (is_min) {                                                            // :1958-1969
    smallest = res[n_obj - 1]; biggest = INT_MIN;
    for (auto& s : lower) biggest = std::max(biggest, s[n_obj - 1]);
    if (biggest == smallest) biggest = INT_MAX;
  } else {                                                                 // :1970-1982
    biggest = res[n_obj - 1]; smallest = INT_MAX;
    for (auto& s : lower) smallest = std::min(smallest, s[n_obj - 1]);
    if (biggest == smallest) smallest = INT_MIN;
  }
  std::vector<double> ss(2 * (size_t)num_threads);
  if ((rc = moip_split_strips(sense, biggest, smallest, num_threads, split_normal, ss.data()))) return rc;
  moip_cache *here = nullptr, *infeasibles = nullptr;
  if ((rc = moip_cache_create(c, &here))) return rc;
  if ((rc = moip_cache_create(c, &infeasibles))) { moip_cache_destroy(here); return rc; }
  for (int t = 0; t < num_threads && !rc; ++t) {                           // :1899-1933
    moip_worker w{};
    w.id = t; w.n_obj = n_obj; w.split = 1;
    for (int i = 0; i < k; ++i) w.perm[i] = i;                             // thread.cpp:124-133
    w.split_start = ss[2 * t]; w.split_stop = ss[2 * t + 1];
    rc = moip_optimise(c, &w, here, infeasibles);
  }
  if (!rc)
    for (auto& r : here->host) if (!r.infeasible) sols.emplace_back(r.result, r.result + k);   // :1934-1942
  moip_cache_destroy(here);
  moip_cache_destroy(infeasibles);
  return rc;
}

}  // namespace

// ------------------------------------------------------------------------------------ worker pool
// The reference runs one std::thread per strip / worker, each with its own CPLEX environment
// (src/aira.cpp:297-308, :1920-1933).  Here the same host threads each own a solver context (stream,
// node pool, staging buffers, caches) on ONE device: a single B&B round keeps only a fraction of the 148 SMs
// busy, concurrent workers fill the rest.  Strips are dealt dynamically; every worker keeps its own
// `here` / `infeasibles` stores (the reference shares them under a mutex, which only changes IP counts).
struct moip_pool {
  moip_model* model = nullptr;
  int device = 0;
  std::vector<moip_ctx*> ctx;
  std::vector<cudaStream_t> streams;
  int max_workers = 0;                       // 0 = every context may draw strips (moip_pool_set_max_workers)
  long long deferrals = 0;                   // boxes postponed at their node budget (run_boxes)
  // the shared stores of the EPP level being solved (nullptr between runs): what the knowledge exchange between the
  // pools of several ranks reads and feeds (moip_pool_export_records / moip_pool_import_records)
  std::mutex run_mu;
  moip_cache* run_here = nullptr;
  moip_cache* run_inf = nullptr;
  size_t exp_cursor[2] = {0, 0};
  int64_t exported = 0, imported = 0, stolen = 0;
  cudaStream_t xstream = nullptr;            // imported records go to the device on this stream (created on first use)
};

extern "C" int moip_ctx_set_sync_mode(moip_ctx* c, int blocking);

extern "C" int moip_pool_create(moip_model* m, int device, int workers, moip_pool** out) {
  if (!m || !out || workers < 1 || workers > 64) return MOIP_ERR_ARG;
  moip_pool* p = new moip_pool();
  p->model = m; p->device = device;
  // MOIP_SYNC=auto (default): the workers sleep on a blocking event instead of spinning when, together with the pools of
  // the other ranks on this host (LOCAL_WORLD_SIZE, set by torchrun), they outnumber the cores
  bool block = false;
  {
    const char* sm = std::getenv("MOIP_SYNC");
    if (!sm || std::strcmp(sm, "auto") == 0) {
      const char* lw = std::getenv("LOCAL_WORLD_SIZE");
      const long ranks = lw ? std::max(1, std::atoi(lw)) : 1;
      long cores = std::max(1u, std::thread::hardware_concurrency());
      cpu_set_t cs;                                   // the cores this process may use (taskset / cpusets), not the machine's
      if (sched_getaffinity(0, sizeof(cs), &cs) == 0 && CPU_COUNT(&cs) > 0) cores = CPU_COUNT(&cs);
      block = (long)workers * ranks + 2 * ranks > cores;
    } else block = std::strcmp(sm, "block") == 0;
  }
  for (int w = 0; w < workers; ++w) {
    cudaStream_t st = nullptr;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
      std::fprintf(stderr, "moip_b200: cannot create a stream on device %d (this library has no CPU fallback)\n", device);
      moip_pool_destroy(p);
      return MOIP_ERR_CUDA;
    }
    p->streams.push_back(st);
    moip_ctx* c = nullptr;
    int rc = moip_ctx_create(m, device, st, &c);
    if (rc) { moip_pool_destroy(p); return rc; }
    moip_ctx_set_sync_mode(c, block ? 1 : 0);
    p->ctx.push_back(c);
  }
  *out = p;
  return MOIP_OK;
}

extern "C" void moip_pool_destroy(moip_pool* p) {
  if (!p) return;
  for (moip_ctx* c : p->ctx) moip_ctx_destroy(c);
  cudaSetDevice(p->device);
  if (p->xstream) cudaStreamDestroy(p->xstream);
  for (cudaStream_t st : p->streams) cudaStreamDestroy(st);
  delete p;
}

extern "C" int moip_pool_workers(const moip_pool* p) { return p ? (int)p->ctx.size() : -1; }

extern "C" int moip_pool_set_max_workers(moip_pool* p, int max_workers) {
  if (!p || max_workers < 0) return MOIP_ERR_ARG;
  p->max_workers = max_workers;
  return MOIP_OK;
}

// ---- knowledge exchange between the pools of several ranks (one rank per GPU).  A cache record states a fact about the
// model -- "the lexicographic optimum under the bounds ip is result" / "nothing lies inside ip" (src/result.h:10-20) -- that
// holds whoever computed it, as long as permutation and stage count agree (all strips of an EPP level use the identity
// permutation and the same n_obj, src/thread.cpp:124-133).  The reference's strip threads share `here` / `infeasibles`
// inside one process (src/aira.cpp:1918-1933); this is the same sharing across processes.  Both calls may come from any
// thread while moip_pool_run_strips* is running; between runs there is nothing to exchange.
extern "C" int moip_pool_export_records(moip_pool* p, int cap, double* ip, int* result, int* infeasible, int* n_out) {
  if (!p || !n_out || cap < 0 || (cap > 0 && (!ip || !result || !infeasible))) return MOIP_ERR_ARG;
  *n_out = 0;
  std::lock_guard<std::mutex> rl(p->run_mu);
  if (!p->run_here || !p->run_inf) return MOIP_OK;
  const int k = p->run_here->k;
  int n = 0;
  moip_cache* stores[2] = {p->run_inf, p->run_here};        // infeasible records first: they prune the most
  for (int w = 0; w < 2 && n < cap; ++w) {
    moip_cache* s = stores[w];
    std::lock_guard<std::mutex> lk(s->mu);
    size_t& cur = p->exp_cursor[w];
    for (; cur < s->host.size() && n < cap; ++cur) {
      const CacheRecord& r = s->host[cur];
      if (r.pad[0]) continue;                               // came from another rank
      for (int j = 0; j < k; ++j) { ip[(size_t)n * k + j] = r.ip[j]; result[(size_t)n * k + j] = r.result[j]; }
      infeasible[n] = r.infeasible;
      ++n;
    }
  }
  p->exported += n;
  *n_out = n;
  return MOIP_OK;
}

extern "C" int moip_pool_import_records(moip_pool* p, int n, const double* ip, const int* result, const int* infeasible) {
  if (!p || n < 0 || (n > 0 && (!ip || !result || !infeasible))) return MOIP_ERR_ARG;
  std::lock_guard<std::mutex> rl(p->run_mu);
  if (!p->run_here || !p->run_inf) return MOIP_OK;
  const int k = p->run_here->k;
  if (!p->xstream) {
    if (cudaSetDevice(p->device) != cudaSuccess || cudaStreamCreateWithFlags(&p->xstream, cudaStreamNonBlocking) != cudaSuccess) {
      p->xstream = nullptr;
      return MOIP_ERR_CUDA;
    }
  }
  // the batch goes to the device right here, on the exchange's own stream: the workers' scans then find it in place
  // instead of each paying for an upload under the store's lock
  for (int w = 0; w < 2; ++w) {
    moip_cache* s = w ? p->run_here : p->run_inf;
    std::lock_guard<std::mutex> lk(s->mu);
    bool any = false;
    for (int i = 0; i < n; ++i) {
      if ((infeasible[i] != 0) != (w == 0)) continue;
      CacheRecord r{};
      for (int j = 0; j < k; ++j) { r.ip[j] = ip[(size_t)i * k + j]; r.result[j] = infeasible[i] ? 0 : result[(size_t)i * k + j]; }
      r.infeasible = infeasible[i] ? 1 : 0;
      r.pad[0] = 1;
      s->host.push_back(r);
      any = true;
    }
    if (any && s->sync_to_device(p->xstream)) return MOIP_ERR_CUDA;
  }
  p->imported += n;
  return MOIP_OK;
}

extern "C" int64_t moip_pool_strips_stolen(const moip_pool* p) { return p ? p->stolen : -1; }
extern "C" int64_t moip_pool_boxes_postponed(const moip_pool* p) { return p ? p->deferrals : -1; }

extern "C" int moip_pool_exchange_counts(const moip_pool* p, int64_t* exported, int64_t* imported) {
  if (!p) return MOIP_ERR_ARG;
  if (exported) *exported = p->exported;
  if (imported) *imported = p->imported;
  return MOIP_OK;
}

extern "C" int moip_pool_stats(const moip_pool* p, moip_stats* out) {
  if (!p || !out) return MOIP_ERR_ARG;
  *out = moip_stats{};
  for (moip_ctx* c : p->ctx) {
    out->ip_solved += c->stats.ip_solved; out->bb_nodes += c->stats.bb_nodes; out->node_lps += c->stats.node_lps;
    out->lp_iterations += c->stats.lp_iterations; out->kernel_launches += c->stats.kernel_launches;
    out->cache_queries += c->stats.cache_queries; out->solver_seconds += c->stats.solver_seconds;
  }
  return MOIP_OK;
}

extern "C" int moip_pool_set_kernel_timing(moip_pool* p, int on) {
  if (!p) return MOIP_ERR_ARG;
  for (moip_ctx* c : p->ctx)
    if (int rc = moip_ctx_set_kernel_timing(c, on)) return rc;
  return MOIP_OK;
}
extern "C" int moip_pool_kernel_times(const moip_pool* p, moip_kernel_times* out) {
  if (!p || !out) return MOIP_ERR_ARG;
  *out = moip_kernel_times{};
  for (moip_ctx* c : p->ctx) {
    out->k1_ms += c->ktimes.k1_ms; out->k2_ms += c->ktimes.k2_ms; out->k3_ms += c->ktimes.k3_ms;
    out->k4_ms += c->ktimes.k4_ms; out->copy_ms += c->ktimes.copy_ms; out->rounds += c->ktimes.rounds;
    out->scans += c->ktimes.scans;
  }
  return MOIP_OK;
}

extern "C" int moip_pool_get_limit(moip_pool* p, int obj, int sense, const double* rhs, int* result, int* mip_status) {
  if (!p || p->ctx.empty()) return MOIP_ERR_ARG;
  return moip_get_limit(p->ctx[0], obj, sense, rhs, result, mip_status);
}

// split_optimise (src/aira.cpp:1886-1943) for an explicit list of strips: start_stop holds nstrips (start, stop)
// pairs; rows_out receives the feasible result vectors found (k ints per row, unsorted, duplicates possible)
extern "C" int moip_pool_run_strips(moip_pool* p, int n_obj, int nstrips, const double* start_stop, int* rows_out, int cap,
                                    int* n_rows) {
  return moip_pool_run_strips_claim(p, n_obj, nstrips, start_stop, nullptr, nullptr, rows_out, cap, n_rows);
}

// Same, but every worker asks `claim` for the index of its next strip (values >= nstrips end the worker): lets
// several pools -- one per GPU / rank -- draw from one global counter, so that no rank idles while another still
// has strips queued.  claim == nullptr: local counter.
static int run_boxes(moip_pool* p, int n_obj, int nstrips, const double* start_stop, const double* windows, moip_claim_fn claim,
                     void* user, int* rows_out, int cap, int* n_rows);

extern "C" int moip_pool_run_strips_claim(moip_pool* p, int n_obj, int nstrips, const double* start_stop, moip_claim_fn claim,
                                          void* user, int* rows_out, int cap, int* n_rows) {
  return run_boxes(p, n_obj, nstrips, start_stop, nullptr, claim, user, rows_out, cap, n_rows);
}

// Boxes: strip t additionally carries a window (windows[2t], windows[2t+1]) on objective 1 (moip_worker::window); several
// boxes may share one range of the last objective.  A strip cut off a busy box by an idle worker inherits its window.
extern "C" int moip_pool_run_boxes_claim(moip_pool* p, int n_obj, int nboxes, const double* start_stop, const double* windows,
                                         moip_claim_fn claim, void* user, int* rows_out, int cap, int* n_rows) {
  if (nboxes > 0 && !windows) return MOIP_ERR_ARG;
  return run_boxes(p, n_obj, nboxes, start_stop, windows, claim, user, rows_out, cap, n_rows);
}

static int run_boxes(moip_pool* p, int n_obj, int nstrips, const double* start_stop, const double* windows, moip_claim_fn claim,
                     void* user, int* rows_out, int cap, int* n_rows) {
  if (!p || p->ctx.empty() || nstrips < 0 || (nstrips > 0 && !start_stop) || !n_rows) return MOIP_ERR_ARG;
  const int k = p->ctx[0]->dm.k;
  if (n_obj < 1 || n_obj > k) return MOIP_ERR_ARG;
  const int sense = p->ctx[0]->model->M.sense;
  const bool is_min = sense == MOIP_SENSE_MIN;
  // Work stealing (MOIP_STEAL=0 disables): a worker that finds no strip left takes over the far half of the widest range
  // a busy strip has not reached yet.  Equal-width strips differ 10x in work (the points crowd in the middle of the last
  // objective's range), and every additional strip pays for a lower-dimensional front of its own before it starts to
  // sweep, so the level starts with few strips and splits only where, and when, a worker would otherwise idle.
  const bool steal = nstrips > 0 && !std::getenv("MOIP_NO_STEAL");
  int W = steal ? (int)p->ctx.size() : std::min<int>((int)p->ctx.size(), std::max(1, nstrips));
  if (p->max_workers > 0) W = std::min(W, p->max_workers);
  const int max_strips = nstrips + (steal ? 8 * std::max(4, W) + 64 : 0);   // (>= nstrips + max_steals below)
  std::vector<StripDyn> dyn((size_t)std::max(1, max_strips));
  std::vector<double> sstart((size_t)std::max(1, max_strips), 0.0);
  std::vector<int> cut_from((size_t)std::max(1, max_strips), -1);
  std::vector<double> swin(windows ? 2 * (size_t)std::max(1, max_strips) : 0, 0.0);   // window of every (claimed or cut) strip
  std::atomic<int> next(0), failed(0), n_dyn(nstrips), n_claiming(W), n_stolen(0);
  std::mutex steal_mu;
  // a cut must leave both halves worth a strip's start-up cost (a lower-dimensional front of its own): at least
  // MOIP_STEAL_MIN units of the last objective (default 2) and 1/MOIP_STEAL_FRAC of the level's range (default 1/512);
  // at most 8 W cuts per level.  An idle worker costs nothing, so the thresholds are low: the start-up of a cut-off
  // strip is paid by a worker that had nothing else to do.
  double lo_edge = HUGE_VAL, hi_edge = -HUGE_VAL;
  for (int t = 0; t < nstrips; ++t)
    for (int e = 0; e < 2; ++e) {
      const double v = start_stop[2 * t + e];
      if (std::fabs(v) < 2147483647.0) { lo_edge = std::min(lo_edge, v); hi_edge = std::max(hi_edge, v); }
    }
  const double level_range = hi_edge > lo_edge ? hi_edge - lo_edge : 0.0;
  const double steal_frac = std::getenv("MOIP_STEAL_FRAC") ? std::atof(std::getenv("MOIP_STEAL_FRAC")) : 512.0;
  const double min_steal = std::max(std::max(1.0, (double)(std::getenv("MOIP_STEAL_MIN") ? std::atoi(std::getenv("MOIP_STEAL_MIN")) : 2)),
                                    level_range / std::max(1.0, steal_frac));
  const int max_steals = 8 * std::max(4, W);
  // returns the index of a new strip cut off a busy one, -1 when nothing is worth cutting (yet), -2 when nothing is running
  auto try_steal = [&]() -> int {
    std::lock_guard<std::mutex> lk(steal_mu);
    const int nd = n_dyn.load();
    int victim = -1, running = 0;
    double widest = 0.0;
    for (int t = 0; t < nd; ++t) {
      if (dyn[t].state.load(std::memory_order_acquire) != 1) continue;
      ++running;
      const double pos = dyn[t].pos.load(std::memory_order_acquire), stop = dyn[t].stop.load(std::memory_order_acquire);
      if (std::fabs(pos) >= 2147483647.0 || std::fabs(stop) >= 2147483647.0) continue;   // open-ended ranges (INT_MIN / INT_MAX edges) are not cut
      const double rem = is_min ? pos - stop : stop - pos;
      if (rem > widest) { widest = rem; victim = t; }
    }
    if (running == 0) return -2;
    if (victim < 0 || widest < 2.0 * min_steal || nd >= max_strips || n_stolen.load() >= max_steals) return -1;
    const double pos = dyn[victim].pos.load(), stop = dyn[victim].stop.load();
    // the strip that ends at the level's far edge -- the single-objective optimum of the last objective -- is not cut into
    // narrow pieces: a cold start under a bound that close to the optimum is the one subproblem plain LP-based B&B is bad
    // at (4KP n=40: 9e6 nodes for "f4 >= max - 8" alone, profiles/r02_fronts.md)
    if (std::fabs(stop - (is_min ? lo_edge : hi_edge)) < 0.5 && widest / 2 < 0.05 * level_range) return -1;
    const double mid = is_min ? pos - std::floor(widest / 2) : pos + std::floor(widest / 2);
    sstart[nd] = mid;
    cut_from[nd] = victim;
    if (windows) { swin[2 * (size_t)nd] = swin[2 * (size_t)victim]; swin[2 * (size_t)nd + 1] = swin[2 * (size_t)victim + 1]; }
    dyn[nd].stop.store(stop); dyn[nd].pos.store(mid); dyn[nd].state.store(1, std::memory_order_release);
    dyn[victim].stop.store(mid, std::memory_order_release);
    n_dyn.store(nd + 1);
    n_stolen.fetch_add(1);
    return nd;
  };
  // `here` and `infeasibles` are shared by the strips of a level, like the reference's threads share them
  // (src/aira.cpp:1918-1933); MOIP_POOL_PRIVATE_CACHES=1 gives every worker its own pair instead
  const bool shared = !std::getenv("MOIP_POOL_PRIVATE_CACHES");
  moip_cache *sh_here = nullptr, *sh_inf = nullptr;
  if (shared) {
    int rc0 = moip_cache_create(p->ctx[0], &sh_here);
    if (!rc0) rc0 = moip_cache_create(p->ctx[0], &sh_inf);
    if (rc0) { moip_cache_destroy(sh_here); moip_cache_destroy(sh_inf); return rc0; }
    std::lock_guard<std::mutex> rl(p->run_mu);
    p->run_here = sh_here; p->run_inf = sh_inf; p->exp_cursor[0] = p->exp_cursor[1] = 0;
  }
  // Boxes whose cold first subproblem ran into its node budget wait here until nothing else is left to claim; by then the
  // records of their neighbours usually answer that subproblem (the box that contains the edge of the feasible region
  // proves "nothing beyond" the cheap way, just past its last point).  Budget: 20 000 nodes, x 8 per further attempt, none
  // from the fourth attempt on (MOIP_BOX_BUDGET; 0 = never postpone).
  const long long box_budget0 = windows ? (std::getenv("MOIP_BOX_BUDGET") ? std::atoll(std::getenv("MOIP_BOX_BUDGET")) : 20000) : 0;
  std::vector<std::pair<int, int>> deferred;            // (box, attempts so far)
  std::vector<int> attempts((size_t)std::max(1, max_strips), 0);
  std::mutex deferred_mu;
  std::atomic<int> n_deferred(0), n_deferrals(0);
  std::vector<std::vector<int>> found(W);
  // MOIP_STRIP_TIMELINE=1: one line per strip at the end of the level (who solved it, when, how many IPs, cut off whom)
  struct StripLog { int worker = -1, from = -1; double t0 = 0, t1 = 0, start = 0, stop0 = 0, stop1 = 0; long long ips = 0; };
  const bool timeline = std::getenv("MOIP_STRIP_TIMELINE") != nullptr;
  std::vector<StripLog> slog(timeline ? (size_t)std::max(1, max_strips) : 0);
  const auto t_level0 = std::chrono::steady_clock::now();
  auto since0 = [&]() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_level0).count(); };
  auto work = [&](int wi) {
    moip_ctx* c = p->ctx[wi];
    moip_cache *here = sh_here, *infeasibles = sh_inf;
    int rc = MOIP_OK;
    if (!shared) {
      rc = moip_cache_create(c, &here);
      if (!rc) rc = moip_cache_create(c, &infeasibles);
    }
    bool claims_left = true;
    while (!rc && !failed.load()) {
      int t = -1;
      if (claims_left) {
        t = claim ? claim(user) : next.fetch_add(1);
        if (t < 0 || t >= nstrips) { claims_left = false; t = -1; n_claiming.fetch_sub(1); }
        else {
          sstart[t] = start_stop[2 * t];
          if (windows) { swin[2 * (size_t)t] = windows[2 * t]; swin[2 * (size_t)t + 1] = windows[2 * t + 1]; }
          dyn[t].stop.store(start_stop[2 * t + 1]); dyn[t].pos.store(start_stop[2 * t]);
          dyn[t].state.store(1, std::memory_order_release);
        }
      }
      if (t < 0 && n_deferred.load() > 0) {               // a postponed box: its turn comes when the claims have run out
        std::lock_guard<std::mutex> lk(deferred_mu);
        if (!deferred.empty()) {
          t = deferred.front().first;
          deferred.erase(deferred.begin());
          n_deferred.fetch_sub(1);
          dyn[t].pos.store(start_stop[2 * t]);             // (its end stays where a cut may have moved it meanwhile)
          dyn[t].state.store(1, std::memory_order_release);
        }
      }
      if (t < 0) {
        if (!steal || !shared || failed.load()) break;
        t = try_steal();
        if (t == -2 && n_claiming.load() == 0 && n_deferred.load() == 0) break;   // nothing running, nothing left to claim: the level is done
        if (t < 0) { std::this_thread::sleep_for(std::chrono::microseconds(200)); continue; }
      }
      moip_worker w{};
      w.id = t; w.n_obj = n_obj; w.split = 1;
      for (int i = 0; i < k; ++i) w.perm[i] = i;                            // thread.cpp:124-133
      w.split_start = sstart[t]; w.split_stop = dyn[t].stop.load();
      if (windows) { w.window = 1; w.win_start = swin[2 * (size_t)t]; w.win_stop = swin[2 * (size_t)t + 1]; }
      c->dbg_strip.store(t, std::memory_order_relaxed);
      if (timeline) { slog[t].worker = wi; slog[t].t0 = since0(); slog[t].start = sstart[t]; slog[t].stop0 = w.split_stop; slog[t].ips = c->stats.ip_solved; slog[t].from = cut_from[t]; }
      const long long budget = (windows && t < nstrips && attempts[t] < 3) ? box_budget0 << (3 * attempts[t]) : 0;
      rc = optimise_strip(c, &w, here, infeasibles, steal && shared ? &dyn[t] : nullptr, budget);
      if (rc == MOIP_ERR_BUDGET) {                        // nothing recorded: put the box back
        rc = MOIP_OK;
        attempts[t] += 1;
        n_deferrals.fetch_add(1);
        dyn[t].state.store(0, std::memory_order_release);
        { std::lock_guard<std::mutex> lk(deferred_mu); deferred.emplace_back(t, attempts[t]); n_deferred.fetch_add(1); }
        c->dbg_strip.store(-1, std::memory_order_relaxed);
        if (!claims_left) std::this_thread::sleep_for(std::chrono::milliseconds(2));   // only postponed boxes are left: give the records time to arrive
        continue;
      }
      if (timeline) { slog[t].t1 = since0(); slog[t].stop1 = dyn[t].stop.load(); slog[t].ips = c->stats.ip_solved - slog[t].ips; }
      c->dbg_strip.store(-1, std::memory_order_relaxed);
      dyn[t].state.store(2, std::memory_order_release);
    }
    if (!rc && here && !shared)
      for (auto& r : here->host) if (!r.infeasible) found[wi].insert(found[wi].end(), r.result, r.result + k);   // :1934-1942
    if (rc) failed.store(rc);
    if (!shared) { moip_cache_destroy(here); moip_cache_destroy(infeasibles); }
  };
  // MOIP_WATCHDOG=<seconds>: a thread that reports what every worker is doing (strip, B&B round of the current IP, open
  // nodes, IPs solved so far) -- the counterpart of the reference's DEBUG prints for a run that seems stuck
  std::atomic<bool> watch_stop(false);
  std::thread watchdog;
  if (const char* ws = std::getenv("MOIP_WATCHDOG")) {
    const double every = std::max(0.5, std::atof(ws));
    watchdog = std::thread([&, every] {
      double waited = 0;
      while (!watch_stop.load()) {
        std::this_thread::sleep_for(std::chrono::milliseconds(100));
        waited += 0.1;
        if (waited < every) continue;
        waited = 0;
        for (int wi = 0; wi < W; ++wi) {
          moip_ctx* c = p->ctx[wi];
          const long long t = c->dbg_strip.load();
          std::fprintf(stderr, "moip_b200: watchdog: worker %d strip %lld [%g -> %g, at %g] %s round %lld open %lld; %lld IPs %lld nodes so far\n", wi, t,
                       t >= 0 ? sstart[t] : 0.0, t >= 0 ? dyn[t].stop.load() : 0.0, t >= 0 ? dyn[t].pos.load() : 0.0,
                       c->dbg_where.load() == 1 ? "cache scan" : (c->dbg_where.load() == 2 ? "B&B" : "generator"), c->dbg_rounds.load(),
                       c->dbg_open.load(), (long long)c->stats.ip_solved, (long long)c->stats.bb_nodes);
        }
      }
    });
  }
  std::vector<std::thread> th;
  for (int wi = 1; wi < W; ++wi) th.emplace_back(work, wi);
  work(0);
  for (auto& t : th) t.join();
  watch_stop.store(true);
  if (watchdog.joinable()) watchdog.join();
  if (timeline) {
    const char* rk = std::getenv("RANK");
    for (int t = 0; t < n_dyn.load(); ++t)
      if (slog[t].worker >= 0)
        std::fprintf(stderr, "moip_b200: timeline rank %s level %d strip %d%s worker %d: %.3f -> %.3f s, %lld IPs, range [%g, %g -> %g)%s\n", rk ? rk : "0", n_obj, t,
                     t >= nstrips ? " (cut)" : "", slog[t].worker, slog[t].t0, slog[t].t1, slog[t].ips, slog[t].start, slog[t].stop0, slog[t].stop1,
                     slog[t].from >= 0 ? (" cut off strip " + std::to_string(slog[t].from)).c_str() : "");
    std::fprintf(stderr, "moip_b200: timeline rank %s level %d done after %.3f s\n", rk ? rk : "0", n_obj, since0());
  }
  if (shared) {
    {
      std::lock_guard<std::mutex> rl(p->run_mu);
      p->run_here = nullptr; p->run_inf = nullptr;
    }
    p->stolen += n_stolen.load();
    p->deferrals += n_deferrals.load();
    if (!failed.load())      // (records imported from other ranks are reported by the rank that found them)
      for (auto& r : sh_here->host) if (!r.infeasible && !r.pad[0]) found[0].insert(found[0].end(), r.result, r.result + k);
    moip_cache_destroy(sh_here);
    moip_cache_destroy(sh_inf);
  }
  if (failed.load()) return failed.load();
  int n = 0;
  for (auto& f : found)
    for (size_t i = 0; i + k <= f.size(); i += k, ++n)
      if (rows_out && n < cap) for (int j = 0; j < k; ++j) rows_out[(size_t)n * k + j] = f[i + j];
  *n_rows = n;
  return MOIP_OK;
}

namespace {

// split_setup (src/aira.cpp:1945-1990) on a pool: the strips of one level run concurrently
int epp_level_pool(moip_pool* p, int n_obj, int num_threads, int split_normal, std::vector<std::vector<int>>& sols) {
  moip_ctx* c = p->ctx[0];
  const int k = c->dm.k, sense = c->model->M.sense;
  const bool is_min = sense == MOIP_SENSE_MIN;
  std::vector<double> free_rhs(k, is_min ? kInf : -kInf);
  std::vector<int> res(k, 0);
  int st = 0, rc;
  if (n_obj == 1) {
    if ((rc = c->get_limit(0, sense, free_rhs.data(), res.data(), &st))) return rc;
    if (st != MOIP_MIP_INFEASIBLE) sols.push_back(res);
    return MOIP_OK;
  }
  std::vector<std::vector<int>> lower;
  if ((rc = epp_level_pool(p, n_obj - 1, num_threads, split_normal, lower))) return rc;
  if (lower.empty()) return MOIP_OK;
  if ((rc = c->get_limit(n_obj - 1, sense, free_rhs.data(), res.data(), &st))) return rc;
  if (st == MOIP_MIP_INFEASIBLE) return MOIP_OK;
  int biggest, smallest;
  if (is_min) {
    smallest = res[n_obj - 1]; biggest = INT_MIN;
    for (auto& s : lower) biggest = std::max(biggest, s[n_obj - 1]);
    if (biggest == smallest) biggest = INT_MAX;
  } else {
    biggest = res[n_obj - 1]; smallest = INT_MAX;
    for (auto& s : lower) smallest = std::min(smallest, s[n_obj - 1]);
    if (biggest == smallest) smallest = INT_MIN;
  }
  // Boxes (moip_worker::window; the same rule as aira.windows_for / aira.window_edges): with 16 or more entries the level is
  // cut both ways -- nwin windows on objective 1 (the largest power of two <= 16 that leaves 8 strips; edges = quantiles of
  // that objective over the level below) times num_threads / nwin strips.  MOIP_WINDOWS=<n> fixes nwin (1 = strips only).
  int nwin = 1;
  if (n_obj >= 3) {
    if (const char* e = std::getenv("MOIP_WINDOWS")) nwin = std::max(1, std::min(std::atoi(e), num_threads));
    else while (nwin < 16 && num_threads / (nwin * 2) >= 8) nwin *= 2;
  }
  std::vector<double> near_edge;                       // near edges of the windows, first one "free"
  if (nwin > 1) {
    std::vector<int> vals;
    for (auto& s : lower) vals.push_back(s[1]);
    std::sort(vals.begin(), vals.end());
    vals.erase(std::unique(vals.begin(), vals.end()), vals.end());
    if ((int)vals.size() < 2 * nwin) nwin = 1;
    else {
      std::vector<int> cuts;
      for (int i = 1; i < nwin; ++i) cuts.push_back(vals[vals.size() * (size_t)i / nwin]);
      cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
      near_edge.push_back(is_min ? kInf : -kInf);
      if (is_min) for (auto it = cuts.rbegin(); it != cuts.rend(); ++it) near_edge.push_back((double)*it);
      else for (int c : cuts) near_edge.push_back((double)c);
      nwin = (int)near_edge.size();
    }
  }
  const int nstrips = nwin > 1 ? std::max(1, num_threads / nwin) : num_threads;
  std::vector<double> ss(2 * (size_t)nstrips);
  if ((rc = moip_split_strips(sense, biggest, smallest, nstrips, split_normal, ss.data()))) return rc;
  std::vector<double> bs, bw;
  if (nwin > 1)
    for (int t = 0; t < nstrips; ++t)
      for (int w = 0; w < nwin; ++w) {
        bs.push_back(ss[2 * t]); bs.push_back(ss[2 * t + 1]);
        bw.push_back(near_edge[w]);                     // a window ends one unit before the next one starts
        bw.push_back(w + 1 < nwin ? near_edge[w + 1] + (is_min ? 1.0 : -1.0) : (is_min ? -kInf : kInf));
      }
  int cap = 1 << 14, nrows = 0;
  std::vector<int> rows((size_t)cap * k);
  for (;;) {
    if (nwin > 1) rc = moip_pool_run_boxes_claim(p, n_obj, nstrips * nwin, bs.data(), bw.data(), nullptr, nullptr, rows.data(), cap, &nrows);
    else rc = moip_pool_run_strips(p, n_obj, num_threads, ss.data(), rows.data(), cap, &nrows);
    if (rc) return rc;
    if (nrows <= cap) break;
    cap = nrows; rows.assign((size_t)cap * k, 0);       // (re-solves; only for fronts beyond 16k rows)
  }
  for (int i = 0; i < nrows; ++i) sols.emplace_back(rows.begin() + (size_t)i * k, rows.begin() + (size_t)(i + 1) * k);
  return MOIP_OK;
}

}  // namespace

extern "C" int moip_pool_pareto_front(moip_pool* p, int num_threads, int split_normal, int* rows_out, int cap, int* n_rows) {
  if (!p || p->ctx.empty() || !n_rows) return MOIP_ERR_ARG;
  const int k = p->ctx[0]->dm.k;
  if (num_threads < 1) num_threads = 1;
  std::vector<std::vector<int>> sols;
  int rc = epp_level_pool(p, k, num_threads, split_normal, sols);
  if (rc) return rc;
  moip_cache* all = nullptr;
  if ((rc = moip_cache_create(p->ctx[0], &all))) return rc;
  std::vector<double> zero(k, 0.0);
  for (auto& s : sols) moip_cache_insert(all, zero.data(), s.data(), 0);
  *n_rows = moip_cache_sort_unique(all, rows_out, cap);
  moip_cache_destroy(all);
  return MOIP_OK;
}

// main() with -t W and no --split (src/aira.cpp:277-308), on the cooperative workers above: worker i runs on solver
// context i of the pool (own stream, same GPU) with the i-th rotation of the objective order; infeasible records are
// shared (valid for every permutation, like the reference's shared `infeasibles`, src/aira.cpp:539), solution records
// are per worker (a lexicographic optimum depends on the permutation).  rows_out = sorted, de-duplicated front.
extern "C" int moip_pool_synergistic_front(moip_pool* p, int n_workers, int* rows_out, int cap, int* n_rows) {
  if (!p || p->ctx.empty() || !n_rows) return MOIP_ERR_ARG;
  const int k = p->ctx[0]->dm.k, sense = p->ctx[0]->model->M.sense;
  n_workers = std::max(1, std::min(n_workers, std::min(k, (int)p->ctx.size())));
  std::vector<moip_worker> ws(n_workers);
  int rc = moip_coop_workers(k, n_workers, ws.data());
  if (rc) return rc;
  CoopShared sh;
  sh.init(k, sense == MOIP_SENSE_MIN);
  for (int i = 0; i < n_workers; ++i) sh.owned[ws[i].perm[k - 1]] = n_workers > 1;
  moip_cache* infeasibles = nullptr;
  std::vector<moip_cache*> sols(n_workers, nullptr);
  rc = moip_cache_create(p->ctx[0], &infeasibles);
  for (int i = 0; i < n_workers && !rc; ++i) rc = moip_cache_create(p->ctx[i], &sols[i]);
  std::vector<int> rcs(n_workers, MOIP_OK);
  if (!rc) {
    auto work = [&](int i) {
      GpuBackend be;
      be.c = p->ctx[i]; be.infeasibles = infeasibles; be.sols = sols[i]; be.sense = sense;
      CoopBackend co;
      co.inner = &be; co.sh = &sh; co.own = ws[i].perm[k - 1];
      rcs[i] = run_worker(co, k, sense, ws[i], nullptr, nullptr);
      if (!rcs[i]) co.finished();
    };
    std::vector<std::thread> th;
    for (int i = 1; i < n_workers; ++i) th.emplace_back(work, i);
    work(0);
    for (auto& t : th) t.join();
    for (int i = 0; i < n_workers && !rc; ++i) rc = rcs[i];
  }
  if (!rc) {
    moip_cache* all = nullptr;
    rc = moip_cache_create(p->ctx[0], &all);
    if (!rc) {
      std::vector<double> zero(k, 0.0);
      for (int i = 0; i < n_workers; ++i)
        for (auto& r : sols[i]->host) if (!r.infeasible) moip_cache_insert(all, zero.data(), r.result, 0);
      *n_rows = moip_cache_sort_unique(all, rows_out, cap);
      moip_cache_destroy(all);
    }
  }
  for (auto* sc : sols) moip_cache_destroy(sc);
  moip_cache_destroy(infeasibles);
  return rc;
}

extern "C" int moip_pareto_front(moip_ctx* c, int split, int num_threads, int split_normal, int* rows_out, int cap,
                                 int* n_rows) {
  if (!c || !n_rows) return MOIP_ERR_ARG;
  const int k = c->dm.k;
  moip_cache* all = nullptr;
  int rc = moip_cache_create(c, &all);
  if (rc) return rc;
  if (split) {                                                             // src/aira.cpp:269-276
    if (num_threads < 1) num_threads = 1;
    std::vector<std::vector<int>> sols;
    rc = epp_level(c, k, num_threads, split_normal, sols);
    std::vector<double> zero(k, 0.0);
    for (auto& s : sols) moip_cache_insert(all, zero.data(), s.data(), 0);
  } else {                                                                 // src/aira.cpp:277-308 with -t 1
    moip_cache* infeasibles = nullptr;
    rc = moip_cache_create(c, &infeasibles);
    if (!rc) {
      moip_worker w{};
      w.id = 0; w.n_obj = k; w.split = 0;
      for (int i = 0; i < k; ++i) w.perm[i] = i;
      rc = moip_optimise(c, &w, all, infeasibles);
      moip_cache_destroy(infeasibles);
    }
  }
  if (!rc) *n_rows = moip_cache_sort_unique(all, rows_out, cap);           // src/aira.cpp:336-346
  moip_cache_destroy(all);
  return rc;
}
