// TEST INFRASTRUCTURE ONLY.  Drives the library's HOST code under ThreadSanitizer on the CPU stand-ins of
// tests/tsan/cuda_double.cpp (built by `make -C oracle tsan`, run by tests/test_tsan_host.py):
//   (1) moip_pool_pareto_front: 12 workers, few strips -> idle workers cut busy strips in two (work stealing), shared
//       `here` / `infeasibles` stores, the model's point store;
//   (2) two pools ("ranks") solving interleaved strips of one level while a third thread carries cache records between
//       them (moip_pool_export_records / moip_pool_import_records, the path aira.RecordExchange drives);
//   (3) moip_pool_synergistic_front: W = k cooperative workers exchanging limits;
//   (5) moip_pool_run_boxes_claim: strips x windows, postponed boxes (MOIP_BOX_BUDGET=1);
//   (4) the link-level CPLEX seam (seam1/cpx_shim.cpp): T threads, one environment + problem object each, the call
//       sequence of the reference's solve() (src/aira.cpp:452-536).
// Every result is compared with the single-context run; the whole programme is repeated argv[2] times.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <thread>
#include <vector>

#include "../../include/moip_b200.h"
#include "../../moip_aira_b200/seam1/include/ilcplex/cplex.h"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "tsan_host: check failed at line %d: %s\n", __LINE__, #c); std::exit(2); } } while (0)

typedef std::set<std::vector<int>> Front;

static Front rows_to_front(const std::vector<int>& rows, int n, int k) {
  Front f;
  for (int i = 0; i < n; ++i) f.insert(std::vector<int>(rows.begin() + (size_t)i * k, rows.begin() + (size_t)(i + 1) * k));
  return f;
}

int main(int argc, char** argv) {
  CHECK(argc >= 2);
  const char* path = argv[1];
  const int reps = argc > 2 ? std::atoi(argv[2]) : 3;
  const int workers = argc > 3 ? std::atoi(argv[3]) : 12;
  moip_model* m = nullptr;
  CHECK(moip_model_load(path, &m) == MOIP_OK);
  moip_model_info info;
  CHECK(moip_model_get_info(m, &info) == MOIP_OK);
  const int k = info.k, n = info.n, cap = 1 << 14;
  std::vector<int> rows((size_t)cap * k);
  int nrows = 0;
  // ---- the answer: one context, sequential generator
  moip_ctx* c0 = nullptr;
  CHECK(moip_ctx_create_own_stream(m, 0, &c0) == MOIP_OK);
  CHECK(moip_pareto_front(c0, 0, 1, 0, rows.data(), cap, &nrows) == MOIP_OK);
  const Front want = rows_to_front(rows, nrows, k);
  CHECK(!want.empty());
  // strip edges of the top level: range of the last objective over the front
  int lo = INT32_MAX, hi = INT32_MIN;
  for (auto& r : want) { lo = std::min(lo, r[k - 1]); hi = std::max(hi, r[k - 1]); }
  long long steals = 0, postponed = 0;
  for (int rep = 0; rep < reps; ++rep) {
    // (1) pool, strips < workers
    moip_pool* p = nullptr;
    CHECK(moip_pool_create(m, 0, workers, &p) == MOIP_OK);
    CHECK(moip_pool_pareto_front(p, 2 + rep % 3, 0, rows.data(), cap, &nrows) == MOIP_OK);
    CHECK(rows_to_front(rows, nrows, k) == want);
    steals += moip_pool_strips_stolen(p);
    // (1b) 16 or more entries: the levels with three or more objectives are cut into boxes (epp_level_pool)
    CHECK(moip_pool_pareto_front(p, 16 + rep % 9, 0, rows.data(), cap, &nrows) == MOIP_OK);
    CHECK(rows_to_front(rows, nrows, k) == want);
    // (3) cooperative workers
    CHECK(moip_pool_synergistic_front(p, k, rows.data(), cap, &nrows) == MOIP_OK);
    CHECK(rows_to_front(rows, nrows, k) == want);
    // (2) two "ranks" + an exchange thread
    moip_pool* q = nullptr;
    CHECK(moip_pool_create(m, 0, workers / 2 + 1, &q) == MOIP_OK);
    const int S = 6;
    std::vector<double> ss(2 * S);
    CHECK(moip_split_strips(info.sense, hi + 1, lo - 1, S, 0, ss.data()) == MOIP_OK);
    struct Claim { std::atomic<int> next; int rank; };
    Claim ca{{0}, 0}, cb{{0}, 1};
    auto claim = [](void* u) -> int { Claim* c = (Claim*)u; const int i = c->next.fetch_add(1); const int s = 2 * i + c->rank; return s < S ? s : S; };
    std::vector<int> ra((size_t)cap * k), rb((size_t)cap * k);
    int na = 0, nb = 0, rca = -1, rcb = -1;
    std::atomic<int> done(0);
    std::thread ta([&] { rca = moip_pool_run_strips_claim(p, k, S, ss.data(), claim, &ca, ra.data(), cap, &na); done.fetch_add(1); });
    std::thread tb([&] { rcb = moip_pool_run_strips_claim(q, k, S, ss.data(), claim, &cb, rb.data(), cap, &nb); done.fetch_add(1); });
    std::thread tx([&] {
      const int X = 64;
      std::vector<double> ip((size_t)X * k);
      std::vector<int> res((size_t)X * k), inf(X);
      for (;;) {
        const bool fin = done.load() == 2;
        moip_pool* pools[2] = {p, q};
        for (int d = 0; d < 2; ++d) {
          int cnt = 0;
          CHECK(moip_pool_export_records(pools[d], X, ip.data(), res.data(), inf.data(), &cnt) == MOIP_OK);
          if (cnt) CHECK(moip_pool_import_records(pools[1 - d], cnt, ip.data(), res.data(), inf.data()) == MOIP_OK);
        }
        if (fin) break;
        std::this_thread::sleep_for(std::chrono::microseconds(200));
      }
    });
    ta.join(); tb.join(); tx.join();
    CHECK(rca == MOIP_OK && rcb == MOIP_OK);
    Front got = rows_to_front(ra, na, k);
    Front gb = rows_to_front(rb, nb, k);
    got.insert(gb.begin(), gb.end());
    CHECK(got == want);
    int64_t ex = 0, im = 0;
    CHECK(moip_pool_exchange_counts(p, &ex, &im) == MOIP_OK);
    // (5) boxes: strips of the last objective x windows on objective 1, more workers than boxes at the end (cuts), and --
    // with MOIP_BOX_BUDGET=1 in the environment -- boxes whose first subproblem "runs out of budget" are put back and retried
    if (k >= 3) {
      int lo1 = INT32_MAX, hi1 = INT32_MIN;
      for (auto& r : want) { lo1 = std::min(lo1, r[1]); hi1 = std::max(hi1, r[1]); }
      const int SB = 2 + rep % 2, WN = 3;
      std::vector<double> sb(2 * SB), bs, bw;
      CHECK(moip_split_strips(info.sense, hi + 1, lo - 1, SB, 0, sb.data()) == MOIP_OK);
      const bool is_min = info.sense == MOIP_SENSE_MIN;
      const double big = 1e20;
      for (int s_ = 0; s_ < SB; ++s_)
        for (int w_ = 0; w_ < WN; ++w_) {
          // window w_ of WN on [lo1, hi1]: near edge free for the first, no far edge for the last
          const double c_hi = lo1 + (double)(hi1 - lo1) * (WN - w_) / WN, c_lo = lo1 + (double)(hi1 - lo1) * (WN - w_ - 1) / WN;
          double near_e, far_e;
          if (is_min) { near_e = w_ == 0 ? big : std::floor(c_hi); far_e = w_ == WN - 1 ? -big : std::floor(c_lo) + 1; }
          else { near_e = w_ == WN - 1 ? -big : std::floor(c_lo) + 1; far_e = w_ == 0 ? big : std::floor(c_hi); }
          bs.push_back(sb[2 * s_]); bs.push_back(sb[2 * s_ + 1]);
          bw.push_back(near_e); bw.push_back(far_e);
        }
      CHECK(moip_pool_run_boxes_claim(p, k, SB * WN, bs.data(), bw.data(), nullptr, nullptr, rows.data(), cap, &nrows) == MOIP_OK);
      CHECK(rows_to_front(rows, nrows, k) == want);
      postponed += moip_pool_boxes_postponed(p);
    }
    moip_pool_destroy(q);
    moip_pool_destroy(p);
  }
  // (4) the CPLEX seam: T threads, one env + problem each, the same sequence of solve() chains
  {
    const int T = 6;
    std::vector<std::vector<double>> objc(k, std::vector<double>(n));
    for (int j = 0; j < k; ++j) CHECK(moip_model_objcoef(m, j, objc[j].data()) == MOIP_OK);
    std::vector<std::vector<int>> results(T);
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
      th.emplace_back([&, t] {
        int st = -1;
        CPXENVptr env = CPXopenCPLEX(&st);
        CHECK(env && st == 0);
        CPXLPptr lp = CPXcreateprob(env, &st, "p");
        CHECK(lp && st == 0);
        CHECK(CPXreadcopyprob(env, lp, path, nullptr) == 0);
        const int nr = CPXgetnumrows(env, lp), nc = CPXgetnumcols(env, lp);
        CHECK(nc == n);
        std::vector<int> ind(n), conind(k);
        for (int j = 0; j < n; ++j) ind[j] = j;
        for (int j = 0; j < k; ++j) conind[j] = nr - k + j;
        std::vector<double> x(n);
        const double free_rhs = info.sense == MOIP_SENSE_MIN ? 1e20 : -1e20;
        int bound = info.sense == MOIP_SENSE_MIN ? hi : lo;
        for (int step = 0; step < 6; ++step) {                 // a few subproblems with a moving bound on the last objective
          std::vector<double> srhs(k, free_rhs);
          srhs[k - 1] = bound;
          bool feasible = true;
          for (int j = 0; j < k && feasible; ++j) {            // the chain of src/aira.cpp:467-517
            CHECK(CPXchgobj(env, lp, n, ind.data(), objc[j].data()) == 0);
            CHECK(CPXchgrhs(env, lp, k, conind.data(), srhs.data()) == 0);
            CHECK(CPXmipopt(env, lp) == 0);
            if (CPXgetstat(env, lp) == CPXMIP_INFEASIBLE) { feasible = false; break; }
            double ov = 0;
            CHECK(CPXgetobjval(env, lp, &ov) == 0);
            srhs[j] = ov;
            results[t].push_back((int)ov);
          }
          if (!feasible) { results[t].push_back(INT32_MIN); break; }
          CHECK(CPXgetx(env, lp, x.data(), 0, n - 1) == 0);
          bound += info.sense == MOIP_SENSE_MIN ? -1 : 1;
        }
        CHECK(CPXfreeprob(env, &lp) == 0);
        CHECK(CPXcloseCPLEX(&env) == 0);
      });
    for (auto& t : th) t.join();
    for (int t = 1; t < T; ++t) CHECK(results[t] == results[0]);
    CHECK(!results[0].empty());
  }
  moip_ctx_destroy(c0);
  moip_model_free(m);
  std::printf("TSAN_HOST_OK front=%zu reps=%d workers=%d strips_cut_by_idle_workers=%lld boxes_postponed=%lld\n", want.size(), reps, workers, steals, postponed);
  return 0;
}
