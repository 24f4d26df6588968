/*
 * moip_b200.h -- C ABI of the B200-native solver core behind moip_aira's hot path.
 *
 * Drop-in boundary ("Seam 2"): every entry point below replaces one
 * reference interface, cited as file:line relative to the reference tree.  Conventions follow
 * the CPLEX callable library the reference binds today: plain pointers and sizes, caller-owned
 * buffers (the library copies), `int` return 0 = OK (nonzero = error, diagnostic on stderr),
 * no exceptions across the boundary, one context per worker thread (a context is used by its
 * owner only; several contexts may share one GPU).
 *
 * The link-level boundary ("Seam 1": the 23 CPX* functions the unmodified reference objects import, src/env.h:4-10)
 * is declared in moip_aira_b200/seam1/include/ilcplex/cplex.h and implemented on top of this header by
 * moip_aira_b200/seam1/cpx_shim.cpp (libcplex_moip_b200.so); see INTEGRATION.md.
 *
 * There is no CPU fallback behind any compute entry point: they return MOIP_ERR_CUDA when no
 * sm_100 device / kernel image is usable.
 */
#ifndef MOIP_B200_H
#define MOIP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status / error codes -------------------------------------------------------------- */
#define MOIP_OK 0
#define MOIP_ERR_ARG 1
#define MOIP_ERR_PARSE 2
#define MOIP_ERR_CUDA 3
#define MOIP_ERR_UNSUPPORTED 4
#define MOIP_ERR_LIMIT 5
#define MOIP_ERR_BUDGET 6   /* an IP was given up at its node budget (moip_ctx_set_ip_node_budget); nothing was recorded */
/* MIP statuses: the values the reference compares against after CPXgetstat
 * (src/aira.cpp:489-492, :644, :840).  Kept numerically equal to CPLEX's so that
 * `solnstat == CPXMIP_INFEASIBLE` keeps working when solve() forwards our status. */
#define MOIP_MIP_OPTIMAL 101
#define MOIP_MIP_INFEASIBLE 103
#define MOIP_MIP_INFORUNBD 119
/* node-LP statuses (no reference counterpart: node LPs live inside CPXmipopt, src/aira.cpp:480) */
#define MOIP_LP_CONVERGED 0   /* relative KKT error <= eps                              */
#define MOIP_LP_CUTOFF 1      /* valid dual bound reached the cutoff (node can be pruned) */
#define MOIP_LP_ITERLIMIT 2   /* iteration cap; dual bound is still valid                */
#define MOIP_LP_INFEASIBLE 3  /* dual bound exceeded the objective's upper bound on the box */

#define MOIP_SENSE_MIN 0 /* reference src/sense.h:4 */
#define MOIP_SENSE_MAX 1
#define MOIP_INFBOUND 1.0E+20 /* CPX_INFBOUND; "no bound" in rhs[] (src/problem.cpp:126) */
#define MOIP_MAX_OBJ 4        /* reference supports objcnt < maxObjCount = 5 (src/aira.cpp:230) */

typedef struct moip_model moip_model; /* replaces Problem + the CPXLPptr model (src/problem.h, src/env.h:6-10) */
typedef struct moip_ctx moip_ctx;     /* replaces one worker's Env (CPXENVptr + CPXLPptr copy, src/aira.cpp:541-585) */
typedef struct moip_cache moip_cache; /* replaces Solutions (src/solutions.h:10-36) */

typedef struct {
  int n;        /* columns                     (CPXgetnumcols, src/problem.cpp:47)            */
  int ms;       /* structural rows (without the k objective-bound rows)                        */
  int k;        /* objcnt                      (src/problem.cpp:61)                            */
  int nnz;      /* structural nonzeros                                                         */
  int sense;    /* MOIP_SENSE_*               (src/problem.cpp:119-120)                        */
  int all_binary;
  int m;        /* ms + k = rows of the LP the kernels see                                     */
  int mask_words; /* uint32 words of one 2-bit/var fixing mask = ceil(n/16)                    */
} moip_model_info;

/* ---- model: replaces Problem::Problem + CPXreadcopyprob (src/problem.cpp:12-154, :157-340) - */
int moip_model_load(const char* path, moip_model** out);
void moip_model_free(moip_model* m);
int moip_model_get_info(const moip_model* m, moip_model_info* info);
/* Problem::objcoef[obj][0..n) (src/problem.cpp:73-107) */
int moip_model_objcoef(const moip_model* m, int obj, double* out_n);
/* structural part, dense row-major ms*n, row sense chars 'L','G','E', rhs, column bounds/integrality */
int moip_model_dense(const moip_model* m, double* a_ms_n, char* row_sense_ms, double* rhs_ms,
                     double* lb_n, double* ub_n, uint8_t* is_int_n);
int moip_model_colname(const moip_model* m, int j, char* buf, int buflen); /* CPXgetcolname, src/problem.cpp:183 */
/* Which K1 kernel the model maps to and whether the packed device images are consistent with the model (every
 * nonzero present exactly once, shared-memory offsets in range and collision free, scalings reproduce the matrix).
 * No device needed.  kernel_path: 0 generic, 1 k1_fast, 2 k1_reg, 3 k1_small, 4 generic in streaming mode.
 * Returns MOIP_OK, or MOIP_ERR_LIMIT with a description in msg when an invariant is violated. */
int moip_model_selfcheck(const moip_model* m, int* kernel_path, char* msg, int msglen);

/* ---- context: one per worker (src/aira.cpp:561-585). `stream` is a cudaStream_t (NULL = default). */
int moip_ctx_create(moip_model* m, int device, void* stream, moip_ctx** out);
/* same on a non-blocking stream created and owned by the context (callers that hold no CUDA handles: the
 * link-level CPLEX seam creates one context per CPXLPptr, src/aira.cpp:561-585) */
int moip_ctx_create_own_stream(moip_model* m, int device, moip_ctx** out);
void moip_ctx_destroy(moip_ctx* c);
/* number of usable CUDA devices (0 when there is none); lets a threaded host place one worker per GPU */
int moip_device_count(void);

/* ---- K1: batched node-LP relaxations (inside CPXmipopt today, src/aira.cpp:480) ------------- */
typedef struct {
  double eps;        /* relative KKT tolerance for MOIP_LP_CONVERGED (default 1e-8)         */
  int max_iter;      /* iteration cap per node (default 100000)                              */
  int check_every;   /* KKT / dual-bound evaluation cadence in iterations (default 32)       */
  int fixed_iters;   /* >0: run exactly this many iterations, no termination tests (roofline) */
  double cutoff;     /* stop with MOIP_LP_CUTOFF once dual bound >= cutoff (internal min-form; use MOIP_INFBOUND to disable) */
} moip_lp_params;
void moip_lp_default_params(moip_lp_params* p);

/* Host-buffer call (the e2e path): B node LPs  min/max objective[cost_idx[b]]
 * s.t. structural rows, objective-bound rows with rhs[b][0..k) (+-1e20 = free),
 * columns fixed by fix_masks[b] (2 bits per column: 0 free, 2 fixed to 0, 3 fixed to 1).
 * Outputs (all optional except status): objective value of the primal iterate and the valid
 * Lagrangian bound in the model's own sense/sign, status (MOIP_LP_*), iterations, x[b][0..n). */
int moip_lp_batch_solve(moip_ctx* c, int B, const int* cost_idx, const double* rhs,
                        const uint32_t* fix_masks, const moip_lp_params* params,
                        double* primal_obj, double* dual_bound, int* status, int* iters,
                        double* x_out);
/* Resident variant: upload once, run many times (bench `value`), download. */
int moip_lp_batch_upload(moip_ctx* c, int B, const int* cost_idx, const double* rhs,
                         const uint32_t* fix_masks);
int moip_lp_batch_run(moip_ctx* c, const moip_lp_params* params);         /* async on the ctx stream */
int moip_lp_batch_download(moip_ctx* c, double* primal_obj, double* dual_bound, int* status,
                           int* iters, double* x_out);                     /* syncs the stream */

/* ---- K3: solution cache (src/solutions.cpp:11-101, src/solutions.h:41-57) ------------------- */
int moip_cache_create(moip_ctx* c, moip_cache** out);
void moip_cache_destroy(moip_cache* s);
/* Solutions::insert (src/solutions.cpp:82-101); result may be NULL when infeasible */
int moip_cache_insert(moip_cache* s, const double* ip, const int* result, int infeasible);
int moip_cache_size(const moip_cache* s);
/* Solutions::find for Q queries at once (src/solutions.cpp:11-81): first_match[q] = insertion
 * index of the first record that is a still-valid relaxation of ip[q][0..k), or -1. */
int moip_cache_find_batch(moip_cache* s, int Q, const double* ip, int sense, int* first_match);
/* read record i back (what the caller does through the returned Result*, src/aira.cpp:826-827) */
int moip_cache_get(const moip_cache* s, int i, double* ip, int* result, int* infeasible);
/* Solutions::merge (src/solutions.h:41-44): splice `other` in front of `s`, leaving it empty */
int moip_cache_merge(moip_cache* s, moip_cache* other);
/* Solutions::sort_unique + the row dump of main (src/solutions.h:54-57, src/aira.cpp:336-346):
 * returns the number of feasible rows, writes up to cap rows of k ints */
int moip_cache_sort_unique(moip_cache* s, int* rows, int cap);

/* ---- K4: exact int64 verification of integer points (round()/Sum c x of src/aira.cpp:517-530) */
int moip_verify_int64(moip_ctx* c, int B, const int32_t* x, const double* rhs /* B*k or NULL */,
                      int64_t* obj_out /* B*k */, uint8_t* feasible_out /* B */);

/* ---- the solve: replaces the bodies of solve() and get_limit() ------------------------------ */
/* int solve(Env&, Problem&, int* result, double* rhs, Thread* t)   (src/aira.cpp:452-536) */
int moip_lex_solve(moip_ctx* c, const int* perm, int n_obj, const double* rhs, int* result,
                   int* mip_status);
/* void get_limit(Env&, Problem&, int obj, double* rhs, int* result, Sense) (src/aira.cpp:367-450);
 * result is left untouched when infeasible, like the reference (:410-412). */
int moip_get_limit(moip_ctx* c, int obj, int sense, const double* rhs, int* result, int* mip_status);

/* Seam 1 primitive: one CPXmipopt + CPXgetstat + CPXgetobjval + CPXgetx (src/aira.cpp:480-521): optimise objective
 * `obj` in the model's sense under the k objective-bound rows at `rhs`; x_out[n] and *objval are written unless the
 * problem is infeasible.  x_start (or NULL) is a point known to satisfy the model and `rhs` (MIP start).  Used by the
 * link-level CPLEX shim in moip_aira_b200/seam1/, which lets the unmodified reference sources run on this library. */
int moip_mip_solve(moip_ctx* c, int obj, const double* rhs, const int32_t* x_start, int32_t* x_out,
                   int64_t* objval, int* mip_status);

/* counters (ipcount of src/aira.cpp:80 and the FINETIMING split of :554-560) */
typedef struct {
  int64_t ip_solved;      /* single-objective IPs solved (CPXmipopt calls in the reference) */
  int64_t bb_nodes;       /* branch-and-bound nodes evaluated                               */
  int64_t node_lps;       /* node LP relaxations solved on the GPU                          */
  int64_t lp_iterations;  /* PDHG iterations summed over node LPs                           */
  int64_t kernel_launches;/* kernels launched by this context                               */
  int64_t cache_queries;  /* K3 queries                                                     */
  double solver_seconds;  /* wall time inside lex_solve/get_limit                           */
} moip_stats;
int moip_ctx_stats(const moip_ctx* c, moip_stats* out);
int moip_ctx_reset_stats(moip_ctx* c);
/* Device-side split of the solver time per kernel class: the counterpart of the reference's -DFINETIMING split
 * (src/aira.cpp:554-560, :1868-1876), measured with CUDA events on the context's stream around each class of a B&B round
 * (K2 propagate + branch, K1 node LPs, K4 round/verify, the round's H2D + D2H blocks) and around each K3 scan.  Off by
 * default (two event records per class and round); MOIP_KERNEL_TIMING=1 switches it on for every new context. */
typedef struct {
  double k1_ms, k2_ms, k3_ms, k4_ms, copy_ms;
  int64_t rounds;          /* B&B rounds timed */
  int64_t scans;           /* K3 launches timed */
} moip_kernel_times;
int moip_ctx_set_kernel_timing(moip_ctx* c, int on);
/* how the context's owner waits for a B&B round: 0 = spin in the driver (lowest latency, one core per worker), 1 = sleep on
 * a blocking event (for more workers than cores).  MOIP_SYNC=spin|block overrides; pools choose by core count (auto). */
int moip_ctx_set_sync_mode(moip_ctx* c, int blocking);
int moip_ctx_kernel_times(const moip_ctx* c, moip_kernel_times* out);
/* Node budget per IP (0 = none, the default): moip_lex_solve / moip_get_limit return MOIP_ERR_BUDGET once the B&B tree of one
 * of their IPs has processed more nodes than this.  The box scheduler (moip_pool_run_boxes_claim) uses it to postpone a box
 * whose cold first subproblem explodes until the records of its neighbours answer it. */
int moip_ctx_set_ip_node_budget(moip_ctx* c, long long nodes);

/* ---- subproblem generator re-hosted on the boundary above (src/aira.cpp:538-1884 without the
 * inter-thread bound cells, and the EPP driver src/aira.cpp:1886-1990) -------------------------- */
typedef int (*moip_solve_fn)(void* user, const int* perm, int n_obj, const double* rhs, int* result,
                             int* mip_status);
/* cache callbacks for the host-logic test hook: find returns 1 (hit: fills infeasible/result), 0 (miss), <0 error */
typedef int (*moip_find_cb)(void* user, const double* ip, int* infeasible, int* result);
typedef int (*moip_insert_cb)(void* user, const double* ip, const int* result, int infeasible);
typedef struct {
  int id;
  int n_obj;                 /* Thread::nObj()  (src/thread.h:41)  */
  int perm[MOIP_MAX_OBJ];    /* Thread::perm(i) (src/thread.h:37)  */
  int split;                 /* global `split`  (src/aira.cpp:38)  */
  double split_start, split_stop; /* Thread::split_start/stop (src/thread.h:25-26) */
  /* Window on the objective of the innermost sweeps, perm[1] (no counterpart in the reference; 0 = off).  An EPP strip is a
   * range of the LAST objective and starts with a whole (n_obj-1)-objective front of its own, so strips alone stop paying
   * once that start-up outweighs a strip's share of the front.  A window cuts the other way: bounds of perm[1] start at
   * win_start instead of "free", and a subproblem whose perm[1] bound lies beyond win_stop is taken as infeasible without
   * being solved.  IPs are still solved WITHOUT a lower limit, so every point found is non-dominated for the whole model,
   * and a point inside the window can only be skipped over by the answer of a solved IP, which the trackers then hold: the
   * boxes (strip x window) of a level together enumerate the level's front.  Needs n_obj >= 3. */
  int window;
  double win_start, win_stop;
} moip_worker;
/* optimise<sense>() for one worker.  `all` / `infeasibles` are the two shared stores of
 * src/aira.cpp:539; found points are merged into `all` (:1877-1879). */
int moip_optimise(moip_ctx* c, const moip_worker* w, moip_cache* all, moip_cache* infeasibles);
/* same state machine with caller-supplied solve/find/insert (host-logic tests without a GPU; the
 * product path is moip_optimise, which binds them to the GPU solve and the K3 scan) */
int moip_optimise_with(int k, int sense, const moip_worker* w, moip_solve_fn solve, moip_find_cb find,
                       moip_insert_cb insert, void* user, int64_t* n_iterations, int64_t* n_hits);
/* strip edges of split_optimise (src/aira.cpp:1886-1917): writes num_threads (start,stop) pairs */
int moip_split_strips(int sense, int biggest, int smallest, int num_threads, int split_normal,
                      double* start_stop);
/* whole program: what main() computes (src/aira.cpp:264-346).  split=0: one synergistic worker
 * with the identity permutation (-t 1); split=1: EPP with num_threads strips solved one after
 * another on this context (multi-GPU runs shard the strips across ranks).
 * rows_out receives the sorted, de-duplicated front (k ints per row). */
int moip_pareto_front(moip_ctx* c, int split, int num_threads, int split_normal, int* rows_out,
                      int cap, int* n_rows);

/* ---- worker pool: the reference's one-thread-per-strip model (src/aira.cpp:1920-1933, one CPLEX
 * environment per thread) with one solver context per host thread on ONE device, so that the B&B rounds
 * of concurrent strips share the GPU.  Every worker keeps its own caches. */
typedef struct moip_pool moip_pool;
int moip_pool_create(moip_model* m, int device, int workers, moip_pool** out);
void moip_pool_destroy(moip_pool* p);
int moip_pool_workers(const moip_pool* p);
int moip_pool_stats(const moip_pool* p, moip_stats* out);                   /* summed over the workers */
int moip_pool_set_kernel_timing(moip_pool* p, int on);
int moip_pool_kernel_times(const moip_pool* p, moip_kernel_times* out);     /* summed over the workers */
/* at most max_workers contexts draw strips in moip_pool_run_strips* (0 = all): several ranks that share one strip counter
 * each take their share instead of the first rank claiming every strip */
int moip_pool_set_max_workers(moip_pool* p, int max_workers);
/* Knowledge exchange between the pools of several ranks while a level's strips are being solved: the records of the
 * shared `here` / `infeasibles` stores (src/aira.cpp:1918-1933) that this pool produced since the last call (export), and
 * records produced elsewhere (import).  A record is a fact about the model (src/result.h:10-20) whoever solved it, so the
 * cache scan (src/solutions.cpp:11-81) stays exact.  Callable from any thread during a run; no-ops between runs. */
int moip_pool_export_records(moip_pool* p, int cap, double* ip /* cap*k */, int* result /* cap*k */,
                             int* infeasible /* cap */, int* n_out);
int moip_pool_import_records(moip_pool* p, int n, const double* ip, const int* result, const int* infeasible);
int moip_pool_exchange_counts(const moip_pool* p, int64_t* exported, int64_t* imported);
/* Strips cut off busy strips by idle workers so far.  moip_pool_run_strips* starts with the strips it is given and lets a
 * worker that finds none left take over the far half of the widest range a busy strip has not reached yet: a strip is
 * only a range of the last objective (src/aira.cpp:1895-1916), so the front is the same (MOIP_NO_STEAL=1 disables). */
int64_t moip_pool_strips_stolen(const moip_pool* p);
/* boxes put back because their first subproblem ran into its node budget (moip_pool_run_boxes_claim), since pool creation */
int64_t moip_pool_boxes_postponed(const moip_pool* p);
int moip_pool_get_limit(moip_pool* p, int obj, int sense, const double* rhs, int* result, int* mip_status);
/* split_optimise (src/aira.cpp:1886-1943) for nstrips explicit (start, stop) pairs, dealt dynamically to
 * the workers; rows_out receives the feasible result vectors (k ints per row, unsorted) */
int moip_pool_run_strips(moip_pool* p, int n_obj, int nstrips, const double* start_stop, int* rows_out, int cap,
                         int* n_rows);
/* same with a caller-supplied strip dispenser (returns the next strip index, anything outside [0, nstrips) ends the
 * calling worker): one global counter can feed the pools of several GPUs / ranks */
typedef int (*moip_claim_fn)(void* user);
int moip_pool_run_strips_claim(moip_pool* p, int n_obj, int nstrips, const double* start_stop, moip_claim_fn claim,
                               void* user, int* rows_out, int cap, int* n_rows);
/* the same with a window on objective 1 per entry (moip_worker::window): windows[2t], windows[2t+1] = near / far edge of
 * box t; boxes that share a range of the last objective split its start-up -- the (n_obj-1)-objective front every strip
 * begins with -- between them.  n_obj >= 3 (ignored below). */
int moip_pool_run_boxes_claim(moip_pool* p, int n_obj, int nboxes, const double* start_stop, const double* windows,
                              moip_claim_fn claim, void* user, int* rows_out, int cap, int* n_rows);
/* main() with --split -t num_threads (src/aira.cpp:269-276, :1945-1990): every level's strips run
 * concurrently on the pool; rows_out = sorted, de-duplicated front */
int moip_pool_pareto_front(moip_pool* p, int num_threads, int split_normal, int* rows_out, int cap, int* n_rows);

/* ---- cooperative ("synergistic") workers: main() with -t W and no --split (src/aira.cpp:277-308), with a defined,
 * race-free replacement for the bound-sharing cells of src/aira.cpp:923-1086 / :1111-1552.  Worker i runs
 * the sequential generator with a permutation whose last objective it OWNS; when the top-level bound of its last
 * stage moves, it publishes that bound as a monotone limit (one 64-bit atomic per objective) and every other worker
 * intersects its subproblems with it.  No waiting, no locks; a stale limit only costs redundant work (csrc/generator.cpp). */
/* W <= k workers, worker i = i-th rotation of the identity permutation (owns objective (k-1-i) mod k) */
int moip_coop_workers(int k, int n_workers, moip_worker* out);
/* host-logic hook like moip_optimise_with: the W workers run on W host threads and call solve/find/insert with
 * users[i] (the callbacks must be thread-safe); n_solves[i] / n_skipped[i] = subproblems worker i handed to solve /
 * answered as infeasible because a partner had finished */
int moip_coop_optimise_with(int k, int sense, int n_workers, const moip_worker* workers, moip_solve_fn solve,
                            moip_find_cb find, moip_insert_cb insert, void* const* users, int64_t* n_solves,
                            int64_t* n_skipped);
/* the limits as a handle, for one worker per process (one rank per GPU): a host thread keeps the local handle in step with
 * the other ranks; publishing is fetch-min / fetch-max, so remote values may be merged in any order, any number of times.
 * owned_mask bit j: some worker of the job owns objective j.  value: right-hand side of f_obj <= value (MIN) / >= (MAX). */
typedef struct moip_coop moip_coop;
int moip_coop_create(int k, int sense, int owned_mask, moip_coop** out);
void moip_coop_destroy(moip_coop* h);
int moip_coop_publish(moip_coop* h, int obj, long long value, int done);
int moip_coop_read(const moip_coop* h, int obj, long long* value, int* state /* 0 none, 1 value, 2 owner done */);
/* one cooperative worker (w->perm's last objective is the one it owns) on a solver context / on callbacks */
int moip_coop_optimise(moip_ctx* c, const moip_worker* w, moip_coop* shared, moip_cache* all, moip_cache* infeasibles);
int moip_coop_optimise_one_with(int k, int sense, const moip_worker* w, moip_coop* shared, moip_solve_fn solve,
                                moip_find_cb find, moip_insert_cb insert, void* user, int64_t* n_solves, int64_t* n_skipped);
/* the product path: worker i on solver context i of the pool (one GPU), shared infeasible records, per-worker solution
 * records; rows_out = sorted, de-duplicated front.  n_workers is clipped to min(k, pool workers). */
int moip_pool_synergistic_front(moip_pool* p, int n_workers, int* rows_out, int cap, int* n_rows);

const char* moip_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MOIP_B200_H */
