// Chained B&B rounds: the tree of ONE single-objective IP advances on the device, round after round, without the host
// in between (replaces the per-round copy-in / synchronise / copy-out of moip_ctx::solve_ip, what CPXmipopt's tree does
// inside one call -- reference src/aira.cpp:480).
//
//   round r  =  [ K2 propagate ] -> K1 node LPs -> [ K4 round/verify ]  ->  K5 bb_advance      (stream order = round barrier)
//
// The open nodes of round r sit in pool rows  (r & 1) * qcap + [0, count[r & 1]) ; K5 reads what K1/K4 left for them,
// prunes against the incumbent (kept on the device, lowered by K1/K4/K2 with atomicMin while the round runs), creates the
// children in the rows of the other parity and publishes the next count.  The host enqueues a few rounds at a time --
// launches of rounds that find the tree exhausted return at once -- and looks at the control block once per chunk.
// Breadth first, every open node every round: when a level outgrows qcap (or no incumbent turns up) the control block says
// so and the host solves that IP with its own round loop instead.
#pragma once
#include <climits>

#include "device.h"
#include "nodepool.h"

namespace moip {

struct alignas(16) BbCtl {
  // ---- the IP (constant while it is solved)
  long long olo[MOIP_MAX_OBJ], ohi[MOIP_MAX_OBJ];   // integer limits of the objective rows (model space)
  double rhs[MOIP_MAX_OBJ];                          // the same rows as the node LPs see them (+-1e20 = free)
  int cost, sense, qcap, bmax, levels_max, pad0;
  // ---- state
  long long plo[MOIP_MAX_OBJ], phi[MOIP_MAX_OBJ];   // olo/ohi with the incumbent cut-off row: what K2 propagates against
  long long inc_val;                                 // min-form incumbent value, LLONG_MAX = none
  long long inc_seen;                                // the value whose point d_inc holds (or that the host brought along)
  double cutoff;                                     // (double)inc_val or +inf; K1 reads it while it runs
  int count[2];                                      // open nodes per parity
  int work_counter, ticket, overflow, rounds_done, root_solved, rounds_live;   // rounds_live: rounds that had nodes
  unsigned long long n_nodes, n_lps, n_iters, n_solved, n_capped, n_children;
};

struct BbInit {                 // by value to bb_init_kernel
  long long olo[MOIP_MAX_OBJ], ohi[MOIP_MAX_OBJ];
  double rhs[MOIP_MAX_OBJ];
  long long inc_val;
  int cost, sense, qcap, bmax, levels_max, warm;
  const double* root_x;        // [n] / [m] warm start of the root (warm != 0)
  const double* root_y;
};

struct BbRound {                // by value to bb_advance_kernel: what K1 / K2 / K4 left for the nodes of this round
  int parity, round, cold_root;
  const int* flag;              // [B] 0 open / 1 infeasible / 2 leaf
  const int* status;            // [B] MOIP_LP_*
  const int* iters;             // [B]
  const int* branch;            // [B][3]
  const double* bval;           // [B][3]
  const int* ff;                // [B][3] first unfixed column, its lb, ub
  const double* dbound;         // [B]
  const long long* leaf;        // [B][k]
  const long long* cobj;        // [B][3][k]
  const unsigned char* cfeas;   // [B][3]
  const int* xr;                // [B][3][n]
  double* node_bound;           // [2 qcap] bound inherited from the parent
  int* node_depth;              // [2 qcap]
  int* inc_x;                   // [n] the incumbent's point
  double* root_x;               // [n] / [m]: the root's iterate is kept as warm start of the next IP on this objective
  double* root_y;
};

int launch_bb_init(const DevModel& dm, const PoolView& pool, BbCtl* ctl, const BbInit& init, double* node_bound,
                   int* node_depth, cudaStream_t st);
int launch_bb_advance(const DevModel& dm, const PoolView& pool, BbCtl* ctl, const BbRound& r, int grid_hint, cudaStream_t st);

}  // namespace moip
