// K5 -- the host's part of a B&B round, on the device (bbchain.h): incumbent bookkeeping, pruning, branching and the
// queue of the next round.  Integer arithmetic and copies; fp64 only in the bound comparison that the host loop makes too
// (moip_ctx::solve_ip: prunable()).  One CTA per node of the finished round.
#include <cfloat>
#include <climits>
#include <cmath>

#include "bbchain.h"

namespace moip {
namespace {

constexpr int kAdvThreads = 256;

__global__ void __launch_bounds__(256) bb_init_kernel(const DevModel dm, const PoolView pool, BbCtl* ctl, const BbInit in,
                                                      double* node_bound, int* node_depth) {
  const int n = dm.n, m = dm.m, tid = threadIdx.x;
  // the root node: pool row 0 (parity 0, position 0)
  for (int j = tid; j < n; j += blockDim.x) {
    pool.lb[j] = dm.lbI[j];
    pool.ub[j] = dm.ubI[j];
    pool.wx[j] = in.warm ? in.root_x[j] : 0.0;
  }
  for (int i = tid; i < m; i += blockDim.x) pool.wy[i] = in.warm ? in.root_y[i] : 0.0;
  if (tid == 0) {
    for (int o = 0; o < MOIP_MAX_OBJ; ++o) {
      ctl->olo[o] = in.olo[o]; ctl->ohi[o] = in.ohi[o]; ctl->rhs[o] = in.rhs[o];
      ctl->plo[o] = in.olo[o]; ctl->phi[o] = in.ohi[o];
    }
    ctl->cost = in.cost; ctl->sense = in.sense; ctl->qcap = in.qcap; ctl->bmax = in.bmax; ctl->levels_max = in.levels_max;
    ctl->inc_val = in.inc_val; ctl->inc_seen = in.inc_val;
    if (in.inc_val != LLONG_MAX) {       // incumbent cut-off row on the optimised objective: strictly better
      if (in.sense == 0) ctl->phi[in.cost] = min(ctl->phi[in.cost], in.inc_val - 1);
      else ctl->plo[in.cost] = max(ctl->plo[in.cost], -in.inc_val + 1);
    }
    ctl->cutoff = in.inc_val != LLONG_MAX ? (double)in.inc_val : HUGE_VAL;
    ctl->count[0] = 1; ctl->count[1] = 0;
    ctl->work_counter = 0; ctl->ticket = 0; ctl->overflow = 0; ctl->rounds_done = 0; ctl->root_solved = 0; ctl->rounds_live = 0;
    ctl->n_nodes = 0; ctl->n_lps = 0; ctl->n_iters = 0; ctl->n_solved = 0; ctl->n_capped = 0; ctl->n_children = 0;
    node_bound[0] = -HUGE_VAL;
    node_depth[0] = 0;
  }
}

__device__ __forceinline__ bool within_limits(const BbCtl* ctl, const long long* ov, int k) {
  for (int o = 0; o < k; ++o)
    if (ov[o] < ctl->olo[o] || ov[o] > ctl->ohi[o]) return false;
  return true;
}

__global__ void __launch_bounds__(kAdvThreads) bb_advance_kernel(const DevModel dm, const PoolView pool, BbCtl* ctl,
                                                                 const BbRound R) {
  __shared__ int s_src, s_nv, s_base, s_depth, s_var[3], s_fl[3], s_last;
  __shared__ double s_lbnd;
  __shared__ unsigned long long s_stat[5];
  const int n = dm.n, m = dm.m, k = dm.k, tid = threadIdx.x;
  const int p = R.parity, Q = ctl->qcap, B = ctl->count[p], cost = ctl->cost;
  const long long sgn = (long long)dm.sgn;
  const long long inc = ctl->inc_val;          // K1/K2/K4 of this round are done: final until the next round runs
  if (tid < 5) s_stat[tid] = 0;
  __syncthreads();

  // ---- block 0: fetch the point of a new incumbent before the next round overwrites the candidates
  if (blockIdx.x == 0 && B > 0) {
    const long long seen = ctl->inc_seen;
    if (tid == 0) s_src = INT_MAX;
    __syncthreads();                            // everybody has read inc_seen
    if (inc < seen) {
      int src = INT_MAX;                        // node * 4 + (candidate | 3 = leaf)
      for (int i = tid; i < B && src == INT_MAX; i += kAdvThreads) {
        const int f = R.flag[i];
        if (f == 2) { if (sgn * R.leaf[(size_t)i * k + cost] == inc) src = i * 4 + 3; }
        else if (f == 0) {
          for (int c3 = 0; c3 < 3 && src == INT_MAX; ++c3) {
            const size_t w3 = (size_t)i * 3 + c3;
            if (R.cfeas[w3] && within_limits(ctl, R.cobj + w3 * k, k) && sgn * R.cobj[w3 * k + cost] == inc) src = i * 4 + c3;
          }
        }
      }
      if (src != INT_MAX) atomicMin(&s_src, src);
      __syncthreads();
      src = s_src;
      if (src == INT_MAX) { if (tid == 0) ctl->overflow = 2; }         // cannot happen; the host then solves the IP itself
      else {
        const int i = src >> 2, c3 = src & 3;
        const int* from = c3 == 3 ? pool.lb + (size_t)(p * Q + i) * n : R.xr + ((size_t)i * 3 + c3) * n;
        for (int j = tid; j < n; j += kAdvThreads) R.inc_x[j] = from[j];
        if (tid == 0) ctl->inc_seen = inc;
      }
    }
    __syncthreads();
  }

  // multi-level expansion while the device is under-filled (solve_ip's rule with the round size in place of the number of
  // nodes that will branch): rounds are latency-bound, idle SMs are free
  int levels = 1;
  while (levels < ctl->levels_max && (long long)B * (2LL << levels) <= (long long)ctl->bmax) ++levels;

  for (int i = blockIdx.x; i < B; i += gridDim.x) {
    const size_t slot = (size_t)p * Q + i;
    if (tid == 0) {
      int nv = 0;
      const int f = R.flag[i];
      s_stat[0] += 1;
      if (f == 0) {
        s_stat[1] += 1; s_stat[2] += (unsigned long long)R.iters[i];
        if (!(R.round == 0 && R.cold_root)) { s_stat[3] += 1; s_stat[4] += R.status[i] == MOIP_LP_ITERLIMIT; }
        const int st = R.status[i];
        const double lbnd = fmax(R.node_bound[slot], R.dbound[i]);
        const double lceil = ceil(lbnd - 1e-6);
        bool open = st != MOIP_LP_CUTOFF && st != MOIP_LP_INFEASIBLE;
        if (open && inc != LLONG_MAX && lbnd > -HUGE_VAL && lceil >= (double)inc) open = false;      // prunable()
        if (open && R.branch[3 * i] < 0) {       // LP point integral: solved when its own rounding reaches the bound
          long long cand = LLONG_MAX;
          for (int c3 = 0; c3 < 3; ++c3) {
            const size_t w3 = (size_t)i * 3 + c3;
            if (R.cfeas[w3] && within_limits(ctl, R.cobj + w3 * k, k)) cand = min(cand, sgn * R.cobj[w3 * k + cost]);
          }
          if (cand != LLONG_MAX && (double)cand <= lceil) open = false;
        }
        if (open) {
          for (int q = 0; q < levels; ++q)
            if (R.branch[3 * i + q] >= 0) { s_var[nv] = R.branch[3 * i + q]; s_fl[nv] = (int)floor(R.bval[3 * i + q]); ++nv; }
          if (nv == 0 && R.ff[3 * i] >= 0) {     // integral but not accepted: first unfixed column at its midpoint
            s_var[0] = R.ff[3 * i];
            s_fl[0] = (int)floor(0.5 * ((double)R.ff[3 * i + 1] + (double)R.ff[3 * i + 2]));
            nv = 1;
          }
        }
        if (nv > 0) {
          const int nchild = 1 << nv;
          const int base = atomicAdd(&ctl->count[p ^ 1], nchild);
          if (base + nchild > Q) { ctl->overflow = 1; nv = 0; }
          s_base = base; s_lbnd = lbnd; s_depth = R.node_depth[slot];
        }
        if (R.round == 0 && i == 0) ctl->root_solved = 1;
      }
      s_nv = nv;
    }
    __syncthreads();
    if (R.round == 0 && i == 0 && R.flag[0] == 0 && R.root_x) {     // warm start of the next IP on this objective
      for (int j = tid; j < n; j += kAdvThreads) R.root_x[j] = pool.wx[slot * n + j];
      for (int r = tid; r < m; r += kAdvThreads) R.root_y[r] = pool.wy[slot * m + r];
    }
    const int nv = s_nv;
    if (nv > 0) {
      const int nchild = 1 << nv;
      const int v0 = s_var[0], v1 = nv > 1 ? s_var[1] : -1, v2 = nv > 2 ? s_var[2] : -1;
      const int f0 = s_fl[0], f1 = nv > 1 ? s_fl[1] : 0, f2 = nv > 2 ? s_fl[2] : 0;
      const size_t cbase = (size_t)(p ^ 1) * Q + s_base;
      for (int j = tid; j < n; j += kAdvThreads) {
        const int a = pool.lb[slot * n + j], b2 = pool.ub[slot * n + j];
        const double x = pool.wx[slot * n + j];
        const int which = j == v0 ? 0 : (j == v1 ? 1 : (j == v2 ? 2 : -1));
        const int fl = which == 0 ? f0 : (which == 1 ? f1 : f2);
        for (int cmb = 0; cmb < nchild; ++cmb) {
          int ca = a, cb = b2;
          if (which >= 0) {
            if ((cmb >> which) & 1) ca = max(a, fl + 1); else cb = min(b2, fl);
          }
          pool.lb[(cbase + cmb) * n + j] = ca;
          pool.ub[(cbase + cmb) * n + j] = cb;
          pool.wx[(cbase + cmb) * n + j] = x;
        }
      }
      for (int r = tid; r < m; r += kAdvThreads) {
        const double y = pool.wy[slot * m + r];
        for (int cmb = 0; cmb < nchild; ++cmb) pool.wy[(cbase + cmb) * m + r] = y;
      }
      if (tid < nchild) { R.node_bound[cbase + tid] = s_lbnd; R.node_depth[cbase + tid] = s_depth + nv; }
    }
    __syncthreads();
  }

  // ---- statistics, then the last CTA closes the round
  if (tid == 0) {
    if (s_stat[0]) {
      atomicAdd(&ctl->n_nodes, s_stat[0]); atomicAdd(&ctl->n_lps, s_stat[1]); atomicAdd(&ctl->n_iters, s_stat[2]);
      atomicAdd(&ctl->n_solved, s_stat[3]); atomicAdd(&ctl->n_capped, s_stat[4]);
    }
    __threadfence();
    s_last = atomicAdd(&ctl->ticket, 1) == (int)gridDim.x - 1;
    if (s_last) {
      __threadfence();
      ctl->count[p] = 0;
      ctl->ticket = 0;
      ctl->work_counter = 0;
      ctl->rounds_done = R.round + 1;
      if (B > 0) ctl->rounds_live = R.round + 1;
      if (ctl->overflow) ctl->count[p ^ 1] = 0;           // the rounds already enqueued find nothing to do
      for (int o = 0; o < MOIP_MAX_OBJ; ++o) { ctl->plo[o] = ctl->olo[o]; ctl->phi[o] = ctl->ohi[o]; }
      if (inc != LLONG_MAX) {
        if (ctl->sense == 0) ctl->phi[cost] = min(ctl->phi[cost], inc - 1);
        else ctl->plo[cost] = max(ctl->plo[cost], -inc + 1);
      }
      ctl->cutoff = inc != LLONG_MAX ? (double)inc : HUGE_VAL;
    }
  }
}

}  // namespace

int launch_bb_init(const DevModel& dm, const PoolView& pool, BbCtl* ctl, const BbInit& init, double* node_bound,
                   int* node_depth, cudaStream_t st) {
  bb_init_kernel<<<1, 256, 0, st>>>(dm, pool, ctl, init, node_bound, node_depth);
  MOIP_CUDA(cudaGetLastError());
  return MOIP_OK;
}

int launch_bb_advance(const DevModel& dm, const PoolView& pool, BbCtl* ctl, const BbRound& r, int grid_hint, cudaStream_t st) {
  static LaunchCfg carve;
  if (set_aux_carveout(bb_advance_kernel, carve)) return MOIP_ERR_CUDA;
  int grid = grid_hint < 1 ? 1 : grid_hint;
  const int cap = 148 * 2;
  if (grid > cap) grid = cap;
  bb_advance_kernel<<<grid, kAdvThreads, 0, st>>>(dm, pool, ctl, r);
  MOIP_CUDA(cudaGetLastError());
  return MOIP_OK;
}

}  // namespace moip
