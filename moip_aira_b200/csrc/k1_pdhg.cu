// K1 -- batched node-LP relaxations (the work hidden inside CPXmipopt in the reference,
// src/aira.cpp:480): one CTA per B&B node runs a reflected, restarted Halpern PDHG
// (r2HPDHG) in fp64 on the scaled LP
//
//      min c^T x   s.t.  lo <= S x <= hi ,  l <= x <= u
//
// with the iterate (x, anchor, reflected point, bounds) resident in shared memory for the
// whole solve; HBM is touched once per node (load bounds / warm start, store result).  The
// matrix is shared by every node of the batch and is read through L1/L2:
//   * S^T y : thread-per-column over a column-ELL image (coalesced, no reduction)
//   * S  x  : the k dense objective-bound rows are accumulated in the column pass
//             (block reduction), the sparse structural rows are warp-per-row CSR with
//             shuffle reductions
//   * projection, reflection, Halpern averaging and the dual step are fused into the two passes.
// Any dual iterate gives a valid Lagrangian bound because every column is boxed; that bound
// (never the primal value) is what B&B prunes with.
#include <cfloat>
#include <cmath>

#include "device.h"

namespace moip {
namespace {

constexpr int kMaxObj = MOIP_MAX_OBJ;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sums NV per-thread values over the CTA; every thread receives the totals.
// Contains one __syncthreads(); `red` must not be reused before another barrier.
template <int NV, int NT>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red, int tid) {
  constexpr int NW = NT / 32;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  if (NW == 1) {
    __syncthreads();
    return;
  }
  if ((tid & 31) == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[(tid >> 5) * NV + i] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += red[w * NV + i];
    v[i] = s;
  }
}

__device__ __forceinline__ double clampd(double v, double a, double b) { return fmin(fmax(v, a), b); }

template <int NT>
__global__ void __launch_bounds__(NT) k1_pdhg_kernel(const DevModel dm, const LpBatch b, const LpParams p) {
  constexpr int NW = NT / 32;
  extern __shared__ double smem[];
  const int n = dm.n, ms = dm.ms, k = dm.k, m = dm.m, ellw = dm.ell_w;
  // Large models (5n + 8m doubles beyond the 227 KB of an SM): the iterate of the node lives in a per-CTA scratch
  // region in HBM / L2 instead of shared memory (streaming mode); only the reduction buffers stay on chip.
  double* base = b.scratch ? b.scratch + (size_t)blockIdx.x * b.scratch_stride : smem;
  double* x = base;
  double* xa = x + n;
  double* xbar = xa + n;
  double* l = xbar + n;
  double* u = l + n;
  double* y = u + n;
  double* ya = y + m;
  double* yt = ya + m;
  double* sx = yt + m;
  double* sxa = sx + m;
  double* sxt = sxa + m;
  double* lo = sxt + m;
  double* hi = lo + m;
  double* redA = b.scratch ? smem : hi + m;        // NW*8
  double* redB = redA + NW * 8; // NW*8
  double* redC = redB + NW * 8; // NW*8
  __shared__ int s_node;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double eta = dm.eta;

  for (;;) {
    if (tid == 0) s_node = atomicAdd(b.work_counter, 1);
    __syncthreads();
    const int node = s_node;
    if (node >= b.B) break;
    if (b.skip && b.skip[node]) {
      if (tid == 0) { b.status[node] = -1; b.iters[node] = 0; }
      __syncthreads();
      continue;
    }

    // ---------------------------------------------------------------- node load
    const size_t srow = b.slot ? (size_t)b.slot[node] : (size_t)node;   // row of the node's arrays
    const int cost = b.cost_idx[(size_t)node * b.cost_stride];
    const double* nrhs = b.rhs + (size_t)node * b.rhs_stride;
    const double inv_dr_cost = 1.0 / dm.dr[ms + cost];
    const double* Dc = dm.D + (size_t)cost * n;
    unsigned act = 0;
    for (int o = 0; o < k; ++o)
      if (fabs(nrhs[o]) < 1e19) act |= 1u << o;
    double a0[6] = {0, 0, 0, 0, 0, 0};  // |c|^2, obj_upper, D.x per objective row
    for (int j = tid; j < n; j += NT) {
      const double idc = 1.0 / dm.dc[j];
      const double lj = (double)b.lb[srow * n + j] * idc;
      const double uj = (double)b.ub[srow * n + j] * idc;
      double xj = b.warm_x ? b.warm_x[srow * n + j] * idc : 0.0;
      xj = clampd(xj, lj, uj);
      l[j] = lj; u[j] = uj; x[j] = xj; xa[j] = xj;
      const double cj = Dc[j] * inv_dr_cost;
      a0[0] += cj * cj;
      a0[1] += fmax(cj * lj, cj * uj);
#pragma unroll
      for (int o = 0; o < kMaxObj; ++o)
        if (o < k) a0[2 + o] += dm.D[(size_t)o * n + j] * xj;
    }
    for (int i = tid; i < m; i += NT) {
      double loi, hii;
      if (i < ms) { loi = dm.s_lo[i]; hii = dm.s_hi[i]; }
      else {
        const int o = i - ms;
        loi = -HUGE_VAL;
        hii = ((act >> o) & 1u) ? dm.sgn * nrhs[o] * dm.dr[i] : HUGE_VAL;
      }
      double yi = b.warm_y ? b.warm_y[srow * m + i] / dm.dr[i] : 0.0;
      if (loi == -HUGE_VAL) yi = fmin(yi, 0.0);
      if (hii == HUGE_VAL) yi = fmax(yi, 0.0);
      lo[i] = loi; hi[i] = hii; y[i] = yi; ya[i] = yi;
    }
    block_sum<6, NT>(a0, redA, tid);   // barrier: x, l, u, y, lo, hi visible
    const double obj_upper = a0[1];
    // S x0 : structural rows (warp per row) + dense rows (from the column pass)
    for (int i = warp; i < ms; i += NW) {
      double q = 0;
      for (int e = dm.s_ptr[i] + lane; e < dm.s_ptr[i + 1]; e += 32) q += dm.s_val[e] * x[dm.s_col[e]];
      q = warp_sum(q);
      if (lane == 0) { sx[i] = q; sxa[i] = q; }
    }
    if (tid < k) { sx[ms + tid] = a0[2 + tid]; sxa[ms + tid] = a0[2 + tid]; }
    // primal weight  w = |c| / |b|  (scaled space), unscaled |b| for the KKT denominator
    double bn2 = 0, bn2_unscaled = dm.norm_row_bounds2;
    for (int i = 0; i < m; ++i) {      // m is small; every thread computes the same value
      const double loi = (i < ms) ? dm.s_lo[i] : -HUGE_VAL;
      double hii;
      if (i < ms) hii = dm.s_hi[i];
      else {
        const int o = i - ms;
        hii = ((act >> o) & 1u) ? dm.sgn * nrhs[o] * dm.dr[i] : HUGE_VAL;
        if ((act >> o) & 1u) { const double r = nrhs[o]; bn2_unscaled += r * r; }
      }
      const double t = (hii != HUGE_VAL) ? hii : ((loi != -HUGE_VAL) ? loi : 0.0);
      bn2 += t * t;
    }
    double w = (a0[0] > 0 && bn2 > 0) ? sqrt(a0[0] / bn2) : 1.0;
    double tau = eta / w, sigma = eta * w;
    const double kkt_bden = 1.0 + sqrt(bn2_unscaled);
    __syncthreads();

    int kk = 0, it = 0, status = MOIP_LP_ITERLIMIT;
    double r0 = 0, rprev = -1.0, best_lb = -HUGE_VAL, pobj = 0;
    const int iter_cap = p.fixed_iters > 0 ? p.fixed_iters : p.max_iter;

    // ---------------------------------------------------------------- PDHG iterations
    for (;;) {
      ++it;
      const bool norm_it = (kk == 0) || (kk % p.norm_every == 0);
      // pass A: columns.  g = S^T y ; xt = proj(x - tau (c - g)) ; xbar = 2 xt - x ; dense rows of S xbar
      double aA[1 + kMaxObj] = {0, 0, 0, 0, 0};
      for (int j = tid; j < n; j += NT) {
        double g = 0;
        for (int e = 0; e < ellw; ++e) g += dm.ellT_val[(size_t)e * n + j] * y[dm.ellT_row[(size_t)e * n + j]];
        double dj[kMaxObj];
#pragma unroll
        for (int o = 0; o < kMaxObj; ++o) {
          dj[o] = 0;
          if (o < k && ((act >> o) & 1u)) { dj[o] = dm.D[(size_t)o * n + j]; g += dj[o] * y[ms + o]; }
        }
        const double cj = Dc[j] * inv_dr_cost;
        const double xj = x[j];
        const double xtj = clampd(xj - tau * (cj - g), l[j], u[j]);
        const double xb = 2.0 * xtj - xj;
        xbar[j] = xb;
        const double d = xtj - xj;
        aA[0] += d * d;
#pragma unroll
        for (int o = 0; o < kMaxObj; ++o) aA[1 + o] += dj[o] * xb;
      }
      block_sum<1 + kMaxObj, NT>(aA, redA, tid);   // barrier: xbar visible
      // pass B: rows.  q = S xbar ; yt = dual step ; S xt by linearity
      double aB[2] = {0, 0};
      auto row_update = [&](int i, double q) {
        const double sxi = sx[i], yi = y[i];
        const double sxti = 0.5 * (q + sxi);
        const double v = yi / sigma - q;
        const double yti = sigma * (v - clampd(v, -hi[i], -lo[i]));
        yt[i] = yti; sxt[i] = sxti;
        const double dy = yti - yi;
        aB[0] += dy * dy;
        aB[1] += dy * (sxti - sxi);
      };
      for (int i = warp; i < ms; i += NW) {
        double q = 0;
        for (int e = dm.s_ptr[i] + lane; e < dm.s_ptr[i + 1]; e += 32) q += dm.s_val[e] * xbar[dm.s_col[e]];
        q = warp_sum(q);
        if (lane == 0) row_update(i, q);
      }
      if (tid < k) {
        if ((act >> tid) & 1u) row_update(ms + tid, aA[1 + tid]);
        else { yt[ms + tid] = 0.0; sxt[ms + tid] = 0.0; }
      }
      bool restart = false;
      if (norm_it) {
        block_sum<2, NT>(aB, redB, tid);            // barrier: yt, sxt visible
        const double fp = fmax(0.0, (w / eta) * aA[0] - 2.0 * aB[1] + aB[0] / (eta * w));   // squared
        if (kk == 0) r0 = fp;
        else if (fp <= 0.04 * r0 || (fp <= 0.64 * r0 && rprev >= 0.0 && fp > rprev) || 25 * kk >= 9 * it)
          restart = true;
        rprev = fp;
      } else {
        __syncthreads();
      }
      // ---- termination tests at (xt, yt)
      bool stop = false;
      if (p.fixed_iters <= 0 && (it % p.check_every) == 0) {
        double aC[4] = {0, 0, 0, 0};   // pobj, dual (columns), dual (rows), primal residual^2 (unscaled)
        for (int j = tid; j < n; j += NT) {
          double g = 0;
          for (int e = 0; e < ellw; ++e) g += dm.ellT_val[(size_t)e * n + j] * yt[dm.ellT_row[(size_t)e * n + j]];
#pragma unroll
          for (int o = 0; o < kMaxObj; ++o)
            if (o < k && ((act >> o) & 1u)) g += dm.D[(size_t)o * n + j] * yt[ms + o];
          const double cj = Dc[j] * inv_dr_cost;
          const double r = cj - g;
          aC[0] += cj * 0.5 * (xbar[j] + x[j]);
          aC[1] += (r > 0) ? r * l[j] : r * u[j];
        }
        for (int i = tid; i < m; i += NT) {
          const double yi = yt[i];
          if (yi > 0) aC[2] += yi * lo[i];
          else if (yi < 0) aC[2] += yi * hi[i];
          const double s = sxt[i];
          const double viol = fmax(0.0, fmax(s - hi[i], lo[i] - s)) / dm.dr[i];
          if (i < ms || ((act >> (i - ms)) & 1u)) aC[3] += viol * viol;
        }
        block_sum<4, NT>(aC, redA, tid);
        pobj = aC[0];
        const double dobj = aC[1] + aC[2];
        if (dobj > best_lb) best_lb = dobj;
        const double gap = fabs(pobj - dobj);
        const double rel = fmax(sqrt(aC[3]) / kkt_bden, gap / (1.0 + fabs(pobj) + fabs(dobj)));
        const double cutoff = b.cutoff ? *((volatile const double*)b.cutoff) : HUGE_VAL;
        if (best_lb >= cutoff - p.cutoff_slack) { status = MOIP_LP_CUTOFF; stop = true; }
        else if (best_lb > obj_upper + 1e-6 * (1.0 + fabs(obj_upper))) { status = MOIP_LP_INFEASIBLE; stop = true; }
        else if (rel <= p.eps) { status = MOIP_LP_CONVERGED; stop = true; }
      }
      if (it >= iter_cap) stop = true;
      if (stop) break;
      // ---- restart or Halpern step
      if (restart) {
        double aR[2] = {0, 0};
        for (int j = tid; j < n; j += NT) {
          const double xn = 0.5 * (xbar[j] + x[j]);
          const double d = xn - xa[j];
          aR[0] += d * d;
          x[j] = xn; xa[j] = xn;
        }
        for (int i = tid; i < m; i += NT) {
          const double d = yt[i] - ya[i];
          aR[1] += d * d;
          y[i] = yt[i]; ya[i] = yt[i]; sx[i] = sxt[i]; sxa[i] = sxt[i];
        }
        block_sum<2, NT>(aR, redC, tid);
        const double dxn = sqrt(aR[0]), dyn = sqrt(aR[1]);
        if (dxn > 1e-10 && dyn > 1e-10) w = exp(0.5 * log(dyn / dxn) + 0.5 * log(w));
        tau = eta / w; sigma = eta * w;
        kk = 0; rprev = -1.0;
      } else {
        const double a = (double)(kk + 1) / (double)(kk + 2), c1 = 1.0 - a;
        for (int j = tid; j < n; j += NT) x[j] = a * xbar[j] + c1 * xa[j];
        for (int i = tid; i < m; i += NT) {
          y[i] = a * (2.0 * yt[i] - y[i]) + c1 * ya[i];
          sx[i] = a * (2.0 * sxt[i] - sx[i]) + c1 * sxa[i];
        }
        ++kk;
        __syncthreads();
      }
    }

    // ---------------------------------------------------------------- node store
    if (p.fixed_iters > 0) {   // fixed-iteration runs report the objective / bound of the last iterate
      double aC[3] = {0, 0, 0};
      for (int j = tid; j < n; j += NT) {
        double g = 0;
        for (int e = 0; e < ellw; ++e) g += dm.ellT_val[(size_t)e * n + j] * yt[dm.ellT_row[(size_t)e * n + j]];
        for (int o = 0; o < k; ++o)
          if ((act >> o) & 1u) g += dm.D[(size_t)o * n + j] * yt[ms + o];
        const double cj = Dc[j] * inv_dr_cost;
        const double r = cj - g;
        aC[0] += cj * 0.5 * (xbar[j] + x[j]);
        aC[1] += (r > 0) ? r * l[j] : r * u[j];
      }
      for (int i = tid; i < m; i += NT) {
        const double yi = yt[i];
        if (yi > 0) aC[2] += yi * lo[i];
        else if (yi < 0) aC[2] += yi * hi[i];
      }
      block_sum<3, NT>(aC, redA, tid);
      pobj = aC[0];
      best_lb = aC[1] + aC[2];
    }
    // most fractional column of the unscaled primal iterate
    double bestf = -1.0; int bestj = -1;
    for (int j = tid; j < n; j += NT) {
      const double v = 0.5 * (xbar[j] + x[j]) * dm.dc[j];
      if (b.out_x) b.out_x[srow * n + j] = v;
      const double f = fabs(v - rint(v));
      if (f > bestf) { bestf = f; bestj = j; }
    }
    if (b.out_y)
      for (int i = tid; i < m; i += NT) b.out_y[srow * m + i] = yt[i] * dm.dr[i];
    if (b.branch_var) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double of = __shfl_xor_sync(0xffffffffu, bestf, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bestj, o);
        if (of > bestf || (of == bestf && oj >= 0 && (bestj < 0 || oj < bestj))) { bestf = of; bestj = oj; }
      }
      __syncthreads();
      if (lane == 0) { redA[warp * 2] = bestf; redA[warp * 2 + 1] = (double)bestj; }
      __syncthreads();
      if (tid == 0) {
        for (int wq = 1; wq < NW; ++wq) {
          const double of = redA[wq * 2]; const int oj = (int)redA[wq * 2 + 1];
          if (of > bestf || (of == bestf && oj >= 0 && (bestj < 0 || oj < bestj))) { bestf = of; bestj = oj; }
        }
        b.branch_var[(size_t)node * 3] = (bestf > 1e-6) ? bestj : -1;
        b.branch_var[(size_t)node * 3 + 1] = -1; b.branch_var[(size_t)node * 3 + 2] = -1;
        if (b.branch_val) b.branch_val[(size_t)node * 3] = (bestj >= 0) ? 0.5 * (xbar[bestj] + x[bestj]) * dm.dc[bestj] : 0.0;
      }
    }
    if (tid == 0) {
      b.primal_obj[node] = pobj;
      b.dual_bound[node] = best_lb;
      b.status[node] = status;
      b.iters[node] = it;
    }
    __syncthreads();
  }
}

// 2-bit/var fixing masks -> integer column bounds of the node (0 free,
// 2 fixed to 0, 3 fixed to 1; for general-integer columns 2 = at lower bound, 3 = at upper bound).
__global__ void expand_masks_kernel(const DevModel dm, int B, const uint32_t* masks, int mask_words, int* lb, int* ub) {
  const size_t total = (size_t)B * dm.n;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int node = (int)(t / dm.n), j = (int)(t % dm.n);
    const uint32_t w = masks ? masks[(size_t)node * mask_words + (j >> 4)] : 0u;
    const uint32_t c = (w >> ((j & 15) * 2)) & 3u;
    int lj = dm.lbI[j], uj = dm.ubI[j];
    if (c == 2u) uj = lj;
    else if (c == 3u) lj = uj;
    lb[t] = lj; ub[t] = uj;
  }
}

template <int NT>
int launch_nt(const DevModel& dm, const LpBatch& b, const LpParams& p, int num_sms, cudaStream_t st) {
  const bool streaming = b.scratch != nullptr;
  const size_t smem = sizeof(double) * ((streaming ? 0 : (size_t)5 * dm.n + (size_t)8 * dm.m) + (size_t)24 * (NT / 32));
  if (!streaming && k1_scratch_stride(dm) != 0) {
    std::fprintf(stderr, "moip_b200: node LP needs the streaming scratch (n=%d) but none was provided\n", dm.n);
    return MOIP_ERR_LIMIT;
  }
  static LaunchCfg cfg;
  std::unique_lock<std::mutex> cfg_lock(launch_cfg_mutex());
  size_t& configured = cfg.configured[current_device()];
  if (smem > configured) {
    MOIP_CUDA(cudaFuncSetAttribute(k1_pdhg_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  cfg_lock.unlock();
  int occ = 1;
  MOIP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k1_pdhg_kernel<NT>, NT, smem));
  if (occ < 1) { std::fprintf(stderr, "moip_b200: node LP does not fit in shared memory (n=%d)\n", dm.n); return MOIP_ERR_LIMIT; }
  if (streaming && occ > 4) occ = 4;
  long long grid = (long long)num_sms * occ;
  if (grid > b.B) grid = b.B;
  if (streaming && grid > b.scratch_slots) grid = b.scratch_slots;
  if (grid < 1) grid = 1;
  k1_pdhg_kernel<NT><<<(unsigned)grid, NT, smem, st>>>(dm, b, p);
  MOIP_CUDA(cudaGetLastError());
  return MOIP_OK;
}

}  // namespace

// doubles of per-CTA scratch the generic kernel needs when the iterate does not fit in shared memory (0 = fits)
size_t k1_scratch_stride(const DevModel& dm) {
  const size_t need = (size_t)5 * dm.n + (size_t)8 * dm.m;
  return (need + 24 * 8) * sizeof(double) > 200 * 1024 ? ((need + 15) & ~(size_t)15) : 0;
}

int launch_k1(const DevModel& dm, const LpBatch& b, const LpParams& p, int num_sms, cudaStream_t st) {
  if (b.B <= 0) return MOIP_OK;
  MOIP_CUDA(cudaMemsetAsync(b.work_counter, 0, sizeof(int), st));
  if (dm.n <= 64) return launch_nt<32>(dm, b, p, num_sms, st);
  if (dm.n <= 256) return launch_nt<64>(dm, b, p, num_sms, st);
  if (dm.n <= 512) return launch_nt<128>(dm, b, p, num_sms, st);
  return launch_nt<256>(dm, b, p, num_sms, st);
}

int launch_expand_masks(const DevModel& dm, int B, const uint32_t* masks, int mask_words, int* lb, int* ub,
                        cudaStream_t st) {
  if (B <= 0) return MOIP_OK;
  const size_t total = (size_t)B * dm.n;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  expand_masks_kernel<<<blocks, 256, 0, st>>>(dm, B, masks, mask_words, lb, ub);
  MOIP_CUDA(cudaGetLastError());
  return MOIP_OK;
}

}  // namespace moip
