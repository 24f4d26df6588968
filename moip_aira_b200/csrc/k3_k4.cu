// K3 -- batched reuse/dominance scan (Solutions::find, reference src/solutions.cpp:11-81)
// K4 -- exact int64 verification of integer points (the round()/Sum c*x of reference
//       src/aira.cpp:517-530, done exactly instead of in fp64).
#include <climits>

#include "device.h"

namespace moip {
namespace {

// One CTA per query; records (64-byte AoS, four 16-byte loads per thread) are visited in
// insertion order in chunks of NT so the first match (lowest index) is found with an early exit.
template <int NT>
__global__ void __launch_bounds__(NT) k3_scan_kernel(const DevCache c0, const DevCache c1, int Q, const double* queries,
                                                     int sense, int* first_match, int* which) {
  __shared__ int s_best;
  const int tid = threadIdx.x;
  const int k = c0.k;
  for (int q = blockIdx.x; q < Q; q += gridDim.x) {
    double ip[MOIP_MAX_OBJ];
#pragma unroll
    for (int i = 0; i < MOIP_MAX_OBJ; ++i) ip[i] = (i < k) ? queries[(size_t)q * k + i] : 0.0;
    int found = -1, found_in = -1;
    for (int st = 0; st < 2 && found < 0; ++st) {
      const DevCache& c = st == 0 ? c0 : c1;
      if (c.size <= 0) continue;
      if (tid == 0) s_best = INT_MAX;
      __syncthreads();
      for (int base = 0; base < c.size; base += NT) {
        const int r = base + tid;
        bool ok = r < c.size;
        if (ok) {
          const int4* p = reinterpret_cast<const int4*>(c.rec + r);
          const int4 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3);
          double rip[4];
          rip[0] = __hiloint2double(v0.y, v0.x); rip[1] = __hiloint2double(v0.w, v0.z);
          rip[2] = __hiloint2double(v1.y, v1.x); rip[3] = __hiloint2double(v1.w, v1.z);
          const int res[4] = {v2.x, v2.y, v2.z, v2.w};
          const bool inf = v3.x != 0;
#pragma unroll
          for (int i = 0; i < MOIP_MAX_OBJ; ++i) {
            if (i < k) {
              if (sense == MOIP_SENSE_MIN) {
                if (rip[i] < ip[i]) ok = false;                         // t1 (src/solutions.cpp:20)
                if (!inf && (double)res[i] > ip[i]) ok = false;         // t3 (:25-30)
              } else {
                if (rip[i] > ip[i]) ok = false;                         // t1 (:35)
                if (!inf && (double)res[i] < ip[i]) ok = false;         // t3 (:40-45)
              }
            }
          }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (bal && (tid & 31) == 0) atomicMin(&s_best, base + (tid & ~31) + (__ffs(bal) - 1));
        __syncthreads();
        const bool hit = s_best != INT_MAX;   // uniform: read between the two barriers
        __syncthreads();
        if (hit) break;
      }
      if (s_best != INT_MAX) { found = s_best; found_in = st; }
      __syncthreads();
    }
    if (tid == 0) { first_match[q] = found; if (which) which[q] = found_in; }
  }
}

// The generator's scan: ONE query against the two stores (infeasibles, then solutions: src/aira.cpp:816-823).  The query
// arrives as a kernel parameter and the answer -- index, store and the matching record itself -- is written straight into
// mapped pinned host memory, published by a sequence number the host polls: no H2D / D2H copy, no stream synchronise.
template <int NT>
__global__ void __launch_bounds__(NT) k3_scan1_kernel(const DevCache c0, const DevCache c1, const K3Query q, const int sense,
                                                      K3Answer* answer, const int seq) {
  __shared__ int s_best;
  const int tid = threadIdx.x;
  const int k = c0.k;
  int found = -1, found_in = -1;
  for (int st = 0; st < 2 && found < 0; ++st) {
    const DevCache& c = st == 0 ? c0 : c1;
    if (c.size <= 0) continue;
    if (tid == 0) s_best = INT_MAX;
    __syncthreads();
    for (int base = 0; base < c.size; base += NT) {
      const int r = base + tid;
      bool ok = r < c.size;
      if (ok) {
        const int4* p = reinterpret_cast<const int4*>(c.rec + r);
        const int4 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2), v3 = __ldg(p + 3);
        double rip[4];
        rip[0] = __hiloint2double(v0.y, v0.x); rip[1] = __hiloint2double(v0.w, v0.z);
        rip[2] = __hiloint2double(v1.y, v1.x); rip[3] = __hiloint2double(v1.w, v1.z);
        const int res[4] = {v2.x, v2.y, v2.z, v2.w};
        const bool inf = v3.x != 0;
#pragma unroll
        for (int i = 0; i < MOIP_MAX_OBJ; ++i) {
          if (i < k) {
            if (sense == MOIP_SENSE_MIN) {
              if (rip[i] < q.ip[i]) ok = false;                         // t1 (src/solutions.cpp:20)
              if (!inf && (double)res[i] > q.ip[i]) ok = false;         // t3 (:25-30)
            } else {
              if (rip[i] > q.ip[i]) ok = false;                         // t1 (:35)
              if (!inf && (double)res[i] < q.ip[i]) ok = false;         // t3 (:40-45)
            }
          }
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (bal && (tid & 31) == 0) atomicMin(&s_best, base + (tid & ~31) + (__ffs(bal) - 1));
      __syncthreads();
      const bool hit = s_best != INT_MAX;   // uniform: read between the two barriers
      __syncthreads();
      if (hit) break;
    }
    if (s_best != INT_MAX) { found = s_best; found_in = st; }
    __syncthreads();
  }
  if (tid < 16 && found >= 0)              // the record itself: the host need not touch the store again
    reinterpret_cast<int*>(&answer->rec)[tid] = reinterpret_cast<const int*>((found_in == 0 ? c0 : c1).rec + found)[tid];
  __syncthreads();
  if (tid == 0) {
    answer->first_match = found;
    answer->which = found_in;
    __threadfence_system();                // answer before its sequence number, all the way to host memory
    *reinterpret_cast<volatile int*>(&answer->seq) = seq;
  }
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per point: column bounds, structural rows (CSR, int64) and the k objectives.
__global__ void __launch_bounds__(128) k4_verify_kernel(const DevModel dm, int B, const int* x, const double* rhs,
                                                        long long* obj_out, unsigned char* feasible_out) {
  const int lane = threadIdx.x & 31;
  const int wglobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int p = wglobal; p < B; p += nwarps) {
    const int* xp = x + (size_t)p * dm.n;
    int bad = 0;
    long long obj[MOIP_MAX_OBJ] = {0, 0, 0, 0};
    for (int j = lane; j < dm.n; j += 32) {
      const int v = xp[j];
      if (v < dm.lbI[j] || v > dm.ubI[j]) bad = 1;
#pragma unroll
      for (int o = 0; o < MOIP_MAX_OBJ; ++o)
        if (o < dm.k) obj[o] += dm.ci[(size_t)o * dm.n + j] * (long long)v;
    }
#pragma unroll
    for (int o = 0; o < MOIP_MAX_OBJ; ++o) obj[o] = warp_sum_ll(obj[o]);
    for (int i = 0; i < dm.ms; ++i) {
      long long a = 0;
      for (int e = dm.s_ptr[i] + lane; e < dm.s_ptr[i + 1]; e += 32) a += dm.ai_val[e] * (long long)xp[dm.s_col[e]];
      a = warp_sum_ll(a);
      if (a < dm.ri_lo[i] || a > dm.ri_hi[i]) bad = 1;
    }
    if (rhs) {
      for (int o = 0; o < dm.k; ++o) {
        const double r = rhs[(size_t)p * dm.k + o];
        if (fabs(r) < 1e19) {
          if (dm.sgn > 0 ? ((double)obj[o] > r) : ((double)obj[o] < r)) bad = 1;
        }
      }
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
      feasible_out[p] = bad ? 0 : 1;
      for (int o = 0; o < dm.k; ++o) obj_out[(size_t)p * dm.k + o] = obj[o];
    }
  }
}

// Rounds a node's LP point three ways (nearest / down / up, clipped to the node's bounds) and verifies
// each candidate exactly; one warp per (node, candidate).  Feeds the incumbent search of the B&B.
__global__ void __launch_bounds__(128) k4_round_verify_kernel(const DevModel dm, int B, const int* slot, const double* wx,
                                                              const int* lb, const int* ub, int* xr, long long* obj_out,
                                                              unsigned char* feasible_out, int* first_free, const int* skip,
                                                              const ChainRef ch) {
  if (ch.B_dev) B = *ch.B_dev;
  const int lane = threadIdx.x & 31;
  const int wglobal = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int n = dm.n;
  for (int w = wglobal; w < B * 3; w += nwarps) {
    const int node = w / 3, mode = w - node * 3;
    if (skip && skip[node]) continue;      // decided by K2 (infeasible / leaf): nobody reads its roundings
    const size_t srow = slot ? (size_t)slot[node] : (size_t)(ch.slot_base + node);
    int* xp = xr + (size_t)w * n;
    long long obj[MOIP_MAX_OBJ] = {0, 0, 0, 0};
    int ff = INT_MAX;                     // first column that is not fixed yet (fallback branching column)
    for (int j = lane; j < n; j += 32) {
      const double v = wx[srow * n + j];
      int r = mode == 0 ? (int)llrint(v) : mode == 1 ? (int)floor(v + 1e-6) : (int)ceil(v - 1e-6);
      const int lj = lb[srow * n + j], uj = ub[srow * n + j];
      if (lj < uj && ff == INT_MAX) ff = j;
      r = max(lj, min(uj, r));
      xp[j] = r;
#pragma unroll
      for (int o = 0; o < MOIP_MAX_OBJ; ++o)
        if (o < dm.k) obj[o] += dm.ci[(size_t)o * n + j] * (long long)r;
    }
#pragma unroll
    for (int o = 0; o < MOIP_MAX_OBJ; ++o) obj[o] = warp_sum_ll(obj[o]);
    __syncwarp();
    int bad = 0;
    for (int i = 0; i < dm.ms; ++i) {
      long long a = 0;
      for (int e = dm.s_ptr[i] + lane; e < dm.s_ptr[i + 1]; e += 32) a += dm.ai_val[e] * (long long)xp[dm.s_col[e]];
      a = warp_sum_ll(a);
      if (a < dm.ri_lo[i] || a > dm.ri_hi[i]) bad = 1;
    }
    if (first_free && mode == 0) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ff = min(ff, __shfl_xor_sync(0xffffffffu, ff, o));
      if (lane == 0) {
        first_free[(size_t)node * 3] = ff == INT_MAX ? -1 : ff;
        first_free[(size_t)node * 3 + 1] = ff == INT_MAX ? 0 : lb[srow * n + ff];
        first_free[(size_t)node * 3 + 2] = ff == INT_MAX ? 0 : ub[srow * n + ff];
      }
    }
    if (lane == 0) {
      feasible_out[w] = bad ? 0 : 1;
      for (int o = 0; o < dm.k; ++o) obj_out[(size_t)w * dm.k + o] = obj[o];
      if (ch.inc && !bad) {                  // chained rounds: a verified candidate within the IP's limits may be the new incumbent
        bool ok = true;
#pragma unroll
        for (int o = 0; o < MOIP_MAX_OBJ; ++o)
          if (o < dm.k) ok = ok && obj[o] >= ch.lim_lo[o] && obj[o] <= ch.lim_hi[o];
        long long v = 0;
        const int cost = *ch.cost;
#pragma unroll
        for (int o = 0; o < MOIP_MAX_OBJ; ++o) if (o == cost) v = (long long)dm.sgn * obj[o];
        if (ok && v < atomicMin(ch.inc, v)) *ch.cutoff = (double)v;
      }
    }
  }
}

}  // namespace

int launch_k4_round(const DevModel& dm, int B, const int* slot, const double* wx, const int* lb, const int* ub, int* xr,
                    long long* obj_out, unsigned char* feasible_out, int* first_free, const int* skip, cudaStream_t st,
                    const ChainRef& ch) {
  if (B <= 0) return MOIP_OK;
  int blocks = (B * 3 + 3) / 4;
  static LaunchCfg carve;
  if (set_aux_carveout(k4_round_verify_kernel, carve)) return MOIP_ERR_CUDA;
  const int cap = aux_grid_cap();
  if (blocks > cap) blocks = cap;
  k4_round_verify_kernel<<<blocks, 128, 0, st>>>(dm, B, slot, wx, lb, ub, xr, obj_out, feasible_out, first_free, skip, ch);
  MOIP_CUDA(cudaGetLastError());
  return MOIP_OK;
}

int launch_k3(const DevCache& c0, const DevCache& c1, int Q, const double* queries, int sense, int* first_match,
              int* which, cudaStream_t st) {
  if (Q <= 0) return MOIP_OK;
  static LaunchCfg carve;
  if (set_aux_carveout(k3_scan_kernel<128>, carve)) return MOIP_ERR_CUDA;
  int grid = Q < 148 * 16 ? Q : 148 * 16;
  k3_scan_kernel<128><<<grid, 128, 0, st>>>(c0, c1, Q, queries, sense, first_match, which);
  MOIP_CUDA(cudaGetLastError());
  return MOIP_OK;
}

int launch_k3_one(const DevCache& c0, const DevCache& c1, const K3Query& q, int sense, K3Answer* answer_dev, int seq,
                  cudaStream_t st) {
  static LaunchCfg carve;
  if (set_aux_carveout(k3_scan1_kernel<512>, carve)) return MOIP_ERR_CUDA;
  k3_scan1_kernel<512><<<1, 512, 0, st>>>(c0, c1, q, sense, answer_dev, seq);
  MOIP_CUDA(cudaGetLastError());
  return MOIP_OK;
}

int launch_k4(const DevModel& dm, int B, const int* x, const double* rhs, long long* obj_out,
              unsigned char* feasible_out, cudaStream_t st) {
  if (B <= 0) return MOIP_OK;
  int blocks = (B + 3) / 4;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k4_verify_kernel<<<blocks, 128, 0, st>>>(dm, B, x, rhs, obj_out, feasible_out);
  MOIP_CUDA(cudaGetLastError());
  return MOIP_OK;
}

}  // namespace moip
