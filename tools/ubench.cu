// Latency micro-benchmarks for the K1 design (dependent chains, one warp / one CTA): build with
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/ubench.cu -o gpurun_out/ubench
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double* out, long long* cyc, int n) {
  double a = out[0], b = out[1], c = out[2];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
  long long t1 = clock64();
  out[3] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dadd(double* out, long long* cyc, int n) {
  double a = out[0], b = out[1];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { a += b; a += b; a += b; a += b; }
  long long t1 = clock64();
  out[3] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_clamp(double* out, long long* cyc, int n) {
  double a = out[0], lo = out[1], hi = out[2];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int q = 0; q < 4; ++q) { a = a * 1.0000001; a = a > lo ? a : lo; a = a < hi ? a : hi; }
  }
  long long t1 = clock64();
  out[3] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_fminmax(double* out, long long* cyc, int n) {
  double a = out[0], lo = out[1], hi = out[2];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int q = 0; q < 4; ++q) { a = a * 1.0000001; a = fmin(fmax(a, lo), hi); }
  }
  long long t1 = clock64();
  out[3] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl(double* out, long long* cyc, int n) {
  double a = out[threadIdx.x & 3];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  }
  long long t1 = clock64();
  out[3] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(double* out, long long* cyc, int n) {
  __shared__ int idx[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) idx[i] = (i * 7 + 1) & 255;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { p = idx[p]; p = idx[p]; p = idx[p]; p = idx[p]; }
  long long t1 = clock64();
  out[3] = p; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_bar(double* out, long long* cyc, int n) {
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { __syncthreads(); __syncthreads(); __syncthreads(); __syncthreads(); }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rcp(double* out, long long* cyc, int n) {
  double a = out[0];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { a = __drcp_rn(a + 1.0); a = __drcp_rn(a + 1.0); a = __drcp_rn(a + 1.0); a = __drcp_rn(a + 1.0); }
  long long t1 = clock64();
  out[3] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// throughput: NT threads, 8 independent DFMA chains each
__global__ void k_dfma_tp(double* out, long long* cyc, int n) {
  double a[8];
  for (int q = 0; q < 8; ++q) a[q] = out[q & 3] + threadIdx.x;
  const double b = out[1], c = out[2];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = fma(a[q], b, c);
  }
  long long t1 = clock64();
  double s = 0; for (int q = 0; q < 8; ++q) s += a[q];
  out[4 + threadIdx.x % 4] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <class F> void run(const char* name, F f, int nt, int n, int ops) {
  double* out; long long* cyc; cudaMalloc(&out, 64 * 8); cudaMalloc(&cyc, 8);
  double h[8] = {1.0, 0.999999, 1e-9, 0, 0, 0, 0, 0}; cudaMemcpy(out, h, sizeof h, cudaMemcpyHostToDevice);
  f<<<1, nt>>>(out, cyc, n); f<<<1, nt>>>(out, cyc, n);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s nt=%4d: %.2f cycles per op\n", name, nt, (double)c / ((double)n * ops));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  const int n = 4096;
  run("dependent DFMA", k_dfma, 32, n, 4);
  run("dependent DADD", k_dadd, 32, n, 4);
  run("DMUL+clamp (setp/sel)", k_clamp, 32, n, 4);
  run("DMUL+fmin(fmax())", k_fminmax, 32, n, 4);
  run("warp sum (5 shfl64 + add)", k_shfl, 32, n, 1);
  run("dependent LDS", k_lds, 32, n, 4);
  run("dependent __drcp_rn(a+1)", k_rcp, 32, n, 4);
  for (int nt : {32, 128, 256, 512, 1024}) run("__syncthreads", k_bar, nt, n, 4);
  for (int nt : {32, 128, 256, 512, 1024}) run("DFMA x8 independent / thread", k_dfma_tp, nt, n, 8);
  return 0;
}
