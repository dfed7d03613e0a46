// DMMA latency / occupancy probe: throughput of mma.m8n8k4.f64 chains vs warps per SM and chains per warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_dmma2.bin tools/probe_dmma2.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double* out, int iters) {
  double a = threadIdx.x * 1e-9, b = 1.0 + threadIdx.x * 1e-9, c0[CH], c1[CH];
  for (int u = 0; u < CH; ++u) c0[u] = c1[u] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int u = 0; u < CH; ++u)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0[u]), "+d"(c1[u]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int u = 0; u < CH; ++u) s += c0[u] + c1[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
static void run(double* out, int sm, int wps) {
  const int iters = 4000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0); k<CH><<<sm, wps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
  }
  const double n = (double)iters * 8 * CH * wps * sm;  // DMMA warp-instructions
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("warps/SM %2d chains %d: %6.2f TFLOP/s   %.1f clk per DMMA per SMSP-warp-chain\n", wps, CH,
         n * 512 / (best * 1e-3) / 1e12, best * 1e-3 * clk * 1e3 / ((double)iters * 8));
}
int main() {
  int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, (size_t)sm * 1024 * 8);
  for (int wps : {4, 8, 16, 32}) { run<1>(out, sm, wps); run<2>(out, sm, wps); run<4>(out, sm, wps); run<8>(out, sm, wps); }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
