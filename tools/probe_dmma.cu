// fp64 tensor-core (DMMA) throughput probe for sm_100a: mma.sync m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16 vs DFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_dmma.bin tools/probe_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE>
__global__ void k_dmma(double* out, int iters) {
  double a[8], b[4], c[4][4];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-9 + i;
  for (int i = 0; i < 4; ++i) b[i] = 1.0 + i * 1e-9;
  for (int u = 0; u < 4; ++u)
    for (int i = 0; i < 4; ++i) c[u][i] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (SHAPE == 0) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[u][0]), "+d"(c[u][1]) : "d"(a[0]), "d"(b[0]));
      } else if (SHAPE == 1) {
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+d"(c[u][0]), "+d"(c[u][1]), "+d"(c[u][2]), "+d"(c[u][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
      } else if (SHAPE == 2) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(c[u][0]), "+d"(c[u][1]), "+d"(c[u][2]), "+d"(c[u][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
      } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                     : "+d"(c[u][0]), "+d"(c[u][1]), "+d"(c[u][2]), "+d"(c[u][3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                       "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
      }
    }
  }
  double s = 0;
  for (int u = 0; u < 4; ++u)
    for (int i = 0; i < 4; ++i) s += c[u][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma(double* out, int iters) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-9 + i;
  const double m = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F>
static float time_it(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r && ms < best) best = ms;
  }
  return best;
}
int main() {
  int sm = 0;
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, (size_t)sm * 8 * 1024 * 8);
  const int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    const int blocks = sm * (1024 / threads) ;
    const double warps = (double)blocks * threads / 32;
    float t;
    t = time_it([&] { k_dfma<<<blocks, threads>>>(out, iters); });
    printf("threads %4d  DFMA        %7.2f TFLOP/s\n", threads, 2.0 * 8 * iters * blocks * threads / (t * 1e-3) / 1e12);
    t = time_it([&] { k_dmma<0><<<blocks, threads>>>(out, iters); });
    printf("threads %4d  m8n8k4      %7.2f TFLOP/s\n", threads, 2.0 * 256 * 4 * iters * warps / (t * 1e-3) / 1e12);
    t = time_it([&] { k_dmma<1><<<blocks, threads>>>(out, iters); });
    printf("threads %4d  m16n8k4     %7.2f TFLOP/s\n", threads, 2.0 * 512 * 4 * iters * warps / (t * 1e-3) / 1e12);
    t = time_it([&] { k_dmma<2><<<blocks, threads>>>(out, iters); });
    printf("threads %4d  m16n8k8     %7.2f TFLOP/s\n", threads, 2.0 * 1024 * 4 * iters * warps / (t * 1e-3) / 1e12);
    t = time_it([&] { k_dmma<3><<<blocks, threads>>>(out, iters); });
    printf("threads %4d  m16n8k16    %7.2f TFLOP/s\n", threads, 2.0 * 2048 * 4 * iters * warps / (t * 1e-3) / 1e12);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
