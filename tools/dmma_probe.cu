// How many resident warps does the FP64 tensor pipe (mma.sync m8n8k4 f64 = DMMA) need, and what do
// interleaved DADDs cost?  Sweeps warps per SM and independent accumulators per warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe.bin dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC, int NADD>
__global__ void k_probe(double* out, int iters) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  double s[NADD > 0 ? NADD : 1];
#pragma unroll
  for (int i = 0; i < (NADD > 0 ? NADD : 1); ++i) s[i] = i;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < NACC; ++u)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[u][0]), "+d"(c[u][1]) : "d"(a), "d"(b));
#pragma unroll
    for (int u = 0; u < NADD; ++u) asm volatile("add.f64 %0, %0, %1;\n" : "+d"(s[u]) : "d"(a));
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) r += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < (NADD > 0 ? NADD : 1); ++i) r += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int NCH>
__global__ void k_dfma_probe(double* out, int iters) {
  double a[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) a[i] = threadIdx.x * 1e-9 + i;
  const double x = 1.0000001, y = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int c = 0; c < NCH; ++c) a[c] = fma(a[c], x, y);
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) r += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int NCH>
void run_dfma(double* out, int sms, int warps_per_sm) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 16384 / NCH;
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    k_dfma_probe<NCH><<<sms, 32 * warps_per_sm>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double ninst = 4.0 * NCH * iters;          // DFMA warp-instructions per warp
  printf("{\"dfma_chains\": %d, \"warps_per_sm\": %d, \"ms\": %.4f, \"clk_per_dfma_per_warp\": %.2f, \"tflops\": %.2f}\n", NCH,
         warps_per_sm, best, best * 1e-3 * 1.965e9 / ninst, 2.0 * 32 * ninst * warps_per_sm * sms / best / 1e9);
}

template <int NACC, int NADD>
void run(double* out, int sms, int warps_per_sm) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096 * 8 / NACC;
  const int threads = 32 * warps_per_sm;          // one CTA per SM
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    k_probe<NACC, NADD><<<sms, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double flops = 2.0 * 256 * NACC * (double)iters * warps_per_sm * sms;
  printf("{\"warps_per_sm\": %d, \"acc\": %d, \"dadd_per_iter\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", warps_per_sm, NACC,
         NADD, best, flops / best / 1e9);
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 2048);
  for (int w : {1, 4, 8, 16}) {
    run_dfma<1>(out, sms, w);
    run_dfma<2>(out, sms, w);
    run_dfma<4>(out, sms, w);
    run_dfma<8>(out, sms, w);
    run_dfma<16>(out, sms, w);
  }
  for (int w : {4, 8, 16, 32}) {
    run<8, 0>(out, sms, w);
    run<16, 0>(out, sms, w);
    run<24, 0>(out, sms, w);
    run<24, 6>(out, sms, w);
    run<24, 12>(out, sms, w);
    run<15, 6>(out, sms, w);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
