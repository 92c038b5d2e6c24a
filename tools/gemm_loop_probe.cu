// Isolates the inner k-loop of the complex DMMA GEMM (csrc/qoc_mma_f64.cu: mma_gemm, RB=2 CB=4) to see how much
// of the FP64 tensor pipe the 3M form (3 DMMAs + operand-sum DADDs) can use with 1 or 2 warps per scheduler,
// and whether the placement of the DADDs matters.  Variants:
//   0: 4M (4 DMMAs per block, no DADD)       1: 3M, sums computed right where the fragments are loaded
//   2: 3M software-pipelined (as shipped)    3: variant 2 + __syncwarp() fences around the sum block
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gemm_loop_probe.bin gemm_loop_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef double2 cplx;
#define DEVINL __device__ __forceinline__
DEVINL void dmma(double& d0, double& d1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
DEVINL double dadd_v(const double a, const double b) {
  double r;
  asm volatile("add.f64 %0, %1, %2;\n" : "=d"(r) : "d"(a), "d"(b));
  return r;
}
DEVINL int sw_mask(int r) { return ((r & 1) * 5) ^ (((r >> 1) & 3) << 1); }

constexpr int NP = 32, RB = 2, CB = 4;

template <int VAR>
__global__ void __launch_bounds__(64) k_loop(double* out, int reps) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* A = reinterpret_cast<cplx*>(smem_raw) + (threadIdx.x >> 5) * 0;   // both warps share the operands (like a CTA of the real kernel)
  cplx* B = A + NP * NP;
  for (int i = threadIdx.x; i < 2 * NP * NP; i += blockDim.x) A[i] = make_double2(1e-3 * (i % 7), 1e-3 * (i % 5));
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const int rb0 = (threadIdx.x >> 5) * RB, cb0 = 0;
  int arow[RB], amask[RB];
  double cr[RB][CB][2], ci[RB][CB][2], t2[RB][CB][2];
#pragma unroll
  for (int i = 0; i < RB; ++i) {
    const int r = 8 * (rb0 + i) + g;
    arow[i] = r * NP; amask[i] = sw_mask(r);
#pragma unroll
    for (int j = 0; j < CB; ++j) cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = t2[i][j][0] = t2[i][j][1] = 0.0;
  }
  const int bc = 8 * cb0 + g;
  const int ksteps = 8;
  for (int rep = 0; rep < reps; ++rep) {
    if (VAR == 0 || VAR == 1) {
#pragma unroll 2
      for (int ks = 0; ks < ksteps; ++ks) {
        const int k = 4 * ks + q;
        cplx a[RB], b[CB];
#pragma unroll
        for (int i = 0; i < RB; ++i) a[i] = A[arow[i] + (k ^ amask[i])];
        const int bm = sw_mask(k);
#pragma unroll
        for (int j = 0; j < CB; ++j) b[j] = B[k * NP + ((bc + 8 * j) ^ bm)];
        if (VAR == 0) {
#pragma unroll
          for (int i = 0; i < RB; ++i)
#pragma unroll
            for (int j = 0; j < CB; ++j) dmma(cr[i][j][0], cr[i][j][1], a[i].x, b[j].x);
#pragma unroll
          for (int i = 0; i < RB; ++i)
#pragma unroll
            for (int j = 0; j < CB; ++j) dmma(ci[i][j][0], ci[i][j][1], a[i].x, b[j].y);
#pragma unroll
          for (int i = 0; i < RB; ++i) {
            const double na = -a[i].y;
#pragma unroll
            for (int j = 0; j < CB; ++j) dmma(cr[i][j][0], cr[i][j][1], na, b[j].y);
          }
#pragma unroll
          for (int i = 0; i < RB; ++i)
#pragma unroll
            for (int j = 0; j < CB; ++j) dmma(ci[i][j][0], ci[i][j][1], a[i].y, b[j].x);
        } else {
          double sa[RB], sb[CB];
#pragma unroll
          for (int i = 0; i < RB; ++i) sa[i] = a[i].x + a[i].y;
#pragma unroll
          for (int j = 0; j < CB; ++j) sb[j] = b[j].x + b[j].y;
#pragma unroll
          for (int i = 0; i < RB; ++i)
#pragma unroll
            for (int j = 0; j < CB; ++j) dmma(cr[i][j][0], cr[i][j][1], a[i].x, b[j].x);
#pragma unroll
          for (int i = 0; i < RB; ++i)
#pragma unroll
            for (int j = 0; j < CB; ++j) dmma(t2[i][j][0], t2[i][j][1], a[i].y, b[j].y);
#pragma unroll
          for (int i = 0; i < RB; ++i)
#pragma unroll
            for (int j = 0; j < CB; ++j) dmma(ci[i][j][0], ci[i][j][1], sa[i], sb[j]);
        }
      }
    } else {
      cplx a[RB], b[CB];
      double sa[RB], sb[CB];
      {
        const int k = q, bm = sw_mask(k);
#pragma unroll
        for (int i = 0; i < RB; ++i) a[i] = A[arow[i] + (k ^ amask[i])];
#pragma unroll
        for (int j = 0; j < CB; ++j) b[j] = B[k * NP + ((bc + 8 * j) ^ bm)];
      }
#pragma unroll
      for (int i = 0; i < RB; ++i) sa[i] = dadd_v(a[i].x, a[i].y);
#pragma unroll
      for (int j = 0; j < CB; ++j) sb[j] = dadd_v(b[j].x, b[j].y);
#pragma unroll 2
      for (int ks = 0; ks < ksteps; ++ks) {
        cplx an[RB], bn[CB];
        {
          const int k = 4 * min(ks + 1, ksteps - 1) + q, bm = sw_mask(k);
#pragma unroll
          for (int i = 0; i < RB; ++i) an[i] = A[arow[i] + (k ^ amask[i])];
#pragma unroll
          for (int j = 0; j < CB; ++j) bn[j] = B[k * NP + ((bc + 8 * j) ^ bm)];
        }
#pragma unroll
        for (int i = 0; i < RB; ++i)
#pragma unroll
          for (int j = 0; j < CB; ++j) dmma(cr[i][j][0], cr[i][j][1], a[i].x, b[j].x);
#pragma unroll
        for (int i = 0; i < RB; ++i)
#pragma unroll
          for (int j = 0; j < CB; ++j) dmma(t2[i][j][0], t2[i][j][1], a[i].y, b[j].y);
        if (VAR == 3) __syncwarp();
        double sna[RB], snb[CB];
#pragma unroll
        for (int i = 0; i < RB; ++i) sna[i] = dadd_v(an[i].x, an[i].y);
#pragma unroll
        for (int j = 0; j < CB; ++j) snb[j] = dadd_v(bn[j].x, bn[j].y);
        if (VAR == 3) __syncwarp();
#pragma unroll
        for (int i = 0; i < RB; ++i)
#pragma unroll
          for (int j = 0; j < CB; ++j) dmma(ci[i][j][0], ci[i][j][1], sa[i], sb[j]);
#pragma unroll
        for (int i = 0; i < RB; ++i) { a[i] = an[i]; sa[i] = sna[i]; }
#pragma unroll
        for (int j = 0; j < CB; ++j) { b[j] = bn[j]; sb[j] = snb[j]; }
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < RB; ++i)
#pragma unroll
    for (int j = 0; j < CB; ++j) s += cr[i][j][0] + cr[i][j][1] + ci[i][j][0] + ci[i][j][1] + t2[i][j][0] + t2[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int VAR>
void run(double* out, int sms, int ctas_per_sm) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = 2000;
  const size_t smem = 2 * NP * NP * sizeof(cplx) + (ctas_per_sm == 2 ? 60 * 1024 : 0);   // pad to pin the CTAs/SM
  cudaFuncSetAttribute(k_loop<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    k_loop<VAR><<<sms * ctas_per_sm, 64, smem>>>(out, reps);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double dm = (VAR == 0 ? 32.0 : 24.0) * 8 * reps;                 // DMMAs per warp
  const double clk = best * 1e-3 * 1.965e9;
  const double warps_per_smsp = ctas_per_sm * 2 / 4.0;
  printf("{\"variant\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, \"clk_per_kstep_per_warp\": %.1f, \"dmma_pipe_util\": %.3f}\n", VAR,
         ctas_per_sm, best, clk / (8.0 * reps), dm * 16.0 * warps_per_smsp / clk);
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 4 * 64);
  for (int c : {2, 4}) {
    run<0>(out, sms, c);
    run<1>(out, sms, c);
    run<2>(out, sms, c);
    run<3>(out, sms, c);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
