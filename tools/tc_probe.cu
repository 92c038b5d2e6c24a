// Probe for the hand-written tcgen05 building block used by csrc/qoc_tc_tf32.cu:
// one CTA, two interleaved M=64 x N=32 x K=64 kind::tf32 products (accumulators at TMEM lane
// offsets 0 and 16), A K-major / B MN-major no-swizzle shared-memory descriptors, tcgen05.commit ->
// mbarrier, tcgen05.ld 32x32b.x32.  Compares with a CPU product.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe.bin tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>

#define DEVINL __device__ __forceinline__

DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

DEVINL uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
  return d;                        // layout_type = 0 (no swizzle), base_offset = 0
}

DEVINL void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

DEVINL void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
DEVINL void mma_tf32_mask(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(z), "r"(z), "r"(z), "r"(z)
      : "memory");
}

DEVINL void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

DEVINL void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (1u << 16) | ((32u >> 3) << 17) | ((64u >> 4) << 24);

__global__ void probe(const float* A /*[2][64][64]*/, const float* B /*[2][64][32] row-major (k, n)*/, float* D /*[2][64][32]*/,
                      int* status, float* Raw, int mode) {
  extern __shared__ __align__(1024) unsigned char smem_raw_[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw_ + 1023) & ~(uintptr_t)1023);
  // per item: A-form 16 KB, B-form 8 KB
  unsigned char* sA = smem;                 // [2][16384]
  unsigned char* sB = smem + 2 * 16384;     // [2][8192]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::); }
  // stage operands into the canonical no-swizzle layouts
  for (int it = 0; it < 2; ++it) {
    for (int idx = tid; idx < 64 * 64; idx += blockDim.x) {
      const int row = idx >> 6, k = idx & 63;
      const uint32_t off = (row >> 3) * 2048 + (k >> 2) * 128 + (row & 7) * 16 + (k & 3) * 4;
      *reinterpret_cast<float*>(sA + it * 16384 + off) = A[it * 4096 + idx];
    }
    for (int idx = tid; idx < 64 * 32; idx += blockDim.x) {
      const int k = idx >> 5, n = idx & 31;
      const uint32_t off = (n >> 2) * 128 + (k >> 3) * 1024 + (k & 7) * 16 + (n & 3) * 4;
      *reinterpret_cast<float*>(sB + it * 8192 + off) = B[it * 2048 + idx];
    }
  }
  if (mode == 31 || mode == 32 || mode == 40) {
    for (int it = 0; it < 2; ++it) {
      for (int idx = tid; idx < 64 * 64; idx += blockDim.x) {
        const int row = idx >> 6, k = idx & 63;
        const uint32_t off = (k >> 5) * 8192 + (row >> 3) * 1024 + (row & 7) * 128 + ((((k & 31) >> 2) ^ (row & 7)) << 4) + (k & 3) * 4;
        *reinterpret_cast<float*>(sA + it * 16384 + off) = A[it * 4096 + idx];
      }
      for (int idx = tid; idx < 64 * 32; idx += blockDim.x) {
        const int k = idx >> 5, n = idx & 31;
        uint32_t off = (k >> 3) * 1024 + (k & 7) * 128 + (((n >> 2) ^ (k & 7)) << 4) + (n & 3) * 4;
        if (mode == 32 || mode == 40) off = (k >> 5) * 4096 + (n >> 3) * 1024 + (n & 7) * 128 + ((((k & 31) >> 2) ^ (n & 7)) << 4) + (k & 3) * 4;
        *reinterpret_cast<float*>(sB + it * 8192 + off) = B[it * 2048 + idx];
      }
    }
  }
  if (mode >= 3 && mode != 31 && mode != 32 && mode != 40) for (int i = tid; i < (2 * 16384 + 2 * 8192) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  const uint32_t taddr = *(volatile uint32_t*)&tmem_base;
  if (tid == 0) status[1] = (int)taddr;
  {  // pre-fill the accumulator region with a pattern through tcgen05.st
    uint32_t pat = __float_as_uint(1000.0f + tid);
    const uint32_t paddr = taddr + ((uint32_t)(32 * warp) << 16);
    for (int c = 0; c < 32; ++c) asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(paddr + c), "r"(pat) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (mode == 21) {   // stage through the async proxy: bulk copy global -> shared, completion on an mbarrier
    __shared__ uint64_t cpbar;
    if (tid == 0) { mbar_init(&cpbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&cpbar)), "r"(32768u) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sA)),
                   "l"(A), "r"(32768u), "r"(smem_u32(&cpbar)) : "memory");
    }
    while (!mbar_try_wait(&cpbar, 0)) {}
    __syncthreads();
  }
  if (mode == 20) {
    for (int i = 0; i < 200; ++i) __nanosleep(1000);
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (mode == 40) {   // A operand in TMEM: every thread stores its own 64-element row at columns 32..95 of its lane
    const int it_ = lane >> 4, row_ = 16 * warp + (lane & 15);
    const uint32_t aaddr = taddr + ((uint32_t)(32 * warp) << 16) + 32;
    for (int c = 0; c < 64; ++c) {
      const uint32_t val = __float_as_uint(A[it_ * 4096 + row_ * 64 + c]);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(aaddr + c), "r"(val) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (tid == 0 && mode != 1) {
    float chk = 0.f;
    for (int i = 0; i < 64; ++i) chk += reinterpret_cast<float*>(sA)[i * 17] + reinterpret_cast<float*>(sB)[i * 13];
    status[2] = (int)chk;
    for (int it = 0; it < (mode == 7 ? 1 : 2); ++it) {
      const uint32_t d = taddr + ((uint32_t)(16 * it) << 16);
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t da = make_desc(smem_u32(sA + it * 16384) + ks * 256, 128, 2048);
        const uint64_t db = make_desc(smem_u32(sB + it * 8192) + ks * 1024, 1024, 128);
        uint64_t da2 = da, db2 = db;
        if (mode == 4) { da2 &= ~((uint64_t)3 << 46); db2 &= ~((uint64_t)3 << 46); }
        if (mode == 40) {
          uint64_t b = make_desc(smem_u32(sB + it * 8192) + (ks & 3) * 32 + (ks >> 2) * 4096, 16, 1024) | ((uint64_t)2 << 61);
          mma_tf32_ts(d, taddr + ((uint32_t)(16 * it) << 16) + 32 + ks * 8, b, IDESC & ~(1u << 16), ks > 0 ? 1u : 0u);
          continue;
        }
        if (mode == 32) {   // A and B both K-major SW128
          uint64_t a = make_desc(smem_u32(sA + it * 16384) + (ks & 3) * 32 + (ks >> 2) * 8192, 16, 1024) | ((uint64_t)2 << 61);
          uint64_t b = make_desc(smem_u32(sB + it * 8192) + (ks & 3) * 32 + (ks >> 2) * 4096, 16, 1024) | ((uint64_t)2 << 61);
          mma_tf32(d, a, b, IDESC & ~(1u << 16), ks > 0 ? 1u : 0u);
          continue;
        }
        if (mode == 31) {   // A: K-major SW128 (2 k-blocks of 32), B: MN-major SW128 (8 atoms of 8 k-rows)
          uint64_t a = make_desc(smem_u32(sA + it * 16384) + (ks & 3) * 32 + (ks >> 2) * 8192, 16, 1024) | ((uint64_t)2 << 61);
          uint64_t b = make_desc(smem_u32(sB + it * 8192) + ks * 1024, 4096, 1024) | ((uint64_t)2 << 61);
          mma_tf32(d, a, b, IDESC, ks > 0 ? 1u : 0u);
          continue;
        }
        if (mode == 30) {   // canonical SWIZZLE_128B K-major operands (all-ones data -> layout details irrelevant)
          uint64_t a = make_desc(smem_u32(sA + it * 16384) + (ks & 3) * 32 + (ks >> 2) * 8192, 16, 1024) | ((uint64_t)2 << 61);
          uint64_t b = make_desc(smem_u32(sB + it * 8192) + (ks & 3) * 32 + (ks >> 2) * 4096, 16, 1024) | ((uint64_t)2 << 61);
          mma_tf32(d, a, b, IDESC & ~(1u << 16), ks > 0 ? 1u : 0u);
          continue;
        }
        uint32_t id = mode == 5 ? (IDESC & ~(1u << 16)) : IDESC;
        if (mode == 10) id = (id & ~(31u << 24)) | (8u << 24);                    // M = 128
        if (mode == 11) { id = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((64u >> 4) << 24); mma_f16(d, da2, db2, id, ks > 0 ? 1u : 0u); }
        else if (mode == 12) mma_tf32_mask(d, da2, db2, id, ks > 0 ? 1u : 0u);
        else if (mode == 13) mma_tf32(d, da2, db2, id | (1u << 15), ks > 0 ? 1u : 0u);
        else mma_tf32(d, da2, db2, id, ks > 0 ? 1u : 0u);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  long long spins = 0;
  while (mode != 1 && !mbar_try_wait(&bar, 0)) {
    if (++spins > 20000000LL) { if (tid == 0) *status = -1; break; }
  }
  if (mode == 6) { for (int i = 0; i < 2000; ++i) __nanosleep(1000); }
  if (tid == 0) status[3] = (int)spins;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[32];
  const uint32_t laddr = taddr + ((uint32_t)(32 * warp) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
      "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(laddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  const int item = lane >> 4, row = 16 * warp + (lane & 15);
  for (int c = 0; c < 32; ++c) D[item * 2048 + row * 32 + c] = __uint_as_float(v[c]);
  for (int c = 0; c < 32; ++c) Raw[tid * 32 + c] = __uint_as_float(v[c]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(128));
}

static float tf32_round(float x) {
  uint32_t u; memcpy(&u, &x, 4);
  u += 0x1000u; u &= 0xFFFFE000u;
  float y; memcpy(&y, &u, 4); return y;
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  std::vector<float> A(2 * 4096), B(2 * 2048), D(2 * 2048, -7.f), R(2 * 2048);
  srand(1);
  for (auto& x : A) x = tf32_round((rand() / (float)RAND_MAX - 0.5f));
  for (auto& x : B) x = tf32_round((rand() / (float)RAND_MAX - 0.5f));
  if (mode == 2 || (mode >= 20 && mode != 31 && mode != 32 && mode != 40)) { for (auto& x : A) x = 1.f; for (auto& x : B) x = 1.f; }
  for (int it = 0; it < 2; ++it)
    for (int r = 0; r < 64; ++r)
      for (int c = 0; c < 32; ++c) {
        double s = 0;
        for (int k = 0; k < 64; ++k) s += (double)A[it * 4096 + r * 64 + k] * B[it * 2048 + k * 32 + c];
        R[it * 2048 + r * 32 + c] = (float)s;
      }
  float *dA, *dB, *dD, *dRaw; int* dS; int hS[4] = {0, 0, 0, 0}; std::vector<float> Raw(128 * 32);
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dS, 16); cudaMalloc(&dRaw, 128 * 32 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dS, 0, 16);
  const int smem = 2 * 16384 + 2 * 8192 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(dA, dB, dD, dS, dRaw, mode);
  printf("launch: %s\n", cudaGetErrorString(cudaGetLastError()));
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hS, dS, 16, cudaMemcpyDeviceToHost); cudaMemcpy(Raw.data(), dRaw, Raw.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (size_t i = 0; i < D.size(); ++i) { double d = fabs(D[i] - R[i]); if (d > maxerr) maxerr = d; if (d > 1e-3) ++bad; }
  printf("taddr=0x%x smemchk=%d spins=%d\n", hS[1], hS[2], hS[3]);
  for (int l = 0; l < 128; l += 8) printf("lane %3d: %10.4f %10.4f %10.4f ... %10.4f\n", l, Raw[l * 32], Raw[l * 32 + 1], Raw[l * 32 + 2], Raw[l * 32 + 31]);
  printf("ref row0: %10.4f %10.4f %10.4f ... %10.4f\n", R[0], R[1], R[2], R[31]);
  printf("{\"cuda\": \"%s\", \"status\": %d, \"max_abs_err\": %.3e, \"bad\": %d, \"D0\": [%f, %f, %f], \"R0\": [%f, %f, %f], \"D1\": [%f, %f], \"R1\": [%f, %f]}\n",
         cudaGetErrorString(e), hS[0], maxerr, bad, D[0], D[1], D[33], R[0], R[1], R[33], D[2048], D[2048 + 63 * 32 + 31], R[2048], R[2048 + 63 * 32 + 31]);
  return (e == cudaSuccess && bad == 0 && hS[0] == 0) ? 0 : 1;
}
