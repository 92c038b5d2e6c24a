// Device-side helpers shared by the tcgen05 engines (qoc_tc_f16.cu: operands streamed by TMA; qoc_tc_small.cu: operands
// resident in shared memory): mbarrier / TMA / tcgen05 wrappers, fp16-pair split and plane-set access.
#pragma once
#include "qoc_tc_f16.cuh"
#include <cuda_fp16.h>
#include <stdint.h>

#define DEVINL __device__ __forceinline__

namespace {

constexpr long long TIMEOUT_CYCLES = 4000000000LL;

DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEVINL void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// pure spin on test_wait: try_wait may suspend the thread for a system-dependent time, which adds wake-up latency to
// every hand-off -- negligible for long products, dominant for the short ones of the small-n kernel
DEVINL bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
DEVINL bool mbar_spin(uint64_t* bar, uint32_t parity, volatile int* dead) {
  if (mbar_test_wait(bar, parity)) return true;
  const long long t0 = clock64();
  int it = 0;
  while (!mbar_test_wait(bar, parity)) {
    if ((++it & 255) == 0 && (*dead || clock64() - t0 > TIMEOUT_CYCLES)) { *dead = 1; return false; }
  }
  return true;
}
// false = timed out (or another role already failed): the caller unwinds to the teardown
DEVINL bool mbar_wait(uint64_t* bar, uint32_t parity, volatile int* dead) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (*dead || clock64() - t0 > TIMEOUT_CYCLES) { *dead = 1; return false; }
  }
  return true;
}
DEVINL void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
DEVINL void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
DEVINL void mma_f16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&u)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
        "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
}
DEVINL uint32_t elect_one() {
  uint32_t e;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(e));
  return e;
}

// shared-memory matrix descriptors (cute::UMMA::SmemDescriptor layout): start >> 4 | LBO >> 4 << 16 |
// SBO >> 4 << 32 | version 1 << 46 | layout type << 61 (SWIZZLE_128B = 2, SWIZZLE_64B = 4)
DEVINL uint64_t make_desc(uint32_t saddr, uint32_t lbo16, uint32_t sbo16, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(lbo16 & 0x3FFF) << 16) | ((uint64_t)(sbo16 & 0x3FFF) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}

// x (stored units) -> fp16 pair
DEVINL void split2(float a, float b, uint32_t& h0, uint32_t& h1) {
  const __half2 x = __floats2half2_rn(a, b);
  const float2 f = __half22float2(x);
  const __half2 y = __floats2half2_rn(a - f.x, b - f.y);
  h0 = *reinterpret_cast<const uint32_t*>(&x);
  h1 = *reinterpret_cast<const uint32_t*>(&y);
}
// 16 consecutive columns of one row: 32 bytes (one L2 sector) per plane, 256-bit accesses
DEVINL void stg256(void* p, const uint32_t (&r)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]),
               "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
DEVINL void ldg256(const void* p, uint32_t (&r)[8]) {     // L2-coherent (.cg): the data was written by this CTA moments ago
  asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p)
               : "memory");
}
DEVINL void store_planes16(__half* mat, size_t plane, int ld, int row, int col, const float (&re)[16], const float (&im)[16]) {
  uint32_t a0[8], a1[8], b0[8], b1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { split2(re[2 * i], re[2 * i + 1], a0[i], a1[i]); split2(im[2 * i], im[2 * i + 1], b0[i], b1[i]); }
  __half* p0 = mat + (size_t)row * ld + col;
  stg256(p0, a0);
  stg256(p0 + plane, a1);
  stg256(p0 + 2 * plane, b0);
  stg256(p0 + 3 * plane, b1);
}
// one component (Re or Im): planes pl0 (h0) and pl0 + 1 (h1)
DEVINL void store_comp16(__half* p0, size_t plane, const float (&v)[16]) {
  uint32_t a0[8], a1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split2(v[2 * i], v[2 * i + 1], a0[i], a1[i]);
  stg256(p0, a0);
  stg256(p0 + plane, a1);
}
// one real -> its fp16 pair
DEVINL void split1(float a, unsigned short& h0, unsigned short& h1) {
  const __half x = __float2half_rn(a);
  const __half y = __float2half_rn(a - __half2float(x));
  h0 = __half_as_ushort(x); h1 = __half_as_ushort(y);
}
// sparse generator assembly (global plane sets): entry e of the union pattern, x = sum_k w_k coef[e][k]
DEVINL void scatter_x_entry(const TcParams& q, __half* X, size_t plane, int ld, int e, const float* wts) {
  const int rc = __ldg(q.pat_rc + e), pr = rc >> 16, pc = rc & 0xffff;
  const float2* cf = q.pat_coef_f + (size_t)e * (q.K + 1);
  float xr = 0.f, xi = 0.f;
  for (int k = 0; k <= q.K; ++k) { const float2 a = __ldg(cf + k); xr = fmaf(wts[k], a.x, xr); xi = fmaf(wts[k], a.y, xi); }
  unsigned short r0, r1, i0, i1;
  split1(xr, r0, r1); split1(xi, i0, i1);
  unsigned short* o = reinterpret_cast<unsigned short*>(X) + (size_t)pr * ld + pc;
  o[0] = r0; o[plane] = r1; o[2 * plane] = i0; o[3 * plane] = i1;
}
DEVINL void unpack16(const uint32_t (&h0)[8], const uint32_t (&h1)[8], float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&h0[i]));
    const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&h1[i]));
    v[2 * i] = fa.x + fb.x; v[2 * i + 1] = fa.y + fb.y;
  }
}

}  // namespace
