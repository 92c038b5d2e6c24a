// tcgen05 / TMEM kernel for the propagator stage in fp32-class arithmetic (QOC_TF32X3), n <= 32.
//
//   k_expm_tc32 : (b,t) -> P_t = (sum_{j<=p} H^j/j!)^(2^s)   (get_matexp, core/tensorflow_state.py:25-46)
//
// Formulation.  The complex n x n product C = X Y is ONE real MMA in the reference's own embedding
// (helper_functions/grape_functions.py:211-220):  [Cr; Ci] = [[Xr, -Xi], [Xi, Xr]] [Yr; Yi], i.e.
// M = 2*32 = 64, N = 32, K = 64 with NO redundant flops (only the left block column of the embedded
// result is formed).  M = 64 is the smallest tcgen05 tile, its accumulator occupies TMEM lanes
// 0-15 of each 32-lane sub-partition, so a CTA runs a PAIR of problem items whose accumulators
// interleave (lane offsets 0 and 16) and all 128 threads own exactly one accumulator row.
//
// Precision.  kind::tf32 keeps 10 mantissa bits; every operand is split x = hi + lo (both tf32) and
// a product is three MMAs  A_lo B_hi + A_hi B_lo + A_hi B_hi  accumulated in fp32 in TMEM
// ("3xTF32"), giving ~fp32 accuracy -- the reference itself is float32.
//
// Data flow per product: the A operand (the 64 x 64 embedding, hi and lo parts) lives in TENSOR MEMORY
// next to the accumulator -- row rho of A sits in the same TMEM lane as row rho of D, so the thread that
// pulled accumulator row rho with tcgen05.ld writes the next A row straight back with tcgen05.st (the
// Im/Re partner row comes through a small shared-memory exchange).  Only the B operand [Yr;Yi]^T goes
// through shared memory (canonical K-major SWIZZLE_128B UMMA layout, generic stores + fence.proxy.async).
// A converged warp issues the 2 x 24 tcgen05.mma (A from TMEM, B from a shared-memory descriptor) with
// elect.sync and a tcgen05.commit onto an mbarrier; all threads wait, tcgen05.ld their row, apply the
// Paterson-Stockmeyer update in registers, re-split and write the next operands.  Two CTAs are resident
// per SM (2 x 256 TMEM columns) so one CTA's epilogue overlaps the other's MMAs.  v1 kept A in shared
// memory too and was shared-memory-bandwidth bound (profiles/r01_ncu_expm_tc32.md).  Layouts, descriptors
// and the A-in-TMEM lane map were validated on hardware with tools/tc_probe.cu (modes 32 and 40).
//
// Output: P[b][t] as fp32 planar padded [2][32][32] (Re plane, Im plane), rows 128-byte aligned.
#include "qoc_internal.cuh"
#include <stdint.h>

#define DEVINL __device__ __forceinline__

namespace {

constexpr int NP = 32;
constexpr uint32_t B_BYTES = 32 * 64 * 4;      // one B-form operand (hi or lo)
constexpr int XLD = 36;                        // exchange-buffer row stride (floats, 16-byte aligned rows)
constexpr uint32_t X_BYTES = 64 * XLD * 4;     // row exchange buffer (9 KB)
constexpr uint32_t ITEM_BYTES = 2 * B_BYTES + 19456;            // B_hi | B_lo | exchange hi rows | exchange lo rows (35 KB, 1024-aligned)
constexpr int STG_LD = 33;                     // staging row stride (floats)
constexpr uint32_t TM_COLS = 256;              // TMEM columns per CTA: D [0,32) | A_hi [32,96) | A_lo [96,160)
constexpr uint32_t TM_AHI = 32, TM_ALO = 96;
constexpr size_t SMEM_REQ = 110 * 1024;        // > 227/3 KB: caps residency at 2 CTAs/SM (2 x 256 TMEM columns)
// instruction descriptor: D=f32 (bits 4-5 = 1), A=B=tf32 (2 at bits 7-9 / 10-12), both K-major,
// N>>3 at bits 17-22, M>>4 at bits 24-28
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((64u >> 4) << 24);

DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (8-row x 128 B atoms, SBO = 1024 B)
DEVINL uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                       // LBO (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;             // SBO
  d |= (uint64_t)1 << 46;                       // descriptor version
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}

// D[tmem] (+)= A[tmem] * B[smem descriptor]
DEVINL void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(IDESC), "r"(accumulate)
      : "memory");
}

DEVINL void tmem_st32(uint32_t taddr, const uint32_t (&u)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
      "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(taddr),
      "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]), "r"(u[9]), "r"(u[10]),
      "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]), "r"(u[16]), "r"(u[17]), "r"(u[18]), "r"(u[19]), "r"(u[20]),
      "r"(u[21]), "r"(u[22]), "r"(u[23]), "r"(u[24]), "r"(u[25]), "r"(u[26]), "r"(u[27]), "r"(u[28]), "r"(u[29]), "r"(u[30]),
      "r"(u[31])
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

// x = hi + lo with hi, lo representable in tf32 (10 mantissa bits), both rounded to nearest (ties away)
// by an integer add on the bit pattern -- 5 instructions; cvt.rna.tf32.f32 expands to ~10 on sm_100a.
DEVINL void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  const float d = x - __uint_as_float(hi);
  lo = (__float_as_uint(d) + 0x1000u) & 0xffffe000u;
}

// byte offset of element (n, k) of a B-form operand (B^T, 32 rows n, K = 64 in two k-blocks)
DEVINL uint32_t b_off(int n, int k) {
  return (uint32_t)((k >> 5) * 4096 + (n >> 3) * 1024 + (n & 7) * 128 + ((((k & 31) >> 2) ^ (n & 7)) << 4) + (k & 3) * 4);
}

DEVINL void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
DEVINL void sts32(uint32_t addr, float a) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(a) : "memory"); }
DEVINL float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

DEVINL void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
DEVINL void sts32u(uint32_t addr, uint32_t a) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(a) : "memory"); }
DEVINL uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// Write the operands derived from this thread's stacked row rho (rho < 32: Re row rho, else Im row rho-32):
//   WB: B-form [Xr; Xi]^T (hi, lo) into shared memory (scattered 4-byte stores into the K-major swizzled tile);
//   WA: A-form row rho of [[Xr,-Xi],[Xi,Xr]] into TMEM: columns 0-31 = own row, 32-63 = the Re/Im partner's row
//       (negated for Re rows), hi part at TM_AHI, lo part at TM_ALO.  The partner's split row comes through the
//       shared exchange buffer (sX: hi rows, then lo rows); contains one __syncthreads.
template <bool WA, bool WB>
DEVINL void write_operands(uint32_t ta_lane, uint32_t sB_hi, uint32_t sX, int rho, int it, const float (&v)[32]) {
  uint32_t hi[32], lo[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) split_tf32(v[c], hi[c], lo[c]);
  if (WB) {
    const uint32_t sB_lo = sB_hi + B_BYTES;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      const int cc = c ^ (it << 2);      // the two items of a pair walk the columns in a different order: no bank clash
      const uint32_t o = b_off(cc, rho);
      // (the value for column cc is element cc of the row)
      sts32u(sB_hi + o, it ? hi[c ^ 4] : hi[c]);
      sts32u(sB_lo + o, it ? lo[c ^ 4] : lo[c]);
    }
  }
  if (WA) {
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      sts128u(sX + (rho * XLD + 4 * cc) * 4, hi[4 * cc], hi[4 * cc + 1], hi[4 * cc + 2], hi[4 * cc + 3]);
      sts128u(sX + X_BYTES + (rho * XLD + 4 * cc) * 4, lo[4 * cc], lo[4 * cc + 1], lo[4 * cc + 2], lo[4 * cc + 3]);
    }
    tmem_st32(ta_lane + TM_AHI, hi);
    tmem_st32(ta_lane + TM_ALO, lo);
    __syncthreads();                     // partner rows are published
    const int prow = rho < 32 ? rho + 32 : rho - 32;
    const uint32_t flip = rho < 32 ? 0x80000000u : 0u;      // -Xi for the Re rows
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      const uint4 ph = lds128u(sX + (prow * XLD + 4 * cc) * 4);
      const uint4 pl = lds128u(sX + X_BYTES + (prow * XLD + 4 * cc) * 4);
      hi[4 * cc] = ph.x ^ flip; hi[4 * cc + 1] = ph.y ^ flip; hi[4 * cc + 2] = ph.z ^ flip; hi[4 * cc + 3] = ph.w ^ flip;
      lo[4 * cc] = pl.x ^ flip; lo[4 * cc + 1] = pl.y ^ flip; lo[4 * cc + 2] = pl.z ^ flip; lo[4 * cc + 3] = pl.w ^ flip;
    }
    tmem_st32(ta_lane + TM_AHI + 32, hi);
    tmem_st32(ta_lane + TM_ALO + 32, lo);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
}

__global__ void __launch_bounds__(128, 2) k_expm_tc32(QocParams p, int* err_flag) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float wts[2][32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int it = lane >> 4;                      // which item of the pair this thread serves
  const int rho = 16 * warp + (lane & 15);       // stacked accumulator row
  const uint32_t sB = smem_u32(smem) + it * ITEM_BYTES;    // B_hi | B_lo | exchange (shared-window addresses)
  const uint32_t sX = sB + 2 * B_BYTES;
  const float* stg = reinterpret_cast<const float*>(smem + it * ITEM_BYTES);   // H staging [64][33] aliases the B-form region
  const int n = p.n, K = p.K, T = p.T;
  const long long items = (long long)p.B * T;
  const long long pairs = (items + 1) >> 1;
  float* Pout = reinterpret_cast<float*>(p.P);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = *(volatile uint32_t*)&tmem_base_s;
  const uint32_t ld_addr = taddr + ((uint32_t)(32 * warp) << 16);     // this thread's TMEM lane, column 0 (D)
  uint32_t parity = 0;
  bool dead = false;

  // one product for both items of the pair: D_it = A_it * B_it (3xTF32), result row -> v[]
  auto product = [&](float (&v)[32], int nvalid) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // operand stores -> async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) {
     // every value below is warp-uniform, so the descriptors live in uniform registers and the
     // elected lane issues back-to-back UTCHMMAs without a divergence "waterfall"
     const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
     const uint32_t tbase = __shfl_sync(0xffffffffu, taddr, 0);
     uint32_t elected;
     asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(elected));
     if (elected) {
      for (int i = 0; i < nvalid; ++i) {
        const uint32_t base = sbase + i * ITEM_BYTES;
        const uint64_t b_hi = make_desc(base), b_lo = make_desc(base + B_BYTES);
        const uint32_t d = tbase + ((uint32_t)(16 * i) << 16);          // accumulator / A rows of item i: lane offset 16*i
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t bo = (uint64_t)((ks & 3) * 2 + (ks >> 2) * (4096 >> 4));
          mma_tf32_ts(d, d + TM_ALO + 8 * ks, b_hi + bo, ks > 0 ? 1u : 0u);
          mma_tf32_ts(d, d + TM_AHI + 8 * ks, b_lo + bo, 1u);
          mma_tf32_ts(d, d + TM_AHI + 8 * ks, b_hi + bo, 1u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
     }
     __syncwarp();
    }
    const long long t_start = clock64();
    bool timed_out = false;
    while (!mbar_try_wait(&bar, parity)) {
      if (clock64() - t_start > 2000000000LL) { timed_out = true; break; }     // never hang the GPU: flag and bail out
    }
    dead = __syncthreads_or(timed_out ? 1 : 0) != 0;                          // CTA-uniform decision
    if (dead) return;
    parity ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t u[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
        "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
          "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]),
          "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]),
          "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
        : "r"(ld_addr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = __uint_as_float(u[c]);
  };

  for (long long pair = blockIdx.x; pair < pairs && !dead; pair += gridDim.x) {
    const long long item = 2 * pair + it;
    const bool valid = item < items;
    const int nvalid = (2 * pair + 1 < items) ? 2 : 1;
    const int b = valid ? (int)(item / T) : 0, t = valid ? (int)(item % T) : 0;

    // ---- H = (A_0 + sum_k u_k A_k) / 2^s assembled through a staging tile -------------------------
    const int j16 = lane & 15;
    if (warp == 0) {
      if (j16 == 0) wts[it][0] = (float)p.inv2s;
      else if (j16 <= K) wts[it][j16] = (float)(p.maxA[j16 - 1] * sin(p.base[((size_t)b * K + j16 - 1) * T + t]) * p.inv2s);
    }
    for (int i = tid; i < 2 * 64 * STG_LD; i += 128) {
      const int w = i / (64 * STG_LD);
      reinterpret_cast<float*>(smem + w * ITEM_BYTES)[i - w * 64 * STG_LD] = 0.0f;
    }
    __syncthreads();
    for (int e = tid; e < p.pat_n; e += 128) {
      const int rc = p.pat_rc[e];
      const int r = rc >> 16, c = rc & 0xffff;
      const float2* cf = p.pat_coef_f + (size_t)e * (K + 1);
#pragma unroll 1
      for (int w = 0; w < 2; ++w) {
        float hx = 0.f, hy = 0.f;
        for (int k = 0; k <= K; ++k) {
          const float2 a = cf[k];
          const float wk = wts[w][k];
          hx = fmaf(wk, a.x, hx); hy = fmaf(wk, a.y, hy);
        }
        float* S = reinterpret_cast<float*>(smem + w * ITEM_BYTES);
        S[r * STG_LD + c] = hx;
        S[(32 + r) * STG_LD + c] = hy;
      }
    }
    __syncthreads();
    float h[32], v[32], r_[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) h[c] = stg[rho * STG_LD + c];
    __syncthreads();                               // staging aliases the B-form region written next

    const int pp = p.p;
    const bool on_diag_row = rho < n;               // identity lives in the Re block, padded rows stay zero
    auto add_block = [&](float c_id, float c_h, bool init) {    // r_ (+)= c_id*I + c_h*H
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float x = c_h * h[c] + ((on_diag_row && c == rho) ? c_id : 0.0f);
        r_[c] = init ? x : r_[c] + x;
      }
    };
    // Paterson-Stockmeyer with block 2 (same polynomial as tensorflow_state.py:37-41)
    if (pp >= 2) {
      write_operands<true, true>(ld_addr, sB, sX, rho, it, h);
      product(v, nvalid);                           // v = H^2
      if (dead) break;
      write_operands<false, true>(ld_addr, sB, sX, rho, it, v);    // H2 as the B operand of every Horner step
    }
    int blk;
    if (pp & 1) {
      add_block((float)p.invfact[pp - 1], (float)p.invfact[pp], true);
      blk = pp / 2 - 1;
    } else {
      const float cp = (float)p.invfact[pp];
#pragma unroll
      for (int c = 0; c < 32; ++c) r_[c] = cp * v[c];
      add_block((float)p.invfact[pp - 2], (float)p.invfact[pp - 1], false);
      blk = pp / 2 - 2;
    }
    for (; blk >= 0; --blk) {
      write_operands<true, false>(ld_addr, sB, sX, rho, it, r_);
      product(v, nvalid);                           // v = R * H2
      if (dead) break;
#pragma unroll
      for (int c = 0; c < 32; ++c) r_[c] = v[c];
      add_block((float)p.invfact[2 * blk], (float)p.invfact[2 * blk + 1], false);
    }
    if (dead) break;
    for (int s = 0; s < p.s; ++s) {                 // squarings (tensorflow_state.py:43-44)
      write_operands<true, true>(ld_addr, sB, sX, rho, it, r_);
      product(v, nvalid);
      if (dead) break;
#pragma unroll
      for (int c = 0; c < 32; ++c) r_[c] = v[c];
    }
    if (dead) break;
    if (valid) {                                    // planar padded fp32 [2][32][32]: this thread's 128-byte row
      float4* dst = reinterpret_cast<float4*>(Pout + (size_t)item * (2 * NP * NP) + (size_t)rho * NP);
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) dst[cc] = make_float4(r_[4 * cc], r_[4 * cc + 1], r_[4 * cc + 2], r_[4 * cc + 3]);
    }
    __syncthreads();                                // wts / staging are rewritten by the next pair
  }
  if (dead && tid == 0 && err_flag) atomicExch(err_flag, 1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(TM_COLS) : "memory");
}

}  // namespace

cudaError_t qoc_launch_expm_tc32(const QocParams& p, int sm_count, int* err_flag, cudaStream_t st, int64_t* launches) {
  ++*launches;
  const size_t smem = SMEM_REQ;
  cudaError_t e = cudaFuncSetAttribute(k_expm_tc32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long long pairs = ((long long)p.B * p.T + 1) / 2;
  long long grid = 2LL * sm_count;
  if (grid > pairs) grid = pairs;
  k_expm_tc32<<<(unsigned)grid, 128, smem, st>>>(p, err_flag);
  return cudaGetLastError();
}
