// tcgen05 / TMEM / TMA complex-GEMM program engine: the fp32-class (QOC_F16X2) arithmetic of the
// propagator stage (get_matexp / matexp_op, core/tensorflow_state.py:25-46,70-75) and of the
// U_final chain (init_tf_propagator, :204-227) for Hilbert dimensions up to 256.
//
// One persistent CTA runs "items"; an item is a sequence of DEPENDENT complex n x n products
// (Taylor / Paterson-Stockmeyer steps and squarings of one (b,t); the 15 products of one chain segment;
// the segment chain of one instance).  Matrices live in global memory (L2-resident scratch or the
// propagator cache in HBM) as split fp16 plane sets (qoc_tc_f16.cuh); a product D = A B is
//     Dr = Ar Br - Ai Bi,  Di = Ar Bi + Ai Br,   each real product = 3 MMAs (h0 h0 + h0 h1 + h1 h0),
// i.e. 12 tcgen05.mma.kind::f16 (M = 128, N <= 128, K = 16) per 16 columns of K, accumulated in fp32 in TMEM; the sign of
// the Ai Bi term is the descriptor's negate-A bit.  An output is cut into 128-row x (<=128)-column tiles ("quadrants");
// products with more than one tile (n > 128) alternate between two TMEM accumulator buffers so that the epilogue of a
// tile overlaps the MMAs of the next one, and completion is tracked per quadrant so that the first tile of the NEXT
// product starts as soon as the quadrants it reads are written.
//
// Roles (320 threads):
//   warp 0     TMA producer: per 32-wide k-block, one A box {32 k x 128 rows x 4 planes} (K-major, SWIZZLE_64B) and one
//              B box {64 n x 32 k x 4 planes} per 64-column group (MN-major, SWIZZLE_128B: B is read in its row-major
//              storage, no transposed copy exists anywhere) into a ring of stages; out-of-range rows / columns are
//              zero-filled by the rank-4 tensor maps (dims {n, n, 4 planes, matrices}).
//   warp 1     MMA issuer: waits full[s], issues 24 MMAs per stage, tcgen05.commit -> empty[s]; after the
//              last k-block commit -> tmem_full[buffer].
//   warps 2-9  epilogue (two per TMEM lane quarter, interleaving 16-column chunks): tcgen05.ld the accumulator rows (thread = row), apply  c0 D + c1 X + c2 I,
//              re-split into h0/h1 and store the plane set(s) of the result straight to global memory
//              (each thread writes whole 32-byte sectors of its row); fence.proxy.async + arrive on op_done
//              so the producer may fetch the result as an operand of the next product.
// Every mbarrier wait carries a clock64 timeout that flags err_flag instead of hanging the GPU.
#include "qoc_tc_dev.cuh"
#include <math.h>
#include <string.h>
#include <stdio.h>

namespace {

constexpr int KB_ELEMS = 32;                     // K elements per stage
constexpr uint32_t A_PLANE_BYTES = 128 * 64;     // box {32 halfs, 128 rows}
constexpr uint32_t B_GROUP_BYTES = 32 * 128;     // box {64 halfs, 32 rows}
constexpr int NEPI = 256;                        // epilogue threads (two warps per TMEM lane quarter)
constexpr int NHALF = NEPI / 128;                // epilogue warps per lane quarter (they interleave 16-column chunks)
constexpr int CSTEP = 16 * NHALF;
constexpr int NTHREADS = 64 + NEPI;
constexpr int MAX_STAGES = 4;

DEVINL void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory"); }

// resolved operands of one product
struct OpR {
  int a_cls, b_cls, e_cls, d1_cls, d2_cls;       // -1 = none
  long long a_idx, b_idx, e_idx, d1_idx, d2_idx; // matrix index within the class
  float c1[3], c2[3];
  int f64out;                                    // last product of a CHAIN item: write U_final, unitary_scale
};

DEVINL int item_nops(const TcParams& q, long long item) {
  switch (q.prog) {
    case TC_PROG_EXPM: return q.nops;
    case TC_PROG_SEG: {
      const int sg = (int)(item % q.S);
      const int t0 = sg * q.L, t1 = min(q.T, t0 + q.L);
      return max(1, t1 - t0 - 1);
    }
    case TC_PROG_CHAIN: return q.chain_len;
    default: return 1;
  }
}

DEVINL void make_op(const TcParams& q, long long item, int j, int nops, int z, int rpar, OpR& o) {
  const long long sb = ((long long)blockIdx.x * q.ilv + z) * TC_NSLOT;
  const int xs = rpar ? 4 : 0;                    // the generator X of odd rounds lives in slot 4 (it is built during the previous round)
  const float cu = 1.0f / (float)(1 << TC_EU);   // unitary x unitary -> unitary scale
  o.e_cls = -1; o.d2_cls = -1; o.d1_cls = -1; o.f64out = 0;
  o.e_idx = o.d1_idx = o.d2_idx = 0;
  o.c1[0] = cu; o.c1[1] = 0.f; o.c1[2] = 0.f;
  o.c2[0] = o.c2[1] = o.c2[2] = 0.f;
  switch (q.prog) {
    case TC_PROG_EXPM: {
      const TcExpmOp e = q.ops[j];
      o.a_cls = TC_CLS_SCR; o.a_idx = sb + (e.sa == 0 ? xs : e.sa);
      o.b_cls = TC_CLS_SCR; o.b_idx = sb + (e.sb == 0 ? xs : e.sb);
      o.e_cls = TC_CLS_SCR; o.e_idx = sb + (e.se == 0 ? xs : e.se);
      if (e.d1 >= 0) {
        if (e.d1 == TC_SLOT_OUT) { o.d1_cls = TC_CLS_P; o.d1_idx = item; } else { o.d1_cls = TC_CLS_SCR; o.d1_idx = sb + e.d1; }
      }
      if (e.d2 >= 0) {
        if (e.d2 == TC_SLOT_OUT) { o.d2_cls = TC_CLS_P; o.d2_idx = item; } else { o.d2_cls = TC_CLS_SCR; o.d2_idx = sb + e.d2; }
      }
      for (int i = 0; i < 3; ++i) { o.c1[i] = e.c1[i]; o.c2[i] = e.c2[i]; }
      break;
    }
    case TC_PROG_SEG: {
      const long long b = item / q.S;
      const int sg = (int)(item % q.S);
      const int t0 = sg * q.L, t1 = min(q.T, t0 + q.L);
      const long long pb = b * q.T;
      if (t1 - t0 == 1) {                          // single propagator: multiply by the identity (CONST 1)
        o.a_cls = TC_CLS_P; o.a_idx = pb + t0; o.b_cls = TC_CLS_CONST; o.b_idx = 1;
      } else {
        o.a_cls = TC_CLS_P; o.a_idx = pb + t0 + j + 1;
        if (j == 0) { o.b_cls = TC_CLS_P; o.b_idx = pb + t0; } else { o.b_cls = TC_CLS_SCR; o.b_idx = sb + 2 + ((j - 1) & 1); }   // slots 2, 3: 0 and 4 hold generators
      }
      if (j == nops - 1) { o.d1_cls = TC_CLS_SEG; o.d1_idx = item; } else { o.d1_cls = TC_CLS_SCR; o.d1_idx = sb + 2 + (j & 1); }
      break;
    }
    case TC_PROG_CHAIN: {
      o.a_cls = q.chain_cls; o.a_idx = item * q.chain_len + j;
      if (j == 0) { o.b_cls = TC_CLS_CONST; o.b_idx = 0; } else { o.b_cls = TC_CLS_SCR; o.b_idx = sb + 2 + ((j - 1) & 1); }
      if (j == nops - 1) o.f64out = 1; else { o.d1_cls = TC_CLS_SCR; o.d1_idx = sb + 2 + (j & 1); }
      break;
    }
    default: {
      o.a_cls = TC_CLS_P; o.a_idx = 2 * item; o.b_cls = TC_CLS_P; o.b_idx = 2 * item + 1;
      o.d1_cls = TC_CLS_SEG; o.d1_idx = item;
      break;
    }
  }
}

// which 128-row block of an operand a k-block of B rows lives in
DEVINL int rb_of_kb(int kb) { return (kb * KB_ELEMS) >> 7; }

__global__ void __launch_bounds__(NTHREADS, 1) k_tc_prog(const TcParams q, const __grid_constant__ TcMaps maps) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_tfull[2], bar_tempty[2];
  __shared__ unsigned int xdone_cnt;               // rounds whose generators X are assembled (EXPM)
  __shared__ unsigned int done_cnt[4];             // completed phases per output quadrant: monotonic counters (a waiter may lag
                                                   // several phases behind with two items interleaved: parity barriers would alias)
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int dead_s;
  __shared__ float wts[32];
  __shared__ double rs[2][128][2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = q.n, ld = q.ld, N16 = q.N16, NT0 = q.NT0, NH = q.NH, NGT = q.NGT, DIOFF = q.DIOFF, NBUF = q.NBUF, RB = q.RB,
            KBLK = q.KBLK, NS = q.stages;
  const int NQ = RB * NH;                                             // tiles (output quadrants) per product
  const uint32_t b_plane_bytes = (uint32_t)NGT * B_GROUP_BYTES;
  const uint32_t stage_bytes = 4 * A_PLANE_BYTES + 4 * b_plane_bytes;
  const size_t plane = (size_t)n * ld, mat = 4 * plane;
  long long t_wait0 = 0, t_wait1 = 0, t_work = 0;                    // per-role cycle counters (q.prof)
  const bool expm = q.prog == TC_PROG_EXPM;
  const int ILV = q.ilv;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&bar_tfull[b], 1); mbar_init(&bar_tempty[b], NEPI); }
    for (int i = 0; i < 4; ++i) done_cnt[i] = 0;
    xdone_cnt = 0;
    dead_s = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)q.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = *(volatile uint32_t*)&tmem_base_s;
  volatile int* dead = &dead_s;
  const bool prof = q.prof != nullptr;

  // Completion phases of the per-quadrant counters done_cnt[rb * NH + nh], counted per CTA: every round contributes one
  // phase for its prologue (EXPM only) and one per product.  Product j of an item may read a quadrant of its scratch
  // operands once phase  base + [EXPM] + j  of that quadrant's barrier has completed.
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;                             // ring fill counter
      uint32_t seen[4] = {0, 0, 0, 0};             // last value of done_cnt[i] observed
      uint32_t base_ph = 0;                        // phases completed by earlier rounds
      bool ok = true;
      long long t_cat[3] = {0, 0, 0};              // done-waits by cause: prologue, previous product, end of round
      int cat = 0;
      uint32_t xseen = 0, round = 0;
      auto need = [&](int i, uint32_t target) {    // block until quadrant i (i < 4) / the generators (i = 4) reached `target`
        uint32_t& sn = i < 4 ? seen[i] : xseen;
        if (sn >= target) return;
        const long long c0 = clock64();
        for (;;) {
          unsigned int v;
          asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(i < 4 ? &done_cnt[i] : &xdone_cnt)) : "memory");
          sn = v;
          if (v >= target) break;
          if (*dead || clock64() - c0 > TIMEOUT_CYCLES) { *dead = 1; ok = false; break; }
        }
        if (prof) { const long long dtw = clock64() - c0; t_wait0 += dtw; t_cat[cat] += dtw; }
        asm volatile("fence.proxy.async;" ::: "memory");       // the epilogue's generic-proxy stores before our async-proxy loads
      };
      for (long long it0 = (long long)blockIdx.x * ILV; it0 < q.items && ok; it0 += (long long)gridDim.x * ILV, ++round) {
        const int nz = (int)min((long long)ILV, q.items - it0);      // items interleaved in this round
        const int nops = item_nops(q, it0);
        for (int j = 0; j < nops && ok; ++j)
         for (int z = 0; z < nz && ok; ++z) {
          const long long item = it0 + z;
          OpR o; make_op(q, item, j, nops, z, (int)(round & 1), o);
          // phases that must be complete before reading scratch: this item's previous product, which sits nz positions
          // earlier in the stream (and, for the first product, the generators of this round)
          const uint32_t tgt = base_ph + (j > 0 ? (uint32_t)((j - 1) * nz + z + 1) : 0u);
          const bool dep = j > 0;
          cat = j == 0 ? 0 : 1;
          if (expm && j == 0) need(4, round + 1);
          const CUtensorMap* ma = &maps.a[o.a_cls];
          const CUtensorMap* mb = &maps.b[o.b_cls];
          const int za = (int)o.a_idx, zb = (int)o.b_idx;
          for (int rb = 0; rb < RB && ok; ++rb)
            for (int nh = 0; nh < NH && ok; ++nh)
              for (int kb = 0; kb < KBLK && ok; ++kb, ++it) {
                if (dep) {
                  need(rb * NH + ((NH == 2 && kb * KB_ELEMS >= NT0) ? 1 : 0), tgt);      // A: rows rb, columns of k-block kb
                  need((RB == 2 ? rb_of_kb(kb) : 0) * NH + nh, tgt);                     // B: rows of k-block kb, columns of half nh
                }
                const int s = it % NS;
                const long long c0 = prof ? clock64() : 0;
                ok = ok && mbar_wait(&bar_empty[s], ((it / NS) & 1) ^ 1, dead);
                if (prof) t_wait1 += clock64() - c0;
                if (!ok) break;
                mbar_expect_tx(&bar_full[s], stage_bytes);
                const uint32_t sa = smem_u32(smem) + s * stage_bytes, sbb = sa + 4 * A_PLANE_BYTES;
                // one instruction per box: all four planes of the A tile, all four planes of each 64-column group of the B tile
                tma_load_4d(sa, ma, kb * KB_ELEMS, rb * 128, 0, za, &bar_full[s]);
                for (int g = 0; g < NGT; ++g)
                  tma_load_4d(sbb + g * 4 * B_GROUP_BYTES, mb, nh * NT0 + g * 64, kb * KB_ELEMS, 0, zb, &bar_full[s]);
              }
        }
        base_ph += (uint32_t)(nops * nz);
        cat = 2;
        for (int i = 0; i < NQ; ++i) need(i, base_ph);   // catch up with the round's last phases (keeps the parity bookkeeping exact)
      }
      if (prof) { q.prof[(size_t)blockIdx.x * 8 + 0] = t_wait0; q.prof[(size_t)blockIdx.x * 8 + 1] = t_wait1;
                  q.prof[(size_t)blockIdx.x * 8 + 6] = t_cat[0]; q.prof[(size_t)blockIdx.x * 8 + 7] = t_cat[1]; }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc0 = (1u << 4) | (1u << 16) | ((128u >> 4) << 24);          // D f32, A/B f16, A K-major, B MN-major, M = 128
    const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    const uint32_t tbase = __shfl_sync(0xffffffffu, taddr, 0);
    uint32_t it = 0, ti = 0;                       // ring fill counter, tile counter (accumulator buffer = ti % NBUF)
    bool ok = true;
    for (long long it0 = (long long)blockIdx.x * ILV; it0 < q.items && ok; it0 += (long long)gridDim.x * ILV) {
      const int nz = (int)min((long long)ILV, q.items - it0);
      const int nops = item_nops(q, it0);
      for (int jz = 0; jz < nops * nz && ok; ++jz)
        for (int rb = 0; rb < RB && ok; ++rb)
          for (int nh = 0; nh < NH && ok; ++nh, ++ti) {
            const int nt = nh == 0 ? NT0 : N16 - NT0;                              // columns of this tile
            const uint32_t idesc = idesc0 | ((uint32_t)(nt >> 3) << 17);
            const uint32_t idesc_na = idesc | (1u << 13);                          // negate A
            const int buf = (int)(ti % NBUF);
            const uint32_t use = ti / NBUF;
            long long c0 = prof ? clock64() : 0;
            ok = __all_sync(0xffffffffu, mbar_wait(&bar_tempty[buf], (use & 1) ^ 1, dead));   // accumulator drained by the epilogue
            if (!ok) break;
            if (prof) t_wait0 += clock64() - c0;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t dr = tbase + (uint32_t)(buf * 2 * DIOFF), di = dr + (uint32_t)DIOFF;
            for (int kb = 0; kb < KBLK && ok; ++kb, ++it) {
              const int s = it % NS;
              c0 = prof ? clock64() : 0;
              ok = __all_sync(0xffffffffu, mbar_wait(&bar_full[s], (it / NS) & 1, dead));
              if (!ok) break;
              if (prof) t_wait1 += clock64() - c0;
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              if (elect_one()) {
                const uint32_t sa = sbase + s * stage_bytes, sbb = sa + 4 * A_PLANE_BYTES;
#pragma unroll
                for (int k16 = 0; k16 < 2; ++k16) {
                  uint64_t da[4], db[4];
#pragma unroll
                  for (int pl = 0; pl < 4; ++pl) {
                    da[pl] = make_desc(sa + pl * A_PLANE_BYTES + k16 * 32, q.a_lbo, q.a_sbo, 4);
                    db[pl] = make_desc(sbb + pl * B_GROUP_BYTES + k16 * 2048, q.b_lbo, q.b_sbo, 2);
                  }
                  const uint32_t first = (kb == 0 && k16 == 0) ? 0u : 1u;
                  // planes: 0 = Re h0, 1 = Re h1, 2 = Im h0, 3 = Im h1
                  mma_f16_ss(dr, da[0], db[1], idesc, first);             // Ar0 Br1
                  mma_f16_ss(dr, da[1], db[0], idesc, 1u);                // Ar1 Br0
                  mma_f16_ss(dr, da[2], db[3], idesc_na, 1u);             // -Ai0 Bi1
                  mma_f16_ss(dr, da[3], db[2], idesc_na, 1u);             // -Ai1 Bi0
                  mma_f16_ss(dr, da[0], db[0], idesc, 1u);                // Ar0 Br0
                  mma_f16_ss(dr, da[2], db[2], idesc_na, 1u);             // -Ai0 Bi0
                  mma_f16_ss(di, da[0], db[3], idesc, first);             // Ar0 Bi1
                  mma_f16_ss(di, da[1], db[2], idesc, 1u);                // Ar1 Bi0
                  mma_f16_ss(di, da[2], db[1], idesc, 1u);                // Ai0 Br1
                  mma_f16_ss(di, da[3], db[0], idesc, 1u);                // Ai1 Br0
                  mma_f16_ss(di, da[0], db[2], idesc, 1u);                // Ar0 Bi0
                  mma_f16_ss(di, da[2], db[0], idesc, 1u);                // Ai0 Br0
                }
                umma_commit(&bar_empty[s]);
                if (kb == KBLK - 1) umma_commit(&bar_tfull[buf]);
              }
              __syncwarp();
            }
          }
    }
    if (prof && lane == 0) { q.prof[(size_t)blockIdx.x * 8 + 2] = t_wait0; q.prof[(size_t)blockIdx.x * 8 + 3] = t_wait1; }
  } else {
    // ------------------------------------------------------------------ epilogue warps (TMEM lane quarter = warp % 4)
    const int qd = warp & 3;
    const int half = (warp - 2) >> 2;              // which of the two warps of this lane quarter: even / odd 16-column chunks
    const int et = (warp - 2) * 32 + lane;
    const int lrow = qd * 32 + lane;
    uint32_t ti = 0, round = 0;
    bool ok = true;
    for (long long it0 = (long long)blockIdx.x * ILV; it0 < q.items && ok; it0 += (long long)gridDim.x * ILV, ++round) {
      const int nz = (int)min((long long)ILV, q.items - it0);
      const int nops = item_nops(q, it0);
      // generators of one round: X' = xscale (A_0 + sum_k u_k A_k), u_k = maxA_k sin(base[b][k][t])  (init_tf_ops_weight,
      // :168-185), into the X slot of that round's parity; round r + 1 is assembled in the middle of round r
      const bool sparse_x = q.pat_n > 0 && q.pat_n * 3 < n * n;
    auto build_x = [&](long long bi0, int bnz, int rpar) {
       for (int z = 0; z < bnz; ++z) {
        const long long item = bi0 + z;
        const long long b = item / q.T;
        const int t = (int)(item % q.T);
        epi_bar();                                   // wts of the previous item are no longer read
        if (et <= q.K) wts[et] = et == 0 ? q.xscale : (float)(q.maxA[et - 1] * sin(q.ctrl[((size_t)b * q.K + et - 1) * q.T + t])) * q.xscale;
        epi_bar();
        __half* X = q.base[TC_CLS_SCR] + (size_t)(((long long)blockIdx.x * ILV + z) * TC_NSLOT + (rpar ? 4 : 0)) * mat;
        const int l16 = ld >> 4;
        const size_t nn = (size_t)n * n;
        if (sparse_x) {
          for (int e = et; e < q.pat_n; e += NEPI) scatter_x_entry(q, X, plane, ld, e, wts);
        } else
        for (int i16 = et; i16 < n * l16; i16 += NEPI) {
          const int r = i16 / l16, c16 = (i16 - r * l16) * 16;
          float re[16], im[16];
#pragma unroll
          for (int cc = 0; cc < 16; ++cc) re[cc] = im[cc] = 0.f;
          // 16 consecutive complex entries of A_k = 128 bytes: eight independent 16-byte loads per k, two k in flight
          // (a per-element loop over k serialises on the load latency: 345k cycles per item at n = 216)
          const bool full = c16 + 16 <= n && (n & 1) == 0;           // 16-byte alignment of the row segment
          for (int k0 = 0; k0 <= q.K; k0 += 2) {
            float4 v[2][8];
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const int k = k0 + kk;
              const float2* src = q.A_f + (size_t)k * nn + (size_t)r * n + c16;
#pragma unroll
              for (int h = 0; h < 8; ++h) {
                if (k <= q.K && full) v[kk][h] = __ldg(reinterpret_cast<const float4*>(src) + h);
                else if (k <= q.K) {
                  const float2 a0 = c16 + 2 * h < n ? __ldg(src + 2 * h) : make_float2(0.f, 0.f);
                  const float2 a1 = c16 + 2 * h + 1 < n ? __ldg(src + 2 * h + 1) : make_float2(0.f, 0.f);
                  v[kk][h] = make_float4(a0.x, a0.y, a1.x, a1.y);
                } else v[kk][h] = make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const float w = k0 + kk <= q.K ? wts[k0 + kk] : 0.f;
#pragma unroll
              for (int h = 0; h < 8; ++h) {
                re[2 * h] = fmaf(w, v[kk][h].x, re[2 * h]); im[2 * h] = fmaf(w, v[kk][h].y, im[2 * h]);
                re[2 * h + 1] = fmaf(w, v[kk][h].z, re[2 * h + 1]); im[2 * h + 1] = fmaf(w, v[kk][h].w, im[2 * h + 1]);
              }
            }
          }
          store_planes16(X, plane, ld, r, c16, re, im);
        }
       }
        asm volatile("fence.proxy.async;" ::: "memory");
        epi_bar();                                   // every thread's stores precede the counter bump; X is read back (elementwise
        if (et == 0)                                 // source) by other threads than its writers
          asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(smem_u32(&xdone_cnt)) : "memory");
      };
      if (expm && round == 0) build_x(it0, nz, 0);
      if (q.prog == TC_PROG_CHAIN) {
        if (et == 0) q.scal[(size_t)it0 * 8 + 5] = 0.0;
        epi_bar();
      }
      for (int j = 0; j < nops && ok; ++j)
       for (int z = 0; z < nz && ok; ++z) {
        if (expm && z == 0 && j == (nops > 2 ? 2 : nops - 1)) {      // next round's generators, off the round boundary
          const long long nx0 = it0 + (long long)gridDim.x * ILV;
          if (nx0 < q.items) build_x(nx0, (int)min((long long)ILV, q.items - nx0), (int)((round + 1) & 1));
        }
        const long long item = it0 + z;
        OpR o; make_op(q, item, j, nops, z, (int)(round & 1), o);
        const __half* E = o.e_cls >= 0 ? q.base[o.e_cls] + (size_t)o.e_idx * mat : nullptr;
        __half* D1 = o.d1_cls >= 0 ? q.base[o.d1_cls] + (size_t)o.d1_idx * mat : nullptr;
        __half* D2 = o.d2_cls >= 0 ? q.base[o.d2_cls] + (size_t)o.d2_idx * mat : nullptr;
        const bool useE = E && (o.c1[1] != 0.f || o.c2[1] != 0.f);
        double sr = 0.0, si = 0.0;                   // row sums of the final product (unitary_scale), over both column halves
        for (int rb = 0; rb < RB && ok; ++rb)
          for (int nh = 0; nh < NH && ok; ++nh, ++ti) {
            const int nt = nh == 0 ? NT0 : N16 - NT0, cbase = nh * NT0;
            const int buf = (int)(ti % NBUF);
            const uint32_t use = ti / NBUF;
            const uint32_t lane_addr = taddr + ((uint32_t)(qd * 32) << 16) + (uint32_t)(buf * 2 * DIOFF);
            const int row = rb * 128 + lrow;
            const bool vrow = row < n;
            if (nh == 0) { sr = 0.0; si = 0.0; }
            // elementwise source (b_j X of a Horner step, 2 E of a squaring): the first chunk is fetched while the MMAs of
            // this tile still run, the next one right after the current one has been unpacked
            uint32_t e0[4][8];
            const bool ldE = useE && vrow;
            auto fetchE = [&](int c0) {
              if (ldE && c0 < nt) {
                const __half* p0 = E + (size_t)row * ld + cbase + c0;
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) ldg256(p0 + pl * plane, e0[pl]);
              }
            };
            fetchE(16 * half);
            long long c0t = prof ? clock64() : 0;
            ok = __all_sync(0xffffffffu, mbar_wait(&bar_tfull[buf], use & 1, dead));
            if (!ok) break;
            if (prof) { const long long c1t = clock64(); t_wait0 += c1t - c0t; c0t = c1t; }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // one 16-column chunk: accumulator rows out of TMEM, elementwise term, split, store
            for (int c0 = 16 * half; c0 < nt; c0 += CSTEP) {
              uint32_t ur[16], ui[16];
              tmem_ld16(lane_addr + (uint32_t)c0, ur);
              tmem_ld16(lane_addr + (uint32_t)(DIOFF + c0), ui);
              float er[16], ei[16];
              if (useE) { unpack16(e0[0], e0[1], er); unpack16(e0[2], e0[3], ei); fetchE(c0 + CSTEP); }
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
              if (!vrow) continue;
              const int col = cbase + c0;
#pragma unroll
              for (int comp = 0; comp < 2; ++comp) {      // Re: planes 0,1; Im: planes 2,3
                float ov[16];
                if (D1) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    const float d = __uint_as_float(comp == 0 ? ur[i] : ui[i]);
                    const float ee = useE ? (comp == 0 ? er[i] : ei[i]) : 0.f;
                    ov[i] = fmaf(o.c1[0], d, o.c1[1] * ee) + ((comp == 0 && col + i == row) ? o.c1[2] : 0.f);
                  }
                  store_comp16(D1 + (size_t)(2 * comp) * plane + (size_t)row * ld + col, plane, ov);
                }
                if (D2) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    const float d = __uint_as_float(comp == 0 ? ur[i] : ui[i]);
                    const float ee = useE ? (comp == 0 ? er[i] : ei[i]) : 0.f;
                    ov[i] = fmaf(o.c2[0], d, o.c2[1] * ee) + ((comp == 0 && col + i == row) ? o.c2[2] : 0.f);
                  }
                  store_comp16(D2 + (size_t)(2 * comp) * plane + (size_t)row * ld + col, plane, ov);
                }
              }
              if (o.f64out) {
                const double sc = 1.0 / ((double)(1 << TC_EU) * (double)(1 << TC_EU));
                double2* U = q.Ufin + (size_t)item * n * n + (size_t)row * n;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (col + i < n) {
                    const double xr = (double)__uint_as_float(ur[i]) * sc, xi = (double)__uint_as_float(ui[i]) * sc;
                    U[col + i] = make_double2(xr, xi);
                    sr += xr; si += xi;
                  }
              }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&bar_tempty[buf]);
            if (o.f64out && nh == NH - 1) {
              // unitary_scale = (1/n) sum_r |sum_c X_rc|^2  (init_tf_propagator, tensorflow_state.py:225 on the real embedding);
              // the two warps of a lane quarter hold the even / odd column chunks of the same rows
              rs[half][lrow][0] = sr; rs[half][lrow][1] = si;
              if (NHALF == 1) { rs[1][lrow][0] = 0.0; rs[1][lrow][1] = 0.0; }
              epi_bar();
              if (half == 0) {
                const double tr = rs[0][lrow][0] + rs[1][lrow][0], tq = rs[0][lrow][1] + rs[1][lrow][1];
                double v = vrow ? (tr * tr + tq * tq) / (double)n : 0.0;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                if (lane == 0) atomicAdd(&q.scal[(size_t)item * 8 + 5], v);
              }
              epi_bar();
            }
            asm volatile("fence.proxy.async;" ::: "memory");
            epi_bar();                                 // this quadrant of the product's outputs is complete: all threads' stores,
            if (et == 0)                               // then one release-increment of its counter
              asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(smem_u32(&done_cnt[rb * NH + nh])) : "memory");
            if (prof) t_work += clock64() - c0t;
          }
      }
    }
    if (prof && et == 0) { q.prof[(size_t)blockIdx.x * 8 + 4] = t_wait0; q.prof[(size_t)blockIdx.x * 8 + 5] = t_work; }
  }
  // ---------------------------------------------------------------------- teardown
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (dead_s && tid == 0 && q.err_flag) atomicExch(q.err_flag, 1);
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"((uint32_t)q.tmem_cols) : "memory");
}

}  // namespace

// ------------------------------------------------------------------------------------------ host side

bool tc_geometry(int n, TcGeom* g) {
  if (n < 1 || n > TC_MAX_N) return false;
  g->n = n;
  g->ld = tc_ld(n);
  g->N16 = (n + 15) / 16 * 16;
  g->NH = g->N16 > 128 ? 2 : 1;
  g->NT0 = g->NH == 2 ? 128 : g->N16;
  g->NGT = (g->NT0 + 63) / 64;
  g->DIOFF = (g->NT0 + 31) / 32 * 32;
  g->RB = n > 128 ? 2 : 1;
  g->NBUF = g->RB * g->NH > 1 ? 2 : 1;
  g->KBLK = (n + KB_ELEMS - 1) / KB_ELEMS;
  int cols = 32;
  while (cols < g->NBUF * 2 * g->DIOFF) cols *= 2;
  g->tmem_cols = cols;
  const size_t stage = 4 * (size_t)A_PLANE_BYTES + 4 * (size_t)g->NGT * B_GROUP_BYTES;
  const size_t budget = 225 * 1024;
  g->ctas_per_sm = 1;
  int st = (int)((budget - 2048) / stage);
  if (st > MAX_STAGES) st = MAX_STAGES;
  if (st < 1) return false;
  g->stages = st;
  g->smem = (size_t)st * stage + 1024;
  g->mat_halfs = (size_t)4 * n * g->ld;
  return true;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

const char* tc_make_map(CUtensorMap* map, const void* base, int n, int ld, unsigned long long n_mats, bool b_form) {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || !p ||
        qr != cudaDriverEntryPointSuccess)
      return "cuTensorMapEncodeTiled is not available from the driver";
    fn = (PFN_encodeTiled)p;
  }
  // rank 4: {column, row, plane, matrix}; a box spans the four planes, so one TMA instruction moves a whole operand tile
  const cuuint64_t dims[4] = {(cuuint64_t)n, (cuuint64_t)n, 4, (cuuint64_t)n_mats};
  const cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)n * ld * 2, (cuuint64_t)4 * n * ld * 2};
  const cuuint32_t box_a[4] = {KB_ELEMS, 128, 4, 1}, box_b[4] = {64, KB_ELEMS, 4, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, b_form ? box_b : box_a, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, b_form ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    static char msg[96];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return msg;
  }
  return nullptr;
}

const char* tc_make_store_map(CUtensorMap* map, const void* base, int n, int ld, unsigned long long n_mats) {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || !p ||
        qr != cudaDriverEntryPointSuccess)
      return "cuTensorMapEncodeTiled is not available from the driver";
    fn = (PFN_encodeTiled)p;
  }
  // columns up to the row pitch: the padding columns of a result are written (zeros), as the thread-per-row stores do
  const cuuint64_t dims[4] = {(cuuint64_t)ld, (cuuint64_t)n, 4, (cuuint64_t)n_mats};
  const cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)n * ld * 2, (cuuint64_t)4 * n * ld * 2};
  const cuuint32_t box[4] = {16, 32, 4, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    static char msg[96];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled (store map) failed with CUresult %d", (int)r);
    return msg;
  }
  return nullptr;
}

void tc_pick_scales(double xmax, double theta, int* eX, int* eY) {
  // stored magnitudes stay below 2^14 = 16384 (fp16 overflows at 65504)
  auto pick = [](double bound) {
    if (!(bound > 0.0)) bound = 1e-30;
    int e = (int)floor(14.0 - log2(bound));
    if (e > 60) e = 60;
    if (e < -20) e = -20;
    return e;
  };
  *eX = pick(xmax);
  *eY = pick(theta * theta);
}

void tc_build_expm_ops(int p, int s, int eX, int eY, std::vector<TcExpmOp>& ops) {
  // S = sum_{j<=p} X^j/j! = sum_{i<=J} Y^i (a_i I + b_i X), Y = X^2, a_i = 1/(2i)!, b_i = 1/(2i+1)! (b_J = 0 for even p),
  // evaluated by Horner in Y (Paterson-Stockmeyer, block 2); then s squarings.  Slots: 0 = X, 1 = Y, 2/3 = Z ping-pong.
  ops.clear();
  double fact[64];
  fact[0] = 1.0;
  for (int i = 1; i < 64; ++i) fact[i] = fact[i - 1] * (double)i;
  const int J = p / 2;
  auto a = [&](int i) { return 1.0 / fact[2 * i]; };
  auto b = [&](int i) { return (2 * i + 1 <= p) ? 1.0 / fact[2 * i + 1] : 0.0; };
  const double sU = ldexp(1.0, TC_EU);
  auto clear = [](TcExpmOp& o) { memset(&o, 0, sizeof(o)); o.d1 = o.d2 = -1; };
  TcExpmOp o;
  clear(o);
  o.sa = 0; o.sb = 0;                                     // D = X X  (units 2^(2 eX))
  o.d1 = 1; o.c1[0] = (float)ldexp(1.0, eY - 2 * eX);     // Y
  o.d2 = 2;                                               // first Horner value Z
  int next_j;
  if ((p & 1) == 0 && J >= 1) {                           // Z = a_J Y + b_{J-1} X + a_{J-1} I
    o.c2[0] = (float)(a(J) * ldexp(1.0, TC_EU - 2 * eX));
    o.c2[1] = (float)(b(J - 1) * ldexp(1.0, TC_EU - eX));
    o.c2[2] = (float)(a(J - 1) * sU);
    next_j = J - 2;
  } else {                                                // Z = b_J X + a_J I
    o.c2[0] = 0.f;
    o.c2[1] = (float)(b(J) * ldexp(1.0, TC_EU - eX));
    o.c2[2] = (float)(a(J) * sU);
    next_j = J - 1;
  }
  ops.push_back(o);
  int cur = 2;
  // With squarings the program tracks E = P - I instead of P: (I + E)^2 = I + (2E + E^2).  The identity never enters a
  // tensor-core accumulation, so the accumulator's truncation error is relative to |E^2| (tiny for the early, most
  // amplified squarings) instead of to the O(1) diagonal -- the fp32 analogue of an expm1-style squaring phase.
  const bool eform = s > 0;
  if (eform && next_j < 0) ops.back().c2[2] -= (float)sU;       // no Horner step: the first value already is S
  for (int j = next_j; j >= 0; --j) {                     // Z <- Z Y + b_j X + a_j I
    clear(o);
    o.sa = (int8_t)cur; o.sb = 1;
    o.d1 = (int8_t)(cur ^ 1);
    o.c1[0] = (float)ldexp(1.0, -eY);
    o.c1[1] = (float)(b(j) * ldexp(1.0, TC_EU - eX));
    o.c1[2] = (float)((a(j) - ((eform && j == 0) ? 1.0 : 0.0)) * sU);
    ops.push_back(o);
    cur ^= 1;
  }
  for (int i = 0; i < s; ++i) {                           // E <- E E + 2 E   (+ I after the last one)
    clear(o);
    o.sa = o.sb = o.se = (int8_t)cur;                     // 2 E is added by the epilogue in fp32 round-to-nearest: riding it on the
    o.d1 = (int8_t)(cur ^ 1);                             // MMA stream (B = E + 2I, or a diagonal tile) costs a decade of accuracy --
    o.c1[0] = (float)ldexp(1.0, -TC_EU);                  // the accumulator would hold |2E| during the truncating accumulations
    o.c1[1] = 2.0f;
    o.c1[2] = (i == s - 1) ? (float)sU : 0.0f;
    ops.push_back(o);
    cur ^= 1;
  }
  // the last value goes to the propagator cache instead of a scratch slot
  TcExpmOp& last = ops.back();
  if (ops.size() == 1) last.d2 = TC_SLOT_OUT; else last.d1 = TC_SLOT_OUT;
}

void tc_pack_host(const double* z, int n, int ld, int e, __half* out) {
  const double sc = ldexp(1.0, e);
  const size_t plane = (size_t)n * ld;
  for (size_t i = 0; i < 4 * plane; ++i) out[i] = __float2half(0.f);
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c)
      for (int ri = 0; ri < 2; ++ri) {
        const float v = (float)(z[((size_t)r * n + c) * 2 + ri] * sc);
        const __half h0 = __float2half_rn(v);
        const __half h1 = __float2half_rn(v - __half2float(h0));
        out[(size_t)(2 * ri) * plane + (size_t)r * ld + c] = h0;
        out[(size_t)(2 * ri + 1) * plane + (size_t)r * ld + c] = h1;
      }
}

cudaError_t tc_launch(const TcParams& q_in, const TcMaps& maps, const TcGeom& g, int grid, cudaStream_t st) {
  TcParams q = q_in;
  q.n = g.n; q.ld = g.ld; q.N16 = g.N16; q.NT0 = g.NT0; q.NH = g.NH; q.NGT = g.NGT; q.DIOFF = g.DIOFF; q.NBUF = g.NBUF; q.RB = g.RB; q.KBLK = g.KBLK; q.stages = g.stages; q.tmem_cols = g.tmem_cols;
  if (!q.a_sbo) { q.a_lbo = 1; q.a_sbo = 512 >> 4; }                    // K-major SWIZZLE_64B: 8-row groups 512 B apart
  if (!q.b_sbo) { q.b_lbo = (4 * B_GROUP_BYTES) >> 4; q.b_sbo = 1024 >> 4; }  // MN-major SWIZZLE_128B: 64-column groups 16 KB apart ([group][plane] tiles), k-atoms 1024 B
  cudaError_t e = cudaFuncSetAttribute(k_tc_prog, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) return e;
  if (grid < 1) grid = 1;
  if (q.ilv < 1 || q.prog != TC_PROG_EXPM) q.ilv = 1;
  k_tc_prog<<<grid, NTHREADS, g.smem, st>>>(q, maps);
  return cudaGetLastError();
}
