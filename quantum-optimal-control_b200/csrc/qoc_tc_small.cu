// tcgen05 / TMEM propagator program for SMALL Hilbert dimensions (n <= 64) in the QOC_F16X2 arithmetic: the whole
// Paterson-Stockmeyer + squaring program of a (b,t) item (get_matexp / matexp_op, core/tensorflow_state.py:25-46,70-75)
// runs out of SHARED MEMORY -- no global round trip between its dependent products, which is what bounds the streaming
// engine (qoc_tc_f16.cu) at these sizes.
//
// Image.  A matrix plane is ONE shared-memory image [RP rows][128 bytes] (64 fp16 columns, RP = n rounded up to 16) with
// the 16-byte chunks of row r XOR-swizzled by (r & 7) -- byte for byte both the canonical K-major SWIZZLE_128B layout of
// an A operand (row = M index, columns = K) and the canonical MN-major SWIZZLE_128B layout of a B operand (row = K
// index, columns = N) of the same row-major matrix.  So X, Y = X^2 and the running value Z exist once each (4 planes:
// Re h0, Re h1, Im h0, Im h1) and are read in place as whichever operand a product needs; Z is updated in place by the
// epilogue (the MMAs that read it have completed by then).
//
// Items in flight.  An MMA has M = 128 accumulator rows = TMEM lanes; an item only needs n of them.  The A descriptor of
// slot s starts roff_s rows BEFORE its image, which puts the item's rows on lanes roff_s .. roff_s + n - 1 (the other
// lanes multiply whatever finite bytes surround the image and are never read).  Slots use disjoint lane ranges and
// disjoint TMEM columns, so 4 (n <= 32) or 2 (n <= 64) items are in flight per CTA: while the epilogue warps of one slot
// re-split and store its result, the tensor pipe works on the others.
//
// Roles: warps 0 .. NSL-1 = MMA issuers, one per slot (one elected lane: per product 12 MMAs per 16 columns of K, commit ->
// acc[slot]); the remaining warps = epilogue, TMEM lane quarter = warp % 4, up to three warps per quarter (each owns every third 16-column chunk:
// a lone warp per scheduler cannot hide its own latencies); thread = accumulator row.  Per product of a slot: wait acc[slot],
// tcgen05.ld the row, out = c0 D + c1 E + c2 I with E read from the X / Z image, fp16-pair split, swizzled 16-byte
// stores into the destination image(s), fence.proxy.async, arrive on ready[slot].  The last product of an item goes to
// the propagator cache in HBM (split fp16 planes, 32-byte sector stores).
#include "qoc_tc_dev.cuh"
#include <math.h>
#include <string.h>
#include <stdlib.h>

namespace {

constexpr int S_MAXWPQ = 3;                      // epilogue warps per TMEM lane quarter (each owns every WPQ-th 16-column chunk)
constexpr uint32_t GUARD_BYTES = 8 * 1024;       // readable finite bytes before the first image (slot 0 starts pad <= 63 rows early)
constexpr uint32_t TAIL_GUARD_BYTES = 4 * 1024;  // ... and after the last one (its 128-row window overshoots by < 2 KB)

struct SmallGeom { int n, N16, K16, RP, NSL, QPS, WPQ, DIOFF, pad, wide; uint32_t plane_bytes, mat_bytes, slot_bytes; size_t smem; int tmem_cols; };

__host__ __device__ inline bool small_geometry(int n, SmallGeom& g) {
  if (n < 1 || n > 64) return false;
  g.n = n;
  g.N16 = (n + 15) / 16 * 16;
  g.K16 = g.N16 / 16;
  g.RP = g.N16;                                  // rows n .. RP-1 stay zero: they are the K padding of the B operand
  g.NSL = n <= 32 ? 4 : 2;
  g.QPS = 4 / g.NSL;
  g.DIOFF = (g.N16 + 31) / 32 * 32;
  g.WPQ = g.K16 < S_MAXWPQ ? g.K16 : S_MAXWPQ;
  g.pad = ((32 * g.QPS - n) / 2 + 4) / 8 * 8;    // centres the item's rows in its lane range (multiple of 8: keeps the swizzle phase)
  if (g.pad + n > 32 * g.QPS) g.pad = (32 * g.QPS - n) / 8 * 8;
  g.plane_bytes = (uint32_t)g.RP * 128;
  g.mat_bytes = 4 * g.plane_bytes;
  g.slot_bytes = 3 * g.mat_bytes;                // X, Y, Z
  g.smem = (size_t)GUARD_BYTES + TAIL_GUARD_BYTES + (size_t)g.NSL * g.slot_bytes + 1024;
  // two slots (32 < n <= 64): "wide" products.  The Re and Im planes of the B operand are two 64-column swizzle atoms
  // 2 plane_bytes apart, so ONE N = 128 MMA multiplies an A plane with [Br | Bi]: D1 = Ar [Br | Bi], D2 = Ai [Br | Bi],
  // Dr = D1[0:64) - D2[64:128), Di = D1[64:128) + D2[0:64) in the epilogue -- 18 instead of 36 MMAs per product (the
  // kernel is bound by the ~45 issue cycles of each tcgen05.mma, not by their width)
  g.wide = g.NSL == 2 ? 1 : 0;
  if (g.wide) g.DIOFF = 128;
  int cols = 32;
  while (cols < g.NSL * 2 * g.DIOFF) cols *= 2;
  g.tmem_cols = cols;
  return g.smem <= 227 * 1024;
}

// byte offset of the 16-byte chunk holding columns [8 c8, 8 c8 + 8) of row r inside a plane image
DEVINL uint32_t img_off(int r, int c8) { return (uint32_t)(r * 128 + ((c8 ^ (r & 7)) << 4)); }

DEVINL void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
DEVINL void lds128(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}
// 16 columns of one component (planes h0 at `p0`, h1 at `p0 + plane_bytes`) of image row r
DEVINL void img_store16(uint32_t p0, uint32_t plane_bytes, int r, int c0, const float (&v)[16]) {
  uint32_t a0[8], a1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split2(v[2 * i], v[2 * i + 1], a0[i], a1[i]);
  const uint32_t o0 = img_off(r, c0 >> 3), o1 = img_off(r, (c0 >> 3) + 1);
  sts128(p0 + o0, a0[0], a0[1], a0[2], a0[3]);
  sts128(p0 + o1, a0[4], a0[5], a0[6], a0[7]);
  sts128(p0 + plane_bytes + o0, a1[0], a1[1], a1[2], a1[3]);
  sts128(p0 + plane_bytes + o1, a1[4], a1[5], a1[6], a1[7]);
}
DEVINL void img_load16(uint32_t p0, uint32_t plane_bytes, int r, int c0, float (&v)[16]) {
  uint32_t a0[8], a1[8];
  const uint32_t o0 = img_off(r, c0 >> 3), o1 = img_off(r, (c0 >> 3) + 1);
  lds128(p0 + o0, a0[0], a0[1], a0[2], a0[3]);
  lds128(p0 + o1, a0[4], a0[5], a0[6], a0[7]);
  lds128(p0 + plane_bytes + o0, a1[0], a1[1], a1[2], a1[3]);
  lds128(p0 + plane_bytes + o1, a1[4], a1[5], a1[6], a1[7]);
  unpack16(a0, a1, v);
}

__global__ void __launch_bounds__(448, 1) k_tc_small_expm(const TcParams q, const SmallGeom g) {
  const int S_NTHREADS = blockDim.x;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_ready[4], bar_acc[4];
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int dead_s;
  __shared__ float wts[4][32];
  __shared__ TcExpmOp ops_s[TC_MAX_OPS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = g.n, N16 = g.N16, NSL = g.NSL, QPS = g.QPS;
  const int ld = tc_ld(n);
  const size_t gplane = (size_t)n * ld, gmat = 4 * gplane;
  const uint32_t img0 = smem_u32(smem) + GUARD_BYTES;                 // slot s: img0 + s * slot_bytes; matrices X, Y, Z
  const int nrounds_total = (int)((q.items + (long long)gridDim.x * NSL - 1) / ((long long)gridDim.x * NSL));
  volatile int* dead = &dead_s;

  for (uint32_t i = tid * 16; i < (uint32_t)(g.smem - 1024); i += S_NTHREADS * 16) sts128(smem_u32(smem) + i, 0u, 0u, 0u, 0u);
  if (tid == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(&bar_ready[s], 32 * QPS * g.WPQ); mbar_init(&bar_acc[s], 1); }
    dead_s = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)g.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = *(volatile uint32_t*)&tmem_base_s;
  const int nops = q.nops;
  for (int i = tid; i < nops; i += S_NTHREADS) ops_s[i] = q.ops[i];
  __syncthreads();

  if (warp < NSL) {
    // ------------------------------------------------------------------ MMA issuers: warp s serves slot s (no head-of-line
    // blocking between slots, and the ~45 issue cycles per tcgen05.mma are spread over NSL warps)
    const int s = warp;
    const uint32_t idesc = (1u << 4) | (1u << 16) | ((uint32_t)(N16 >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_na = idesc | (1u << 13);
    const uint32_t tbase = __shfl_sync(0xffffffffu, taddr, 0);
    const uint32_t ibase = __shfl_sync(0xffffffffu, img0, 0);
    const uint32_t sl = ibase + (uint32_t)s * g.slot_bytes;
    const uint32_t roff = (uint32_t)(s * 32 * QPS + g.pad);
    const uint32_t dr = tbase + (uint32_t)(s * 2 * g.DIOFF), di = dr + (uint32_t)g.DIOFF;
    uint32_t phase = 0;
    bool ok = true;
    for (int round = 0; round < nrounds_total && ok; ++round)
      for (int j = 0; j < nops && ok; ++j, ++phase) {
        const TcExpmOp e = ops_s[j];
        const int ma = e.sa == 0 ? 0 : e.sa == 1 ? 1 : 2, mb = e.sb == 0 ? 0 : e.sb == 1 ? 1 : 2;     // X, Y, Z images
        ok = __all_sync(0xffffffffu, mbar_wait(&bar_ready[s], phase & 1, dead));      // operands written, accumulator drained
        if (!ok) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t abase = sl + (uint32_t)ma * g.mat_bytes - roff * 128u;       // image row r <-> tile row r + roff
          const uint32_t bbase = sl + (uint32_t)mb * g.mat_bytes;
          uint64_t da[4], db[4];
#pragma unroll
          for (int pl = 0; pl < 4; ++pl) {
            da[pl] = make_desc(abase + pl * g.plane_bytes, 1, 1024 >> 4, 2);          // K-major SWIZZLE_128B
            db[pl] = make_desc(bbase + pl * g.plane_bytes, 1, 1024 >> 4, 2);          // MN-major SWIZZLE_128B
          }
          if (g.wide) {
            const uint32_t idw = (1u << 4) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);   // N = 128: [Re atom | Im atom]
            uint64_t dw[2];
            dw[0] = make_desc(bbase, (2 * g.plane_bytes) >> 4, 1024 >> 4, 2);                     // [Br h0 | Bi h0]
            dw[1] = make_desc(bbase + g.plane_bytes, (2 * g.plane_bytes) >> 4, 1024 >> 4, 2);     // [Br h1 | Bi h1]
            for (int k = 0; k < g.K16; ++k) {
              const uint32_t first = k == 0 ? 0u : 1u;
              mma_f16_ss(dr, da[0], dw[1], idw, first);        // D1 += Ar0 [B1]
              mma_f16_ss(di, da[2], dw[1], idw, first);        // D2 += Ai0 [B1]
              mma_f16_ss(dr, da[1], dw[0], idw, 1u);           // D1 += Ar1 [B0]
              mma_f16_ss(di, da[3], dw[0], idw, 1u);           // D2 += Ai1 [B0]
              mma_f16_ss(dr, da[0], dw[0], idw, 1u);           // D1 += Ar0 [B0]
              mma_f16_ss(di, da[2], dw[0], idw, 1u);           // D2 += Ai0 [B0]
#pragma unroll
              for (int pl = 0; pl < 4; ++pl) da[pl] += 32 >> 4;
              dw[0] += 2048 >> 4; dw[1] += 2048 >> 4;
            }
          } else
          for (int k = 0; k < g.K16; ++k) {
            const uint32_t first = k == 0 ? 0u : 1u;
            // Dr and Di alternate (independent accumulation chains)
            mma_f16_ss(dr, da[0], db[1], idesc, first);
            mma_f16_ss(di, da[0], db[3], idesc, first);
            mma_f16_ss(dr, da[1], db[0], idesc, 1u);
            mma_f16_ss(di, da[1], db[2], idesc, 1u);
            mma_f16_ss(dr, da[2], db[3], idesc_na, 1u);
            mma_f16_ss(di, da[2], db[1], idesc, 1u);
            mma_f16_ss(dr, da[3], db[2], idesc_na, 1u);
            mma_f16_ss(di, da[3], db[0], idesc, 1u);
            mma_f16_ss(dr, da[0], db[0], idesc, 1u);
            mma_f16_ss(di, da[0], db[2], idesc, 1u);
            mma_f16_ss(dr, da[2], db[2], idesc_na, 1u);
            mma_f16_ss(di, da[2], db[0], idesc, 1u);
#pragma unroll
            for (int pl = 0; pl < 4; ++pl) { da[pl] += 32 >> 4; db[pl] += 2048 >> 4; }   // next 16 columns of K: start address fields
          }
          umma_commit(&bar_acc[s]);
        }
        __syncwarp();
      }
  } else {
    // ------------------------------------------------------------------ epilogue warps: quarter = warp % 4, slot = quarter / QPS
    const int qd = warp & 3;
    const int cg = (warp - NSL) >> 2, WPQ = g.WPQ;                   // chunk group: this warp owns chunks cg, cg + WPQ, ...
    const int s = qd / QPS;
    const int rho = qd * 32 + lane;                                  // accumulator row (TMEM lane)
    const int roff = s * 32 * QPS + g.pad;
    const int r = rho - roff;                                        // matrix row of this thread
    const bool vrow = r >= 0 && r < n;
    const int et = ((qd - s * QPS) * WPQ + cg) * 32 + lane;          // thread index within the slot
    const uint32_t sl = img0 + (uint32_t)s * g.slot_bytes;
    const uint32_t imgX = sl, imgY = sl + g.mat_bytes, imgZ = sl + 2 * g.mat_bytes;
    const uint32_t lane_addr = taddr + ((uint32_t)(qd * 32) << 16) + (uint32_t)(s * 2 * g.DIOFF);
    const uint32_t pb = g.plane_bytes;
    uint32_t phase = 0;
    bool ok = true;
    const size_t nn = (size_t)n * n;
    const bool prof = q.prof != nullptr && warp == NSL && lane == 0;
    long long tp[6] = {0, 0, 0, 0, 0, 0};            // cycles: X assembly, wait acc, tmem ld, math + stores, fences + arrive
    // generator X' = xscale (A_0 + sum_k u_k A_k), u_k = maxA_k sin(base[b][k][t]) (init_tf_ops_weight, tensorflow_state.py:168-185)
    // of the item this slot processes in round `rnd`, into the X image
    const bool sparse_x = q.pat_n > 0 && q.pat_n * 3 < n * n;
    auto assemble_x = [&](int rnd) {
      const long long it = ((long long)rnd * gridDim.x + blockIdx.x) * NSL + s;
      const bool v = it < q.items;
      if (v) {
        const long long b = it / q.T;
        const int t = (int)(it % q.T);
        if (et <= q.K) wts[s][et] = et == 0 ? q.xscale : (float)(q.maxA[et - 1] * sin(q.ctrl[((size_t)b * q.K + et - 1) * q.T + t])) * q.xscale;
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + s), "r"(32 * QPS * WPQ) : "memory");
      if (v && sparse_x) {
        // sparse generators: one thread per entry of the union pattern (the rest of the X image is zero and stays zero)
        for (int e = et; e < q.pat_n; e += 32 * QPS * WPQ) {
          const int rc = __ldg(q.pat_rc + e), pr = rc >> 16, pc = rc & 0xffff;
          const float2* cf = q.pat_coef_f + (size_t)e * (q.K + 1);
          float xr = 0.f, xi = 0.f;
          for (int k = 0; k <= q.K; ++k) { const float2 a = __ldg(cf + k); xr = fmaf(wts[s][k], a.x, xr); xi = fmaf(wts[s][k], a.y, xi); }
          unsigned short r0, r1, i0, i1;
          split1(xr, r0, r1); split1(xi, i0, i1);
          const uint32_t o = imgX + img_off(pr, pc >> 3) + (uint32_t)((pc & 7) << 1);
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(o), "h"(r0) : "memory");
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(o + pb), "h"(r1) : "memory");
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(o + 2 * pb), "h"(i0) : "memory");
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(o + 3 * pb), "h"(i1) : "memory");
        }
      } else if (v && vrow) {
        for (int c0 = 16 * cg; c0 < N16; c0 += 16 * WPQ) {
          float re[16], im[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) re[i] = im[i] = 0.f;
          for (int k = 0; k <= q.K; ++k) {
            const float w = wts[s][k];
            const float2* src = q.A_f + (size_t)k * nn + (size_t)r * n + c0;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c0 + i < n) { const float2 a = __ldg(src + i); re[i] = fmaf(w, a.x, re[i]); im[i] = fmaf(w, a.y, im[i]); }
          }
          img_store16(imgX, pb, r, c0, re);
          img_store16(imgX + 2 * pb, pb, r, c0, im);
        }
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + s), "r"(32 * QPS * WPQ) : "memory");      // wts may be rewritten after this
    };
    // the last product (a squaring whenever s >= 1) does not touch X: the next round's generator is assembled while the
    // tensor pipe works on it
    const TcExpmOp elast = ops_s[nops - 1];
    const bool early_x = nops >= 2 && elast.sa != 0 && elast.sb != 0 && (elast.se != 0 || (elast.c1[1] == 0.f && elast.c2[1] == 0.f));
    for (int round = 0; round < nrounds_total && ok; ++round) {
      const long long item = ((long long)round * gridDim.x + blockIdx.x) * NSL + s;
      const bool valid = item < q.items;
      long long c0p = prof ? clock64() : 0;
      if (round == 0 || !early_x) assemble_x(round);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&bar_ready[s]);
      if (prof) { const long long c = clock64(); tp[0] += c - c0p; c0p = c; }
      for (int j = 0; j < nops && ok; ++j, ++phase) {
        const TcExpmOp e = ops_s[j];
        if (early_x && j == nops - 1 && round + 1 < nrounds_total) {
          assemble_x(round + 1);                     // hidden behind the MMAs of the last product
          if (prof) { const long long c = clock64(); tp[0] += c - c0p; c0p = c; }
        }
        ok = __all_sync(0xffffffffu, mbar_wait(&bar_acc[s], phase & 1, dead));
        if (!ok) break;
        if (prof) { const long long c = clock64(); tp[1] += c - c0p; c0p = c; }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t imgE = e.se == 0 ? imgX : imgZ;
        const bool useE = e.c1[1] != 0.f || e.c2[1] != 0.f;
        auto dst_of = [&](int d) -> uint32_t { return d == 1 ? imgY : imgZ; };
        for (int c0 = 16 * cg; c0 < N16; c0 += 16 * WPQ) {
          uint32_t ur[16], ui[16];
          if (g.wide) {                              // Dr = D1[c0] - D2[64 + c0], Di = D1[64 + c0] + D2[c0]
            uint32_t t1[16], t2[16];
            tmem_ld16(lane_addr + (uint32_t)c0, ur);
            tmem_ld16(lane_addr + (uint32_t)(128 + 64 + c0), t1);
            tmem_ld16(lane_addr + (uint32_t)(64 + c0), ui);
            tmem_ld16(lane_addr + (uint32_t)(128 + c0), t2);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              ur[i] = __float_as_uint(__uint_as_float(ur[i]) - __uint_as_float(t1[i]));
              ui[i] = __float_as_uint(__uint_as_float(ui[i]) + __uint_as_float(t2[i]));
            }
          } else {
            tmem_ld16(lane_addr + (uint32_t)c0, ur);
            tmem_ld16(lane_addr + (uint32_t)(g.DIOFF + c0), ui);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          }
          if (prof) { const long long c = clock64(); tp[2] += c - c0p; c0p = c; }
          if (!(valid && vrow)) continue;
#pragma unroll
          for (int comp = 0; comp < 2; ++comp) {
            float ev[16], ov[16];
            if (useE) img_load16(imgE + 2 * comp * pb, pb, r, c0, ev);
            else {
#pragma unroll
              for (int i = 0; i < 16; ++i) ev[i] = 0.f;
            }
            if (e.d1 >= 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float d = __uint_as_float(comp == 0 ? ur[i] : ui[i]);
                ov[i] = fmaf(e.c1[0], d, e.c1[1] * ev[i]) + ((comp == 0 && c0 + i == r) ? e.c1[2] : 0.f);
              }
              if (e.d1 == TC_SLOT_OUT)
                store_comp16(q.base[TC_CLS_P] + (size_t)item * gmat + (size_t)(2 * comp) * gplane + (size_t)r * ld + c0, gplane, ov);
              else img_store16(dst_of(e.d1) + 2 * comp * pb, pb, r, c0, ov);
            }
            if (e.d2 >= 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float d = __uint_as_float(comp == 0 ? ur[i] : ui[i]);
                ov[i] = fmaf(e.c2[0], d, e.c2[1] * ev[i]) + ((comp == 0 && c0 + i == r) ? e.c2[2] : 0.f);
              }
              if (e.d2 == TC_SLOT_OUT)
                store_comp16(q.base[TC_CLS_P] + (size_t)item * gmat + (size_t)(2 * comp) * gplane + (size_t)r * ld + c0, gplane, ov);
              else img_store16(dst_of(e.d2) + 2 * comp * pb, pb, r, c0, ov);
            }
          }
        }
        if (prof) { const long long c = clock64(); tp[3] += c - c0p; c0p = c; }
        if (j < nops - 1) {                        // the next product of this item may start (the last one hands over to the
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generator assembly of the next round)
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          mbar_arrive(&bar_ready[s]);
        }
        if (prof) { const long long c = clock64(); tp[4] += c - c0p; c0p = c; }
      }
    }
    if (prof) for (int i = 0; i < 5; ++i) q.prof[(size_t)blockIdx.x * 8 + i] = (unsigned long long)tp[i];
    // phases: per round the MMA warp waits ready[s] once per product; arrivals: one after X, one after each product but
    // the last -- equal counts, and the final arrival of a round is the X of the next one
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (dead_s && tid == 0 && q.err_flag) atomicExch(q.err_flag, 1);
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"((uint32_t)g.tmem_cols) : "memory");
}

}  // namespace

bool tc_small_supported(int n) {
  SmallGeom g;
  return small_geometry(n, g);
}

cudaError_t tc_small_launch_expm(const TcParams& q, int n, int sm_count, cudaStream_t st) {
  SmallGeom g;
  if (!small_geometry(n, g)) return cudaErrorNotSupported;
  if (g.wide && getenv("QOC_B200_SMALL_WIDE") && atoi(getenv("QOC_B200_SMALL_WIDE")) == 0) {   // A/B knob: 36 narrow MMAs per product
    g.wide = 0; g.DIOFF = (g.N16 + 31) / 32 * 32;
    int cols = 32;
    while (cols < g.NSL * 2 * g.DIOFF) cols *= 2;
    g.tmem_cols = cols;
  }
  cudaError_t e = cudaFuncSetAttribute(k_tc_small_expm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) return e;
  long long groups = (q.items + g.NSL - 1) / g.NSL;
  int grid = (int)(groups < sm_count ? groups : sm_count);
  if (grid < 1) grid = 1;
  k_tc_small_expm<<<grid, 32 * g.NSL + 128 * g.WPQ, g.smem, st>>>(q, g);
  return cudaGetLastError();
}
