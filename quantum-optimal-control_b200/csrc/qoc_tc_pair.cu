// CTA-pair (tcgen05 cta_group::2) form of the propagator program for 128 < n <= 256 (QOC_F16X2; C4: n = 216).
//
// Same arithmetic, plane-set layout, op table and scratch slots as k_tc_prog (qoc_tc_f16.cu; get_matexp / matexp_op,
// core/tensorflow_state.py:25-46,70-75); what changes is who holds what:
//   * a thread-block cluster of 2 or 4 CTAs works on ONE (b,t) item (two interleaved).  CTA rank c = 2 p + r: pair p,
//     row block r.  A pair issues M = 256 MMAs (tcgen05.mma.cta_group::2): each CTA stages ITS 128 rows of A and HALF of the
//     tile's B columns, so a k-block costs 48 KB of L2 -> shared-memory traffic per SM instead of 64 KB and the ring holds
//     four stages instead of three.  With 2 CTAs the pair computes both column halves of a product one after the other
//     (two TMEM accumulator buffers); with 4 CTAs pair p computes column half p only, so 37 clusters keep 74 items in
//     flight instead of 148: their scratch matrices stay inside the 126 MB L2.
//   * the leader (r = 0) of a pair issues the MMAs; tcgen05.commit multicasts the "stage free" / "accumulator full"
//     arrivals to both CTAs; both producers signal the leader's full barrier (cp.async.bulk.tensor .cta_group::2).
//   * completion of an output quadrant is published to every CTA of the cluster (red.release.cluster on the peers'
//     monotonic counters through mapa addresses): the B operand of the next product spans rows written by the peer.
// Every wait carries the clock64 timeout of the single-CTA engine.
#include "qoc_tc_dev.cuh"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>

namespace {

constexpr int KB_ELEMS = 32;                     // K elements per stage
constexpr uint32_t A_PLANE_BYTES = 128 * 64;     // box {32 halfs, 128 rows}
constexpr uint32_t B_PLANE_BYTES = 32 * 128;     // box {64 halfs, 32 rows}: this CTA's half of the tile's columns
constexpr uint32_t STAGE_BYTES = 4 * A_PLANE_BYTES + 4 * B_PLANE_BYTES;   // 48 KB
constexpr int NSTAGE = 3;
constexpr int NEPI = 256;
constexpr int NHALF = NEPI / 128;
constexpr int CSTEP = 16 * NHALF;
constexpr int NTHREADS = 64 + NEPI;
constexpr int NEPIW = NEPI / 32;

DEVINL void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory"); }
DEVINL uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
DEVINL uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
DEVINL uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
DEVINL uint32_t ncluster_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
DEVINL uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the full barrier lives in the leader CTA of the pair: .cta_group::2 lets either CTA's copy signal it
DEVINL void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
DEVINL void mma_f16_ss2(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
DEVINL void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
DEVINL void mbar_arrive_remote(uint32_t bar_cluster_addr) {
  // relaxed: the arrival only orders TMEM reads (tcgen05.fence::before_thread_sync), no memory is published through it --
  // a release at cluster scope is a MEMBAR.ALL.GPU per warp and tile
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
DEVINL void red_release_cluster(uint32_t cluster_addr) {
  asm volatile("red.release.cluster.shared::cluster.add.u32 [%0], 1;" ::"r"(cluster_addr) : "memory");
}

// epilogue staging of one 16-column chunk of a warp's 32 rows: [4 planes][32 rows][32 bytes] in the SWIZZLE_32B pattern of
// the store tensor map (address bit 4 ^= bit 7), written thread-per-row without bank conflicts and shipped by ONE TMA store
// -- the thread-per-row st.global it replaces costs the LSU 32 passes per instruction
constexpr uint32_t STG_BYTES = 4 * 32 * 32;
DEVINL void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
DEVINL void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1),
               "r"(c2), "r"(c3)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

struct PairOp {
  long long a_idx, b_idx, e_idx, d1_idx, d2_idx;   // matrix indices; d?_cls < 0: none
  int a_cls, b_cls, d1_cls, d2_cls;
  float c1[3], c2[3];
};

// products per item: EXPM = the op table; SEG = L - 1 products of a full segment (the launcher only sends problems with T % L == 0)
DEVINL int pair_nops(const TcParams& q) { return q.prog == TC_PROG_SEG ? q.L - 1 : q.nops; }

DEVINL void make_pair_op(const TcParams& q, long long item, int j, long long sb, int rpar, PairOp& o) {
  o.a_cls = o.b_cls = TC_CLS_SCR;
  if (q.prog == TC_PROG_SEG) {
    // seg[b][sg] = P[b][t0 + L - 1] ... P[b][t0]  (re-associated chain, tensorflow_state.py:214-220): product j multiplies the
    // running value (slots 2, 3) from the left by P[t0 + j + 1]
    const long long bb = item / q.S;
    const int sg = (int)(item % q.S);
    const long long pb = bb * q.T + (long long)sg * q.L;
    o.a_cls = TC_CLS_P; o.a_idx = pb + j + 1;
    if (j == 0) { o.b_cls = TC_CLS_P; o.b_idx = pb; } else o.b_idx = sb + 2 + ((j - 1) & 1);
    o.e_idx = 0;
    o.d2_cls = -1; o.d2_idx = 0;
    if (j == q.L - 2) { o.d1_cls = TC_CLS_SEG; o.d1_idx = item; } else { o.d1_cls = TC_CLS_SCR; o.d1_idx = sb + 2 + (j & 1); }
    o.c1[0] = 1.0f / (float)(1 << TC_EU); o.c1[1] = o.c1[2] = 0.f;
    o.c2[0] = o.c2[1] = o.c2[2] = 0.f;
    return;
  }
  const int xs = rpar ? 4 : 0;                    // the generator X of odd rounds lives in slot 4
  const TcExpmOp e = q.ops[j];
  o.a_idx = sb + (e.sa == 0 ? xs : e.sa);
  o.b_idx = sb + (e.sb == 0 ? xs : e.sb);
  o.e_idx = sb + (e.se == 0 ? xs : e.se);
  o.d1_cls = o.d2_cls = -1; o.d1_idx = o.d2_idx = 0;
  if (e.d1 >= 0) {
    if (e.d1 == TC_SLOT_OUT) { o.d1_cls = TC_CLS_P; o.d1_idx = item; } else { o.d1_cls = TC_CLS_SCR; o.d1_idx = sb + e.d1; }
  }
  if (e.d2 >= 0) {
    if (e.d2 == TC_SLOT_OUT) { o.d2_cls = TC_CLS_P; o.d2_idx = item; } else { o.d2_cls = TC_CLS_SCR; o.d2_idx = sb + e.d2; }
  }
  for (int i = 0; i < 3; ++i) { o.c1[i] = e.c1[i]; o.c2[i] = e.c2[i]; }
}

__global__ void __launch_bounds__(NTHREADS, 1) k_tc_pair_expm(const TcParams q, const __grid_constant__ TcMaps maps,
                                                              const __grid_constant__ TcStoreMaps smaps) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_full[NSTAGE], bar_empty[NSTAGE], bar_tfull[2], bar_tempty[2];
  __shared__ unsigned int xdone_cnt;               // generator row slices assembled (CS per round)
  __shared__ unsigned int done_cnt[4];             // completed products per output quadrant (row block, column half)
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int dead_s;
  __shared__ float wts[32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = q.n, ld = q.ld, N16 = q.N16, KBLK = q.KBLK;
  const int NT1 = ((N16 - 128) + 31) / 32 * 32;                       // second column half, padded to a multiple of 32
  const int CS = (int)cluster_nctarank();
  const int crank = (int)cluster_ctarank();
  const int pr = crank >> 1, r = crank & 1;                           // pair, row block
  const bool leader = r == 0;
  const int nh_lo = CS == 4 ? pr : 0, nh_hi = CS == 4 ? pr + 1 : 2;    // column halves this pair computes
  const long long cid = cluster_id_x(), ncl = ncluster_x();
  const size_t plane = (size_t)n * ld, mat = 4 * plane;
  long long t_wait0 = 0, t_wait1 = 0, t_work = 0;
  const int ILV = q.ilv;
  const bool expm = q.prog == TC_PROG_EXPM;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&bar_tfull[b], 1); mbar_init(&bar_tempty[b], 2 * NEPIW); }
    for (int i = 0; i < 4; ++i) done_cnt[i] = 0;
    xdone_cnt = 0;
    dead_s = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                              // the peers' barriers and counters are initialised before anything remote lands
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = *(volatile uint32_t*)&tmem_base_s;
  volatile int* dead = &dead_s;
  const bool prof = q.prof != nullptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs of a pair)
    if (lane == 0) {
      uint32_t it = 0;
      uint32_t seen[4] = {0, 0, 0, 0};
      uint32_t base_ph = 0;
      bool ok = true;
      long long t_cat[3] = {0, 0, 0};
      int cat = 0;
      uint32_t xseen = 0, round = 0;
      auto need = [&](int i, uint32_t target) {
        uint32_t& sn = i < 4 ? seen[i] : xseen;
        if (sn >= target) return;
        const long long c0 = clock64();
        for (;;) {
          unsigned int v;
          asm volatile("ld.acquire.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(i < 4 ? &done_cnt[i] : &xdone_cnt)) : "memory");
          sn = v;
          if (v >= target) break;
          if (*dead || clock64() - c0 > TIMEOUT_CYCLES) { *dead = 1; ok = false; break; }
        }
        if (prof) { const long long dtw = clock64() - c0; t_wait0 += dtw; t_cat[cat] += dtw; }
        asm volatile("fence.proxy.async;" ::: "memory");
      };
      const uint32_t full0 = mapa(smem_u32(&bar_full[0]), (uint32_t)(crank & ~1));      // the leader's full barriers
      for (long long it0 = cid * ILV; it0 < q.items && ok; it0 += ncl * ILV, ++round) {
        const int nz = (int)min((long long)ILV, q.items - it0);
        const int nops = pair_nops(q);
        for (int j = 0; j < nops && ok; ++j)
          for (int z = 0; z < nz && ok; ++z) {
            PairOp o; make_pair_op(q, it0 + z, j, (cid * ILV + z) * TC_NSLOT, (int)(round & 1), o);
            const uint32_t tgt = base_ph + (j > 0 ? (uint32_t)((j - 1) * nz + z + 1) : 0u);
            const bool dep = j > 0;
            cat = j == 0 ? 0 : 1;
            if (expm && j == 0) need(4, (round + 1) * (uint32_t)CS);
            const int za = (int)o.a_idx, zb = (int)o.b_idx;
            for (int nh = nh_lo; nh < nh_hi && ok; ++nh) {
              const int nt = nh == 0 ? 128 : NT1;
              const int col0 = nh * 128 + r * (nt >> 1);                 // this CTA's half of the tile's B columns
              for (int kb = 0; kb < KBLK && ok; ++kb, ++it) {
                if (dep) {
                  need(r * 2 + (kb * KB_ELEMS >= 128 ? 1 : 0), tgt);      // A: own row block, columns of k-block kb
                  need(((kb * KB_ELEMS) >> 7) * 2 + nh, tgt);             // B: rows of k-block kb, columns of half nh
                }
                const int s = it % NSTAGE;
                const long long c0 = prof ? clock64() : 0;
                ok = ok && mbar_wait(&bar_empty[s], ((it / NSTAGE) & 1) ^ 1, dead);
                if (prof) t_wait1 += clock64() - c0;
                if (!ok) break;
                if (leader) mbar_expect_tx(&bar_full[s], 2 * STAGE_BYTES);
                const uint32_t sa = smem_u32(smem) + s * STAGE_BYTES, sbb = sa + 4 * A_PLANE_BYTES;
                const uint32_t fb = full0 + (uint32_t)(s * sizeof(uint64_t));
                tma_load_4d_pair(sa, &maps.a[o.a_cls], kb * KB_ELEMS, r * 128, 0, za, fb);
                tma_load_4d_pair(sbb, &maps.b[o.b_cls], col0, kb * KB_ELEMS, 0, zb, fb);
              }
            }
          }
        base_ph += (uint32_t)(nops * nz);
        cat = 2;
        for (int i = 0; i < 4; ++i) need(i, base_ph);
      }
      if (prof) { q.prof[(size_t)blockIdx.x * 8 + 0] = t_wait0; q.prof[(size_t)blockIdx.x * 8 + 1] = t_wait1;
                  q.prof[(size_t)blockIdx.x * 8 + 6] = t_cat[0]; q.prof[(size_t)blockIdx.x * 8 + 7] = t_cat[1]; }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA of the pair)
    if (leader) {
      const uint32_t idesc0 = (1u << 4) | (1u << 16) | ((256u >> 4) << 24);          // D f32, A/B f16, A K-major, B MN-major, M = 256
      const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint32_t tbase = __shfl_sync(0xffffffffu, taddr, 0);
      const uint16_t pmask = (uint16_t)(3u << (2 * pr));
      uint32_t it = 0, ti = 0;
      bool ok = true;
      for (long long it0 = cid * ILV; it0 < q.items && ok; it0 += ncl * ILV) {
        const int nz = (int)min((long long)ILV, q.items - it0);
        const int nops = pair_nops(q);
        for (int jz = 0; jz < nops * nz && ok; ++jz)
          for (int nh = nh_lo; nh < nh_hi && ok; ++nh, ++ti) {
            const int nt = nh == 0 ? 128 : NT1;
            const uint32_t idesc = idesc0 | ((uint32_t)(nt >> 3) << 17);
            const uint32_t idesc_na = idesc | (1u << 13);
            const int buf = (int)(ti & 1);
            const uint32_t use = ti >> 1;
            long long c0 = prof ? clock64() : 0;
            ok = __all_sync(0xffffffffu, mbar_wait(&bar_tempty[buf], (use & 1) ^ 1, dead));
            if (!ok) break;
            if (prof) t_wait0 += clock64() - c0;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t dr = tbase + (uint32_t)(buf * 256), di = dr + 128u;
            for (int kb = 0; kb < KBLK && ok; ++kb, ++it) {
              const int s = it % NSTAGE;
              c0 = prof ? clock64() : 0;
              ok = __all_sync(0xffffffffu, mbar_wait(&bar_full[s], (it / NSTAGE) & 1, dead));
              if (!ok) break;
              if (prof) t_wait1 += clock64() - c0;
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              if (elect_one()) {
                const uint32_t sa = sbase + s * STAGE_BYTES, sbb = sa + 4 * A_PLANE_BYTES;
#pragma unroll
                for (int k16 = 0; k16 < 2; ++k16) {
                  uint64_t da[4], db[4];
#pragma unroll
                  for (int pl = 0; pl < 4; ++pl) {
                    da[pl] = make_desc(sa + pl * A_PLANE_BYTES + k16 * 32, 1, 512 >> 4, 4);
                    db[pl] = make_desc(sbb + pl * B_PLANE_BYTES + k16 * 2048, B_PLANE_BYTES >> 4, 1024 >> 4, 2);
                  }
                  const uint32_t first = (kb == 0 && k16 == 0) ? 0u : 1u;
                  mma_f16_ss2(dr, da[0], db[1], idesc, first);             // Ar0 Br1
                  mma_f16_ss2(dr, da[1], db[0], idesc, 1u);                // Ar1 Br0
                  mma_f16_ss2(dr, da[2], db[3], idesc_na, 1u);             // -Ai0 Bi1
                  mma_f16_ss2(dr, da[3], db[2], idesc_na, 1u);             // -Ai1 Bi0
                  mma_f16_ss2(dr, da[0], db[0], idesc, 1u);                // Ar0 Br0
                  mma_f16_ss2(dr, da[2], db[2], idesc_na, 1u);             // -Ai0 Bi0
                  mma_f16_ss2(di, da[0], db[3], idesc, first);             // Ar0 Bi1
                  mma_f16_ss2(di, da[1], db[2], idesc, 1u);                // Ar1 Bi0
                  mma_f16_ss2(di, da[2], db[1], idesc, 1u);                // Ai0 Br1
                  mma_f16_ss2(di, da[3], db[0], idesc, 1u);                // Ai1 Br0
                  mma_f16_ss2(di, da[0], db[2], idesc, 1u);                // Ar0 Bi0
                  mma_f16_ss2(di, da[2], db[0], idesc, 1u);                // Ai0 Br0
                }
                umma_commit_mc(&bar_empty[s], pmask);
                if (kb == KBLK - 1) umma_commit_mc(&bar_tfull[buf], pmask);
              }
              __syncwarp();
            }
          }
      }
      if (prof && lane == 0) { q.prof[(size_t)blockIdx.x * 8 + 2] = t_wait0; q.prof[(size_t)blockIdx.x * 8 + 3] = t_wait1; }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (TMEM lane quarter = warp % 4)
    const int qd = warp & 3;
    const int half = (warp - 2) >> 2;
    const int et = (warp - 2) * 32 + lane;
    const int lrow = qd * 32 + lane;
    const int row = r * 128 + lrow;
    const bool vrow = row < n;
    const uint32_t tempty_leader0 = mapa(smem_u32(&bar_tempty[0]), (uint32_t)(crank & ~1));
    const uint32_t stg = smem_u32(smem) + NSTAGE * STAGE_BYTES + (uint32_t)(warp - 2) * STG_BYTES;
    const int row0 = r * 128 + qd * 32;              // first row of this warp's boxes
    uint32_t ti = 0, round = 0;
    bool ok = true;
    // generators of one round: X' = xscale (A_0 + sum_k u_k A_k), u_k = maxA_k sin(base[b][k][t]) (init_tf_ops_weight, :168-185);
    // every CTA of the cluster assembles a slice of 256 / CS rows and tells all of them
    const int xr0 = crank * (256 / CS), xr1 = min(n, xr0 + 256 / CS);
    const bool sparse_x = q.pat_n > 0 && q.pat_n * 3 < n * n;
    auto build_x = [&](long long bi0, int bnz, int rpar) {
      for (int z = 0; z < bnz; ++z) {
        const long long item = bi0 + z;
        const long long b = item / q.T;
        const int t = (int)(item % q.T);
        epi_bar();
        if (et <= q.K) wts[et] = et == 0 ? q.xscale : (float)(q.maxA[et - 1] * sin(q.ctrl[((size_t)b * q.K + et - 1) * q.T + t])) * q.xscale;
        epi_bar();
        __half* X = q.base[TC_CLS_SCR] + (size_t)((cid * ILV + z) * TC_NSLOT + (rpar ? 4 : 0)) * mat;
        const int l16 = ld >> 4;
        const size_t nn = (size_t)n * n;
        if (sparse_x) {                              // entries dealt round-robin to the CTAs of the cluster
          for (int e = crank * NEPI + et; e < q.pat_n; e += CS * NEPI) scatter_x_entry(q, X, plane, ld, e, wts);
        } else
        for (int i16 = et; i16 < max(0, xr1 - xr0) * l16; i16 += NEPI) {
          const int rr = xr0 + i16 / l16, c16 = (i16 % l16) * 16;
          float re[16], im[16];
#pragma unroll
          for (int cc = 0; cc < 16; ++cc) re[cc] = im[cc] = 0.f;
          const bool full = c16 + 16 <= n && (n & 1) == 0;
          for (int k0 = 0; k0 <= q.K; k0 += 2) {
            float4 v[2][8];
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const int k = k0 + kk;
              const float2* src = q.A_f + (size_t)k * nn + (size_t)rr * n + c16;
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
          store_planes16(X, plane, ld, rr, c16, re, im);
        }
      }
      asm volatile("fence.proxy.async;" ::: "memory");
      epi_bar();
      if (et < CS) red_release_cluster(mapa(smem_u32(&xdone_cnt), (uint32_t)et));
    };
    for (long long it0 = cid * ILV; it0 < q.items && ok; it0 += ncl * ILV, ++round) {
      const int nz = (int)min((long long)ILV, q.items - it0);
      const int nops = pair_nops(q);
      if (expm && round == 0) build_x(it0, nz, 0);
      for (int j = 0; j < nops && ok; ++j)
        for (int z = 0; z < nz && ok; ++z) {
          if (expm && z == 0 && j == (nops > 2 ? 2 : nops - 1)) {
            const long long nx0 = it0 + ncl * ILV;
            if (nx0 < q.items) build_x(nx0, (int)min((long long)ILV, q.items - nx0), (int)((round + 1) & 1));
          }
          const long long item = it0 + z;
          PairOp o; make_pair_op(q, item, j, (cid * ILV + z) * TC_NSLOT, (int)(round & 1), o);
          const __half* E = q.base[TC_CLS_SCR] + (size_t)o.e_idx * mat;
          const bool useE = o.c1[1] != 0.f || o.c2[1] != 0.f;
          for (int nh = nh_lo; nh < nh_hi && ok; ++nh, ++ti) {
            const int nt = min(nh == 0 ? 128 : NT1, ld - nh * 128), cbase = nh * 128;   // columns that exist in storage
            const int buf = (int)(ti & 1);
            const uint32_t use = ti >> 1;
            const uint32_t lane_addr = taddr + ((uint32_t)(qd * 32) << 16) + (uint32_t)(buf * 256);
            uint32_t e0[4][8];
            const bool ldE = useE && vrow;
            auto fetchE = [&](int c0) {
              if (ldE && c0 < nt && !(q.tma_store & 4)) {
                const __half* p0 = E + (size_t)row * ld + cbase + c0;
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) ldg256(p0 + pl * plane, e0[pl]);
              }
            };
            // the generators read as elementwise source may have been assembled by another CTA: the producer's acquire of
            // xdone_cnt precedes the loads whose MMAs precede bar_tfull; the data is read from L2 (ld.global.cg)
            long long c0t = prof ? clock64() : 0;
            if (j > 0) fetchE(16 * half);
            ok = __all_sync(0xffffffffu, mbar_wait(&bar_tfull[buf], use & 1, dead));
            if (!ok) break;
            if (prof) { const long long c1t = clock64(); t_wait0 += c1t - c0t; c0t = c1t; }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (j == 0) fetchE(16 * half);
            // 16-column chunks, software-pipelined: the accumulator rows of the NEXT chunk are requested as soon as the current
            // ones have been turned into fp16 pairs, so the TMEM read and the elementwise-source fetch fly during the staging,
            // the proxy fence and the TMA store of the current chunk
            uint32_t ur[16], ui[16];
            if (16 * half < nt) {
              tmem_ld16(lane_addr + (uint32_t)(16 * half), ur);
              tmem_ld16(lane_addr + (uint32_t)(128 + 16 * half), ui);
            }
            for (int c0 = 16 * half; c0 < nt; c0 += CSTEP) {
              const int col = cbase + c0;
              const bool last = c0 + CSTEP >= nt;
              const bool diag = col + 15 >= row0 && col <= row0 + 31;      // warp-uniform: the chunk meets the diagonal
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
              if (last) {                              // the accumulator buffer may be refilled
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(tempty_leader0 + (uint32_t)(buf * sizeof(uint64_t)));
              }
              for (int dd = 0; dd < 2; ++dd) {
                const int dcls = dd == 0 ? o.d1_cls : o.d2_cls;
                if (dcls < 0) continue;
                const bool final_dst = dd == 1 || o.d2_cls < 0;            // ur / ui / e0 are dead after this destination
                const float k0 = dd == 0 ? o.c1[0] : o.c2[0], k1 = dd == 0 ? o.c1[1] : o.c2[1], k2 = dd == 0 ? o.c1[2] : o.c2[2];
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous box has left the staging
                __syncwarp();
                const uint32_t sw = (uint32_t)((lane >> 2) & 1) << 4;
#pragma unroll
                for (int comp = 0; comp < 2; ++comp) {       // one component at a time: half the live registers
                  float ov[16];
                  if (useE) {
                    float ee[16];
                    unpack16(e0[2 * comp], e0[2 * comp + 1], ee);
#pragma unroll
                    for (int i = 0; i < 16; ++i) ov[i] = fmaf(k0, __uint_as_float(comp == 0 ? ur[i] : ui[i]), k1 * ee[i]);
                  } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) ov[i] = k0 * __uint_as_float(comp == 0 ? ur[i] : ui[i]);
                  }
                  if (comp == 0 && diag && k2 != 0.f) {
                    const int dc = row - col;
#pragma unroll
                    for (int i = 0; i < 16; ++i) if (i == dc) ov[i] += k2;
                  }
                  uint32_t p0[8], p1[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) split2(ov[2 * i], ov[2 * i + 1], p0[i], p1[i]);
                  const uint32_t b0 = stg + (uint32_t)(2 * comp) * 1024u + (uint32_t)lane * 32u, b1 = b0 + 1024u;
                  sts128(b0 + sw, p0[0], p0[1], p0[2], p0[3]); sts128(b0 + (16u ^ sw), p0[4], p0[5], p0[6], p0[7]);
                  sts128(b1 + sw, p1[0], p1[1], p1[2], p1[3]); sts128(b1 + (16u ^ sw), p1[4], p1[5], p1[6], p1[7]);
                }
                if (final_dst && !last) {
                  if (useE) fetchE(c0 + CSTEP);
                  tmem_ld16(lane_addr + (uint32_t)(c0 + CSTEP), ur);
                  tmem_ld16(lane_addr + (uint32_t)(128 + c0 + CSTEP), ui);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0 && row0 < n && !(q.tma_store & 2)) tma_store_4d(&smaps.st[dcls], stg, col, row0, 0, (int)(dd == 0 ? o.d1_idx : o.d2_idx));
              }
            }
            if (lane == 0 && !(q.tma_store & 8)) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this tile's boxes are written, not just read
            __syncwarp();
            if (16 * half >= nt) {                     // a warp without any chunk in this tile still owes its arrival
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              __syncwarp();
              if (lane == 0) mbar_arrive_remote(tempty_leader0 + (uint32_t)(buf * sizeof(uint64_t)));
            }
            epi_bar();                                 // quadrant (r, nh) of the product's outputs is complete in this CTA:
            if (et < CS)                               // one release-increment of its counter in every CTA of the cluster
              red_release_cluster(mapa(smem_u32(&done_cnt[r * 2 + nh]), (uint32_t)et));
            if (prof) t_work += clock64() - c0t;
          }
          if (CS == 4) {
            // the other pair's quadrants of this product: nothing to do here, their CTAs bump our counters
          }
        }
    }
    if (prof && et == 0) { q.prof[(size_t)blockIdx.x * 8 + 4] = t_wait0; q.prof[(size_t)blockIdx.x * 8 + 5] = t_work; }
  }
  // ---------------------------------------------------------------------- teardown
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                              // no remote arrival may target a CTA that has exited
  if (dead_s && tid == 0 && q.err_flag) atomicExch(q.err_flag, 1);
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(512u) : "memory");
}

}  // namespace

bool tc_pair_supported(int n) { return n > 128 && n <= TC_MAX_N; }

int tc_pair_max_clusters(int cs) {
  static int cached[5] = {0, 0, 0, 0, 0};
  if (cs != 2 && cs != 4) return 0;
  if (cached[cs]) return cached[cs];
  const size_t smem = (size_t)NSTAGE * STAGE_BYTES + (size_t)NEPIW * STG_BYTES + 1024;
  if (cudaFuncSetAttribute(k_tc_pair_expm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(sms / cs * cs)); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int ncl = 0;
  if (cudaOccupancyMaxActiveClusters(&ncl, k_tc_pair_expm, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (ncl > sms / cs) ncl = sms / cs;
  cached[cs] = ncl;
  return ncl;
}

// q: an EXPM program (tc_launch's fields); cs = CTAs per cluster (2 or 4); scratch must hold (clusters * ilv * TC_NSLOT) matrices
cudaError_t tc_pair_launch_expm(const TcParams& q_in, const TcMaps& maps, const TcStoreMaps& smaps, const TcGeom& g, int cs, cudaStream_t st) {
  if (!tc_pair_supported(g.n) || (q_in.prog != TC_PROG_EXPM && q_in.prog != TC_PROG_SEG)) return cudaErrorInvalidValue;
  if (q_in.prog == TC_PROG_SEG && (q_in.L < 2 || q_in.T % q_in.L != 0)) return cudaErrorInvalidValue;
  TcParams q = q_in;
  q.n = g.n; q.ld = g.ld; q.N16 = g.N16; q.KBLK = g.KBLK;
  if (q.ilv < 1) q.ilv = 1;
  int ncl = tc_pair_max_clusters(cs);
  if (ncl < 1) return cudaErrorInvalidConfiguration;
  if (getenv("QOC_B200_PAIR_CLUSTERS")) ncl = std::max(1, std::min(ncl, atoi(getenv("QOC_B200_PAIR_CLUSTERS"))));   // A/B knob: fewer items in flight
  const long long rounds_items = (q.items + q.ilv - 1) / q.ilv;
  if (ncl > rounds_items) ncl = (int)rounds_items;
  const size_t smem = (size_t)NSTAGE * STAGE_BYTES + (size_t)NEPIW * STG_BYTES + 1024;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(ncl * cs)); cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k_tc_pair_expm, q, maps, smaps);
}
