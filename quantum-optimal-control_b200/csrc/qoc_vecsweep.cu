// State / costate sweeps for problems with FEW concerned states (m < NP/2), fp64, n <= 64.
//
//   k_vec_sweep<false>: psi_j(t+1) = P_t psi_j(t),  t = 0..T-1      (init_tf_inter_vectors,
//                       core/tensorflow_state.py:229-242, restricted to the m columns X_t V_j)
//   k_vec_sweep<true> : lambda_j(t) = P_t^dagger lambda_j(t+1) + source_j(t),  t = T-1..1
//                       (the reverse sweep TF autodiff performs through :214-220)
//
// Both are T dependent n x n x m matrix-vector products per instance: 16 n^2 bytes of P_t against
// 8 n^2 m flops per step, i.e. HBM-bound once the per-step latency is out of the way -- and that
// latency is instruction issue (a lone warp runs at ~0.3 IPC), so a step is spread over WK x CP
// warps and stripped to the bone:
//   * P_t arrives by ONE bulk-copy instruction (TMA, cp.async.bulk 1-D, 16 n^2 contiguous bytes)
//     into an NST-deep ring, completion on an mbarrier per stage -- no per-thread copy loop;
//   * warp (wk, cp) owns the summation slice k = wk (mod WK) and the column pairs cp, cp+CP, ...;
//     lane l owns rows l (and l+32); every shared-memory offset of the k loop is loop-invariant;
//   * barrier | partial[wk][j][l] = sum_{k in slice} P(k,l) v_j[k] | barrier | threads (j,l): sum the
//     WK partials, add the regulariser source, write v_j(next) to shared memory and psi/lambda to HBM.
// The ring keeps the row-major layout of HBM: the reverse sweep (P^dagger) reads rows (lanes
// contiguous, conflict-free); the forward sweep reads columns (lane stride n: 2-way bank conflicts
// for n = 30, irrelevant at 8 loads per step).  The four real partial sums per complex product live
// in separate accumulators (independent DFMA chains, 8.4-cycle latency: tools/dmma_probe.cu).
#include "qoc_internal.cuh"
#include <math.h>
#include <stdlib.h>

#define DEVINL __device__ __forceinline__

namespace {

DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEVINL void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
// 1-D bulk copy global -> shared (TMA), completion counted in bytes on `bar`
DEVINL void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

DEVINL cplx cmul(const cplx a, const cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

constexpr int WK = 4;                            // summation slices (warps along k)

struct SweepShape {
  int cp;                                        // warps along the column pairs
  int npairs;                                    // ceil(m / 2)
  size_t smem;
};

template <int NR, int NST, int CPW = 2>
SweepShape sweep_shape(int n, int m, bool pf32 = false) {
  SweepShape s;
  s.npairs = (m + CPW - 1) / CPW;                // column groups (CPW columns per warp)
  s.cp = s.npairs < 4 ? s.npairs : 4;
  const int mp = CPW * s.npairs;                 // columns rounded up to whole groups
  // ring (+ slack for the padded lanes of the last row) + vectors [2][mp][VL] + partials [WK][mp][VL] + mbarriers
  s.smem = ((size_t)NST * (pf32 ? 512 : n * n) + 64 + (size_t)(2 + WK) * mp * NR * 32) * sizeof(cplx) + 8 * NST;
  return s;
}

// PF32: the propagators are the tcgen05 path's fp32 planar padded [2][32][32] tiles (csrc/qoc_tc_tf32.cu), 8 KB per
// step; they are widened to double on the fly.  The forward sweep reads tile COLUMNS (lane stride 32 floats = one bank),
// so there every lane walks its warp's k-slice in a rotated order, k = 8 wk + ((i + lane) & 7): 4-way instead of 32-way
// bank conflicts, no transposition.
template <bool REV, int NR, int NST, int CPW, bool PF32>
__global__ void __launch_bounds__(512) k_vec_sweep(QocParams p, int CP) {
  static_assert(!PF32 || NR == 1, "fp32 tiles are 32 x 32");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int VL = NR * 32;
  constexpr int KI = VL / WK;                    // k iterations per warp (upper bound, predicated on k < n)
  const int n = p.n, m = p.m, T = p.T, nn = n * n, mn = m * n;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int wk = warp % WK, cp0 = warp / WK;
  const int npairs = (m + CPW - 1) / CPW, mp = CPW * npairs;
  const int b = blockIdx.x;
  const int stage_elems = PF32 ? 512 : nn;                              // in cplx units (PF32: 2048 floats)
  cplx* ring = reinterpret_cast<cplx*>(smem_raw);                       // [NST][n][n] row-major, as in HBM
  cplx* vecs = ring + (size_t)NST * stage_elems + 64;                   // [2][mp][VL]
  cplx* part = vecs + (size_t)2 * mp * VL;                              // [WK][mp][VL]
  uint64_t* full = reinterpret_cast<uint64_t*>(part + (size_t)WK * mp * VL);   // [NST]
  const cplx* Pg = reinterpret_cast<const cplx*>(p.P) + (size_t)b * T * stage_elems;
  cplx* psi_b = p.psi + (size_t)b * (T + 1) * mn;
  cplx* lam_b = p.lam + (size_t)b * (T + 1) * mn;
  const int nsteps = REV ? T - 1 : T;
  const uint32_t stage_bytes = (uint32_t)stage_elems * sizeof(cplx);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // P_step -> ring stage (step index in sweep order); issued by one thread
  auto prefetch = [&](int step) {
    if (step < nsteps) {
      const int t = REV ? T - 1 - step : step;
      uint64_t* bar = &full[step % NST];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of the stage (ordered by the barrier) before the async write
      mbar_expect_tx(bar, stage_bytes);
      bulk_g2s(ring + (size_t)(step % NST) * stage_elems, Pg + (size_t)t * stage_elems, stage_bytes, bar);
    }
  };
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NST - 1; ++i) prefetch(i);
  }

  // regulariser sources of the costate recursion (core/regularization_functions.py:71-95)
  const bool forb = REV && p.reg.has_forbidden && p.fw != nullptr;
  const bool spd = REV && p.reg.has_speed_up != 0;
  const bool has_src = forb || spd;
  const double* sc = p.scal + (size_t)b * 8;
  const double spdfac = REV ? sc[4] : 0.0;
  auto source = [&](int t, int j, int r) -> cplx {
    cplx s = make_double2(0.0, 0.0);
    if (forb && p.dressW) {
      s = p.psid[((size_t)b * (T + 1) + t) * mn + (size_t)j * n + r];          // precomputed by k_dress phase 1
    } else if (forb) {
      const cplx x = psi_b[(size_t)t * mn + (size_t)j * n + r];
      const double c = p.fw[r] / (double)T * 2.0 * (x.x * x.x + x.y * x.y);
      s.x = c * x.x; s.y = c * x.y;
    }
    if (spd) {
      const cplx q = cmul(p.ot[(size_t)b * (T + 1) + t], p.phi[(size_t)j * n + r]);
      s.x += spdfac * q.x; s.y += spdfac * q.y;
    }
    return s;
  };

  // initial vectors: element e = (j, r) of the padded [mp][VL] block is owned by thread e (mod nt)
  for (int e = tid; e < mp * VL; e += nt) {
    const int j = e / VL, r = e - j * VL;
    cplx v = make_double2(0.0, 0.0);
    if (j < m && r < n) {
      if (REV) {                                 // lambda(T) = -(2/m^2) o phi + source(T)
        const double f = -2.0 / ((double)m * (double)m);
        v = cmul(make_double2(sc[0] * f, sc[1] * f), p.phi[(size_t)j * n + r]);
        const cplx s = source(T, j, r);
        v.x += s.x; v.y += s.y;
        lam_b[(size_t)T * mn + (size_t)j * n + r] = v;
      } else {                                   // psi(0) = V is stored as is (:233-234); the chain starts from U0 V
        for (int c = 0; c < n; ++c) {
          const cplx u = p.U0[(size_t)r * n + c], a = p.V[(size_t)j * n + c];
          v.x += u.x * a.x - u.y * a.y; v.y += u.x * a.y + u.y * a.x;
        }
        psi_b[(size_t)j * n + r] = p.V[(size_t)j * n + r];
      }
    }
    vecs[e] = v;
    vecs[(size_t)mp * VL + e] = make_double2(0.0, 0.0);
  }

  // loop-invariant element offsets of the k loop: reverse P[k][l] -> k*n + l, forward P[l][k] -> l*n + k;
  // rows / columns beyond n are clamped (their products are discarded or multiplied by v = 0)
  int moff[KI][NR], kidx[KI];
#pragma unroll
  for (int i = 0; i < KI; ++i) {
    if (PF32) {                                  // contiguous k-slices; forward: rotated per lane (see above)
      const int k = 8 * wk + (REV ? i : ((i + lane) & 7));
      kidx[i] = k;
      moff[i][0] = REV ? k * 32 + lane : lane * 32 + k;                 // float index into the Re plane
    } else {
      kidx[i] = wk + WK * i;
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) {
        const int k = min(wk + WK * i, n - 1), l = min(lane + 32 * rr, n - 1);
        moff[i][rr] = REV ? k * n + l : l * n + k;
      }
    }
  }
  // the element this thread finalises
  const int fj = tid / VL, fr = tid - fj * VL;
  const bool fin = tid < mp * VL, fval = fin && fj < m && fr < n;
  const size_t fout = (size_t)fj * n + fr;

  int cur = 0;
  for (int step = 0; step < nsteps; ++step) {
    __syncthreads();                             // v(cur) complete; stage step-1 and the partials are free
    if (tid == 0) prefetch(step + NST - 1);
    const int t = REV ? T - 1 - step : step;     // propagator index; result is lambda(t) / psi(t+1)
    cplx src0 = make_double2(0.0, 0.0);
    if (has_src && fval) src0 = source(t, fj, fr);       // issued early, consumed after the second barrier
    {
      uint64_t* bar = &full[step % NST];
      const uint32_t parity = (uint32_t)(step / NST) & 1u;
      while (!mbar_try_wait(bar, parity)) {}
    }
    const cplx* M = ring + (size_t)(step % NST) * stage_elems;
    const float* Mf = reinterpret_cast<const float*>(M);
    for (int pr = cp0; pr < npairs; pr += CP) {
      const cplx* va = vecs + ((size_t)cur * mp + CPW * pr) * VL;
      const cplx* vb = va + (CPW - 1) * VL;
      double a0[NR][4], a1[NR][4];               // partial sums  xx, yy, xy, yx
#pragma unroll
      for (int rr = 0; rr < NR; ++rr)
#pragma unroll
        for (int e = 0; e < 4; ++e) a0[rr][e] = a1[rr][e] = 0.0;
#pragma unroll
      for (int i = 0; i < KI; ++i) {
        if (PF32 || wk + WK * i < n) {           // warp-uniform (fp32 tiles are zero-padded: no predicate)
          const cplx x = va[kidx[i]], y = CPW == 2 ? vb[kidx[i]] : x;
#pragma unroll
          for (int rr = 0; rr < NR; ++rr) {
            cplx e;
            if (PF32) e = make_double2((double)Mf[moff[i][0]], (double)Mf[1024 + moff[i][0]]);
            else e = M[moff[i][rr]];
            a0[rr][0] = fma(e.x, x.x, a0[rr][0]); a0[rr][1] = fma(e.y, x.y, a0[rr][1]);
            a0[rr][2] = fma(e.x, x.y, a0[rr][2]); a0[rr][3] = fma(e.y, x.x, a0[rr][3]);
            if (CPW == 2) {
              a1[rr][0] = fma(e.x, y.x, a1[rr][0]); a1[rr][1] = fma(e.y, y.y, a1[rr][1]);
              a1[rr][2] = fma(e.x, y.y, a1[rr][2]); a1[rr][3] = fma(e.y, y.x, a1[rr][3]);
            }
          }
        }
      }
      cplx* pa = part + ((size_t)wk * mp + CPW * pr) * VL + lane;
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) {
        if (REV) {                               // conj(e) * v
          pa[32 * rr] = make_double2(a0[rr][0] + a0[rr][1], a0[rr][2] - a0[rr][3]);
          if (CPW == 2) pa[VL + 32 * rr] = make_double2(a1[rr][0] + a1[rr][1], a1[rr][2] - a1[rr][3]);
        } else {
          pa[32 * rr] = make_double2(a0[rr][0] - a0[rr][1], a0[rr][2] + a0[rr][3]);
          if (CPW == 2) pa[VL + 32 * rr] = make_double2(a1[rr][0] - a1[rr][1], a1[rr][2] + a1[rr][3]);
        }
      }
    }
    __syncthreads();
    const int nxt = cur ^ 1;
    cplx* out = REV ? lam_b + (size_t)t * mn : psi_b + (size_t)(t + 1) * mn;
    if (fval) {
      cplx v = src0;
#pragma unroll
      for (int w = 0; w < WK; ++w) { const cplx q = part[(size_t)w * mp * VL + tid]; v.x += q.x; v.y += q.y; }
      out[fout] = v;
      vecs[(size_t)nxt * mp * VL + tid] = v;
    }
    for (int e = tid + nt; e < mp * VL; e += nt) {           // more than nt elements (many columns, n > 32)
      const int j = e / VL, r = e - j * VL;
      if (j < m && r < n) {
        cplx v = has_src ? source(t, j, r) : make_double2(0.0, 0.0);
#pragma unroll
        for (int w = 0; w < WK; ++w) { const cplx q = part[(size_t)w * mp * VL + e]; v.x += q.x; v.y += q.y; }
        out[(size_t)j * n + r] = v;
        vecs[(size_t)nxt * mp * VL + e] = v;
      }
    }
    cur = nxt;
  }
}

template <bool REV, int NR, int NST, int CPW, bool PF32>
cudaError_t launch_cpw(const QocParams& p, cudaStream_t st) {
  const SweepShape s = sweep_shape<NR, NST, CPW>(p.n, p.m, PF32);
  cudaError_t e = cudaFuncSetAttribute(k_vec_sweep<REV, NR, NST, CPW, PF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem);
  if (e != cudaSuccess) return e;
  // same (maximal) shared-memory carve-out as the kernels it runs beside: CTAs of kernels that ask
  // for different carve-outs cannot share an SM
  e = cudaFuncSetAttribute(k_vec_sweep<REV, NR, NST, CPW, PF32>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  k_vec_sweep<REV, NR, NST, CPW, PF32><<<p.B, 32 * WK * s.cp, s.smem, st>>>(p, s.cp);
  return cudaGetLastError();
}

// two columns per warp; QOC_B200_SWEEP_CPW=1 gives every column its own warp (half the dependent DFMAs per warp
// and step: the sweeps alone get ~5 % faster, the step does not -- the U_final branch becomes the longer one)
template <bool REV, int NR, int NST>
cudaError_t launch(const QocParams& p, cudaStream_t st) {
  static const int force = getenv("QOC_B200_SWEEP_CPW") ? atoi(getenv("QOC_B200_SWEEP_CPW")) : 0;
  const bool one = force == 1;
  return one ? launch_cpw<REV, NR, NST, 1, false>(p, st) : launch_cpw<REV, NR, NST, 2, false>(p, st);
}

}  // namespace

bool qoc_vec_sweep_supported(const QocParams& p) {
  if (p.n > 64 || p.m > 32) return false;
  const size_t smem = p.n > 32 ? sweep_shape<2, 3>(p.n, p.m).smem : sweep_shape<1, 3>(p.n, p.m).smem;
  return smem <= 200 * 1024;
}

// ring depth: 4 stages when two CTAs of that size still leave an SM room for the U_final branch, else 3
cudaError_t qoc_launch_vec_sweep(const QocParams& p, int reverse, int p_is_f32, cudaStream_t st, int64_t* launches) {
  if (!qoc_vec_sweep_supported(p) || (p_is_f32 && p.n > 32)) return cudaErrorNotSupported;
  ++*launches;
  if (p_is_f32) return reverse ? launch_cpw<true, 1, 4, 2, true>(p, st) : launch_cpw<false, 1, 4, 2, true>(p, st);
  static const int force = getenv("QOC_B200_SWEEP_STAGES") ? atoi(getenv("QOC_B200_SWEEP_STAGES")) : 0;
  if (p.n > 32) {
    const bool deep = force ? force >= 4 : sweep_shape<2, 4>(p.n, p.m).smem <= 84 * 1024;
    if (deep) return reverse ? launch<true, 2, 4>(p, st) : launch<false, 2, 4>(p, st);
    return reverse ? launch<true, 2, 3>(p, st) : launch<false, 2, 3>(p, st);
  }
  const bool deep = force ? force >= 4 : sweep_shape<1, 4>(p.n, p.m).smem <= 84 * 1024;
  if (deep) return reverse ? launch<true, 1, 4>(p, st) : launch<false, 1, 4>(p, st);
  return reverse ? launch<true, 1, 3>(p, st) : launch<false, 1, 3>(p, st);
}
