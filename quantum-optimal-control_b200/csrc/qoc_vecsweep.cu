// State / costate sweeps for problems with FEW concerned states (m < NP/2), fp64, n <= 64.
//
//   k_vec_sweep<false>: psi_j(t+1) = P_t psi_j(t),  t = 0..T-1      (init_tf_inter_vectors,
//                       core/tensorflow_state.py:229-242, restricted to the m columns X_t V_j)
//   k_vec_sweep<true> : lambda_j(t) = P_t^dagger lambda_j(t+1) + source_j(t),  t = T-1..1
//                       (the reverse sweep TF autodiff performs through :214-220)
//
// Both are T dependent n x n x m matrix-vector products per instance: 16 n^2 bytes of P_t against
// 8 n^2 m flops per step, i.e. HBM-bound once the per-step latency is out of the way -- and a lone
// warp issues at ~0.3 IPC, so the step is spread over many warps.  One CTA per instance, WK x CP
// warps: warp (wk, cp) owns the summation slice k = wk (mod WK) and the column pairs cp, cp+CP, ...;
// lane l owns rows l and l+32 of the result.  A step is
//   barrier | partial[wk][j][l] = sum_{k in slice} ring[k][l] * v_j[k] | barrier | sum the WK
//   partials, add the regulariser source, write v_j(next) to shared memory and psi / lambda to HBM
// P_t is streamed by cp.async into an NST-deep ring, laid out so that the 32 lanes of a warp read 32
// consecutive 16-byte elements (conflict-free):
//   forward : ring[k][l] = P_t[l][k]   (transposing scatter; 16-byte granularity makes it free)
//   reverse : ring[k][l] = P_t[k][l]   (straight copy), used conjugated
// The four real partial sums per complex product live in separate accumulators (independent DFMA
// chains, 8.4-cycle latency at 2.3 cycles per warp instruction: tools/dmma_probe.cu).
#include "qoc_internal.cuh"
#include <math.h>
#include <stdlib.h>

#define DEVINL __device__ __forceinline__

namespace {

DEVINL void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

DEVINL cplx cmul(const cplx a, const cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

constexpr int WK = 4;                            // summation slices (warps along k)

struct SweepShape {
  int cp;                                        // warps along the column pairs
  int npairs;                                    // ceil(m / 2)
  size_t smem;
};

template <int NR, int NST>
SweepShape sweep_shape(int n, int m) {
  SweepShape s;
  s.npairs = (m + 1) / 2;
  s.cp = s.npairs < 4 ? s.npairs : 4;
  const int mp = 2 * s.npairs;                   // columns rounded up to pairs
  // ring + vectors [2][mp][VL] + partials [WK][mp][VL]
  s.smem = ((size_t)NST * n * (n | 1) + 32 + (size_t)(2 + WK) * mp * NR * 32) * sizeof(cplx);
  return s;
}

template <bool REV, int NR, int NST>
__global__ void __launch_bounds__(512) k_vec_sweep(QocParams p, int CP) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int VL = NR * 32;
  const int n = p.n, m = p.m, T = p.T, nn = n * n, mn = m * n;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int wk = warp % WK, cp0 = warp / WK;
  const int npairs = (m + 1) >> 1, mp = 2 * npairs;
  const int b = blockIdx.x;
  // ring stage: n rows of LD elements; LD odd so that the transposing scatter of the forward sweep (lanes
  // along k, stride LD) and the row reads (lanes along l) are both bank-conflict free
  const int LD = n | 1, stage = n * LD;
  cplx* ring = reinterpret_cast<cplx*>(smem_raw);                       // [NST][n][LD] (+32 slack for the padded lanes)
  cplx* vecs = ring + (size_t)NST * stage + 32;                         // [2][mp][VL]
  cplx* part = vecs + (size_t)2 * mp * VL;                              // [WK][mp][VL]
  const cplx* Pg = reinterpret_cast<const cplx*>(p.P) + (size_t)b * T * nn;
  cplx* psi_b = p.psi + (size_t)b * (T + 1) * mn;
  cplx* lam_b = p.lam + (size_t)b * (T + 1) * mn;
  const int nsteps = REV ? T - 1 : T;

  // P_step -> ring stage (step index in sweep order); fwd: transposing scatter
  const int dr = nt / n, dc = nt - dr * n, r_first = tid / n, c_first = tid - r_first * n;
  auto prefetch = [&](int step) {
    if (step < nsteps) {
      const int t = REV ? T - 1 - step : step;
      const cplx* src = Pg + (size_t)t * nn + tid;
      cplx* dst = ring + (size_t)(step % NST) * stage;
      int r = r_first, c = c_first;
      if (REV) {
        for (int idx = tid; idx < nn; idx += nt, src += nt) {
          cp_async16(dst + r * LD + c, src);
          r += dr; c += dc;
          if (c >= n) { c -= n; ++r; }
        }
      } else {
        for (int idx = tid; idx < nn; idx += nt, src += nt) {
          cp_async16(dst + c * LD + r, src);
          r += dr; c += dc;
          if (c >= n) { c -= n; ++r; }
        }
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int i = 0; i < NST - 1; ++i) prefetch(i);

  // regulariser sources of the costate recursion (core/regularization_functions.py:71-95)
  const bool forb = REV && p.reg.has_forbidden && p.fw != nullptr;
  const bool spd = REV && p.reg.has_speed_up != 0;
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
  }

  int cur = 0;
  for (int step = 0; step < nsteps; ++step) {
    cp_async_wait<NST - 2>();
    __syncthreads();                             // stage `step` landed; v(cur) complete; stage step-1 and partials free
    prefetch(step + NST - 1);
    const int t = REV ? T - 1 - step : step;     // propagator index; result is lambda(t) / psi(t+1)
    // the source of the element this thread finalises (first pass of the loop below), issued early
    cplx src0 = make_double2(0.0, 0.0);
    {
      const int j = tid / VL, r = tid - j * VL;
      if (REV && tid < mp * VL && j < m && r < n) src0 = source(t, j, r);
    }
    const cplx* M = ring + (size_t)(step % NST) * stage + lane;
    for (int pr = cp0; pr < npairs; pr += CP) {
      const cplx* va = vecs + ((size_t)cur * mp + 2 * pr) * VL;
      const cplx* vb = va + VL;
      double a0[NR][4], a1[NR][4];               // partial sums  xx, yy, xy, yx
#pragma unroll
      for (int rr = 0; rr < NR; ++rr)
#pragma unroll
        for (int e = 0; e < 4; ++e) a0[rr][e] = a1[rr][e] = 0.0;
      const cplx* Mk = M + (size_t)wk * LD;
#pragma unroll 4
      for (int k = wk; k < n; k += WK, Mk += WK * LD) {
        const cplx x = va[k], y = vb[k];
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) {
          cplx e = make_double2(0.0, 0.0);
          if (NR == 1 || lane + 32 * rr < n) e = Mk[32 * rr];
          a0[rr][0] = fma(e.x, x.x, a0[rr][0]); a0[rr][1] = fma(e.y, x.y, a0[rr][1]);
          a0[rr][2] = fma(e.x, x.y, a0[rr][2]); a0[rr][3] = fma(e.y, x.x, a0[rr][3]);
          a1[rr][0] = fma(e.x, y.x, a1[rr][0]); a1[rr][1] = fma(e.y, y.y, a1[rr][1]);
          a1[rr][2] = fma(e.x, y.y, a1[rr][2]); a1[rr][3] = fma(e.y, y.x, a1[rr][3]);
        }
      }
      cplx* pa = part + ((size_t)wk * mp + 2 * pr) * VL + lane;
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) {
        if (REV) {                               // conj(e) * v
          pa[32 * rr] = make_double2(a0[rr][0] + a0[rr][1], a0[rr][2] - a0[rr][3]);
          pa[VL + 32 * rr] = make_double2(a1[rr][0] + a1[rr][1], a1[rr][2] - a1[rr][3]);
        } else {
          pa[32 * rr] = make_double2(a0[rr][0] - a0[rr][1], a0[rr][2] + a0[rr][3]);
          pa[VL + 32 * rr] = make_double2(a1[rr][0] - a1[rr][1], a1[rr][2] + a1[rr][3]);
        }
      }
    }
    __syncthreads();
    const int nxt = cur ^ 1;
    cplx* out = REV ? lam_b + (size_t)t * mn : psi_b + (size_t)(t + 1) * mn;
    for (int e = tid; e < mp * VL; e += nt) {
      const int j = e / VL, r = e - j * VL;
      cplx v = make_double2(0.0, 0.0);
      if (j < m && r < n) {
        v = (REV && e != tid) ? source(t, j, r) : src0;
#pragma unroll
        for (int w = 0; w < WK; ++w) { const cplx q = part[(size_t)w * mp * VL + e]; v.x += q.x; v.y += q.y; }
        out[(size_t)j * n + r] = v;
      }
      vecs[(size_t)nxt * mp * VL + e] = v;
    }
    cur = nxt;
  }
  cp_async_wait<0>();
}

template <bool REV, int NR, int NST>
cudaError_t launch(const QocParams& p, cudaStream_t st) {
  const SweepShape s = sweep_shape<NR, NST>(p.n, p.m);
  cudaError_t e = cudaFuncSetAttribute(k_vec_sweep<REV, NR, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem);
  if (e != cudaSuccess) return e;
  // same (maximal) shared-memory carve-out as the kernels it runs beside: CTAs of kernels that ask
  // for different carve-outs cannot share an SM
  e = cudaFuncSetAttribute(k_vec_sweep<REV, NR, NST>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  k_vec_sweep<REV, NR, NST><<<p.B, 32 * WK * s.cp, s.smem, st>>>(p, s.cp);
  return cudaGetLastError();
}

}  // namespace

bool qoc_vec_sweep_supported(const QocParams& p) {
  if (p.n > 64 || p.m > 32) return false;
  const size_t smem = p.n > 32 ? sweep_shape<2, 3>(p.n, p.m).smem : sweep_shape<1, 3>(p.n, p.m).smem;
  return smem <= 200 * 1024;
}

// ring depth: the sweeps are limited by the bytes they keep in flight (HBM latency under load is ~2 us),
// so 5 stages when two CTAs of that size still leave an SM room for the U_final branch, else 3
cudaError_t qoc_launch_vec_sweep(const QocParams& p, int reverse, cudaStream_t st, int64_t* launches) {
  if (!qoc_vec_sweep_supported(p)) return cudaErrorNotSupported;
  ++*launches;
  static const int force = getenv("QOC_B200_SWEEP_STAGES") ? atoi(getenv("QOC_B200_SWEEP_STAGES")) : 0;
  if (p.n > 32) {
    const bool deep = force ? force >= 5 : sweep_shape<2, 5>(p.n, p.m).smem <= 84 * 1024;
    if (deep) return reverse ? launch<true, 2, 5>(p, st) : launch<false, 2, 5>(p, st);
    return reverse ? launch<true, 2, 3>(p, st) : launch<false, 2, 3>(p, st);
  }
  const bool deep = force ? force >= 5 : sweep_shape<1, 5>(p.n, p.m).smem <= 84 * 1024;
  if (deep) return reverse ? launch<true, 1, 5>(p, st) : launch<false, 1, 5>(p, st);
  return reverse ? launch<true, 1, 3>(p, st) : launch<false, 1, 3>(p, st);
}
