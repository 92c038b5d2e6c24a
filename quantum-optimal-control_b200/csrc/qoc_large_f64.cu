// FP64 path for n > 64 (configs C4: n = 216, C5: n = 128): the matrices no longer fit in shared
// memory, so each CTA runs a tiled DMMA GEMM with operands streamed from global memory / L2
// (cp.async, double-buffered, zero-filled at the edges) and keeps its intermediates in a per-CTA
// global scratch area.  Same algorithm as qoc_mma_f64.cu: Paterson-Stockmeyer Taylor polynomial +
// squarings (core/tensorflow_state.py:25-46), chain X_t = P_t X_{t-1} (:204-242), costate sweep.
#include "qoc_internal.cuh"
#include <math.h>
#include <stdlib.h>
#include <algorithm>

#define DEVINL __device__ __forceinline__

namespace {

constexpr int KT = 32;    // k-tile

DEVINL double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
DEVINL void cp_async16_zfill(void* smem_dst, const void* gmem_src, bool valid) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "r"(sz));
}
DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
DEVINL int sw_mask(int r) { return ((r & 1) * 5) ^ (((r >> 1) & 3) << 1); }
DEVINL void dmma(double& d0, double& d1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

template <int TS, int RB, int CB>
struct LT {
  static constexpr int NBLK = TS / 8;
  static constexpr int WR = NBLK / RB;
  static constexpr int WC = NBLK / CB;
  static constexpr int WARPS = WR * WC;
  static constexpr int THREADS = 32 * WARPS;
  static constexpr int A_ELEMS = TS * KT;       // A tile [TS][KT]
  static constexpr int B_ELEMS = KT * TS;       // B tile [KT][TS]
  static constexpr size_t SMEM = (size_t)2 * (A_ELEMS + B_ELEMS) * sizeof(cplx);
  static_assert(NBLK % RB == 0 && NBLK % CB == 0, "tile must divide the block grid");
};

// C (n x n, ldc) = A (lda) * B (ldb)  [+ c_id * I + c_h * Hadd]   -- one CTA, all operands in global memory
// mrows < n: A and C have only mrows rows (dense-m costate: Lambda is [m][n]); CONJB: B is used conjugated
template <int TS, int RB, int CB, bool CONJB = false>
DEVINL void cta_gemm(const cplx* __restrict__ A, int lda, const cplx* __restrict__ B, int ldb, cplx* __restrict__ C, int ldc,
                     int n, double c_id, double c_h, const cplx* Hadd, int ldh, cplx* sm, int mrows = -1) {
  if (mrows < 0) mrows = n;
  typedef LT<TS, RB, CB> L;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int rb0 = (warp / L::WC) * RB, cb0 = (warp % L::WC) * CB;
  cplx* As = sm;                       // [2][TS*KT]
  cplx* Bs = sm + 2 * L::A_ELEMS;      // [2][KT*TS]
  const int nt = (n + TS - 1) / TS, nk = (n + KT - 1) / KT, ntr = (mrows + TS - 1) / TS;

  auto load_tiles = [&](int ti, int tj, int kt, int buf) {
    cplx* a = As + buf * L::A_ELEMS;
    cplx* b = Bs + buf * L::B_ELEMS;
    for (int idx = tid; idx < TS * KT; idx += L::THREADS) {
      const int r = idx / KT, k = idx - r * KT;
      const int gr = ti * TS + r, gk = kt * KT + k;
      const bool ok = gr < mrows && gk < n;
      cp_async16_zfill(a + r * KT + (k ^ sw_mask(r)), A + (ok ? (size_t)gr * lda + gk : 0), ok);
    }
    for (int idx = tid; idx < KT * TS; idx += L::THREADS) {
      const int k = idx / TS, c = idx - k * TS;
      const int gk = kt * KT + k, gc = tj * TS + c;
      const bool ok = gk < n && gc < n;
      cp_async16_zfill(b + k * TS + (c ^ sw_mask(k)), B + (ok ? (size_t)gk * ldb + gc : 0), ok);
    }
    cp_async_commit();
  };

  for (int ti = 0; ti < ntr; ++ti)
    for (int tj = 0; tj < nt; ++tj) {
      // 3M (Gauss) complex product, as in qoc_mma_f64.cu: cr = Ar Br, t2 = Ai Bi, ci = (Ar+Ai)(Br+Bi); combined below
      double cr[RB][CB][2], ci[RB][CB][2], t2[RB][CB][2];
#pragma unroll
      for (int i = 0; i < RB; ++i)
#pragma unroll
        for (int j = 0; j < CB; ++j) cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = t2[i][j][0] = t2[i][j][1] = 0.0;
      __syncthreads();                                 // previous tile's readers are done with the buffers
      load_tiles(ti, tj, 0, 0);
      for (int kt = 0; kt < nk; ++kt) {
        if (kt + 1 < nk) { load_tiles(ti, tj, kt + 1, (kt + 1) & 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const cplx* a = As + (kt & 1) * L::A_ELEMS;
        const cplx* b = Bs + (kt & 1) * L::B_ELEMS;
        const int krem = n - kt * KT;
        const int ksteps = krem >= KT ? KT / 4 : (krem + 3) / 4;
#pragma unroll 2
        for (int ks = 0; ks < ksteps; ++ks) {
          const int k = 4 * ks + q;
          cplx av[RB], bv[CB];
#pragma unroll
          for (int i = 0; i < RB; ++i) { const int r = 8 * (rb0 + i) + g; av[i] = a[r * KT + (k ^ sw_mask(r))]; }
          const int bm = sw_mask(k);
#pragma unroll
          for (int j = 0; j < CB; ++j) { bv[j] = b[k * TS + ((8 * (cb0 + j) + g) ^ bm)]; if (CONJB) bv[j].y = -bv[j].y; }
          double sa[RB], sb[CB];
#pragma unroll
          for (int i = 0; i < RB; ++i) sa[i] = av[i].x + av[i].y;
#pragma unroll
          for (int j = 0; j < CB; ++j) sb[j] = bv[j].x + bv[j].y;
#pragma unroll
          for (int i = 0; i < RB; ++i)
#pragma unroll
            for (int j = 0; j < CB; ++j) dmma(cr[i][j][0], cr[i][j][1], av[i].x, bv[j].x);
#pragma unroll
          for (int i = 0; i < RB; ++i)
#pragma unroll
            for (int j = 0; j < CB; ++j) dmma(t2[i][j][0], t2[i][j][1], av[i].y, bv[j].y);
#pragma unroll
          for (int i = 0; i < RB; ++i)
#pragma unroll
            for (int j = 0; j < CB; ++j) dmma(ci[i][j][0], ci[i][j][1], sa[i], sb[j]);
        }
        __syncthreads();                               // buffer (kt & 1) may be refilled two iterations later
      }
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int r = ti * TS + 8 * (rb0 + i) + g;
#pragma unroll
        for (int j = 0; j < CB; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = tj * TS + 8 * (cb0 + j) + 2 * q + e;
            if (r < mrows && c < n) {
              double vr = cr[i][j][e] - t2[i][j][e], vi = ci[i][j][e] - cr[i][j][e] - t2[i][j][e];
              if (Hadd) { const cplx h = Hadd[(size_t)r * ldh + c]; vr += c_h * h.x; vi += c_h * h.y; }
              if (r == c) vr += c_id;
              C[(size_t)r * ldc + c] = make_double2(vr, vi);
            }
          }
      }
    }
  __syncthreads();                                     // C is complete and visible to the whole CTA
}

// ---------------------------------------------------------------------------------------------
template <int TS, int RB, int CB>
__global__ void __launch_bounds__(LT<TS, RB, CB>::THREADS) k_expm_large(QocParams p, cplx* scratch) {
  typedef LT<TS, RB, CB> L;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* sm = reinterpret_cast<cplx*>(smem_raw);
  __shared__ double wts[32];
  const int tid = threadIdx.x;
  const int n = p.n, K = p.K, T = p.T;
  const size_t nn = (size_t)n * n;
  cplx* Hg = scratch + (size_t)blockIdx.x * 4 * nn;
  cplx* H2 = Hg + nn;
  cplx* R0 = H2 + nn;
  cplx* R1 = R0 + nn;
  const long long items = (long long)p.B * T;
  cplx* Pout = reinterpret_cast<cplx*>(p.P);
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = (int)(item / T), t = (int)(item % T);
    if (tid == 0) wts[0] = p.inv2s;
    if (tid >= 1 && tid <= K) wts[tid] = p.maxA[tid - 1] * sin(p.base[((size_t)b * K + tid - 1) * T + t]) * p.inv2s;
    __syncthreads();
    for (size_t idx = tid; idx < nn; idx += L::THREADS) {
      double hx = 0.0, hy = 0.0;
      for (int k = 0; k <= K; ++k) {
        const cplx a = p.A[(size_t)k * nn + idx];
        hx = fma(wts[k], a.x, hx); hy = fma(wts[k], a.y, hy);
      }
      Hg[idx] = make_double2(hx, hy);
    }
    __syncthreads();
    cplx* dstP = Pout + (size_t)item * nn;
    const int pp = p.p;
    cplx* R = R0;
    cplx* Tm = R1;
    // Paterson-Stockmeyer, block 2 (see qoc_mma_f64.cu)
    if (pp >= 2) cta_gemm<TS, RB, CB>(Hg, n, Hg, n, H2, n, n, 0.0, 0.0, nullptr, 0, sm);
    int blk;
    if (pp & 1) {
      const double cid = p.invfact[pp - 1], ch = p.invfact[pp];
      for (size_t idx = tid; idx < nn; idx += L::THREADS) {
        const cplx h = Hg[idx];
        const int r = (int)(idx / n), c = (int)(idx - (size_t)r * n);
        R[idx] = make_double2(ch * h.x + (r == c ? cid : 0.0), ch * h.y);
      }
      blk = pp / 2 - 1;
    } else {
      const double cp = p.invfact[pp], cid = p.invfact[pp - 2], ch = p.invfact[pp - 1];
      for (size_t idx = tid; idx < nn; idx += L::THREADS) {
        const cplx h = Hg[idx], h2 = H2[idx];
        const int r = (int)(idx / n), c = (int)(idx - (size_t)r * n);
        R[idx] = make_double2(cp * h2.x + ch * h.x + (r == c ? cid : 0.0), cp * h2.y + ch * h.y);
      }
      blk = pp / 2 - 2;
    }
    __syncthreads();
    for (; blk >= 0; --blk) {
      cplx* out = (blk == 0 && p.s == 0) ? dstP : Tm;
      cta_gemm<TS, RB, CB>(R, n, H2, n, out, n, n, p.invfact[2 * blk], p.invfact[2 * blk + 1], Hg, n, sm);
      Tm = R; R = out;
    }
    if (R == dstP) continue;
    if (p.s == 0) {
      for (size_t idx = tid; idx < nn; idx += L::THREADS) dstP[idx] = R[idx];
      __syncthreads();
      continue;
    }
    if (Tm == dstP) Tm = (R == R0) ? R1 : R0;
    for (int s = 0; s < p.s; ++s) {                    // squarings; the last one lands in the propagator cache
      cplx* out = (s == p.s - 1) ? dstP : Tm;
      cta_gemm<TS, RB, CB>(R, n, R, n, out, n, n, 0.0, 0.0, nullptr, 0, sm);
      Tm = R; R = out;
    }
  }
}

// chain: X_{t+1} = P_t X_t, one CTA per instance, X ping-pong in global scratch
template <int TS, int RB, int CB>
__global__ void __launch_bounds__(LT<TS, RB, CB>::THREADS) k_chain_large(QocParams p, cplx* scratch) {
  typedef LT<TS, RB, CB> L;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* sm = reinterpret_cast<cplx*>(smem_raw);
  __shared__ double red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = p.n, T = p.T, m = p.m, b = blockIdx.x;
  const size_t nn = (size_t)n * n;
  cplx* X0 = scratch + (size_t)b * 2 * nn;
  cplx* X1 = X0 + nn;
  const cplx* Pg = reinterpret_cast<const cplx*>(p.P) + (size_t)b * T * nn;
  cplx* psi_b = p.psi + (size_t)b * (T + 1) * m * n;
  for (size_t idx = tid; idx < nn; idx += L::THREADS) X0[idx] = p.U0[idx];
  for (int idx = tid; idx < m * n; idx += L::THREADS) psi_b[idx] = p.V[idx];
  __syncthreads();
  auto extract = [&](const cplx* X, int t) {
    cplx* out = psi_b + (size_t)t * m * n;
    for (int idx = tid; idx < m * n; idx += L::THREADS) {
      const int j = idx / n, i = idx - j * n;
      if (p.has_cidx) out[idx] = X[(size_t)i * n + p.cidx[j]];
      else {
        double ax = 0.0, ay = 0.0;
        for (int c = 0; c < n; ++c) {
          const cplx x = X[(size_t)i * n + c], v = p.V[j * n + c];
          ax += x.x * v.x - x.y * v.y; ay += x.x * v.y + x.y * v.x;
        }
        out[idx] = make_double2(ax, ay);
      }
    }
  };
  cplx* Xc = X0;
  cplx* Xn = X1;
  for (int t = 0; t < T; ++t) {
    cta_gemm<TS, RB, CB>(Pg + (size_t)t * nn, n, Xc, n, Xn, n, n, 0.0, 0.0, nullptr, 0, sm);
    extract(Xn, t + 1);
    cplx* tmp = Xc; Xc = Xn; Xn = tmp;
  }
  __syncthreads();
  cplx* Uf = p.Ufin + (size_t)b * nn;
  for (size_t idx = tid; idx < nn; idx += L::THREADS) Uf[idx] = Xc[idx];
  double v = 0.0;
  for (int r = tid; r < n; r += L::THREADS) {
    double sr = 0.0, si = 0.0;
    for (int c = 0; c < n; ++c) { const cplx x = Xc[(size_t)r * n + c]; sr += x.x; si += x.y; }
    v += sr * sr + si * si;
  }
  v = warp_sum_d(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < L::WARPS; ++w) s += red[w];
    p.scal[(size_t)b * 8 + 5] = s / (double)n;
  }
}

// costate for large n: lambda chunk in shared memory, P_t read straight from global / L2
__global__ void k_costate_large(QocParams p, int mc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = p.n, m = p.m, T = p.T, mn = m * n;
  const size_t nn = (size_t)n * n;
  const int j0 = blockIdx.y * mc;
  const int mloc = min(mc, m - j0);
  const int ln = mloc * n;
  cplx* lam_s = reinterpret_cast<cplx*>(smem_raw);        // [2][mc*n]
  const int tid = threadIdx.x, nt = blockDim.x;
  const int b = blockIdx.x;
  const cplx* Pg = reinterpret_cast<const cplx*>(p.P) + (size_t)b * T * nn;
  const cplx* psi_b = p.psi + (size_t)b * (T + 1) * mn + (size_t)j0 * n;
  cplx* lam_b = p.lam + (size_t)b * (T + 1) * mn + (size_t)j0 * n;
  const cplx* phi = p.phi + (size_t)j0 * n;
  const double* sc = p.scal + (size_t)b * 8;
  const double o_re = sc[0], o_im = sc[1], spdfac = sc[4];
  const bool forb = p.reg.has_forbidden && p.fw != nullptr;
  const bool spd = p.reg.has_speed_up != 0;
  const double m2 = (double)m * (double)m;
  auto source = [&](int t, int idx) -> cplx {
    cplx s = make_double2(0.0, 0.0);
    if (forb && p.dressW) {
      const cplx d = p.psid[((size_t)b * (T + 1) + t) * mn + (size_t)j0 * n + idx];
      s.x += d.x; s.y += d.y;
    } else if (forb) {
      const cplx x = psi_b[(size_t)t * mn + idx];
      const double pop = x.x * x.x + x.y * x.y;
      const double c = p.fw[idx % n] / (double)T * 2.0 * pop;
      s.x += c * x.x; s.y += c * x.y;
    }
    if (spd) {
      const cplx o = p.ot[(size_t)b * (T + 1) + t];
      const cplx ph = phi[idx];
      s.x += spdfac * (o.x * ph.x - o.y * ph.y); s.y += spdfac * (o.x * ph.y + o.y * ph.x);
    }
    return s;
  };
  for (int idx = tid; idx < ln; idx += nt) {
    const cplx ph = phi[idx];
    cplx l = make_double2((o_re * ph.x - o_im * ph.y) * (-2.0 / m2), (o_re * ph.y + o_im * ph.x) * (-2.0 / m2));
    const cplx s = source(T, idx);
    l.x += s.x; l.y += s.y;
    lam_s[idx] = l;
    lam_b[(size_t)T * mn + idx] = l;
  }
  int cur = 0;
  for (int t = T - 1; t >= 1; --t) {
    __syncthreads();
    const cplx* Pt = Pg + (size_t)t * nn;
    const cplx* lc = lam_s + cur * mc * n;
    cplx* lnx = lam_s + (cur ^ 1) * mc * n;
    for (int idx = tid; idx < ln; idx += nt) {
      const int j = idx / n, i = idx - j * n;
      double ax = 0.0, ay = 0.0;
      for (int r = 0; r < n; ++r) {                    // conj(P[r][i]) * lam[j][r]; coalesced over i
        const cplx a = __ldg(Pt + (size_t)r * n + i);
        const cplx l = lc[j * n + r];
        ax += a.x * l.x + a.y * l.y; ay += a.x * l.y - a.y * l.x;
      }
      const cplx s = source(t, idx);
      const cplx l = make_double2(ax + s.x, ay + s.y);
      lnx[idx] = l;
      lam_b[(size_t)t * mn + idx] = l;
    }
    cur ^= 1;
  }
}


// ---------------------------------------------------------------------------------------------
// k_costate_large_mma: dense-m reverse sweep for n > 64 on the DMMA tiles.  With the states as rows,
//   L(t) = L(t+1) conj(P_t) + S(t),   L = [m][n]  (lambda_j(t) = P_t^dagger lambda_j(t+1) + s_j(t), tensorflow_state.py:214-220 reversed)
// is one m x n x n GEMM per step; one CTA per instance walks t = T-1 .. 1 (the scalar k_costate_large needs n dependent
// FMAs per element and step).  Sources are written into L(t) first and added by the GEMM epilogue.
// ---------------------------------------------------------------------------------------------
template <int TS, int RB, int CB>
__global__ void __launch_bounds__(LT<TS, RB, CB>::THREADS) k_costate_large_mma(QocParams p) {
  typedef LT<TS, RB, CB> L;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* sm = reinterpret_cast<cplx*>(smem_raw);
  const int n = p.n, m = p.m, T = p.T, mn = m * n;
  const size_t nn = (size_t)n * n;
  const int tid = threadIdx.x, nth = L::THREADS;
  const int b = blockIdx.x;
  const cplx* Pg = reinterpret_cast<const cplx*>(p.P) + (size_t)b * T * nn;
  const cplx* psi_b = p.psi + (size_t)b * (T + 1) * mn;
  cplx* lam_b = p.lam + (size_t)b * (T + 1) * mn;
  const double* sc = p.scal + (size_t)b * 8;
  const double o_re = sc[0], o_im = sc[1], spdfac = sc[4];
  const bool forb = p.reg.has_forbidden && p.fw != nullptr;
  const bool spd = p.reg.has_speed_up != 0;
  const bool has_src = forb || spd;
  const double m2 = (double)m * (double)m;
  auto source = [&](int t, int idx) -> cplx {
    cplx s = make_double2(0.0, 0.0);
    if (forb && p.dressW) {
      const cplx d = p.psid[((size_t)b * (T + 1) + t) * mn + idx];
      s.x += d.x; s.y += d.y;
    } else if (forb) {
      const cplx x = psi_b[(size_t)t * mn + idx];
      const double pop = x.x * x.x + x.y * x.y;
      const double c = p.fw[idx % n] / (double)T * 2.0 * pop;
      s.x += c * x.x; s.y += c * x.y;
    }
    if (spd) {
      const cplx o = p.ot[(size_t)b * (T + 1) + t];
      const cplx ph = p.phi[idx];
      s.x += spdfac * (o.x * ph.x - o.y * ph.y); s.y += spdfac * (o.x * ph.y + o.y * ph.x);
    }
    return s;
  };
  for (int idx = tid; idx < mn; idx += nth) {
    const cplx ph = p.phi[idx];
    cplx l = make_double2((o_re * ph.x - o_im * ph.y) * (-2.0 / m2), (o_re * ph.y + o_im * ph.x) * (-2.0 / m2));
    const cplx s = source(T, idx);
    l.x += s.x; l.y += s.y;
    lam_b[(size_t)T * mn + idx] = l;
  }
  __syncthreads();
  for (int t = T - 1; t >= 1; --t) {
    cplx* Lt = lam_b + (size_t)t * mn;
    if (has_src) {
      for (int idx = tid; idx < mn; idx += nth) Lt[idx] = source(t, idx);
      __syncthreads();
    }
    cta_gemm<TS, RB, CB, true>(lam_b + (size_t)(t + 1) * mn, n, Pg + (size_t)t * nn, n, Lt, n, n, 0.0, has_src ? 1.0 : 0.0,
                               has_src ? Lt : nullptr, n, sm, m);
  }
}

// ---------------------------------------------------------------------------------------------
// k_grad_large: dense-m, dense-control gradient for n > 64 as ONE GEMM per (b,t) (matexp_op_grad, tensorflow_state.py:49-65):
//   W[a][c] = sum_j conj(lambda_j(t+1)[a]) psi_j(t+1)[c]   (n x n x m, A = Lambda^H read transposed from its [m][n] storage),
//   g_k = Re sum_ac A_k[a][c] W[a][c]                        (K dense operators, reduced in the epilogue, W never stored)
// instead of K n^2 length-m dot products (k_grad): 1/K of the flops, on the DMMA pipe.
// ---------------------------------------------------------------------------------------------
constexpr int GRAD_MAXK = 8;
template <int TS, int RB, int CB>
__global__ void __launch_bounds__(LT<TS, RB, CB>::THREADS) k_grad_large(QocParams p) {
  typedef LT<TS, RB, CB> L;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* sm = reinterpret_cast<cplx*>(smem_raw);
  __shared__ double red[GRAD_MAXK][L::WARPS];
  const int n = p.n, m = p.m, T = p.T, K = p.K, mn = m * n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int rb0 = (warp / L::WC) * RB, cb0 = (warp % L::WC) * CB;
  cplx* As = sm;                       // [2][KT*TS]: Lambda tile [j][a] (k-major, like a B tile)
  cplx* Bs = sm + 2 * L::A_ELEMS;      // [2][KT*TS]: Psi tile [j][c]
  const int nt = (n + TS - 1) / TS, nk = (m + KT - 1) / KT;
  const long long items = (long long)p.B * T;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = (int)(item / T), t = (int)(item % T);
    const size_t off = ((size_t)b * (T + 1) + (t + 1)) * mn;
    const cplx* __restrict__ lam = p.lam + off;
    const cplx* __restrict__ psi = p.psi + off;
    double gk[GRAD_MAXK];
#pragma unroll
    for (int k = 0; k < GRAD_MAXK; ++k) gk[k] = 0.0;
    auto load_tiles = [&](int ti, int tj, int kt, int buf) {
      cplx* a = As + buf * L::A_ELEMS;
      cplx* bb = Bs + buf * L::B_ELEMS;
      for (int idx = tid; idx < KT * TS; idx += L::THREADS) {
        const int k = idx / TS, c = idx - k * TS;
        const int gj = kt * KT + k;
        const int ga = ti * TS + c, gc = tj * TS + c;
        const bool oka = gj < m && ga < n, okb = gj < m && gc < n;
        cp_async16_zfill(a + k * TS + (c ^ sw_mask(k)), lam + (oka ? (size_t)gj * n + ga : 0), oka);
        cp_async16_zfill(bb + k * TS + (c ^ sw_mask(k)), psi + (okb ? (size_t)gj * n + gc : 0), okb);
      }
      cp_async_commit();
    };
    for (int ti = 0; ti < nt; ++ti)
      for (int tj = 0; tj < nt; ++tj) {
        double cr[RB][CB][2], ci[RB][CB][2], t2[RB][CB][2];
#pragma unroll
        for (int i = 0; i < RB; ++i)
#pragma unroll
          for (int j = 0; j < CB; ++j) cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = t2[i][j][0] = t2[i][j][1] = 0.0;
        __syncthreads();
        load_tiles(ti, tj, 0, 0);
        for (int kt = 0; kt < nk; ++kt) {
          if (kt + 1 < nk) { load_tiles(ti, tj, kt + 1, (kt + 1) & 1); cp_async_wait<1>(); }
          else cp_async_wait<0>();
          __syncthreads();
          const cplx* a = As + (kt & 1) * L::A_ELEMS;
          const cplx* bb = Bs + (kt & 1) * L::B_ELEMS;
          const int krem = m - kt * KT;
          const int ksteps = krem >= KT ? KT / 4 : (krem + 3) / 4;
#pragma unroll 2
          for (int ks = 0; ks < ksteps; ++ks) {
            const int k = 4 * ks + q;
            const int km = sw_mask(k);
            cplx av[RB], bv[CB];
#pragma unroll
            for (int i = 0; i < RB; ++i) { av[i] = a[k * TS + ((8 * (rb0 + i) + g) ^ km)]; av[i].y = -av[i].y; }
#pragma unroll
            for (int j = 0; j < CB; ++j) bv[j] = bb[k * TS + ((8 * (cb0 + j) + g) ^ km)];
            double sa[RB], sb[CB];
#pragma unroll
            for (int i = 0; i < RB; ++i) sa[i] = av[i].x + av[i].y;
#pragma unroll
            for (int j = 0; j < CB; ++j) sb[j] = bv[j].x + bv[j].y;
#pragma unroll
            for (int i = 0; i < RB; ++i)
#pragma unroll
              for (int j = 0; j < CB; ++j) dmma(cr[i][j][0], cr[i][j][1], av[i].x, bv[j].x);
#pragma unroll
            for (int i = 0; i < RB; ++i)
#pragma unroll
              for (int j = 0; j < CB; ++j) dmma(t2[i][j][0], t2[i][j][1], av[i].y, bv[j].y);
#pragma unroll
            for (int i = 0; i < RB; ++i)
#pragma unroll
              for (int j = 0; j < CB; ++j) dmma(ci[i][j][0], ci[i][j][1], sa[i], sb[j]);
          }
          __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < RB; ++i) {
          const int r = ti * TS + 8 * (rb0 + i) + g;
#pragma unroll
          for (int j = 0; j < CB; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = tj * TS + 8 * (cb0 + j) + 2 * q + e;
              if (r < n && c < n) {
                const double vr = cr[i][j][e] - t2[i][j][e], vi = ci[i][j][e] - cr[i][j][e] - t2[i][j][e];
#pragma unroll
                for (int k = 0; k < GRAD_MAXK; ++k)
                  if (k < K) {
                    const cplx ak = __ldg(p.A + ((size_t)(k + 1) * n + r) * n + c);
                    gk[k] += ak.x * vr - ak.y * vi;
                  }
              }
            }
        }
      }
#pragma unroll
    for (int k = 0; k < GRAD_MAXK; ++k)
      if (k < K) {
        const double v = warp_sum_d(gk[k]);
        if (lane == 0) red[k][warp] = v;
      }
    __syncthreads();
    if (tid < K) {
      double v = 0.0;
      for (int w = 0; w < L::WARPS; ++w) v += red[tid][w];
      p.gctrl[((size_t)b * K + tid) * T + t] = v;
    }
    __syncthreads();
  }
}

}  // namespace

// scratch sizes (complex elements) -----------------------------------------------------------------
size_t qoc_large_scratch_elems(int n, int B, int sm_count) {
  const size_t nn = (size_t)n * n;
  const size_t expm = (size_t)sm_count * 4 * nn, chain = (size_t)B * 2 * nn;
  return expm > chain ? expm : chain;
}

static int pick_ts(int n) {   // the tile edge with the least padding
  const int p64 = (n + 63) / 64 * 64, p72 = (n + 71) / 72 * 72;
  return p72 < p64 ? 72 : 64;
}

cudaError_t qoc_launch_expm_large(const QocParams& p, int sm_count, void* scratch, cudaStream_t st, int64_t* launches) {
  ++*launches;
  const long long items = (long long)p.B * p.T;
  const int grid = (int)(items < sm_count ? items : sm_count);
  cudaError_t e;
  if (pick_ts(p.n) == 72) {
    typedef LT<72, 3, 3> L;
    e = cudaFuncSetAttribute(k_expm_large<72, 3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
    if (e != cudaSuccess) return e;
    k_expm_large<72, 3, 3><<<grid, L::THREADS, L::SMEM, st>>>(p, (cplx*)scratch);
  } else {
    typedef LT<64, 2, 4> L;
    e = cudaFuncSetAttribute(k_expm_large<64, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
    if (e != cudaSuccess) return e;
    k_expm_large<64, 2, 4><<<grid, L::THREADS, L::SMEM, st>>>(p, (cplx*)scratch);
  }
  return cudaGetLastError();
}

cudaError_t qoc_launch_chain_large(const QocParams& p, void* scratch, cudaStream_t st, int64_t* launches) {
  ++*launches;
  cudaError_t e;
  if (pick_ts(p.n) == 72) {
    typedef LT<72, 3, 3> L;
    e = cudaFuncSetAttribute(k_chain_large<72, 3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
    if (e != cudaSuccess) return e;
    k_chain_large<72, 3, 3><<<p.B, L::THREADS, L::SMEM, st>>>(p, (cplx*)scratch);
  } else {
    typedef LT<64, 2, 4> L;
    e = cudaFuncSetAttribute(k_chain_large<64, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
    if (e != cudaSuccess) return e;
    k_chain_large<64, 2, 4><<<p.B, L::THREADS, L::SMEM, st>>>(p, (cplx*)scratch);
  }
  return cudaGetLastError();
}

cudaError_t qoc_launch_costate_large(const QocParams& p, cudaStream_t st, int64_t* launches) {
  ++*launches;
  if (2 * p.m >= 64 && !getenv("QOC_B200_NO_COSTATE_GEMM")) {            // dense-m: one GEMM per step on the DMMA tiles
    cudaError_t e;
    if (pick_ts(p.n) == 72) {
      typedef LT<72, 3, 3> L;
      e = cudaFuncSetAttribute(k_costate_large_mma<72, 3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
      if (e != cudaSuccess) return e;
      k_costate_large_mma<72, 3, 3><<<p.B, L::THREADS, L::SMEM, st>>>(p);
    } else {
      typedef LT<64, 2, 4> L;
      e = cudaFuncSetAttribute(k_costate_large_mma<64, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
      if (e != cudaSuccess) return e;
      k_costate_large_mma<64, 2, 4><<<p.B, L::THREADS, L::SMEM, st>>>(p);
    }
    return cudaGetLastError();
  }
  int mc = p.m;
  while ((size_t)2 * mc * p.n * sizeof(cplx) > 96 * 1024 && mc > 1) mc = (mc + 1) / 2;
  if (mc > 4) mc = 4;                                   // more CTAs: the sweep is latency bound
  const size_t smem = (size_t)2 * mc * p.n * sizeof(cplx);
  cudaError_t e = cudaFuncSetAttribute(k_costate_large, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const dim3 grid(p.B, (p.m + mc - 1) / mc);
  k_costate_large<<<grid, 256, smem, st>>>(p, mc);
  return cudaGetLastError();
}

bool qoc_grad_large_supported(const QocParams& p) { return p.n > 64 && p.K <= GRAD_MAXK; }

cudaError_t qoc_launch_grad_large(const QocParams& p, int sm_count, cudaStream_t st, int64_t* launches) {
  ++*launches;
  const long long items = (long long)p.B * p.T;
  cudaError_t e;
  if (pick_ts(p.n) == 72) {
    typedef LT<72, 3, 3> L;
    e = cudaFuncSetAttribute(k_grad_large<72, 3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
    if (e != cudaSuccess) return e;
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_grad_large<72, 3, 3>, L::THREADS, L::SMEM);
    const long long grid = std::min<long long>(items, (long long)sm_count * (occ < 1 ? 1 : occ));
    k_grad_large<72, 3, 3><<<(unsigned)grid, L::THREADS, L::SMEM, st>>>(p);
  } else {
    typedef LT<64, 2, 4> L;
    e = cudaFuncSetAttribute(k_grad_large<64, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
    if (e != cudaSuccess) return e;
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_grad_large<64, 2, 4>, L::THREADS, L::SMEM);
    const long long grid = std::min<long long>(items, (long long)sm_count * (occ < 1 ? 1 : occ));
    k_grad_large<64, 2, 4><<<(unsigned)grid, L::THREADS, L::SMEM, st>>>(p);
  }
  return cudaGetLastError();
}
