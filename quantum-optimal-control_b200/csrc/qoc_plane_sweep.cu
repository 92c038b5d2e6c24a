// State / costate sweeps over split-plane (QOC_F16X2) propagators, few concerned states (m <= 8), n <= 256.
//
//   k_plane_sweep<false>: psi_j(t+1) = P_t psi_j(t)                          (init_tf_inter_vectors, core/tensorflow_state.py:229-242)
//   k_plane_sweep<true> : lambda_j(t) = P_t^dagger lambda_j(t+1) + source_j(t)   (TF autodiff through :214-220)
//
// One CTA per instance walks the T steps; a step reads the 8 n ld bytes of P_t exactly once from HBM, in
// R-row chunks (4 planes x R rows, one cp.async.bulk per plane) through an NST-deep shared-memory ring with
// an mbarrier per stage.  2^13 P = h0 + h1 is exact in fp32 and widened to double; the states stay fp64.
//   forward: warp = row of the chunk, lane = pair of k (4-byte loads of P, 16-byte loads of the even-k / odd-k halves
//            of the state vectors: all conflict-free); the 2 m partial sums meet in a shuffle reduce-scatter;
//   reverse: thread = (column pair, state pair) keeps its four complex sums in registers over all chunks of the
//            step, every shared-memory read of P is a conflict-free 4-byte load, lambda is a broadcast.
#include "qoc_internal.cuh"
#include "qoc_tc_f16.cuh"
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
DEVINL void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
DEVINL cplx cmul(const cplx a, const cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
[[maybe_unused]] DEVINL void widen8(const uint4 h0, const uint4 h1, double (&v)[8]) {
  const uint32_t a[4] = {h0.x, h0.y, h0.z, h0.w}, b[4] = {h1.x, h1.y, h1.z, h1.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&a[i]));
    const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&b[i]));
    v[2 * i] = (double)(fa.x + fb.x); v[2 * i + 1] = (double)(fa.y + fb.y);
  }
}
DEVINL double2 widen2(uint32_t h0, uint32_t h1) {
  const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&h0));
  const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&h1));
  return make_double2((double)(fa.x + fb.x), (double)(fa.y + fb.y));
}

// CTA shape: NW warps (4 for n <= 64, 8 for n <= 128, else 16); R rows of P per chunk (a multiple of NW; the whole matrix
// when it is small, else ~24 KB of planes); NST ring stages within a per-CTA shared-memory budget chosen so that small
// problems keep several CTAs (instances) resident per SM -- a sweep is a chain of T dependent steps, its latency is hidden
// by running many instances side by side.
struct PlaneSweepShape { int nw, R, nst; size_t stage_halfs, smem; };

static int sweep_vec_rows(int m) { return m <= 2 ? 2 : m <= 4 ? 4 : 8; }

// B, sms: when ceil(B / sms) instances fit on an SM at once the whole batch runs in ONE wave (C3: 1024 instances on 148 SMs
// need 7 per SM; at 6 per SM a second, nearly empty wave doubles the sweep time)
PlaneSweepShape plane_sweep_shape(int n, int m, int B = 0, int sms = 0) {
  PlaneSweepShape s;
  const int ld = tc_ld(n);
  s.nw = n <= 64 ? 4 : n <= 128 ? 8 : 16;
  const size_t stage_target = n <= 64 ? 8 * 1024 : n <= 128 ? 16 * 1024 : 24 * 1024;
  int r = (int)(stage_target / (8 * (size_t)ld)) / s.nw * s.nw;
  if (r < s.nw) r = s.nw;
  {                                                              // two rows per warp and pass (forward sweep) when two stages fit
    const size_t vec0 = (size_t)2 * sweep_vec_rows(m) * ld * sizeof(cplx);
    if (r < 2 * s.nw && vec0 + 256 + 2 * (size_t)8 * 2 * s.nw * ld <= 220 * 1024) r = 2 * s.nw;
  }
  const int rfull = (n + s.nw - 1) / s.nw * s.nw;
  s.R = r < rfull ? r : rfull;
  s.stage_halfs = (size_t)4 * s.R * ld;
  const size_t vec = (size_t)2 * sweep_vec_rows(m) * ld * sizeof(cplx);          // [2][mv][2][ld/2]
  const size_t fixed = vec + 64 + 128;
  size_t budget = n <= 64 ? 36 * 1024 : n <= 128 ? 100 * 1024 : 220 * 1024;
  if (B > 0 && sms > 0) {
    const int per_sm = (B + sms - 1) / sms;
    if (per_sm >= 2 && per_sm <= 12) {
      const size_t fit = (size_t)227 * 1024 / per_sm - 1024 - 256;      // 1 KB per CTA is reserved by the driver
      if (fit < budget && fit >= fixed + 2 * s.stage_halfs * sizeof(__half)) budget = fit;
    }
  }
  int nst = budget > fixed ? (int)((budget - fixed) / (s.stage_halfs * sizeof(__half))) : 0;
  if (nst > 8) nst = 8;
  if (nst < 2 && fixed + 2 * s.stage_halfs * sizeof(__half) <= 220 * 1024) nst = 2;
  s.nst = nst;
  s.smem = fixed + (size_t)(nst > 0 ? nst : 0) * s.stage_halfs * sizeof(__half);
  return s;
}

// butterfly reduce-scatter of 2*MS doubles over the 32 lanes: afterwards lane L (L even) holds the warp sum of element
// idx(L) = L >> (5 - log2(2 MS)) ... expressed below by halving the live set at every step
template <int NV>
DEVINL double reduce_scatter(double (&a)[NV], int lane, int& idx) {
  int base = 0;
#pragma unroll
  for (int cnt = NV, off = 16; cnt > 1; cnt >>= 1, off >>= 1) {
    const bool up = (lane & off) != 0;
    const int h = cnt >> 1;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const double send = up ? a[i] : a[i + h];
      const double keep = up ? a[i + h] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    base = 2 * base + (up ? 1 : 0);
  }
  // remaining lane bits hold replicas of partial sums: finish with plain butterflies
  double v = a[0];
  int off = 16;
#pragma unroll
  for (int cnt = NV; cnt > 1; cnt >>= 1) off >>= 1;
  for (; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  idx = base;          // element index in "bit-reversed halving" order: see caller
  return v;
}

template <bool REV, int MS, int NW>
__global__ void __launch_bounds__(32 * NW) k_plane_sweep(QocParams p, const __half* __restrict__ Pp, int NST, int R) {
  constexpr int NTH = 32 * NW;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int n = p.n, m = p.m, T = p.T, mn = m * n;
  const int ld = tc_ld(n), lh = ld >> 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const size_t stage_halfs = (size_t)4 * R * ld;
  __half* ring = reinterpret_cast<__half*>(smem_raw);                                    // [NST][4][R][ld]
  cplx* vecs = reinterpret_cast<cplx*>(smem_raw + (size_t)NST * stage_halfs * sizeof(__half));   // [2][8][2][ld/2]: even / odd k split
  const int mv = m <= 2 ? 2 : m <= 4 ? 4 : 8;                                             // rows of a vector buffer
  uint64_t* full = reinterpret_cast<uint64_t*>(vecs + (size_t)2 * mv * ld);               // [NST]
  const size_t plane = (size_t)n * ld, mat = 4 * plane;
  const __half* Pb = Pp + (size_t)b * T * mat;
  cplx* psi_b = p.psi + (size_t)b * (T + 1) * mn;
  cplx* lam_b = p.lam + (size_t)b * (T + 1) * mn;
  const int NC = (n + R - 1) / R;
  const int nsteps = REV ? T - 1 : T;
  const long long nchunks = (long long)nsteps * NC;
  const double pscale = 1.0 / (double)(1 << TC_EU);
  auto vidx = [&](int j, int k) { return (size_t)(2 * j + (k & 1)) * lh + (k >> 1); };

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto prefetch = [&](long long g) {                       // chunk g (sweep order) -> stage g % NST; one thread, 4 bulk copies
    if (g < nchunks) {
      const int step = (int)(g / NC), c = (int)(g - (long long)step * NC);
      const int t = REV ? T - 1 - step : step;
      const int rows = min(R, n - c * R);
      uint64_t* bar = &full[g % NST];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const uint32_t bytes = (uint32_t)(rows * ld * sizeof(__half));
      mbar_expect_tx(bar, 4 * bytes);
      __half* dst = ring + (size_t)(g % NST) * stage_halfs;
      const __half* src = Pb + (size_t)t * mat + (size_t)c * R * ld;
#pragma unroll
      for (int pl = 0; pl < 4; ++pl) bulk_g2s(dst + (size_t)pl * R * ld, src + (size_t)pl * plane, bytes, bar);
    }
  };
  if (tid == 0)
    for (int i = 0; i < NST - 1; ++i) prefetch(i);

  const bool forb = REV && p.reg.has_forbidden && p.fw != nullptr;
  const bool spd = REV && p.reg.has_speed_up != 0;
  const double* sc = p.scal + (size_t)b * 8;
  const double spdfac = REV ? sc[4] : 0.0;
  auto source = [&](int t, int j, int r) -> cplx {         // regulariser sources (core/regularization_functions.py:71-95)
    cplx s = make_double2(0.0, 0.0);
    if (forb && p.dressW) {
      s = p.psid[((size_t)b * (T + 1) + t) * mn + (size_t)j * n + r];
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

  // initial vectors, zero-padded
  for (int e = tid; e < 2 * mv * ld; e += NTH) vecs[e] = make_double2(0.0, 0.0);
  __syncthreads();
  for (int e = tid; e < m * n; e += NTH) {
    const int j = e / n, r = e - j * n;
    cplx v = make_double2(0.0, 0.0);
    if (REV) {                                             // lambda(T) = -(2/m^2) o phi + source(T)
      const double f = -2.0 / ((double)m * (double)m);
      v = cmul(make_double2(sc[0] * f, sc[1] * f), p.phi[(size_t)j * n + r]);
      const cplx s = source(T, j, r);
      v.x += s.x; v.y += s.y;
      lam_b[(size_t)T * mn + (size_t)j * n + r] = v;
    } else {                                               // psi(0) = V is stored as is (:233-234); the chain starts from U0 V
      for (int c = 0; c < n; ++c) {
        const cplx u = p.U0[(size_t)r * n + c], a = p.V[(size_t)j * n + c];
        v.x += u.x * a.x - u.y * a.y; v.y += u.x * a.y + u.y * a.x;
      }
      psi_b[(size_t)j * n + r] = p.V[(size_t)j * n + r];
    }
    vecs[vidx(j, r)] = v;
  }

  int cur = 0;
  long long g = 0;
  if (!REV) {
    // warp = row (rows warp, warp + 16 of the chunk), lane = pair of k; every shared-memory access is conflict-free
    // (P: consecutive 4-byte words; v: consecutive 16-byte elements of the even-k / odd-k halves)
    for (int step = 0; step < nsteps; ++step) {
      const cplx* vc = vecs + (size_t)cur * mv * ld;
      cplx* vn = vecs + (size_t)(cur ^ 1) * mv * ld;
      for (int c = 0; c < NC; ++c, ++g) {
        __syncthreads();                                   // previous chunk's stage is free; for c = 0: v(cur) complete
        if (tid == 0) prefetch(g + NST - 1);
        while (!mbar_try_wait(&full[g % NST], (uint32_t)(g / NST) & 1u)) {}
        const uint32_t* St = reinterpret_cast<const uint32_t*>(ring + (size_t)(g % NST) * stage_halfs);
        const size_t pw = (size_t)R * lh;                  // plane stride in 32-bit words
        // two rows per pass (rows rr and rr + NW of the chunk): the vector elements are read from shared memory once for both
        // rows and 4 MS independent accumulation chains are in flight per lane
#pragma unroll 1
        for (int rr = warp; rr < R; rr += 2 * NW) {
          const int row = c * R + rr;
          if (row >= n) break;                             // warp-uniform
          const bool two = rr + NW < R && row + NW < n;    // warp-uniform
          double a[2 * MS], a2[2 * MS];
#pragma unroll
          for (int i = 0; i < 2 * MS; ++i) { a[i] = 0.0; a2[i] = 0.0; }
          const uint32_t* wr = St + (size_t)rr * lh;
          const uint32_t* wr2 = wr + (two ? (size_t)NW * lh : 0);
          for (int w = lane; w < lh; w += 32) {
            const double2 er = widen2(wr[w], wr[pw + w]);  // Re of k = 2w, 2w+1
            const double2 ei = widen2(wr[2 * pw + w], wr[3 * pw + w]);
            const double2 fr = widen2(wr2[w], wr2[pw + w]);
            const double2 fi = widen2(wr2[2 * pw + w], wr2[3 * pw + w]);
#pragma unroll
            for (int j = 0; j < MS; ++j) {
              const cplx x = vc[(size_t)(2 * j) * lh + w], y = vc[(size_t)(2 * j + 1) * lh + w];
              a[2 * j] = fma(er.x, x.x, a[2 * j]); a[2 * j] = fma(-ei.x, x.y, a[2 * j]);
              a[2 * j] = fma(er.y, y.x, a[2 * j]); a[2 * j] = fma(-ei.y, y.y, a[2 * j]);
              a[2 * j + 1] = fma(er.x, x.y, a[2 * j + 1]); a[2 * j + 1] = fma(ei.x, x.x, a[2 * j + 1]);
              a[2 * j + 1] = fma(er.y, y.y, a[2 * j + 1]); a[2 * j + 1] = fma(ei.y, y.x, a[2 * j + 1]);
              if (two) {
                a2[2 * j] = fma(fr.x, x.x, a2[2 * j]); a2[2 * j] = fma(-fi.x, x.y, a2[2 * j]);
                a2[2 * j] = fma(fr.y, y.x, a2[2 * j]); a2[2 * j] = fma(-fi.y, y.y, a2[2 * j]);
                a2[2 * j + 1] = fma(fr.x, x.y, a2[2 * j + 1]); a2[2 * j + 1] = fma(fi.x, x.x, a2[2 * j + 1]);
                a2[2 * j + 1] = fma(fr.y, y.y, a2[2 * j + 1]); a2[2 * j + 1] = fma(fi.y, y.x, a2[2 * j + 1]);
              }
            }
          }
          // reduce-scatter: at every step the lower / upper half of the live elements stays with (lane & off) == 0 / != 0,
          // so the element index is read off the lane bits from the top
          int e, e2 = 0;
          const double v = reduce_scatter<2 * MS>(a, lane, e) * pscale;
          const double v2 = two ? reduce_scatter<2 * MS>(a2, lane, e2) * pscale : 0.0;
          // after log2(2 MS) halvings the surviving lanes are those with the low (5 - log2(2 MS)) bits arbitrary: let the
          // lane whose low bits are zero write
          constexpr int LOWBITS = 5 - (MS == 8 ? 4 : MS == 4 ? 3 : MS == 2 ? 2 : 1);
          if ((lane & ((1 << LOWBITS) - 1)) == 0) {
            const int j = e >> 1, ri = e & 1;
            if (j < m) {
              double* dv = reinterpret_cast<double*>(&vn[vidx(j, row)]);
              dv[ri] = v;
              double* dp = reinterpret_cast<double*>(&psi_b[(size_t)(step + 1) * mn + (size_t)j * n + row]);
              dp[ri] = v;
              if (two) {
                double* dv2 = reinterpret_cast<double*>(&vn[vidx(j, row + NW)]);
                dv2[ri] = v2;
                double* dp2 = reinterpret_cast<double*>(&psi_b[(size_t)(step + 1) * mn + (size_t)j * n + row + NW]);
                dp2[ri] = v2;
              }
            }
          }
        }
      }
      cur ^= 1;
    }
  } else {
    constexpr int CPN = 8 * NW;                            // column pairs per state pair: 32 / 64 / 128 >= ld / 2
    const int sg = tid / CPN, cp = tid % CPN;              // state pair, column pair (columns 2 cp, 2 cp + 1)
    const bool act = 2 * sg < m && 2 * cp < ld;
    for (int step = 0; step < nsteps; ++step) {
      const int t = T - 1 - step;
      const cplx* vc = vecs + (size_t)cur * mv * ld;
      cplx* vn = vecs + (size_t)(cur ^ 1) * mv * ld;
      double a[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};      // [state][col][re/im]
      for (int c = 0; c < NC; ++c, ++g) {
        __syncthreads();                                   // previous chunk's stage is free (and, for c = 0, v(cur) complete)
        if (tid == 0) prefetch(g + NST - 1);
        while (!mbar_try_wait(&full[g % NST], (uint32_t)(g / NST) & 1u)) {}
        const __half* St = ring + (size_t)(g % NST) * stage_halfs;
        const int rows = min(R, n - c * R);
        if (act) {
          const uint32_t* w0 = reinterpret_cast<const uint32_t*>(St) + cp;
          const size_t pw = (size_t)R * lh;                // plane stride in 32-bit words
          for (int rr = 0; rr < rows; ++rr) {
            const uint32_t* wr = w0 + (size_t)rr * lh;
            const double2 er = widen2(wr[0], wr[pw]);      // Re of columns 2cp, 2cp+1
            const double2 ei = widen2(wr[2 * pw], wr[3 * pw]);
            const int r = c * R + rr;
            const cplx x = vc[vidx(2 * sg, r)], y = vc[vidx(2 * sg + 1, r)];
            // conj(e) * lambda = (er x.x + ei x.y) + i (er x.y - ei x.x)
            a[0][0][0] = fma(er.x, x.x, a[0][0][0]); a[0][0][0] = fma(ei.x, x.y, a[0][0][0]);
            a[0][0][1] = fma(er.x, x.y, a[0][0][1]); a[0][0][1] = fma(-ei.x, x.x, a[0][0][1]);
            a[0][1][0] = fma(er.y, x.x, a[0][1][0]); a[0][1][0] = fma(ei.y, x.y, a[0][1][0]);
            a[0][1][1] = fma(er.y, x.y, a[0][1][1]); a[0][1][1] = fma(-ei.y, x.x, a[0][1][1]);
            a[1][0][0] = fma(er.x, y.x, a[1][0][0]); a[1][0][0] = fma(ei.x, y.y, a[1][0][0]);
            a[1][0][1] = fma(er.x, y.y, a[1][0][1]); a[1][0][1] = fma(-ei.x, y.x, a[1][0][1]);
            a[1][1][0] = fma(er.y, y.x, a[1][1][0]); a[1][1][0] = fma(ei.y, y.y, a[1][1][0]);
            a[1][1][1] = fma(er.y, y.y, a[1][1][1]); a[1][1][1] = fma(-ei.y, y.x, a[1][1][1]);
          }
        }
      }
      if (act) {
#pragma unroll
        for (int js = 0; js < 2; ++js)
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int j = 2 * sg + js, col = 2 * cp + cc;
            if (j < m && col < n) {
              cplx v = make_double2(a[js][cc][0] * pscale, a[js][cc][1] * pscale);
              const cplx s = source(t, j, col);
              v.x += s.x; v.y += s.y;
              vn[vidx(j, col)] = v;
              lam_b[(size_t)t * mn + (size_t)j * n + col] = v;
            }
          }
      }
      cur ^= 1;
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// DMMA form of the sweeps.  A step is a GEMM with the states as N = 8 columns (m <= 8, zero-padded):
//   forward : Psi(t+1)[row][j] = sum_k P[row][k] Psi(t)[k][j]          A = P rows of the chunk (fp16 pairs widened on the fly)
//   reverse : Lam(t)[c][j]     = sum_r conj(P[r][c]) Lam(t+1)[r][j]    A = P^H, read transposed from the same row-major chunk
// on mma.sync.m8n8k4.f64 with the 3M complex product (as qoc_mma_f64.cu).  The DFMA form above issues ~4 instructions per
// useful FMA (shuffle reduce-scatter, vector re-reads per row, idle lanes when n/2 < 32); here a warp owns 8 x 8 output blocks.
// K is permuted so that a lane's four consecutive k-steps read four consecutive columns of P (one 8-byte load per plane):
// in the group of 16 k-values kk, DMMA step i of lane quarter-index q uses k = 16 kk + 4 q + i for A and B alike.
// State vectors: vec[k][8] complex (k-major), slot of state j stored at j ^ (((k >> 2) & 3) << 1): the B-fragment loads
// (lanes (q, g) read vec[16 kk + 4 q + i][g]) are then conflict-free.
//   forward: the chunk has R rows = R/8 row blocks; with fewer row blocks than warps the k-groups are split over
//            KS = NW / (R/8) warps per block and the partial blocks meet in shared memory;
//   reverse: a warp owns column blocks (all of them stay in registers over the chunks of a step).
// ------------------------------------------------------------------------------------------------------------------
DEVINL void dmma884(double& d0, double& d1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
DEVINL int vslot(int k, int j) { return k * 8 + (j ^ (((k >> 2) & 3) << 1)); }

struct MmaSweepShape { int nw, R, nst; size_t stage_halfs, smem; };

MmaSweepShape mma_sweep_shape(int n, int B, int sms) {
  MmaSweepShape s;
  const int ld = tc_ld(n);
  s.nw = n <= 64 ? 4 : n <= 128 ? 8 : 16;
  const size_t stage_target = n <= 64 ? 10 * 1024 : n <= 128 ? 20 * 1024 : 58 * 1024;
  int r = (int)(stage_target / (8 * (size_t)ld)) / 8 * 8;
  if (r < 8) r = 8;
  const int rfull = (n + 7) / 8 * 8;
  s.R = r < rfull ? r : rfull;
  s.stage_halfs = (size_t)4 * s.R * ld;
  const size_t vec = (size_t)2 * ld * 8 * sizeof(cplx);
  const size_t fixed = vec + (size_t)s.nw * 64 * sizeof(cplx) + 64 + 128;
  size_t budget = n <= 64 ? 40 * 1024 : n <= 128 ? 100 * 1024 : 220 * 1024;
  if (B > 0 && sms > 0) {
    const int per_sm = (B + sms - 1) / sms;
    if (per_sm >= 2 && per_sm <= 12) {
      const size_t fit = (size_t)227 * 1024 / per_sm - 1024 - 256;
      if (fit < budget && fit >= fixed + 2 * s.stage_halfs * sizeof(__half)) budget = fit;
    }
  }
  int nst = budget > fixed ? (int)((budget - fixed) / (s.stage_halfs * sizeof(__half))) : 0;
  if (nst > 8) nst = 8;
  if (nst < 2 && fixed + 2 * s.stage_halfs * sizeof(__half) <= 220 * 1024) nst = 2;
  s.nst = nst;
  s.smem = fixed + (size_t)(nst > 0 ? nst : 0) * s.stage_halfs * sizeof(__half);
  return s;
}

template <bool REV, int NW>
__global__ void __launch_bounds__(32 * NW) k_plane_sweep_mma(QocParams p, const __half* __restrict__ Pp, int NST, int R) {
  constexpr int NTH = 32 * NW;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int n = p.n, m = p.m, T = p.T, mn = m * n;
  const int ld = tc_ld(n);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q4 = lane & 3;
  const int b = blockIdx.x;
  const size_t stage_halfs = (size_t)4 * R * ld;
  __half* ring = reinterpret_cast<__half*>(smem_raw);                                    // [NST][4][R][ld]
  cplx* vecs = reinterpret_cast<cplx*>(smem_raw + (size_t)NST * stage_halfs * sizeof(__half));   // [2][ld][8]
  cplx* part = vecs + (size_t)2 * ld * 8;                                                // [NW][64] partial blocks (forward)
  uint64_t* full = reinterpret_cast<uint64_t*>(part + (size_t)NW * 64);                  // [NST]
  const size_t plane = (size_t)n * ld, mat = 4 * plane;
  const __half* Pb = Pp + (size_t)b * T * mat;
  cplx* psi_b = p.psi + (size_t)b * (T + 1) * mn;
  cplx* lam_b = p.lam + (size_t)b * (T + 1) * mn;
  const int NC = (n + R - 1) / R;
  const int nsteps = REV ? T - 1 : T;
  const long long nchunks = (long long)nsteps * NC;
  const double pscale = 1.0 / (double)(1 << TC_EU);
  const int KG = ld >> 4;                                 // groups of 16 k-values

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto prefetch = [&](long long gi) {
    if (gi < nchunks) {
      const int step = (int)(gi / NC), c = (int)(gi - (long long)step * NC);
      const int t = REV ? T - 1 - step : step;
      const int rows = min(R, n - c * R);
      uint64_t* bar = &full[gi % NST];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const uint32_t bytes = (uint32_t)(rows * ld * sizeof(__half));
      mbar_expect_tx(bar, 4 * bytes);
      __half* dst = ring + (size_t)(gi % NST) * stage_halfs;
      const __half* src = Pb + (size_t)t * mat + (size_t)c * R * ld;
#pragma unroll
      for (int pl = 0; pl < 4; ++pl) bulk_g2s(dst + (size_t)pl * R * ld, src + (size_t)pl * plane, bytes, bar);
    }
  };
  const bool forb = REV && p.reg.has_forbidden && p.fw != nullptr;
  const bool spd = REV && p.reg.has_speed_up != 0;
  const double* sc = p.scal + (size_t)b * 8;
  const double spdfac = REV ? sc[4] : 0.0;
  auto source = [&](int t, int j, int r) -> cplx {         // regulariser sources (core/regularization_functions.py:71-95)
    cplx s = make_double2(0.0, 0.0);
    if (forb && p.dressW) {
      s = p.psid[((size_t)b * (T + 1) + t) * mn + (size_t)j * n + r];
    } else if (forb) {
      const cplx x = psi_b[(size_t)t * mn + (size_t)j * n + r];
      const double c = p.fw[r] / (double)T * 2.0 * (x.x * x.x + x.y * x.y);
      s.x = c * x.x; s.y = c * x.y;
    }
    if (spd) {
      const cplx qq = cmul(p.ot[(size_t)b * (T + 1) + t], p.phi[(size_t)j * n + r]);
      s.x += spdfac * qq.x; s.y += spdfac * qq.y;
    }
    return s;
  };

  // stage rows beyond the matrix (last chunk) and the stage tails are read as A operands of rows / columns that are
  // never stored: make them finite once
  for (size_t e = tid; e < (size_t)NST * stage_halfs / 2; e += NTH) reinterpret_cast<uint32_t*>(ring)[e] = 0u;
  for (int e = tid; e < 2 * ld * 8; e += NTH) vecs[e] = make_double2(0.0, 0.0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0)                                            // only now: the bulk copies must not race with the zero fill
    for (int i = 0; i < NST - 1; ++i) prefetch(i);
  for (int e = tid; e < m * n; e += NTH) {
    const int j = e / n, r = e - j * n;
    cplx v = make_double2(0.0, 0.0);
    if (REV) {                                             // lambda(T) = -(2/m^2) o phi + source(T)
      const double f = -2.0 / ((double)m * (double)m);
      v = cmul(make_double2(sc[0] * f, sc[1] * f), p.phi[(size_t)j * n + r]);
      const cplx s = source(T, j, r);
      v.x += s.x; v.y += s.y;
      lam_b[(size_t)T * mn + (size_t)j * n + r] = v;
    } else {                                               // psi(0) = V is stored as is (:233-234); the chain starts from U0 V
      for (int c = 0; c < n; ++c) {
        const cplx u = p.U0[(size_t)r * n + c], a = p.V[(size_t)j * n + c];
        v.x += u.x * a.x - u.y * a.y; v.y += u.x * a.y + u.y * a.x;
      }
      psi_b[(size_t)j * n + r] = p.V[(size_t)j * n + r];
    }
    vecs[vslot(r, j)] = v;
  }

  // four consecutive columns (k = c4 .. c4+3) of one row of the chunk: Re / Im as doubles (stored units)
  auto load4 = [&](const __half* St, int rl, int c4, double (&ar)[4], double (&ai)[4]) {
    const size_t ps = (size_t)R * ld;
    const __half* pr = St + (size_t)rl * ld + c4;
    const uint2 a0 = *reinterpret_cast<const uint2*>(pr), a1 = *reinterpret_cast<const uint2*>(pr + ps);
    const uint2 b0 = *reinterpret_cast<const uint2*>(pr + 2 * ps), b1 = *reinterpret_cast<const uint2*>(pr + 3 * ps);
    const double2 r01 = widen2(a0.x, a1.x), r23 = widen2(a0.y, a1.y), i01 = widen2(b0.x, b1.x), i23 = widen2(b0.y, b1.y);
    ar[0] = r01.x; ar[1] = r01.y; ar[2] = r23.x; ar[3] = r23.y;
    ai[0] = i01.x; ai[1] = i01.y; ai[2] = i23.x; ai[3] = i23.y;
  };

  int cur = 0;
  long long gi = 0;
  if (!REV) {
    for (int step = 0; step < nsteps; ++step) {
      const cplx* vc = vecs + (size_t)cur * ld * 8;
      cplx* vn = vecs + (size_t)(cur ^ 1) * ld * 8;
      for (int c = 0; c < NC; ++c, ++gi) {
        __syncthreads();                                   // previous chunk's stage and partial blocks are free; c = 0: v(cur) complete
        if (tid == 0) prefetch(gi + NST - 1);
        while (!mbar_try_wait(&full[gi % NST], (uint32_t)(gi / NST) & 1u)) {}
        const __half* St = ring + (size_t)(gi % NST) * stage_halfs;
        const int rows = min(R, n - c * R);
        const int nblk = (rows + 7) >> 3;
        const int KS = max(1, NW / nblk);                  // warps per row block (k-groups dealt round-robin)
        const int rb = warp % nblk, ksl = warp / nblk;
        const bool active = ksl < KS;
        double cr0 = 0.0, cr1 = 0.0, ci0 = 0.0, ci1 = 0.0, t0 = 0.0, t1 = 0.0;
        if (active) {
          const int rl = min(rb * 8 + g, R - 1);
          for (int kk = ksl; kk < KG; kk += KS) {
            double ar[4], ai[4];
            load4(St, rl, 16 * kk + 4 * q4, ar, ai);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int k = 16 * kk + 4 * q4 + i;
              const cplx bv = vc[vslot(k, g)];
              dmma884(cr0, cr1, ar[i], bv.x);
              dmma884(t0, t1, ai[i], bv.y);
              dmma884(ci0, ci1, ar[i] + ai[i], bv.x + bv.y);
            }
          }
        }
        // C fragment: rows 8 rb + g, states 2 q4, 2 q4 + 1
        cplx o0 = make_double2((cr0 - t0) * pscale, (ci0 - cr0 - t0) * pscale);
        cplx o1 = make_double2((cr1 - t1) * pscale, (ci1 - cr1 - t1) * pscale);
        if (KS > 1) {
          if (active) { part[(size_t)warp * 64 + g * 8 + 2 * q4] = o0; part[(size_t)warp * 64 + g * 8 + 2 * q4 + 1] = o1; }
          __syncthreads();
          for (int e = tid; e < nblk * 64; e += NTH) {
            const int bb = e >> 6, rr = (e >> 3) & 7, j = e & 7;
            cplx v = make_double2(0.0, 0.0);
            for (int s2 = 0; s2 < KS; ++s2) { const cplx x = part[(size_t)(s2 * nblk + bb) * 64 + rr * 8 + j]; v.x += x.x; v.y += x.y; }
            const int row = c * R + bb * 8 + rr;
            if (row < n) {
              vn[vslot(row, j)] = v;
              if (j < m) psi_b[(size_t)(step + 1) * mn + (size_t)j * n + row] = v;
            }
          }
        } else if (active) {
          const int row = c * R + rb * 8 + g;
          if (row < n && rb * 8 + g < rows) {
            const int j0 = 2 * q4;
            vn[vslot(row, j0)] = o0; vn[vslot(row, j0 + 1)] = o1;
            if (j0 < m) psi_b[(size_t)(step + 1) * mn + (size_t)j0 * n + row] = o0;
            if (j0 + 1 < m) psi_b[(size_t)(step + 1) * mn + (size_t)(j0 + 1) * n + row] = o1;
          }
        }
      }
      cur ^= 1;
    }
  } else {
    // a warp owns 16 columns (NW x 16 >= ld): lane g loads the two adjacent columns 16 w + 2 g, + 1 of its row with one 4-byte
    // read per plane and feeds two "logical" 8-column blocks e = 0, 1 (block e = columns {16 w + 2 g + e}: a permutation of
    // the output columns, resolved when the results are stored)
    const int colb = 16 * warp + 2 * g;
    const bool wact = 16 * warp < ld;
    for (int step = 0; step < nsteps; ++step) {
      const int t = T - 1 - step;
      const cplx* vc = vecs + (size_t)cur * ld * 8;
      cplx* vn = vecs + (size_t)(cur ^ 1) * ld * 8;
      double cr[2][2], ci[2][2], t2[2][2];
#pragma unroll
      for (int u = 0; u < 2; ++u) cr[u][0] = cr[u][1] = ci[u][0] = ci[u][1] = t2[u][0] = t2[u][1] = 0.0;
      // the regulariser sources of this step (global reads of psi / psid) are requested now and consumed after the GEMM
      cplx src[2][2];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = colb + u, j = 2 * q4 + e;
          src[u][e] = (wact && col < n && j < m && (forb || spd)) ? source(t, j, col) : make_double2(0.0, 0.0);
        }
      for (int c = 0; c < NC; ++c, ++gi) {
        __syncthreads();                                   // previous chunk's stage is free (and, for c = 0, v(cur) complete)
        if (tid == 0) prefetch(gi + NST - 1);
        while (!mbar_try_wait(&full[gi % NST], (uint32_t)(gi / NST) & 1u)) {}
        const __half* St = ring + (size_t)(gi % NST) * stage_halfs;
        const size_t ps = (size_t)R * ld;
        const int rows = min(R, n - c * R);
        const int kq = (rows + 3) >> 2;                    // k-steps of this chunk: k = row of P = index of the state vector
        if (wact) {
          for (int ks = 0; ks < kq; ++ks) {
            const int rl = 4 * ks + q4;                    // row inside the chunk
            double2 ar = make_double2(0.0, 0.0), ai = make_double2(0.0, 0.0);
            if (rl < rows) {
              const __half* pr = St + (size_t)rl * ld + colb;
              ar = widen2(*reinterpret_cast<const uint32_t*>(pr), *reinterpret_cast<const uint32_t*>(pr + ps));
              ai = widen2(*reinterpret_cast<const uint32_t*>(pr + 2 * ps), *reinterpret_cast<const uint32_t*>(pr + 3 * ps));
              ai.x = -ai.x; ai.y = -ai.y;                  // conj
            }
            const cplx bv = vc[vslot(min(c * R + rl, ld - 1), g)];
            const double sb = bv.x + bv.y;
            dmma884(cr[0][0], cr[0][1], ar.x, bv.x);
            dmma884(t2[0][0], t2[0][1], ai.x, bv.y);
            dmma884(ci[0][0], ci[0][1], ar.x + ai.x, sb);
            dmma884(cr[1][0], cr[1][1], ar.y, bv.x);
            dmma884(t2[1][0], t2[1][1], ai.y, bv.y);
            dmma884(ci[1][0], ci[1][1], ar.y + ai.y, sb);
          }
        }
      }
      if (wact) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int col = colb + u;
          if (col < n) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int j = 2 * q4 + e;
              cplx v = make_double2((cr[u][e] - t2[u][e]) * pscale, (ci[u][e] - cr[u][e] - t2[u][e]) * pscale);
              if (j < m) {
                v.x += src[u][e].x; v.y += src[u][e].y;
                lam_b[(size_t)t * mn + (size_t)j * n + col] = v;
              }
              vn[vslot(col, j)] = v;
            }
          }
        }
      }
      cur ^= 1;
    }
  }
}

}  // namespace

bool qoc_plane_sweep_supported(int n, int m) {
  if (n > TC_MAX_N || m > 8) return false;
  return plane_sweep_shape(n, m).nst >= 2;
}

template <bool REV, int MS, int NW>
static cudaError_t launch_ps3(const QocParams& p, const void* planes, const PlaneSweepShape& s, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(k_plane_sweep<REV, MS, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_plane_sweep<REV, MS, NW>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  k_plane_sweep<REV, MS, NW><<<p.B, 32 * NW, s.smem, st>>>(p, reinterpret_cast<const __half*>(planes), s.nst, s.R);
  return cudaGetLastError();
}
template <bool REV, int MS>
static cudaError_t launch_ps(const QocParams& p, const void* planes, const PlaneSweepShape& s, cudaStream_t st) {
  if (s.nw == 4) return launch_ps3<REV, MS, 4>(p, planes, s, st);
  if (s.nw == 8) return launch_ps3<REV, MS, 8>(p, planes, s, st);
  return launch_ps3<REV, MS, 16>(p, planes, s, st);
}

template <bool REV, int NW>
static cudaError_t launch_psm(const QocParams& p, const void* planes, const MmaSweepShape& s, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(k_plane_sweep_mma<REV, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_plane_sweep_mma<REV, NW>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  k_plane_sweep_mma<REV, NW><<<p.B, 32 * NW, s.smem, st>>>(p, reinterpret_cast<const __half*>(planes), s.nst, s.R);
  return cudaGetLastError();
}

cudaError_t qoc_launch_plane_sweep(const QocParams& p, const void* planes, int reverse, cudaStream_t st, int64_t* launches) {
  if (!qoc_plane_sweep_supported(p.n, p.m)) return cudaErrorNotSupported;
  ++*launches;
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  if (!(getenv("QOC_B200_SWEEP_MMA") && atoi(getenv("QOC_B200_SWEEP_MMA")) == 0)) {      // DMMA form (default); =0: the DFMA form
    const MmaSweepShape ms = mma_sweep_shape(p.n, p.B, sms);
    if (ms.nst >= 2) {
      if (ms.nw == 4) return reverse ? launch_psm<true, 4>(p, planes, ms, st) : launch_psm<false, 4>(p, planes, ms, st);
      if (ms.nw == 8) return reverse ? launch_psm<true, 8>(p, planes, ms, st) : launch_psm<false, 8>(p, planes, ms, st);
      return reverse ? launch_psm<true, 16>(p, planes, ms, st) : launch_psm<false, 16>(p, planes, ms, st);
    }
  }
  const PlaneSweepShape s = plane_sweep_shape(p.n, p.m, p.B, sms);
  if (reverse) return launch_ps<true, 8>(p, planes, s, st);
  if (p.m <= 2) return launch_ps<false, 2>(p, planes, s, st);
  if (p.m <= 4) return launch_ps<false, 4>(p, planes, s, st);
  return launch_ps<false, 8>(p, planes, s, st);
}
