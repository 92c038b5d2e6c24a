// FP64 kernels of the GRAPE hot path (sm_100a).
//
//   k_expm      : (b,t) -> P_t = (sum_{j<=p} H^j/j!)^(2^s),  H = (A_0 + sum_k u_k(t) A_k)/2^s
//                 replaces get_matexp / matexp_op   (core/tensorflow_state.py:25-46,70-75)
//   k_chain     : b -> X_t = P_t X_{t-1}, psi_j(t+1) = X_t V_j, U_final, unitary_scale
//                 replaces init_tf_propagator / init_tf_inter_vectors (:204-242)
//   k_fwd_reduce: b -> overlap, loss, forbidden / speed_up values (:282-321,:323-329,
//                 core/regularization_functions.py:71-95)
//   k_costate   : b -> lambda(t) = P_t^dagger lambda(t+1) + sources(t); the reverse sweep TF autodiff
//                 performs through :214-220 (restricted to the m concerned columns)
//   k_grad      : (b,t) -> sum_j Re<lambda_j(t+1), A_k psi_j(t+1)>; matexp_op_grad (:49-65)
//   k_finalize  : b -> pulse regularisers (+ analytic gradients), sin/maxA chain rule, grad_squared
//                 (core/regularization_functions.py:15-45, core/tensorflow_state.py:176-178,348-353)
//
// Complex numbers are double2 (re,im). All matrices row-major.
#include "qoc_internal.cuh"
#include <math.h>

#define DEVINL __device__ __forceinline__

DEVINL void cfma(cplx& c, const cplx a, const cplx b) {
  c.x = fma(a.x, b.x, c.x);
  c.x = fma(-a.y, b.y, c.x);
  c.y = fma(a.x, b.y, c.y);
  c.y = fma(a.y, b.x, c.y);
}
// c += conj(a) * b
DEVINL void cfma_conj(cplx& c, const cplx a, const cplx b) {
  c.x = fma(a.x, b.x, c.x);
  c.x = fma(a.y, b.y, c.x);
  c.y = fma(a.x, b.y, c.y);
  c.y = fma(-a.y, b.x, c.y);
}
DEVINL cplx cmul(const cplx a, const cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

DEVINL double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of NV doubles per thread; result valid in thread 0. red must hold NV*32 doubles.
template <int NV>
DEVINL void block_sum(double (&v)[NV], double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[i * 32 + w] = v[i];
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double x = lane < nw ? red[i * 32 + lane] : 0.0;
      v[i] = warp_sum(x);
    }
  }
}

DEVINL void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---------------------------------------------------------------------------------------------
// Register-tiled complex GEMM on shared-memory operands.
// A group of G = (NP/TM)*(NP/TN) threads owns one NP x NP product; thread (ty,tx) owns rows
// ty + i*TY and columns tx + j*TX (interleaved so that b-loads of a warp are contiguous and
// a-loads are broadcasts).  Operand leading dimension LD = NP+1 complex (bank spread).
// ---------------------------------------------------------------------------------------------
template <int NP, int TM, int TN>
struct Tile {
  static constexpr int TY = NP / TM;
  static constexpr int TX = NP / TN;
  static constexpr int G = TX * TY;
  static constexpr int LD = NP + 1;
  static constexpr int MAT = NP * LD;   // complex elements per padded matrix
};

template <int NP, int TM, int TN>
DEVINL void gemm_tile(const cplx* __restrict__ As, const cplx* __restrict__ Bs, cplx (&acc)[TM][TN],
                      int ty, int tx, int kdim) {
  typedef Tile<NP, TM, TN> TL;
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = make_double2(0.0, 0.0);
#pragma unroll 2
  for (int k = 0; k < kdim; ++k) {
    cplx a[TM], b[TN];
#pragma unroll
    for (int i = 0; i < TM; ++i) a[i] = As[(ty + i * TL::TY) * TL::LD + k];
#pragma unroll
    for (int j = 0; j < TN; ++j) b[j] = Bs[k * TL::LD + tx + j * TL::TX];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) cfma(acc[i][j], a[i], b[j]);
  }
}

template <int G>
DEVINL void group_sync() {
  if (G <= 32) __syncwarp(); else __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// k_expm: persistent over (b,t) items.  CTA = GPC groups of G threads; each group has 3 padded
// matrices in shared memory (H, ping, pong) + K weights.
// ---------------------------------------------------------------------------------------------
template <int NP, int TM, int TN>
__global__ void k_expm(QocParams p) {
  typedef Tile<NP, TM, TN> TL;
  constexpr int G = TL::G;
  constexpr int LD = TL::LD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int gpc = blockDim.x / G;                 // groups per CTA
  const int grp = threadIdx.x / G;
  const int gt = threadIdx.x % G;
  const int ty = gt / TL::TX, tx = gt % TL::TX;
  const size_t grp_bytes = (size_t)3 * TL::MAT * sizeof(cplx) + 32 * sizeof(double);
  cplx* Hs = reinterpret_cast<cplx*>(smem_raw + grp * grp_bytes);
  cplx* buf1 = Hs + TL::MAT;
  cplx* buf2 = buf1 + TL::MAT;
  double* wts = reinterpret_cast<double*>(buf2 + TL::MAT);
  const int n = p.n, K = p.K, T = p.T;
  const int nn = n * n;
  const long long items = (long long)p.B * T;
  cplx* Pout = reinterpret_cast<cplx*>(p.P);

  // zero the three buffers once: padding stays zero for the whole kernel
  for (int i = gt; i < 3 * TL::MAT; i += G) Hs[i] = make_double2(0.0, 0.0);
  group_sync<G>();

  for (long long item = (long long)blockIdx.x * gpc + grp; item < items; item += (long long)gridDim.x * gpc) {
    const int b = (int)(item / T), t = (int)(item % T);
    // control amplitudes u_k(t) = maxA_k sin(base) (tensorflow_state.py:176-178), pre-divided by 2^s (:31)
    if (gt < K) wts[gt] = p.maxA[gt] * sin(p.base[((size_t)b * K + gt) * T + t]) * p.inv2s;
    group_sync<G>();
    for (int idx = gt; idx < nn; idx += G) {
      const int r = idx / n, c = idx - r * n;
      cplx v = p.A[idx];
      v.x *= p.inv2s; v.y *= p.inv2s;
      for (int k = 0; k < K; ++k) {
        const cplx a = p.A[(size_t)(k + 1) * nn + idx];
        const double w = wts[k];
        v.x = fma(w, a.x, v.x); v.y = fma(w, a.y, v.y);
      }
      Hs[r * LD + c] = v;
    }
    group_sync<G>();

    // Taylor: S = I + H + sum_{j=2..p} H^j/j!, term_j = H * term_{j-1} / j
    cplx S[TM][TN], C[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int r = ty + i * TL::TY, c = tx + j * TL::TX;
        S[i][j] = Hs[r * LD + c];
        if (r == c && r < n) S[i][j].x += 1.0;
      }
    const cplx* cur = Hs;
    cplx* nxt = buf1;
    for (int j = 2; j <= p.p; ++j) {
      gemm_tile<NP, TM, TN>(Hs, cur, C, ty, tx, n);
      const double inv = 1.0 / (double)j;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int jj = 0; jj < TN; ++jj) {
          C[i][jj].x *= inv; C[i][jj].y *= inv;
          S[i][jj].x += C[i][jj].x; S[i][jj].y += C[i][jj].y;
        }
      if (j < p.p) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int jj = 0; jj < TN; ++jj) nxt[(ty + i * TL::TY) * LD + tx + jj * TL::TX] = C[i][jj];
        group_sync<G>();
        cur = nxt;
        nxt = (nxt == buf1) ? buf2 : buf1;
      }
    }
    // squarings (tensorflow_state.py:43-44)
    if (p.s > 0) {
      cplx* X = nxt;                       // not read by the last Taylor product
      cplx* Y = (X == buf1) ? buf2 : buf1;
      if (p.p < 2) group_sync<G>();
      for (int q = 0; q < p.s; ++q) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int jj = 0; jj < TN; ++jj) X[(ty + i * TL::TY) * LD + tx + jj * TL::TX] = S[i][jj];
        group_sync<G>();
        gemm_tile<NP, TM, TN>(X, X, S, ty, tx, n);
        cplx* tmp = X; X = Y; Y = tmp;
      }
    }
    cplx* dst = Pout + (size_t)item * nn;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int r = ty + i * TL::TY;
#pragma unroll
      for (int jj = 0; jj < TN; ++jj) {
        const int c = tx + jj * TL::TX;
        if (r < n && c < n) dst[r * n + c] = S[i][jj];
      }
    }
    group_sync<G>();
  }
}

// ---------------------------------------------------------------------------------------------
// k_chain: one CTA (= one group) per instance; X resident in shared memory, P_t streamed with
// cp.async (3 buffers, prefetch distance 2).
// ---------------------------------------------------------------------------------------------
template <int NP, int TM, int TN, int NPB, int NXB>
__global__ void k_chain(QocParams p) {
  typedef Tile<NP, TM, TN> TL;
  constexpr int G = TL::G;
  constexpr int LD = TL::LD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* Pb = reinterpret_cast<cplx*>(smem_raw);           // [NPB][MAT]
  cplx* Xb = Pb + NPB * TL::MAT;                          // [NXB][MAT]
  __shared__ double red[64];
  const int tid = threadIdx.x;
  const int ty = tid / TL::TX, tx = tid % TL::TX;
  const int n = p.n, T = p.T, m = p.m;
  const int nn = n * n;
  const int b = blockIdx.x;
  const cplx* Pg = reinterpret_cast<const cplx*>(p.P) + (size_t)b * T * nn;
  cplx* psi_b = p.psi + (size_t)b * (T + 1) * m * n;

  for (int i = tid; i < (NPB + NXB) * TL::MAT; i += G) Pb[i] = make_double2(0.0, 0.0);
  __syncthreads();
  for (int idx = tid; idx < nn; idx += G) {
    const int r = idx / n, c = idx - r * n;
    Xb[r * LD + c] = p.U0[idx];
  }
  for (int idx = tid; idx < m * n; idx += G) psi_b[idx] = p.V[idx];     // inter_vecs[0] = V (:233-234)

  auto prefetch = [&](int t) {
    if (t < T) {
      const cplx* src = Pg + (size_t)t * nn;
      cplx* dst = Pb + (t % NPB) * TL::MAT;
      for (int idx = tid; idx < nn; idx += G) {
        const int r = idx / n, c = idx - r * n;
        cp_async16(dst + r * LD + c, src + idx);
      }
    }
    cp_async_commit();
  };
  auto extract = [&](const cplx* X, int t) {     // psi[t][j][i] = (X V)_ij
    cplx* out = psi_b + (size_t)t * m * n;
    if (p.has_cidx) {
      for (int idx = tid; idx < m * n; idx += G) {
        const int j = idx / n, i = idx - j * n;
        out[idx] = X[i * LD + p.cidx[j]];
      }
    } else {
      for (int idx = tid; idx < m * n; idx += G) {
        const int j = idx / n, i = idx - j * n;
        cplx acc = make_double2(0.0, 0.0);
        for (int c = 0; c < n; ++c) cfma(acc, X[i * LD + c], p.V[j * n + c]);
        out[idx] = acc;
      }
    }
  };

#pragma unroll
  for (int i = 0; i < NPB - 1; ++i) prefetch(i);
  for (int t = 0; t < T; ++t) {
    cp_async_wait<NPB - 2>();
    __syncthreads();                                  // P_t landed; X_t complete; step t-1 reads done
    const cplx* Xc = Xb + (NXB == 2 ? (t & 1) : 0) * TL::MAT;
    cplx* Xn = Xb + (NXB == 2 ? ((t + 1) & 1) : 0) * TL::MAT;
    if (t > 0) extract(Xc, t);
    prefetch(t + NPB - 1);
    cplx C[TM][TN];
    gemm_tile<NP, TM, TN>(Pb + (t % NPB) * TL::MAT, Xc, C, ty, tx, n);
    if (NXB == 1) __syncthreads();                    // in-place update: everyone has finished reading X
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int jj = 0; jj < TN; ++jj) Xn[(ty + i * TL::TY) * LD + tx + jj * TL::TX] = C[i][jj];
  }
  cp_async_wait<0>();
  __syncthreads();
  const cplx* Xf = Xb + (NXB == 2 ? (T & 1) : 0) * TL::MAT;
  extract(Xf, T);
  cplx* Uf = p.Ufin + (size_t)b * nn;
  for (int idx = tid; idx < nn; idx += G) {
    const int r = idx / n, c = idx - r * n;
    Uf[idx] = Xf[r * LD + c];
  }
  // unitary_scale = (0.5/n) sum_ab (X^T X)_ab over the real embedding = (1/n) sum_r |sum_c X_rc|^2 (:225)
  double v[1] = {0.0};
  for (int r = tid; r < n; r += G) {
    double sr = 0.0, si = 0.0;
    for (int c = 0; c < n; ++c) { sr += Xf[r * LD + c].x; si += Xf[r * LD + c].y; }
    v[0] += sr * sr + si * si;
  }
  block_sum<1>(v, red);
  if (tid == 0) p.scal[(size_t)b * 8 + 5] = v[0] / (double)n;
}

// ---------------------------------------------------------------------------------------------
// k_fwd_reduce: one CTA per instance; a warp per time step.
// ---------------------------------------------------------------------------------------------
__global__ void k_fwd_reduce(QocParams p) {
  __shared__ double red[4 * 32];
  const int b = blockIdx.x, n = p.n, m = p.m, T = p.T;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const cplx* psi_b = p.psi + (size_t)b * (T + 1) * m * n;
  const int mn = m * n;
  const bool forb = p.reg.has_forbidden && p.fw != nullptr;
  const bool spd = p.reg.has_speed_up != 0;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};       // o.re o.im forb S
  const int t_lo = (forb || spd) ? 0 : T;
  for (int t = t_lo + w; t <= T; t += nw) {
    const cplx* ps = psi_b + (size_t)t * mn;
    double orr = 0.0, oi = 0.0, f = 0.0;
    const bool need_o = spd || t == T;
    for (int idx = lane; idx < mn; idx += 32) {
      const cplx x = ps[idx];
      if (need_o) {                               // <phi|psi> = sum conj(phi) psi (:282-300)
        const cplx ph = p.phi[idx];
        orr += ph.x * x.x + ph.y * x.y;
        oi += ph.x * x.y - ph.y * x.x;
      }
      if (forb) {
        const double pop = x.x * x.x + x.y * x.y;
        f += p.fw[idx % n] * pop * pop;
      }
    }
    orr = warp_sum(orr); oi = warp_sum(oi);
    acc[2] += f;
    if (lane == 0) {
      if (spd) { p.ot[(size_t)b * (T + 1) + t] = make_double2(orr, oi); acc[3] += orr * orr + oi * oi; }
      if (t == T) { acc[0] = orr; acc[1] = oi; }
    }
  }
  block_sum<4>(acc, red);
  if (threadIdx.x == 0) {
    const double m2 = (double)m * (double)m;
    const double loss = 1.0 - (acc[0] * acc[0] + acc[1] * acc[1]) / m2;
    double statereg = 0.0, spdfac = 0.0;
    if (forb) statereg += 0.5 * acc[2] / (double)T;                   // sum_f (c_f/T) * l2_loss(pop)
    if (spd) {
      const double c = p.reg.speed_up / (double)T;
      const double S = acc[3] / m2;
      const double d = (double)(T + 1) - S;
      statereg += c * 0.5 * d * d;
      spdfac = -c * d * (2.0 / m2);
    }
    double* sc = p.scal + (size_t)b * 8;
    sc[0] = acc[0]; sc[1] = acc[1]; sc[2] = loss; sc[3] = statereg; sc[4] = spdfac;
  }
}

// ---------------------------------------------------------------------------------------------
// k_costate: one CTA per instance, reverse sweep.  PT = storage type of P (double2 | float2).
// ---------------------------------------------------------------------------------------------
template <typename PT>
DEVINL cplx load_p(const PT* s, int i);
template <>
DEVINL cplx load_p<double2>(const double2* s, int i) { return s[i]; }
template <>
DEVINL cplx load_p<float2>(const float2* s, int i) { const float2 v = s[i]; return make_double2((double)v.x, (double)v.y); }

template <typename PT>
__global__ void k_costate(QocParams p, int parts, int nbuf, int mc) {
  // grid = (B, ceil(m/mc)): the m costate columns are independent chains, a CTA sweeps mc of them
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = p.n, m = p.m, T = p.T, nn = n * n, mn = m * n;
  const int j0 = blockIdx.y * mc;
  const int mloc = min(mc, m - j0);
  const int ln = mloc * n;                                                       // local outputs
  PT* Pb = reinterpret_cast<PT*>(smem_raw);                                      // [nbuf][nn]
  cplx* lam_s = reinterpret_cast<cplx*>(smem_raw + (size_t)nbuf * nn * sizeof(PT));   // [mc*n]
  cplx* part_s = lam_s + mc * n;                                                 // [parts][mc*n]
  const int tid = threadIdx.x, nt = blockDim.x;
  const int b = blockIdx.x;
  const PT* Pg = reinterpret_cast<const PT*>(p.P) + (size_t)b * T * nn;
  const cplx* psi_b = p.psi + (size_t)b * (T + 1) * mn + (size_t)j0 * n;
  cplx* lam_b = p.lam + (size_t)b * (T + 1) * mn + (size_t)j0 * n;
  const cplx* phi = p.phi + (size_t)j0 * n;
  const double* sc = p.scal + (size_t)b * 8;
  const double o_re = sc[0], o_im = sc[1], spdfac = sc[4];
  const bool forb = p.reg.has_forbidden && p.fw != nullptr;
  const bool spd = p.reg.has_speed_up != 0;
  const double m2 = (double)m * (double)m;
  constexpr int EPV = 16 / sizeof(PT) > 0 ? 16 / sizeof(PT) : 1;                 // elements per 16 B

  auto source = [&](int t, int idx) -> cplx {          // regulariser source at time t for local element idx=(jl,i)
    cplx s = make_double2(0.0, 0.0);
    if (forb) {
      const cplx x = psi_b[(size_t)t * mn + idx];
      const double pop = x.x * x.x + x.y * x.y;
      const double c = p.fw[idx % n] / (double)T * 2.0 * pop;
      s.x += c * x.x; s.y += c * x.y;
    }
    if (spd) {
      const cplx o = p.ot[(size_t)b * (T + 1) + t];
      const cplx q = cmul(o, phi[idx]);
      s.x += spdfac * q.x; s.y += spdfac * q.y;
    }
    return s;
  };
  auto prefetch = [&](int t) {                          // P_t into buffer (t % nbuf); 16-byte chunks
    if (t >= 1) {
      const PT* src = Pg + (size_t)t * nn;
      PT* dst = Pb + (size_t)(t % nbuf) * nn;
      if (sizeof(PT) == 16 || (nn % EPV) == 0) {
        const int chunks = nn / EPV;
        for (int c = tid; c < chunks; c += nt) cp_async16(dst + c * EPV, src + c * EPV);
      } else {
        for (int c = tid; c < nn; c += nt) dst[c] = src[c];
      }
    }
    cp_async_commit();
  };

  // lambda(T) = -(2/m^2) * o * phi + source(T)
  for (int idx = tid; idx < ln; idx += nt) {
    cplx l = cmul(make_double2(o_re, o_im), phi[idx]);
    l.x *= -2.0 / m2; l.y *= -2.0 / m2;
    const cplx s = source(T, idx);
    l.x += s.x; l.y += s.y;
    lam_s[idx] = l;
    lam_b[(size_t)T * mn + idx] = l;
  }
  prefetch(T - 1);
  if (nbuf > 2) prefetch(T - 2);
  const int rchunk = (n + parts - 1) / parts;
  for (int t = T - 1; t >= 1; --t) {
    if (nbuf > 2) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();                                   // P_t landed, lam_s(t+1) complete
    const PT* Pt = Pb + (size_t)(t % nbuf) * nn;
    for (int w = tid; w < parts * ln; w += nt) {       // partial products: work item = (part q, output idx)
      const int q = w / ln, idx = w - q * ln;
      const int j = idx / n, i = idx - j * n;
      const int r0 = q * rchunk, r1 = min(n, r0 + rchunk);
      cplx acc = make_double2(0.0, 0.0);
      for (int r = r0; r < r1; ++r) cfma_conj(acc, load_p<PT>(Pt, r * n + i), lam_s[j * n + r]);
      part_s[w] = acc;
    }
    __syncthreads();
    prefetch(t - (nbuf > 2 ? 2 : 1));                  // buffer of P_{t+1} (nbuf=3) / P_t (nbuf=2) is free now
    for (int idx = tid; idx < ln; idx += nt) {
      cplx l = source(t, idx);
      for (int q = 0; q < parts; ++q) { const cplx v = part_s[q * ln + idx]; l.x += v.x; l.y += v.y; }
      lam_s[idx] = l;
      lam_b[(size_t)t * mn + idx] = l;
    }
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// k_grad: a warp per (b,t); sparse (COO) control operators.
//   gctrl[b][k][t] = sum_j Re< lambda_j(t+1), A_{k+1} psi_j(t+1) >
// ---------------------------------------------------------------------------------------------
__global__ void k_grad(QocParams p) {
  const int n = p.n, m = p.m, T = p.T, K = p.K, mn = m * n;
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long items = (long long)p.B * T;
  for (long long item = warp; item < items; item += nwarps) {
    const int b = (int)(item / T), t = (int)(item % T);
    const size_t off = ((size_t)b * (T + 1) + (t + 1)) * mn;
    const cplx* __restrict__ lam = p.lam + off;
    const cplx* __restrict__ psi = p.psi + off;
    for (int k = 0; k < K; ++k) {
      double g = 0.0;
      const int e0 = p.coo_off[k], e1 = p.coo_off[k + 1];
      for (int e = e0 + lane; e < e1; e += 32) {
        const int r = p.coo_r[e], c = p.coo_c[e];
        cplx w = make_double2(0.0, 0.0);
        for (int j = 0; j < m; ++j) cfma_conj(w, lam[j * n + r], psi[j * n + c]);
        const cplx a = p.coo_v[e];
        g += a.x * w.x - a.y * w.y;
      }
      g = warp_sum(g);
      if (lane == 0) p.gctrl[((size_t)b * K + k) * T + t] = g;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k_finalize: one CTA per instance.
// ---------------------------------------------------------------------------------------------
__global__ void k_finalize(QocParams p) {
  __shared__ double red[2 * 32];
  const int b = blockIdx.x, K = p.K, T = p.T;
  const double* base = p.base + (size_t)b * K * T;
  const double* gc = p.gctrl + (size_t)b * K * T;
  double* grad = p.grad + (size_t)b * K * T;
  const double Tf = (double)T;
  const double dt = p.dt, idt2 = 1.0 / (dt * dt), idt4 = idt2 * idt2;
  const double c_amp = p.reg.has_amplitude ? p.reg.amplitude / Tf : 0.0;
  const double c_env = (p.reg.has_envelope && p.env) ? p.reg.envelope / Tf : 0.0;
  const double c_d1 = p.reg.has_dwdt ? p.reg.dwdt / Tf : 0.0;
  const double c_d2 = p.reg.has_d2wdt2 ? p.reg.d2wdt2 / Tf : 0.0;
  const bool nb = p.reg.has_dwdt || p.reg.has_d2wdt2;
  double acc[2] = {0.0, 0.0};                      // grad^2 sum, pulse-reg value
  // index space (k, t) with t in [0, T+2) so the two tail terms of the d2wdt2 sum are covered
  for (int idx = threadIdx.x; idx < K * (T + 2); idx += blockDim.x) {
    const int k = idx / (T + 2), t = idx - k * (T + 2);
    const double* bk = base + (size_t)k * T;
    auto W = [&](int i) -> double { return (i >= 0 && i < T) ? sin(bk[i]) : 0.0; };
    if (t < T) {
      const double x = bk[t];
      const double w0 = sin(x);
      double gw = p.maxA[k] * gc[(size_t)k * T + t];
      double val = 0.0;
      if (p.reg.has_amplitude) { gw += c_amp * w0; val += c_amp * 0.5 * w0 * w0; }
      if (p.reg.has_envelope && p.env) {
        const double e = p.env[(size_t)k * T + t];
        gw += c_env * e * e * w0; val += c_env * 0.5 * e * e * w0 * w0;
      }
      if (nb) {
        const double wm1 = W(t - 1), wm2 = W(t - 2), wp1 = W(t + 1), wp2 = W(t + 2);
        if (p.reg.has_dwdt) {
          gw += c_d1 * idt2 * (2.0 * w0 - wm1 - wp1);
          const double d = wp1 - w0;                  // (z_{j+1}-z_j) for j = t+2
          val += c_d1 * 0.5 * idt2 * d * d;
          if (t == 0) val += c_d1 * 0.5 * idt2 * w0 * w0;      // j = 1 term
        }
        if (p.reg.has_d2wdt2) {
          gw += c_d2 * idt4 * (wm2 - 4.0 * wm1 + 6.0 * w0 - 4.0 * wp1 + wp2);
          const double e2 = w0 - 2.0 * wm1 + wm2;     // e_j for j = t
          val += c_d2 * 0.5 * idt4 * e2 * e2;
        }
      }
      const double g = gw * cos(x);
      grad[(size_t)k * T + t] = g;
      acc[0] += g * g;
      acc[1] += val;
    } else if (p.reg.has_d2wdt2) {                    // j = T, T+1 tail terms of the d2wdt2 sum
      const double e2 = W(t) - 2.0 * W(t - 1) + W(t - 2);
      acc[1] += c_d2 * 0.5 * idt4 * e2 * e2;
    }
  }
  block_sum<2>(acc, red);
  if (threadIdx.x == 0) {
    const double* sc = p.scal + (size_t)b * 8;
    if (p.loss) p.loss[b] = sc[2];
    if (p.reg_loss) p.reg_loss[b] = sc[2] + sc[3] + acc[1];
    if (p.grad_squared) p.grad_squared[b] = 0.5 * acc[0];
    if (p.unitary_scale) p.unitary_scale[b] = sc[5];
  }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
template <int NP, int TM, int TN>
static cudaError_t launch_expm_cfg(const QocParams& p, int sm_count, cudaStream_t st) {
  typedef Tile<NP, TM, TN> TL;
  constexpr int G = TL::G;
  const int gpc = G >= 64 ? 1 : 128 / G;
  const size_t grp_bytes = (size_t)3 * TL::MAT * sizeof(cplx) + 32 * sizeof(double);
  const size_t smem = grp_bytes * gpc;
  cudaError_t e = cudaFuncSetAttribute(k_expm<NP, TM, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_expm<NP, TM, TN>, G * gpc, smem);
  if (e != cudaSuccess) return e;
  if (occ < 1) occ = 1;
  const long long items = (long long)p.B * p.T;
  long long grid = (long long)sm_count * occ;
  const long long need = (items + gpc - 1) / gpc;
  if (grid > need) grid = need;
  k_expm<NP, TM, TN><<<(unsigned)grid, G * gpc, smem, st>>>(p);
  return cudaGetLastError();
}

cudaError_t qoc_launch_expm_f64(const QocParams& p, int NP, int sm_count, cudaStream_t st, int64_t* launches) {
  ++*launches;
  switch (NP) {
    case 8: return launch_expm_cfg<8, 1, 2>(p, sm_count, st);
    case 16: return launch_expm_cfg<16, 2, 4>(p, sm_count, st);
    case 32: return launch_expm_cfg<32, 4, 4>(p, sm_count, st);
    case 48: return launch_expm_cfg<48, 3, 4>(p, sm_count, st);
    case 64: return launch_expm_cfg<64, 4, 4>(p, sm_count, st);
  }
  return cudaErrorInvalidValue;
}

template <int NP, int TM, int TN, int NPB, int NXB>
static cudaError_t launch_chain_cfg(const QocParams& p, cudaStream_t st) {
  typedef Tile<NP, TM, TN> TL;
  const size_t smem = (size_t)(NPB + NXB) * TL::MAT * sizeof(cplx);
  cudaError_t e = cudaFuncSetAttribute(k_chain<NP, TM, TN, NPB, NXB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_chain<NP, TM, TN, NPB, NXB><<<p.B, TL::G, smem, st>>>(p);
  return cudaGetLastError();
}

cudaError_t qoc_launch_chain_f64(const QocParams& p, int NP, cudaStream_t st, int64_t* launches) {
  ++*launches;
  switch (NP) {
    case 8: return launch_chain_cfg<8, 1, 2, 3, 2>(p, st);
    case 16: return launch_chain_cfg<16, 1, 2, 3, 2>(p, st);
    case 32: return launch_chain_cfg<32, 2, 2, 3, 2>(p, st);
    case 48: return launch_chain_cfg<48, 3, 2, 3, 2>(p, st);
    case 64: return launch_chain_cfg<64, 4, 4, 2, 1>(p, st);   // 3 x 66.5 KB: P double-buffered, X updated in place
  }
  return cudaErrorInvalidValue;
}

cudaError_t qoc_launch_fwd_reduce(const QocParams& p, cudaStream_t st, int64_t* launches) {
  ++*launches;
  k_fwd_reduce<<<p.B, 256, 0, st>>>(p);
  return cudaGetLastError();
}

cudaError_t qoc_launch_costate(const QocParams& p, int p_is_f32, cudaStream_t st, int64_t* launches) {
  ++*launches;
  const int nn = p.n * p.n;
  const size_t psz = p_is_f32 ? sizeof(float2) : sizeof(cplx);
  // columns per CTA: all m when small, else chunks of <= 8 columns (more CTAs, less shared memory)
  int mc = p.m;
  if ((size_t)mc * p.n > 512) mc = 512 / p.n > 0 ? 512 / p.n : 1;
  const int ln = mc * p.n;
  const int threads = ln >= 256 ? 256 : 128;
  int parts = threads / ln;
  if (parts < 1) parts = 1;
  if (parts > 8) parts = 8;
  if (parts > p.n) parts = p.n;
  int nbuf = 3;
  size_t smem = (size_t)nbuf * nn * psz + (size_t)(1 + parts) * ln * sizeof(cplx);
  if (smem > 110 * 1024) { nbuf = 2; smem = (size_t)nbuf * nn * psz + (size_t)(1 + parts) * ln * sizeof(cplx); }
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  const dim3 grid(p.B, (p.m + mc - 1) / mc);
  cudaError_t e;
  if (p_is_f32) {
    e = cudaFuncSetAttribute(k_costate<float2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_costate<float2><<<grid, threads, smem, st>>>(p, parts, nbuf, mc);
  } else {
    e = cudaFuncSetAttribute(k_costate<double2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_costate<double2><<<grid, threads, smem, st>>>(p, parts, nbuf, mc);
  }
  return cudaGetLastError();
}

cudaError_t qoc_launch_grad(const QocParams& p, int sm_count, cudaStream_t st, int64_t* launches) {
  ++*launches;
  const long long items = (long long)p.B * p.T;
  long long blocks = (items + 7) / 8;                 // 8 warps per CTA
  const long long cap = (long long)sm_count * 8;
  if (blocks > cap) blocks = cap;
  k_grad<<<(unsigned)blocks, 256, 0, st>>>(p);
  return cudaGetLastError();
}

cudaError_t qoc_launch_finalize(const QocParams& p, cudaStream_t st, int64_t* launches) {
  ++*launches;
  k_finalize<<<p.B, 256, 0, st>>>(p);
  return cudaGetLastError();
}
