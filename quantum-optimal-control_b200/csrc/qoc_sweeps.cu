// Sweep / reduction kernels of the GRAPE hot path (sm_100a); the GEMM-shaped stages (propagator
// exponentials and the forward chain) live in qoc_mma_f64.cu.
//
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
// k_dress: forbid_dressed support, a warp per (b, t, j) state vector.
//   phase 0: psid = W psi                                  (regularization_functions.py:79)
//   phase 1: psid <- W^dagger ( (fw/T) * 2 |psid|^2 psid )  = the costate source of the dressed term
// ---------------------------------------------------------------------------------------------
__global__ void k_dress(QocParams p, int phase) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = p.n, m = p.m, T = p.T;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  cplx* vec = reinterpret_cast<cplx*>(smem_raw) + (size_t)wib * n;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long items = (long long)p.B * (T + 1) * m;
  for (long long item = warp; item < items; item += nwarps) {
    const cplx* src = (phase == 0 ? p.psi : p.psid) + (size_t)item * n;
    cplx* dst = p.psid + (size_t)item * n;
    for (int c = lane; c < n; c += 32) {
      cplx x = src[c];
      if (phase == 1) {
        const double pop = x.x * x.x + x.y * x.y;
        const double f = p.fw[c] / (double)T * 2.0 * pop;
        x.x *= f; x.y *= f;
      }
      vec[c] = x;
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      double ax = 0.0, ay = 0.0;
      for (int c = 0; c < n; ++c) {
        const cplx v = vec[c];
        if (phase == 0) { const cplx w = p.dressW[(size_t)i * n + c]; ax += w.x * v.x - w.y * v.y; ay += w.x * v.y + w.y * v.x; }
        else { const cplx w = p.dressW[(size_t)c * n + i]; ax += w.x * v.x + w.y * v.y; ay += w.x * v.y - w.y * v.x; }
      }
      dst[i] = make_double2(ax, ay);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// k_fwd_reduce: one CTA per instance; a warp per time step.
// ---------------------------------------------------------------------------------------------
__global__ void k_fwd_reduce(QocParams p) {
  __shared__ double red[4 * 32];
  const int b = blockIdx.x, n = p.n, m = p.m, T = p.T;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const cplx* psi_b = p.psi + (size_t)b * (T + 1) * m * n;
  const cplx* psif_b = (p.dressW ? p.psid : p.psi) + (size_t)b * (T + 1) * m * n;   // states the forbidden term looks at
  const int mn = m * n;
  const bool forb = p.reg.has_forbidden && p.fw != nullptr;
  const bool spd = p.reg.has_speed_up != 0;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};       // o.re o.im forb S
  double nrm = 0.0;                           // state transfer: sum_j |psi_j(T)|^2
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
        const cplx y = p.dressW ? psif_b[(size_t)t * mn + idx] : x;
        const double pop = y.x * y.x + y.y * y.y;
        f += p.fw[idx % n] * pop * pop;
      }
      if (p.state_transfer && t == T) nrm += x.x * x.x + x.y * x.y;
    }
    orr = warp_sum(orr); oi = warp_sum(oi);
    acc[2] += f;
    if (lane == 0) {
      if (spd) { p.ot[(size_t)b * (T + 1) + t] = make_double2(orr, oi); acc[3] += orr * orr + oi * oi; }
      if (t == T) { acc[0] = orr; acc[1] = oi; }
    }
  }
  block_sum<4>(acc, red);
  double nv[1] = {nrm};
  if (p.state_transfer) block_sum<1>(nv, red);
  if (threadIdx.x == 0) {
    const double m2 = (double)m * (double)m;
    if (p.state_transfer) p.scal[(size_t)b * 8 + 5] = nv[0] * nv[0] / m2;      // tensorflow_state.py:335
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
// PT = double2: interleaved [n][n] complex (fp64 path).  PT = float: planar padded [2][32][32] fp32
// (tcgen05 path, csrc/qoc_tc_tf32.cu).
template <typename PT>
struct PLay;
template <>
struct PLay<double2> {
  DEVINL static int item_elems(int n) { return n * n; }
  DEVINL static cplx load(const double2* s, int r, int i, int n) { return s[r * n + i]; }
};
template <>
struct PLay<float> {
  DEVINL static int item_elems(int) { return 2 * 32 * 32; }
  DEVINL static cplx load(const float* s, int r, int i, int) {
    return make_double2((double)s[r * 32 + i], (double)s[1024 + r * 32 + i]);
  }
};

template <typename PT>
__global__ void k_costate(QocParams p, int parts, int nbuf, int mc) {
  // grid = (B, ceil(m/mc)): the m costate columns are independent chains, a CTA sweeps mc of them
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = p.n, m = p.m, T = p.T, mn = m * n;
  const int nn = PLay<PT>::item_elems(n);        // elements of PT per cached propagator
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
    if (forb && p.dressW) {
      const cplx d = p.psid[((size_t)b * (T + 1) + t) * mn + (size_t)j0 * n + idx];     // precomputed by k_dress phase 1
      s.x += d.x; s.y += d.y;
    } else if (forb) {
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
      for (int r = r0; r < r1; ++r) cfma_conj(acc, PLay<PT>::load(Pt, r, i, n), lam_s[j * n + r]);
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
cudaError_t qoc_launch_dress(const QocParams& p, int phase, cudaStream_t st, int64_t* launches) {
  ++*launches;
  const long long items = (long long)p.B * (p.T + 1) * p.m;
  long long blocks = (items + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const size_t smem = (size_t)8 * p.n * sizeof(cplx);
  cudaError_t e = cudaFuncSetAttribute(k_dress, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_dress<<<(unsigned)blocks, 256, smem, st>>>(p, phase);
  return cudaGetLastError();
}

cudaError_t qoc_launch_fwd_reduce(const QocParams& p, cudaStream_t st, int64_t* launches) {
  ++*launches;
  cudaFuncSetAttribute(k_fwd_reduce, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  k_fwd_reduce<<<p.B, 256, 0, st>>>(p);
  return cudaGetLastError();
}

cudaError_t qoc_launch_costate(const QocParams& p, int p_is_f32, cudaStream_t st, int64_t* launches) {
  ++*launches;
  const int nn = p_is_f32 ? 2 * 32 * 32 : p.n * p.n;
  const size_t psz = p_is_f32 ? sizeof(float) : sizeof(cplx);
  // columns per CTA: all m when small, else chunks of <= 8 columns (more CTAs, less shared memory)
  int mc = p.m;
  if ((size_t)mc * p.n > 512) mc = 512 / p.n > 0 ? 512 / p.n : 1;
  const int ln = mc * p.n;
  const int threads = 256;
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
    e = cudaFuncSetAttribute(k_costate<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_costate<float><<<grid, threads, smem, st>>>(p, parts, nbuf, mc);
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
