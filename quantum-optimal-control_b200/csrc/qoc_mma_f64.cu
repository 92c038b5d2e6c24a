// FP64 tensor-pipe kernels (mma.sync.m8n8k4.f64 = DMMA) for the two GEMM-shaped stages of the
// GRAPE hot path on sm_100a.  tcgen05 has no f64 kind, so the fp64 configuration is bounded by the
// FP64 pipe (DFMA and DMMA both measure ~36.7-37.0 TFLOP/s on B200, profiles/r01_fp64_peaks.json);
// DMMA is used because per FMA it needs 8x fewer issue slots and ~4x fewer shared-memory
// wavefronts than a register-tiled DFMA loop (profiles/r01_ncu_expm_v1.md).
//
//   k_expm_mma : (b,t) -> P_t = (sum_{j<=p} H^j/j!)^(2^s), H = (A_0 + sum_k u_k(t) A_k)/2^s
//                (get_matexp / matexp_op, core/tensorflow_state.py:25-46,70-75)
//   k_chain_mma: b -> X_t = P_t X_{t-1}; psi_j(t+1) = X_t V_j; U_final; unitary_scale
//                (init_tf_propagator / init_tf_inter_vectors, :204-242)
//   k_segprod  : (b,seg) -> product of 16 consecutive propagators (U_final re-associated; few-state problems)
//   k_costate_mma / k_grad_mma: dense-m reverse sweep and control gradient
//
// Shared-memory matrix layout: NP x NP complex (double2), row-major, NP = n rounded up to 8, no
// padding, 16-byte columns XOR-swizzled by row:  phys(r,c) = r*NP + (c ^ sw(r)),
// sw(r) = 5*(r&1) ^ 2*((r>>1)&3).  With it every access pattern below is conflict-free per
// quarter-warp: A fragments (rows g, 4 consecutive k), B fragments straight from the row-major
// operand (rows k0+q, column c0+g) and the C-fragment stores.
#include "qoc_internal.cuh"
#include <math.h>

#define DEVINL __device__ __forceinline__

namespace {

DEVINL double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

DEVINL void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

DEVINL int sw_mask(int r) { return ((r & 1) * 5) ^ (((r >> 1) & 3) << 1); }

template <int NP>
DEVINL int swz(int r, int c) { return r * NP + (c ^ sw_mask(r)); }

// D(8x8) += A(8x4) * B(4x8), fp64.  a: A[g][q], b: B[q][g], d0/d1: D[g][2q], D[g][2q+1]
DEVINL void dmma(double& d0, double& d1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

template <int NP, int RB, int CB>
struct MT {
  static constexpr int NBLK = NP / 8;
  static constexpr int WR = NBLK / RB;
  static constexpr int WC = NBLK / CB;
  static constexpr int WARPS = WR * WC;
  static constexpr int THREADS = 32 * WARPS;
  static constexpr int MAT = NP * NP;
  static_assert(NBLK % RB == 0 && NBLK % CB == 0, "tile must divide the block grid");
};

// Complex C = A * B on swizzled shared operands; the warp owns block rows [rb0, rb0+RB) and block
// columns [cb0, cb0+CB).  cr/ci[i][j][e] = Re/Im C[8(rb0+i)+g][8(cb0+j)+2q+e].
// 3M (Gauss) complex product: with t1 = Ar Br, t2 = Ai Bi, t3 = (Ar+Ai)(Br+Bi),
//   Re C = t1 - t2,  Im C = t3 - t1 - t2   -> 3 real DMMAs per 8x8x4 block instead of 4.
// The operand sums cost one DADD per loaded fragment element (RB + CB per k-step against
// 3 RB CB DMMAs); the result differs from the 4-product form by O(eps |A||B|) (normwise stable).
// QOC_CMUL_3M=0 restores the 4-product form.
#ifndef QOC_CMUL_3M
#define QOC_CMUL_3M 1
#endif
#ifndef QOC_KUNROLL
#define QOC_KUNROLL 1            // k-steps unrolled in the GEMM loops: C2 expm 5.30 / 5.36 / 5.37 / 5.52 ms at 1 / 2 / 4 / 8
#endif
#if QOC_KUNROLL == 2
#define QOC_PRAGMA_KUNROLL _Pragma("unroll 2")
#elif QOC_KUNROLL == 4
#define QOC_PRAGMA_KUNROLL _Pragma("unroll 4")
#elif QOC_KUNROLL == 8
#define QOC_PRAGMA_KUNROLL _Pragma("unroll 8")
#else
#define QOC_PRAGMA_KUNROLL _Pragma("unroll 1")
#endif

template <int NP, int RB, int CB>
DEVINL void mma_gemm(const cplx* __restrict__ A, const cplx* __restrict__ B, double (&cr)[RB][CB][2],
                     double (&ci)[RB][CB][2], int rb0, int cb0, int ksteps, int lane) {
  const int g = lane >> 2, q = lane & 3;
  int arow[RB], amask[RB];
#if QOC_CMUL_3M
  double t2[RB][CB][2];
#endif
#pragma unroll
  for (int i = 0; i < RB; ++i) {
    const int r = 8 * (rb0 + i) + g;
    arow[i] = r * NP;
    amask[i] = sw_mask(r);
#pragma unroll
    for (int j = 0; j < CB; ++j) {
      cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0;
#if QOC_CMUL_3M
      t2[i][j][0] = t2[i][j][1] = 0.0;
#endif
    }
  }
  const int bc = 8 * cb0 + g;
#if QOC_CMUL_3M
  // operand sums right where the fragments are loaded: a hand-pipelined variant (sums one k-step ahead) measured
  // 2.5 % slower in the isolated loop (tools/gemm_loop_probe.cu) and in the kernel, and needs more registers
QOC_PRAGMA_KUNROLL
  for (int ks = 0; ks < ksteps; ++ks) {
    const int k = 4 * ks + q;
    cplx a[RB], b[CB];
#pragma unroll
    for (int i = 0; i < RB; ++i) a[i] = A[arow[i] + (k ^ amask[i])];
    const int bm = sw_mask(k);
    const cplx* Brow = B + k * NP;
#pragma unroll
    for (int j = 0; j < CB; ++j) b[j] = Brow[(bc + 8 * j) ^ bm];
    double sa[RB], sb[CB];
#pragma unroll
    for (int i = 0; i < RB; ++i) sa[i] = a[i].x + a[i].y;
#pragma unroll
    for (int j = 0; j < CB; ++j) sb[j] = b[j].x + b[j].y;
#pragma unroll
    for (int i = 0; i < RB; ++i)
#pragma unroll
      for (int j = 0; j < CB; ++j) dmma(cr[i][j][0], cr[i][j][1], a[i].x, b[j].x);
#pragma unroll
    for (int i = 0; i < RB; ++i)
#pragma unroll
      for (int j = 0; j < CB; ++j) dmma(t2[i][j][0], t2[i][j][1], a[i].y, b[j].y);
#pragma unroll
    for (int i = 0; i < RB; ++i)
#pragma unroll
      for (int j = 0; j < CB; ++j) dmma(ci[i][j][0], ci[i][j][1], sa[i], sb[j]);
  }
#else
QOC_PRAGMA_KUNROLL
  for (int ks = 0; ks < ksteps; ++ks) {
    const int k = 4 * ks + q;
    cplx a[RB], b[CB];
#pragma unroll
    for (int i = 0; i < RB; ++i) a[i] = A[arow[i] + (k ^ amask[i])];
    const int bm = sw_mask(k);
    const cplx* Brow = B + k * NP;
#pragma unroll
    for (int j = 0; j < CB; ++j) b[j] = Brow[(bc + 8 * j) ^ bm];
#pragma unroll
    for (int i = 0; i < RB; ++i)
#pragma unroll
      for (int j = 0; j < CB; ++j) dmma(cr[i][j][0], cr[i][j][1], a[i].x, b[j].x);
#pragma unroll
    for (int i = 0; i < RB; ++i)
#pragma unroll
      for (int j = 0; j < CB; ++j) dmma(ci[i][j][0], ci[i][j][1], a[i].x, b[j].y);
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const double nai = -a[i].y;
#pragma unroll
      for (int j = 0; j < CB; ++j) dmma(cr[i][j][0], cr[i][j][1], nai, b[j].y);
    }
#pragma unroll
    for (int i = 0; i < RB; ++i)
#pragma unroll
      for (int j = 0; j < CB; ++j) dmma(ci[i][j][0], ci[i][j][1], a[i].y, b[j].x);
  }
#endif
#if QOC_CMUL_3M
#pragma unroll
  for (int i = 0; i < RB; ++i)
#pragma unroll
    for (int j = 0; j < CB; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const double t1 = cr[i][j][e], u = t2[i][j][e];
        cr[i][j][e] = t1 - u;
        ci[i][j][e] = ci[i][j][e] - t1 - u;
      }
#endif
}

template <int NP, int RB, int CB>
DEVINL void store_tile(cplx* __restrict__ M, const double (&cr)[RB][CB][2], const double (&ci)[RB][CB][2], int rb0,
                       int cb0, int lane) {
  const int g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int i = 0; i < RB; ++i) {
    const int r = 8 * (rb0 + i) + g;
    const int m = sw_mask(r);
#pragma unroll
    for (int j = 0; j < CB; ++j) {
      const int c = 8 * (cb0 + j) + 2 * q;
      M[r * NP + (c ^ m)] = make_double2(cr[i][j][0], ci[i][j][0]);
      M[r * NP + ((c + 1) ^ m)] = make_double2(cr[i][j][1], ci[i][j][1]);
    }
  }
}

template <int WARPS>
DEVINL void item_sync() {
  if (WARPS == 1) __syncwarp(); else __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Hermitian-structure evaluation of the Taylor polynomial (used when every generator A_k is
// anti-Hermitian, i.e. the Hamiltonians are Hermitian -- checked on the host):
//   S = E(H2) + H O(H2),  H2 = H H,  E = sum_i c_{2i} H2^i,  O = sum_i c_{2i+1} H2^i.
// H2, its powers, E and O are Hermitian and H O is anti-Hermitian, so every product only needs
// its upper block triangle (NBLK(NBLK+1)/2 of NBLK^2 8x8 blocks); the lower triangle is written by
// mirroring.  Warp W owns block rows W and NBLK-1-W, which balances the triangle exactly.
// ---------------------------------------------------------------------------------------------
template <int NP, int W>
struct HT {
  static constexpr int NBLK = NP / 8;
  static constexpr int RA = W, RB_ = NBLK - 1 - W;
  static constexpr int CA = NBLK - RA;       // blocks (RA, RA..NBLK-1)
  static constexpr int CB_ = NBLK - RB_;     // blocks (RB_, RB_..NBLK-1)
  static constexpr int NB = CA + CB_;        // = NBLK + 1
};

// upper-triangle product: acc[b] over this warp's blocks; b < CA -> (RA, RA+b), else (RB_, RB_+b-CA)
template <int NP, int W>
DEVINL void tri_gemm(const cplx* __restrict__ A, const cplx* __restrict__ B, double (&cr)[HT<NP, W>::NB][2],
                     double (&ci)[HT<NP, W>::NB][2], int ksteps, int lane) {
  typedef HT<NP, W> H_;
  const int g = lane >> 2, q = lane & 3;
  const int ra = 8 * H_::RA + g, rb = 8 * H_::RB_ + g;
  const int ma = sw_mask(ra), mb = sw_mask(rb);
#if QOC_CMUL_3M
  double t2[H_::NB][2];
#pragma unroll
  for (int b = 0; b < H_::NB; ++b) t2[b][0] = t2[b][1] = 0.0;
#endif
#pragma unroll
  for (int b = 0; b < H_::NB; ++b) cr[b][0] = cr[b][1] = ci[b][0] = ci[b][1] = 0.0;
#if QOC_CMUL_3M
QOC_PRAGMA_KUNROLL
  for (int ks = 0; ks < ksteps; ++ks) {
    const int k = 4 * ks + q;
    const cplx a0 = A[ra * NP + (k ^ ma)], a1 = A[rb * NP + (k ^ mb)];
    const int bm = sw_mask(k);
    const cplx* Brow = B + k * NP;
    cplx bv[H_::CA];                           // columns RA..NBLK-1 (superset of RB_..NBLK-1)
#pragma unroll
    for (int j = 0; j < H_::CA; ++j) bv[j] = Brow[(8 * (H_::RA + j) + g) ^ bm];
    const double s0 = a0.x + a0.y, s1 = a1.x + a1.y;
    double sb[H_::CA];
#pragma unroll
    for (int j = 0; j < H_::CA; ++j) sb[j] = bv[j].x + bv[j].y;
#pragma unroll
    for (int j = 0; j < H_::CA; ++j) dmma(cr[j][0], cr[j][1], a0.x, bv[j].x);
#pragma unroll
    for (int j = 0; j < H_::CB_; ++j) dmma(cr[H_::CA + j][0], cr[H_::CA + j][1], a1.x, bv[H_::RB_ - H_::RA + j].x);
#pragma unroll
    for (int j = 0; j < H_::CA; ++j) dmma(t2[j][0], t2[j][1], a0.y, bv[j].y);
#pragma unroll
    for (int j = 0; j < H_::CB_; ++j) dmma(t2[H_::CA + j][0], t2[H_::CA + j][1], a1.y, bv[H_::RB_ - H_::RA + j].y);
#pragma unroll
    for (int j = 0; j < H_::CA; ++j) dmma(ci[j][0], ci[j][1], s0, sb[j]);
#pragma unroll
    for (int j = 0; j < H_::CB_; ++j) dmma(ci[H_::CA + j][0], ci[H_::CA + j][1], s1, sb[H_::RB_ - H_::RA + j]);
  }
#else
QOC_PRAGMA_KUNROLL
  for (int ks = 0; ks < ksteps; ++ks) {
    const int k = 4 * ks + q;
    const cplx a0 = A[ra * NP + (k ^ ma)], a1 = A[rb * NP + (k ^ mb)];
    const int bm = sw_mask(k);
    const cplx* Brow = B + k * NP;
    cplx bv[H_::CA];                           // columns RA..NBLK-1 (superset of RB_..NBLK-1)
#pragma unroll
    for (int j = 0; j < H_::CA; ++j) bv[j] = Brow[(8 * (H_::RA + j) + g) ^ bm];
    const double na0 = -a0.y, na1 = -a1.y;
#pragma unroll
    for (int j = 0; j < H_::CA; ++j) dmma(cr[j][0], cr[j][1], a0.x, bv[j].x);
#pragma unroll
    for (int j = 0; j < H_::CB_; ++j) dmma(cr[H_::CA + j][0], cr[H_::CA + j][1], a1.x, bv[H_::RB_ - H_::RA + j].x);
#pragma unroll
    for (int j = 0; j < H_::CA; ++j) dmma(ci[j][0], ci[j][1], a0.x, bv[j].y);
#pragma unroll
    for (int j = 0; j < H_::CB_; ++j) dmma(ci[H_::CA + j][0], ci[H_::CA + j][1], a1.x, bv[H_::RB_ - H_::RA + j].y);
#pragma unroll
    for (int j = 0; j < H_::CA; ++j) dmma(cr[j][0], cr[j][1], na0, bv[j].y);
#pragma unroll
    for (int j = 0; j < H_::CB_; ++j) dmma(cr[H_::CA + j][0], cr[H_::CA + j][1], na1, bv[H_::RB_ - H_::RA + j].y);
#pragma unroll
    for (int j = 0; j < H_::CA; ++j) dmma(ci[j][0], ci[j][1], a0.y, bv[j].x);
#pragma unroll
    for (int j = 0; j < H_::CB_; ++j) dmma(ci[H_::CA + j][0], ci[H_::CA + j][1], a1.y, bv[H_::RB_ - H_::RA + j].x);
  }
#endif
#if QOC_CMUL_3M
#pragma unroll
  for (int b = 0; b < H_::NB; ++b)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const double t1 = cr[b][e], u = t2[b][e];
      cr[b][e] = t1 - u;
      ci[b][e] = ci[b][e] - t1 - u;
    }
#endif
}

// Store the warp's upper-triangle blocks of  U = X + Y  and mirror the strictly-upper blocks as
// L[c][r] = conj(X[r][c]) * sx + conj(Y[r][c]) * sy  (sx/sy = +1 Hermitian part, -1 anti-Hermitian part).
template <int NP, int W>
DEVINL void tri_store(cplx* __restrict__ M, const double (&xr)[HT<NP, W>::NB][2], const double (&xi)[HT<NP, W>::NB][2],
                      const double (&yr)[HT<NP, W>::NB][2], const double (&yi)[HT<NP, W>::NB][2], bool use_y, double sx,
                      double sy, int lane) {
  typedef HT<NP, W> H_;
  const int g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int b = 0; b < H_::NB; ++b) {
    const int bi = b < H_::CA ? H_::RA : H_::RB_;
    const int bj = b < H_::CA ? H_::RA + b : H_::RB_ + (b - H_::CA);
    const int r = 8 * bi + g;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = 8 * bj + 2 * q + e;
      const double ur = xr[b][e] + (use_y ? yr[b][e] : 0.0), ui = xi[b][e] + (use_y ? yi[b][e] : 0.0);
      M[swz<NP>(r, c)] = make_double2(ur, ui);
      if (bi != bj) {
        const double lr = sx * xr[b][e] + (use_y ? sy * yr[b][e] : 0.0);
        const double li = -(sx * xi[b][e] + (use_y ? sy * yi[b][e] : 0.0));
        M[swz<NP>(c, r)] = make_double2(lr, li);
      }
    }
  }
}

// Taylor phase for warp W; leaves the full S in buf1.  Needs p >= 2.
template <int NP, int W, int WARPS>
DEVINL void herm_taylor(const QocParams& p, const cplx* Hs, cplx* buf1, cplx* buf2, int n, int ksteps, int lane) {
  typedef HT<NP, W> H_;
  constexpr int NB = H_::NB;
  const int g = lane >> 2, q = lane & 3;
  const int pp = p.p;
  auto coef = [&](int j) -> double { return j <= pp ? p.invfact[j] : 0.0; };
  double ar[NB][2], ai[NB][2], er[NB][2], ei[NB][2], orr[NB][2], oi[NB][2];
  tri_gemm<NP, W>(Hs, Hs, ar, ai, ksteps, lane);                  // H2
  {
    const double c0 = coef(0), c1 = coef(1), c2 = coef(2), c3 = coef(3);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const int bi = b < H_::CA ? H_::RA : H_::RB_;
      const int bj = b < H_::CA ? H_::RA + b : H_::RB_ + (b - H_::CA);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int r = 8 * bi + g, c = 8 * bj + 2 * q + e;
        const double id = (r == c && r < n) ? 1.0 : 0.0;
        er[b][e] = c0 * id + c2 * ar[b][e]; ei[b][e] = c2 * ai[b][e];
        orr[b][e] = c1 * id + c3 * ar[b][e]; oi[b][e] = c3 * ai[b][e];
      }
    }
  }
  tri_store<NP, W>(buf1, ar, ai, ar, ai, false, 1.0, 1.0, lane);  // H2 (Hermitian), full
  item_sync<WARPS>();
  const int imax = pp / 2;                                         // highest power of H2 needed
  for (int i = 2; i <= imax; ++i) {
    tri_gemm<NP, W>(buf1, i == 2 ? buf1 : buf2, ar, ai, ksteps, lane);      // H2^i
    const double ce = coef(2 * i), co = coef(2 * i + 1);
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        er[b][e] += ce * ar[b][e]; ei[b][e] += ce * ai[b][e];
        orr[b][e] += co * ar[b][e]; oi[b][e] += co * ai[b][e];
      }
    if (i < imax) {
      item_sync<WARPS>();                                          // everyone has finished reading buf2
      tri_store<NP, W>(buf2, ar, ai, ar, ai, false, 1.0, 1.0, lane);
      item_sync<WARPS>();
    }
  }
  item_sync<WARPS>();
  tri_store<NP, W>(buf2, orr, oi, orr, oi, false, 1.0, 1.0, lane);   // O (Hermitian), full
  item_sync<WARPS>();
  tri_gemm<NP, W>(Hs, buf2, ar, ai, ksteps, lane);                  // A = H O  (anti-Hermitian)
  // S = E + A above the diagonal, conj(E) - conj(A) = (E - A)^H below; buf1 (H2) is no longer read
  tri_store<NP, W>(buf1, er, ei, ar, ai, true, 1.0, -1.0, lane);
  item_sync<WARPS>();
}

// ---------------------------------------------------------------------------------------------
// k_expm_mma: persistent over (b,t) items.  An item is owned by WARPS warps; when WARPS == 1 a
// CTA carries 4 independent items (one per warp), else exactly one.  Shared memory per item:
// H, ping, pong (3 x NP^2 x 16 B) + the control weights of the next 16 items.
// ---------------------------------------------------------------------------------------------
constexpr int EXPM_WB = 16;                      // items whose control weights are evaluated together
__host__ __device__ inline size_t expm_item_bytes(int mat, int K) {
  return (size_t)3 * mat * sizeof(cplx) + (((size_t)EXPM_WB * (K + 1) * sizeof(double) + 15) & ~(size_t)15);
}

// HERM selects the Hermitian-structure Taylor evaluation at compile time (the host checks the generators), so that
// each instantiation carries one path only: smaller code for the 8 warps of an SM that sit in different phases of it.
template <int NP, int RB, int CB, bool HERM>
__global__ void __launch_bounds__(MT<NP, RB, CB>::WARPS == 1 ? 128 : MT<NP, RB, CB>::THREADS)
k_expm_mma(QocParams p) {
  typedef MT<NP, RB, CB> T_;
  constexpr int WARPS = T_::WARPS;
  constexpr int G = T_::THREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ipc = WARPS == 1 ? 4 : 1;                 // items per CTA
  const int slot = WARPS == 1 ? (threadIdx.x >> 5) : 0;
  const int gt = WARPS == 1 ? (threadIdx.x & 31) : threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int warp = WARPS == 1 ? 0 : (threadIdx.x >> 5);
  const int rb0 = (warp / T_::WC) * RB, cb0 = (warp % T_::WC) * CB;
  const int n = p.n, K = p.K, T = p.T, nn = n * n;
  const size_t item_bytes = expm_item_bytes(T_::MAT, K);
  cplx* Hs = reinterpret_cast<cplx*>(smem_raw + slot * item_bytes);
  cplx* buf1 = Hs + T_::MAT;
  cplx* buf2 = buf1 + T_::MAT;
  double* wtab = reinterpret_cast<double*>(buf2 + T_::MAT);       // [EXPM_WB][K+1] control weights of the next items
  const int ksteps = (n + 3) >> 2;
  const long long items = (long long)p.B * T;
  const bool small = items < 0x7fffffffLL;                        // 32-bit index arithmetic (the common case)
  cplx* Pout = reinterpret_cast<cplx*>(p.P);
  const int g = lane >> 2, q = lane & 3;
  const int pat_n = p.pat_n;

  for (int i = gt; i < 3 * T_::MAT; i += G) Hs[i] = make_double2(0.0, 0.0);   // padding stays zero forever
  item_sync<WARPS>();

  const long long stride = (long long)gridDim.x * ipc;
  int it = 0;
  for (long long item = (long long)blockIdx.x * ipc + slot; item < items; item += stride, ++it) {
    int b, t;
    if (small) { b = (int)((unsigned)item / (unsigned)T); t = (int)item - b * T; }
    else { b = (int)(item / T); t = (int)(item % T); }
    // u_k(t)/2^s with u_k = maxA_k sin(base) (tensorflow_state.py:31,176-178); weight 0 = drift.
    // The sines of the next EXPM_WB items of this CTA are evaluated together by all its threads
    // (K of 64 lanes busy per item otherwise: ~1100 cycles of an item's ~50 000).
    const int wslot = it % EXPM_WB;
    if (wslot == 0) {
      for (int idx = gt; idx < EXPM_WB * (K + 1); idx += G) {
        const int j = idx / (K + 1), k = idx - j * (K + 1);
        const long long itj = item + (long long)j * stride;
        double w = p.inv2s;
        if (k > 0) {
          w = 0.0;
          if (itj < items) {
            const int bj = (int)(itj / T), tj = (int)(itj - (long long)bj * T);
            w = p.maxA[k - 1] * sin(p.base[((size_t)bj * K + k - 1) * T + tj]) * p.inv2s;
          }
        }
        wtab[idx] = w;
      }
      item_sync<WARPS>();
    }
    const double* wts = wtab + wslot * (K + 1);
    // H assembly over the union sparsity pattern of A_0..A_K (entries outside it are never written);
    // coefficients are k-major ([K+1][pat_n]: coalesced).  Keeping several entries per thread in flight was
    // tried and lost: the extra live registers spill in this 255-register kernel (5.41 -> 5.59 ms).
    for (int e = gt; e < pat_n; e += G) {
      const int rc = __ldg(p.pat_rc + e);
      double hx = 0.0, hy = 0.0;
      for (int k = 0; k <= K; ++k) {
        const cplx a = __ldg(p.pat_coef + (size_t)k * pat_n + e);
        const double w = wts[k];
        hx = fma(w, a.x, hx); hy = fma(w, a.y, hy);
      }
      Hs[swz<NP>(rc >> 16, rc & 0xffff)] = make_double2(hx, hy);
    }
    item_sync<WARPS>();

    // Taylor polynomial S = sum_{j<=p} H^j/j! (tensorflow_state.py:37-41), evaluated in
    // Paterson-Stockmeyer form with block size 2: S = (..(B_r H2 + B_{r-1}) H2 + ..) H2 + B_0,
    // B_i = c_{2i} I + c_{2i+1} H, c_j = 1/j!  -> 1 + floor(p/2) - [p even] products instead of p-1.
    double sr[RB][CB][2], si[RB][CB][2];
    // Hermitian Hamiltonians (all generators anti-Hermitian): triangle-only products, see herm_taylor
    if constexpr (HERM) {
      if (NP == 32) {
        if (warp == 0) herm_taylor<NP, 0, WARPS>(p, Hs, buf1, buf2, n, ksteps, lane);
        else herm_taylor<NP, (NP == 32 ? 1 : 0), WARPS>(p, Hs, buf1, buf2, n, ksteps, lane);
      } else {
        herm_taylor<NP, 0, WARPS>(p, Hs, buf1, buf2, n, ksteps, lane);
      }
      // S is complete in buf1: squarings read it directly, or it is copied out when s == 0
      cplx* X = buf1;
      cplx* Y = buf2;
      if (p.s == 0) {
#pragma unroll
        for (int i = 0; i < RB; ++i)
#pragma unroll
          for (int j = 0; j < CB; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const cplx v = X[swz<NP>(8 * (rb0 + i) + g, 8 * (cb0 + j) + 2 * q + e)];
              sr[i][j][e] = v.x; si[i][j][e] = v.y;
            }
      }
      for (int s = 0; s < p.s; ++s) {
        if (s > 0) { store_tile<NP, RB, CB>(X, sr, si, rb0, cb0, lane); item_sync<WARPS>(); }
        mma_gemm<NP, RB, CB>(X, X, sr, si, rb0, cb0, ksteps, lane);
        cplx* tmp = X; X = Y; Y = tmp;
      }
    } else {
    auto add_block = [&](double c_id, double c_h, bool init) {     // S (+)= c_id*I + c_h*H on this lane's fragments
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int r = 8 * (rb0 + i) + g;
#pragma unroll
        for (int j = 0; j < CB; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 8 * (cb0 + j) + 2 * q + e;
            const cplx h = Hs[swz<NP>(r, c)];
            const double vr = c_h * h.x + ((r == c && r < n) ? c_id : 0.0), vi = c_h * h.y;
            if (init) { sr[i][j][e] = vr; si[i][j][e] = vi; }
            else { sr[i][j][e] += vr; si[i][j][e] += vi; }
          }
      }
    };
    const int pp = p.p;
    cplx* H2 = buf1;
    cplx* Rb = buf2;
    if (pp >= 2) {
      mma_gemm<NP, RB, CB>(Hs, Hs, sr, si, rb0, cb0, ksteps, lane);
      store_tile<NP, RB, CB>(H2, sr, si, rb0, cb0, lane);
      item_sync<WARPS>();
    }
    int blk;
    if (pp & 1) {                                   // top block B_r = c_{p-1} I + c_p H
      add_block(p.invfact[pp - 1], p.invfact[pp], true);
      blk = pp / 2 - 1;
    } else {                                        // top block is c_p I: first Horner step needs no product
      const double cp = p.invfact[pp];
#pragma unroll
      for (int i = 0; i < RB; ++i)
#pragma unroll
        for (int j = 0; j < CB; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) { sr[i][j][e] *= cp; si[i][j][e] *= cp; }     // sr/si hold H2 here
      add_block(p.invfact[pp - 2], p.invfact[pp - 1], false);
      blk = pp / 2 - 2;
    }
    for (; blk >= 0; --blk) {
      // R <- R * H2 + B_blk.  R is the A operand: a warp only reads the rows it owns when it spans
      // all block columns (WC == 1), so the in-place round trip through Rb needs no CTA barrier then.
      if (T_::WC == 1) __syncwarp(); else __syncthreads();
      store_tile<NP, RB, CB>(Rb, sr, si, rb0, cb0, lane);
      if (T_::WC == 1) __syncwarp(); else __syncthreads();
      mma_gemm<NP, RB, CB>(Rb, H2, sr, si, rb0, cb0, ksteps, lane);
      add_block(p.invfact[2 * blk], p.invfact[2 * blk + 1], false);
    }
    // squarings (tensorflow_state.py:43-44)
    if (p.s > 0) {
      cplx* X = Rb;
      cplx* Y = H2;
      if (T_::WC != 1) __syncthreads();             // other warps may still read Rb rows they do not own
      for (int s = 0; s < p.s; ++s) {
        if (T_::WC == 1 && s == 0) __syncwarp();
        store_tile<NP, RB, CB>(X, sr, si, rb0, cb0, lane);
        item_sync<WARPS>();
        mma_gemm<NP, RB, CB>(X, X, sr, si, rb0, cb0, ksteps, lane);
        cplx* tmp = X; X = Y; Y = tmp;
      }
    }
    }  // general (non-Hermitian) path
    cplx* dst = Pout + (size_t)item * nn;
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int r = 8 * (rb0 + i) + g;
#pragma unroll
      for (int j = 0; j < CB; ++j) {
        const int c = 8 * (cb0 + j) + 2 * q;
        if (r < n && c < n) dst[r * n + c] = make_double2(sr[i][j][0], si[i][j][0]);
        if (r < n && c + 1 < n) dst[r * n + c + 1] = make_double2(sr[i][j][1], si[i][j][1]);
      }
    }
    item_sync<WARPS>();
  }
}

// ---------------------------------------------------------------------------------------------
// k_chain_mma: one CTA per instance; X resident in shared memory (ping-pong), P_t streamed with
// cp.async into NPB swizzled buffers (prefetch distance NPB-1).
// ---------------------------------------------------------------------------------------------
// PF32: propagators are the tcgen05 path's fp32 planar padded [2][32][32] tiles; they are streamed
// raw into a staging ring and widened to double2 in the swizzled operand buffer each step.
template <int NP, int RB, int CB, int NPB, int NXB, bool PF32>
__global__ void __launch_bounds__(MT<NP, RB, CB>::THREADS) k_chain_mma(QocParams p) {
  typedef MT<NP, RB, CB> T_;
  constexpr int G = T_::THREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NPBUF = PF32 ? 1 : NPB;                   // swizzled double2 operand buffers for P
  cplx* Pb = reinterpret_cast<cplx*>(smem_raw);           // [NPBUF][MAT]
  cplx* Xb = Pb + NPBUF * T_::MAT;                        // [NXB][MAT]
  float* Pstage = reinterpret_cast<float*>(Xb + NXB * T_::MAT);   // PF32: [NPB][2048] raw fp32 tiles
  __shared__ double red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rb0 = (warp / T_::WC) * RB, cb0 = (warp % T_::WC) * CB;
  const int n = p.n, T = p.T, m = p.m, nn = n * n;
  const int ksteps = (n + 3) >> 2;
  const int b = blockIdx.x;
  const cplx* Pg = reinterpret_cast<const cplx*>(p.P) + (size_t)b * T * nn;
  const float* Pgf = reinterpret_cast<const float*>(p.P) + (size_t)b * T * 2048;
  cplx* psi_b = p.psi + (size_t)b * (T + 1) * m * n;

  for (int i = tid; i < (NPBUF + NXB) * T_::MAT; i += G) Pb[i] = make_double2(0.0, 0.0);
  __syncthreads();
  for (int idx = tid; idx < nn; idx += G) {
    const int r = idx / n, c = idx - r * n;
    Xb[swz<NP>(r, c)] = p.U0[idx];
  }
  const bool want_psi = !p.chain_no_psi;
  if (want_psi)
    for (int idx = tid; idx < m * n; idx += G) psi_b[idx] = p.V[idx];   // inter_vecs[0] = V (:233-234)

  auto prefetch = [&](int t) {
    if (t < T) {
      if (PF32) {
        const float* src = Pgf + (size_t)t * 2048;
        float* dst = Pstage + (t % NPB) * 2048;
        for (int c = tid; c < 512; c += G) cp_async16(dst + 4 * c, src + 4 * c);
      } else {
        const cplx* src = Pg + (size_t)t * nn;
        cplx* dst = Pb + (t % NPB) * T_::MAT;
        for (int idx = tid; idx < nn; idx += G) {
          const int r = idx / n, c = idx - r * n;
          cp_async16(dst + swz<NP>(r, c), src + idx);
        }
      }
    }
    cp_async_commit();
  };
  auto extract = [&](const cplx* X, int t) {              // psi[t][j][i] = (X V)_ij
    cplx* out = psi_b + (size_t)t * m * n;
    if (p.has_cidx) {
      for (int idx = tid; idx < m * n; idx += G) {
        const int j = idx / n, i = idx - j * n;
        out[idx] = X[swz<NP>(i, p.cidx[j])];
      }
    } else {
      for (int idx = tid; idx < m * n; idx += G) {
        const int j = idx / n, i = idx - j * n;
        double ax = 0.0, ay = 0.0;
        for (int c = 0; c < n; ++c) {
          const cplx x = X[swz<NP>(i, c)], v = p.V[j * n + c];
          ax += x.x * v.x - x.y * v.y; ay += x.x * v.y + x.y * v.x;
        }
        out[idx] = make_double2(ax, ay);
      }
    }
  };

#pragma unroll
  for (int i = 0; i < NPB - 1; ++i) prefetch(i);
  for (int t = 0; t < T; ++t) {
    cp_async_wait<NPB - 2>();
    __syncthreads();                                  // P_t landed; X_t complete; step t-1 reads done
    const cplx* Xc = Xb + (NXB == 2 ? (t & 1) : 0) * T_::MAT;
    cplx* Xn = Xb + (NXB == 2 ? ((t + 1) & 1) : 0) * T_::MAT;
    if (t > 0 && want_psi) extract(Xc, t);
    prefetch(t + NPB - 1);
    if (PF32) {                                       // widen the raw fp32 tile into the swizzled operand buffer
      const float* src = Pstage + (t % NPB) * 2048;
      for (int idx = tid; idx < nn; idx += G) {
        const int r = idx / n, c = idx - r * n;
        Pb[swz<NP>(r, c)] = make_double2((double)src[r * 32 + c], (double)src[1024 + r * 32 + c]);
      }
      __syncthreads();
    }
    double cr[RB][CB][2], ci[RB][CB][2];
    mma_gemm<NP, RB, CB>(Pb + (PF32 ? 0 : (t % NPB)) * T_::MAT, Xc, cr, ci, rb0, cb0, ksteps, lane);
    if (NXB == 1) __syncthreads();                    // in-place update: every warp has finished reading X
    store_tile<NP, RB, CB>(Xn, cr, ci, rb0, cb0, lane);
  }
  cp_async_wait<0>();
  __syncthreads();
  const cplx* Xf = Xb + (NXB == 2 ? (T & 1) : 0) * T_::MAT;
  if (want_psi) extract(Xf, T);
  cplx* Uf = p.Ufin + (size_t)b * nn;
  for (int idx = tid; idx < nn; idx += G) {
    const int r = idx / n, c = idx - r * n;
    Uf[idx] = Xf[swz<NP>(r, c)];
  }
  // unitary_scale = (0.5/n) sum_ab (X^T X)_ab over the real embedding = (1/n) sum_r |sum_c X_rc|^2 (:225)
  double v = 0.0;
  for (int r = tid; r < n; r += G) {
    double sr = 0.0, si = 0.0;
    for (int c = 0; c < n; ++c) { const cplx x = Xf[swz<NP>(r, c)]; sr += x.x; si += x.y; }
    v += sr * sr + si * si;
  }
  v = warp_sum_d(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int w = 0; w < T_::WARPS; ++w) s += red[w];
    p.scal[(size_t)b * 8 + 5] = s / (double)n;
  }
}

// ---------------------------------------------------------------------------------------------
// k_segprod: (b, seg) -> Q_seg = P_{t1-1} ... P_{t0},  t0 = seg*L, t1 = min(T, t0+L).
// U_final = X_T = P_{T-1} ... P_0 U0 and unitary_scale are NOT on the critical path of the loss or
// the gradient when the states are propagated by k_vec_sweep, and only the total product is needed,
// so the T-1 products are re-associated: B*ceil(T/L) independent segment products (this kernel,
// tensor-pipe bound and perfectly balanced, unlike one sequential chain per instance) followed by a
// ceil(T/L)-step chain over the segment matrices (k_chain_mma on the segment buffer).
// One short-lived CTA per segment (so that higher-priority sweep CTAs get SM slots as segments
// retire); X is updated in place, P_t double-buffered with cp.async.
// ---------------------------------------------------------------------------------------------
// PF32: the propagators are the tcgen05 path's fp32 planar padded [2][32][32] tiles; they are staged raw (double
// buffered) and widened into the one swizzled operand buffer each step.
// PFMT 2: the propagators are QOC_F16X2 plane sets [4][n][ld] fp16 (qoc_tc_f16.cuh; scale 2^13): staged raw, widened h0 + h1.
template <int NP, int RB, int CB, int PFMT>
__global__ void __launch_bounds__(MT<NP, RB, CB>::THREADS) k_segprod(QocParams p, int L, int S, cplx* __restrict__ seg_out) {
  constexpr bool PF32 = PFMT != 0;                         // any raw-staged narrow format
  typedef MT<NP, RB, CB> T_;
  constexpr int G = T_::THREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* Xb = reinterpret_cast<cplx*>(smem_raw);            // [MAT]
  cplx* Pb = Xb + T_::MAT;                                 // [2][MAT]   (PF32: [1][MAT] + raw staging [2][2048] floats)
  float* Pstage = reinterpret_cast<float*>(Pb + T_::MAT);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rb0 = (warp / T_::WC) * RB, cb0 = (warp % T_::WC) * CB;
  const int n = p.n, T = p.T, nn = n * n;
  const int ksteps = (n + 3) >> 2;
  // segment-major dispatch: the CTAs of one time segment of ALL instances run together, at the pace at which the
  // forward sweep (all instances in lock step) moves through the same propagators -> the second reader hits L2
  const int seg = blockIdx.x / p.B, b = blockIdx.x - seg * p.B;
  const int t0 = seg * L, len = min(L, T - t0);
  const cplx* Pg = reinterpret_cast<const cplx*>(p.P) + ((size_t)b * T + t0) * nn;
  const int ldh = (n + 15) / 16 * 16;                      // plane-set row pitch (tc_ld)
  const int raw_floats = PFMT == 2 ? 2 * n * ldh : 2048;   // one raw propagator, in 4-byte units
  const float* Pgf = reinterpret_cast<const float*>(p.P) + ((size_t)b * T + t0) * raw_floats;

  for (int i = tid; i < (PFMT ? 2 : 3) * T_::MAT; i += G) Xb[i] = make_double2(0.0, 0.0);   // padding rows / columns stay zero
  __syncthreads();
  const int dr = G / n, dc = G - dr * n, r_first = tid / n, c_first = tid - r_first * n;
  auto fetch = [&](int l, cplx* dst) {                     // P_{t0+l} -> swizzled operand buffer (PF32: raw staging)
    if (l < len) {
      if (PF32) {
        const float* src = Pgf + (size_t)l * raw_floats;
        float* stg = Pstage + (PFMT == 2 ? 0 : (l & 1)) * raw_floats;       // plane sets: ONE staging buffer (refilled right after its widening)
        for (int c = tid; c < raw_floats / 4; c += G) cp_async16(stg + 4 * c, src + 4 * c);
      } else {
        const cplx* src = Pg + (size_t)l * nn + tid;
        int r = r_first, c = c_first;
        for (int idx = tid; idx < nn; idx += G, src += G) {
          cp_async16(dst + swz<NP>(r, c), src);
          r += dr; c += dc;
          if (c >= n) { c -= n; ++r; }
        }
      }
    }
    cp_async_commit();
  };
  auto widen = [&](int l, cplx* dst) {                     // PF32: raw tile of step l -> swizzled double2 operand
    const float* stg = Pstage + (PFMT == 2 ? 0 : (l & 1)) * raw_floats;
    const __half* sh = reinterpret_cast<const __half*>(stg);
    const int pl = n * ldh;
    const double sc = 1.0 / 8192.0;                        // 2^-TC_EU
    int r = r_first, c = c_first;
    for (int idx = tid; idx < nn; idx += G) {
      if (PFMT == 2) {
        const int o = r * ldh + c;
        dst[swz<NP>(r, c)] = make_double2(((double)__half2float(sh[o]) + (double)__half2float(sh[pl + o])) * sc,
                                          ((double)__half2float(sh[2 * pl + o]) + (double)__half2float(sh[3 * pl + o])) * sc);
      } else
        dst[swz<NP>(r, c)] = make_double2((double)stg[r * 32 + c], (double)stg[1024 + r * 32 + c]);
      r += dr; c += dc;
      if (c >= n) { c -= n; ++r; }
    }
  };
  fetch(0, Xb);
  if (PFMT == 2) {
    cp_async_wait<0>();
    __syncthreads();
    widen(0, Xb);
    __syncthreads();                                       // the staging buffer is free again
    fetch(1, nullptr);
  } else {
    fetch(1, Pb);
    if (PF32) {
      cp_async_wait<1>();
      __syncthreads();
      widen(0, Xb);
    }
  }
  for (int l = 1; l < len; ++l) {
    if (PF32) {
      cp_async_wait<0>();
      __syncthreads();                                     // raw P_{t0+l} landed; X complete; Pb free (step l-1 done)
      widen(l, Pb);
      if (PFMT == 2) __syncthreads();                      // single staging buffer: everybody has read it
      fetch(l + 1, nullptr);                               // fp32 tiles: staging half (l+1)&1 == (l-1)&1 was consumed a step ago
      if (PFMT != 2) __syncthreads();
    } else {
      fetch(l + 1, Pb + (l & 1) * T_::MAT);                // the buffer step l-1 used
      cp_async_wait<1>();
      __syncthreads();                                     // P_{t0+l} landed; X complete
    }
    double cr[RB][CB][2], ci[RB][CB][2];
    mma_gemm<NP, RB, CB>(PF32 ? Pb : Pb + ((l - 1) & 1) * T_::MAT, Xb, cr, ci, rb0, cb0, ksteps, lane);
    __syncthreads();                                       // in-place update: every warp has finished reading X
    store_tile<NP, RB, CB>(Xb, cr, ci, rb0, cb0, lane);
  }
  cp_async_wait<0>();
  __syncthreads();
  cplx* out = seg_out + ((size_t)b * S + seg) * nn;
  {
    int r = r_first, c = c_first;
    for (int idx = tid; idx < nn; idx += G) {
      out[idx] = Xb[swz<NP>(r, c)];
      r += dr; c += dc;
      if (c >= n) { c -= n; ++r; }
    }
  }
}

template <int NP, int RB, int CB, int PFMT = 0>
cudaError_t launch_segprod(const QocParams& p, int L, int S, cplx* seg_out, cudaStream_t st) {
  typedef MT<NP, RB, CB> T_;
  constexpr int PF32 = PFMT;
  const size_t raw = PFMT == 2 ? (size_t)8 * p.n * ((p.n + 15) / 16 * 16) : (size_t)2048 * sizeof(float);
  const size_t smem = PFMT == 2 ? (size_t)2 * T_::MAT * sizeof(cplx) + raw : PFMT ? (size_t)2 * T_::MAT * sizeof(cplx) + 2 * raw : (size_t)3 * T_::MAT * sizeof(cplx);
  cudaError_t e = cudaFuncSetAttribute(k_segprod<NP, RB, CB, PF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_segprod<NP, RB, CB, PF32>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  k_segprod<NP, RB, CB, PF32><<<(unsigned)((size_t)p.B * S), T_::THREADS, smem, st>>>(p, L, S, seg_out);
  return cudaGetLastError();
}

template <int NP, int RB, int CB, bool HERM>
cudaError_t launch_expm_h(const QocParams& p, int sm_count, cudaStream_t st) {
  typedef MT<NP, RB, CB> T_;
  const int ipc = T_::WARPS == 1 ? 4 : 1;
  const int threads = T_::WARPS == 1 ? 128 : T_::THREADS;
  const size_t smem = expm_item_bytes(T_::MAT, p.K) * ipc;
  cudaError_t e = cudaFuncSetAttribute(k_expm_mma<NP, RB, CB, HERM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_expm_mma<NP, RB, CB, HERM>, threads, smem);
  if (e != cudaSuccess) return e;
  if (occ < 1) occ = 1;
  const long long items = (long long)p.B * p.T;
  long long grid = (long long)sm_count * occ;
  const long long need = (items + ipc - 1) / ipc;
  if (grid > need) grid = need;
  k_expm_mma<NP, RB, CB, HERM><<<(unsigned)grid, threads, smem, st>>>(p);
  return cudaGetLastError();
}

// Hermitian Hamiltonians (all generators anti-Hermitian, checked on the host) take the triangle-only Taylor
// evaluation where the tile shape supports it (NP = 16, 32)
template <int NP, int RB, int CB>
cudaError_t launch_expm(const QocParams& p, int sm_count, cudaStream_t st) {
  constexpr bool HERM_OK = (NP == 32 && RB == 2 && CB == 4) || (NP == 16 && RB == 2 && CB == 2);
  if constexpr (HERM_OK) {
    if (p.herm && p.p >= 2) return launch_expm_h<NP, RB, CB, true>(p, sm_count, st);
  }
  return launch_expm_h<NP, RB, CB, false>(p, sm_count, st);
}

template <int NP, int RB, int CB, int NPB, int NXB, bool PF32 = false>
cudaError_t launch_chain(const QocParams& p, cudaStream_t st) {
  typedef MT<NP, RB, CB> T_;
  const size_t smem = PF32 ? (size_t)(1 + NXB) * T_::MAT * sizeof(cplx) + (size_t)NPB * 2048 * sizeof(float)
                           : (size_t)(NPB + NXB) * T_::MAT * sizeof(cplx);
  cudaError_t e = cudaFuncSetAttribute(k_chain_mma<NP, RB, CB, NPB, NXB, PF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_chain_mma<NP, RB, CB, NPB, NXB, PF32>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  k_chain_mma<NP, RB, CB, NPB, NXB, PF32><<<p.B, T_::THREADS, smem, st>>>(p);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Dense-m variants (m >= NP/2, e.g. config C5 with m = n): the costate sweep and the control
// gradient are GEMM-shaped too and run on the same DMMA machinery.
//   k_costate_mma : Lambda(t) = P_t^dagger Lambda(t+1) + S(t), Lambda = [n x m] (columns = states)
//   k_grad_mma    : W = Lambda(t+1)^dagger-ish outer sum  W[a][c] = sum_j conj(lam_j[a]) psi_j[c],
//                   g_k = Re sum_ac A_k[a][c] W[a][c]      (matexp_op_grad, tensorflow_state.py:49-65)
// ---------------------------------------------------------------------------------------------
// C = A^H * B: the A fragment (row r, k) = conj(A[k][r]) is read with the B-fragment pattern
template <int NP, int RB, int CB>
DEVINL void mma_gemm_ah(const cplx* __restrict__ A, const cplx* __restrict__ B, double (&cr)[RB][CB][2],
                        double (&ci)[RB][CB][2], int rb0, int cb0, int ksteps, int lane) {
  const int g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int i = 0; i < RB; ++i)
#pragma unroll
    for (int j = 0; j < CB; ++j) { cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0; }
#if QOC_CMUL_3M
  double t2[RB][CB][2];
#pragma unroll
  for (int i = 0; i < RB; ++i)
#pragma unroll
    for (int j = 0; j < CB; ++j) t2[i][j][0] = t2[i][j][1] = 0.0;
#endif
  const int ac = 8 * rb0 + g, bc = 8 * cb0 + g;
QOC_PRAGMA_KUNROLL
  for (int ks = 0; ks < ksteps; ++ks) {
    const int k = 4 * ks + q;
    const int km = sw_mask(k);
    const cplx* Arow = A + k * NP;
    const cplx* Brow = B + k * NP;
    cplx a[RB], b[CB];
#pragma unroll
    for (int i = 0; i < RB; ++i) { a[i] = Arow[(ac + 8 * i) ^ km]; a[i].y = -a[i].y; }
#pragma unroll
    for (int j = 0; j < CB; ++j) b[j] = Brow[(bc + 8 * j) ^ km];
#if QOC_CMUL_3M
    double sa[RB], sb[CB];
#pragma unroll
    for (int i = 0; i < RB; ++i) sa[i] = a[i].x + a[i].y;
#pragma unroll
    for (int j = 0; j < CB; ++j) sb[j] = b[j].x + b[j].y;
#pragma unroll
    for (int i = 0; i < RB; ++i)
#pragma unroll
      for (int j = 0; j < CB; ++j) dmma(cr[i][j][0], cr[i][j][1], a[i].x, b[j].x);
#pragma unroll
    for (int i = 0; i < RB; ++i)
#pragma unroll
      for (int j = 0; j < CB; ++j) dmma(t2[i][j][0], t2[i][j][1], a[i].y, b[j].y);
#pragma unroll
    for (int i = 0; i < RB; ++i)
#pragma unroll
      for (int j = 0; j < CB; ++j) dmma(ci[i][j][0], ci[i][j][1], sa[i], sb[j]);
#else
#pragma unroll
    for (int i = 0; i < RB; ++i)
#pragma unroll
      for (int j = 0; j < CB; ++j) dmma(cr[i][j][0], cr[i][j][1], a[i].x, b[j].x);
#pragma unroll
    for (int i = 0; i < RB; ++i)
#pragma unroll
      for (int j = 0; j < CB; ++j) dmma(ci[i][j][0], ci[i][j][1], a[i].x, b[j].y);
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const double nai = -a[i].y;
#pragma unroll
      for (int j = 0; j < CB; ++j) dmma(cr[i][j][0], cr[i][j][1], nai, b[j].y);
    }
#pragma unroll
    for (int i = 0; i < RB; ++i)
#pragma unroll
      for (int j = 0; j < CB; ++j) dmma(ci[i][j][0], ci[i][j][1], a[i].y, b[j].x);
#endif
  }
#if QOC_CMUL_3M
#pragma unroll
  for (int i = 0; i < RB; ++i)
#pragma unroll
    for (int j = 0; j < CB; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const double t1 = cr[i][j][e], u = t2[i][j][e];
        cr[i][j][e] = t1 - u;
        ci[i][j][e] = ci[i][j][e] - t1 - u;
      }
#endif
}

template <int NP, int RB, int CB, int NPB, int NLB>
__global__ void __launch_bounds__(MT<NP, RB, CB>::THREADS) k_costate_mma(QocParams p) {
  typedef MT<NP, RB, CB> T_;
  constexpr int G = T_::THREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* Pb = reinterpret_cast<cplx*>(smem_raw);           // [NPB][MAT]
  cplx* Lb = Pb + NPB * T_::MAT;                          // [NLB][MAT]  Lambda[r][j]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int rb0 = (warp / T_::WC) * RB, cb0 = (warp % T_::WC) * CB;
  const int n = p.n, T = p.T, m = p.m, nn = n * n, mn = m * n;
  const int ksteps = (n + 3) >> 2;
  const int b = blockIdx.x;
  const cplx* Pg = reinterpret_cast<const cplx*>(p.P) + (size_t)b * T * nn;
  const cplx* psi_b = p.psi + (size_t)b * (T + 1) * mn;
  cplx* lam_b = p.lam + (size_t)b * (T + 1) * mn;
  const double* sc = p.scal + (size_t)b * 8;
  const double o_re = sc[0], o_im = sc[1], spdfac = sc[4];
  const bool forb = p.reg.has_forbidden && p.fw != nullptr;
  const bool spd = p.reg.has_speed_up != 0;
  const double m2 = (double)m * (double)m;

  auto source = [&](int t, int j, int r) -> cplx {       // S(t)[r][j]
    cplx s = make_double2(0.0, 0.0);
    if (forb && p.dressW) {
      const cplx d = p.psid[((size_t)b * (T + 1) + t) * mn + (size_t)j * n + r];
      s.x += d.x; s.y += d.y;
    } else if (forb) {
      const cplx x = psi_b[(size_t)t * mn + (size_t)j * n + r];
      const double pop = x.x * x.x + x.y * x.y;
      const double c = p.fw[r] / (double)T * 2.0 * pop;
      s.x += c * x.x; s.y += c * x.y;
    }
    if (spd) {
      const cplx o = p.ot[(size_t)b * (T + 1) + t];
      const cplx ph = p.phi[(size_t)j * n + r];
      s.x += spdfac * (o.x * ph.x - o.y * ph.y); s.y += spdfac * (o.x * ph.y + o.y * ph.x);
    }
    return s;
  };
  auto prefetch = [&](int t) {
    if (t >= 1) {
      const cplx* src = Pg + (size_t)t * nn;
      cplx* dst = Pb + (t % NPB) * T_::MAT;
      for (int idx = tid; idx < nn; idx += G) {
        const int r = idx / n, c = idx - r * n;
        cp_async16(dst + swz<NP>(r, c), src + idx);
      }
    }
    cp_async_commit();
  };

  for (int i = tid; i < (NPB + NLB) * T_::MAT; i += G) Pb[i] = make_double2(0.0, 0.0);
  __syncthreads();
  for (int idx = tid; idx < mn; idx += G) {                // Lambda(T) = -(2/m^2) o phi + S(T)
    const int j = idx / n, r = idx - j * n;
    const cplx ph = p.phi[idx];
    cplx l = make_double2((o_re * ph.x - o_im * ph.y) * (-2.0 / m2), (o_re * ph.y + o_im * ph.x) * (-2.0 / m2));
    const cplx s = source(T, j, r);
    l.x += s.x; l.y += s.y;
    Lb[swz<NP>(r, j)] = l;
    lam_b[(size_t)T * mn + idx] = l;
  }
#pragma unroll
  for (int i = 0; i < NPB - 1; ++i) prefetch(T - 1 - i);
  int cur = 0;
  for (int t = T - 1; t >= 1; --t) {
    cp_async_wait<NPB - 2>();
    __syncthreads();
    prefetch(t - (NPB - 1));
    const cplx* Lc = Lb + (NLB == 2 ? cur : 0) * T_::MAT;
    cplx* Ln = Lb + (NLB == 2 ? (cur ^ 1) : 0) * T_::MAT;
    double cr[RB][CB][2], ci[RB][CB][2];
    mma_gemm_ah<NP, RB, CB>(Pb + (t % NPB) * T_::MAT, Lc, cr, ci, rb0, cb0, ksteps, lane);
    if (NLB == 1) __syncthreads();                    // in-place update: every warp has finished reading Lambda
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int r = 8 * (rb0 + i) + g;
#pragma unroll
      for (int jj = 0; jj < CB; ++jj)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = 8 * (cb0 + jj) + 2 * q + e;
          cplx l = make_double2(cr[i][jj][e], ci[i][jj][e]);
          if (r < n && j < m) {
            const cplx s = source(t, j, r);
            l.x += s.x; l.y += s.y;
            lam_b[(size_t)t * mn + (size_t)j * n + r] = l;
          } else {
            l = make_double2(0.0, 0.0);
          }
          Ln[swz<NP>(r, j)] = l;
        }
    }
    cur ^= 1;
  }
  cp_async_wait<0>();
}

template <int NP, int RB, int CB>
__global__ void __launch_bounds__(MT<NP, RB, CB>::THREADS) k_grad_mma(QocParams p) {
  typedef MT<NP, RB, CB> T_;
  constexpr int G = T_::THREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* Ls = reinterpret_cast<cplx*>(smem_raw);           // [MAT]  L[j][a] = lam_j(t+1)[a]
  cplx* Ys = Ls + T_::MAT;                                // [MAT]  Y[j][c] = psi_j(t+1)[c]
  __shared__ double red[32 * 8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int rb0 = (warp / T_::WC) * RB, cb0 = (warp % T_::WC) * CB;
  const int n = p.n, T = p.T, m = p.m, K = p.K, nn = n * n, mn = m * n;
  const int ksteps = (m + 3) >> 2;
  const long long items = (long long)p.B * T;
  for (int i = tid; i < 2 * T_::MAT; i += G) Ls[i] = make_double2(0.0, 0.0);
  __syncthreads();
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = (int)(item / T), t = (int)(item % T);
    const size_t off = ((size_t)b * (T + 1) + (t + 1)) * mn;
    for (int idx = tid; idx < mn; idx += G) {
      const int j = idx / n, i = idx - j * n;
      cp_async16(Ls + swz<NP>(j, i), p.lam + off + idx);
      cp_async16(Ys + swz<NP>(j, i), p.psi + off + idx);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    double wr[RB][CB][2], wi[RB][CB][2];
    mma_gemm_ah<NP, RB, CB>(Ls, Ys, wr, wi, rb0, cb0, ksteps, lane);
    for (int k0 = 0; k0 < K; k0 += 8) {                   // up to 8 controls per reduction round
      double part[8];
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) part[kk] = 0.0;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        if (k0 + kk < K) {
          const cplx* Ak = p.A + (size_t)(k0 + kk + 1) * nn;
          double s = 0.0;
#pragma unroll
          for (int i = 0; i < RB; ++i) {
            const int a = 8 * (rb0 + i) + g;
#pragma unroll
            for (int jj = 0; jj < CB; ++jj)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int c = 8 * (cb0 + jj) + 2 * q + e;
                if (a < n && c < n) {
                  const cplx av = __ldg(Ak + (size_t)a * n + c);
                  s += av.x * wr[i][jj][e] - av.y * wi[i][jj][e];
                }
              }
          }
          part[kk] = warp_sum_d(s);
        }
      }
      if (lane == 0) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) red[warp * 8 + kk] = part[kk];
      }
      __syncthreads();
      if (tid < 8 && k0 + tid < K) {
        double s = 0.0;
        for (int w = 0; w < T_::WARPS; ++w) s += red[w * 8 + tid];
        p.gctrl[((size_t)b * K + k0 + tid) * T + t] = s;
      }
      __syncthreads();
    }
  }
}

template <int NP, int RB, int CB, int NPB, int NLB>
cudaError_t launch_costate_mma(const QocParams& p, cudaStream_t st) {
  typedef MT<NP, RB, CB> T_;
  static_assert(NPB >= 2, "prefetch needs two propagator buffers");
  const size_t smem = (size_t)(NPB + NLB) * T_::MAT * sizeof(cplx);
  cudaError_t e = cudaFuncSetAttribute(k_costate_mma<NP, RB, CB, NPB, NLB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_costate_mma<NP, RB, CB, NPB, NLB><<<p.B, T_::THREADS, smem, st>>>(p);
  return cudaGetLastError();
}

template <int NP, int RB, int CB>
cudaError_t launch_grad_mma(const QocParams& p, int sm_count, cudaStream_t st) {
  typedef MT<NP, RB, CB> T_;
  const size_t smem = (size_t)2 * T_::MAT * sizeof(cplx);
  cudaError_t e = cudaFuncSetAttribute(k_grad_mma<NP, RB, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int occ = 1;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_grad_mma<NP, RB, CB>, T_::THREADS, smem);
  if (e != cudaSuccess) return e;
  if (occ < 1) occ = 1;
  const long long items = (long long)p.B * p.T;
  long long grid = (long long)sm_count * occ;
  if (grid > items) grid = items;
  k_grad_mma<NP, RB, CB><<<(unsigned)grid, T_::THREADS, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t qoc_launch_expm_f64(const QocParams& p, int NP, int sm_count, cudaStream_t st, int64_t* launches) {
  ++*launches;
  switch (NP) {
    case 8: return launch_expm<8, 1, 1>(p, sm_count, st);
    case 16: return launch_expm<16, 2, 2>(p, sm_count, st);
    case 24: return launch_expm<24, 1, 3>(p, sm_count, st);
    case 32: return launch_expm<32, 2, 4>(p, sm_count, st);
    case 40: return launch_expm<40, 1, 5>(p, sm_count, st);
    case 48: return launch_expm<48, 2, 3>(p, sm_count, st);
    case 56: return launch_expm<56, 1, 7>(p, sm_count, st);
    case 64: return launch_expm<64, 2, 4>(p, sm_count, st);
  }
  return cudaErrorInvalidValue;
}

cudaError_t qoc_launch_segprod_f64(const QocParams& p, int NP, int p_is_f32, int L, int S, cplx* seg_out, cudaStream_t st,
                                   int64_t* launches) {
  ++*launches;
  if (p_is_f32 == 2) {                                // QOC_F16X2 plane sets, n <= 64
    switch (NP) {
      case 8: return launch_segprod<8, 1, 1, 2>(p, L, S, seg_out, st);
      case 16: return launch_segprod<16, 2, 2, 2>(p, L, S, seg_out, st);
      case 24: return launch_segprod<24, 1, 3, 2>(p, L, S, seg_out, st);
      case 32: return launch_segprod<32, 2, 4, 2>(p, L, S, seg_out, st);
      case 40: return launch_segprod<40, 1, 5, 2>(p, L, S, seg_out, st);
      case 48: return launch_segprod<48, 2, 3, 2>(p, L, S, seg_out, st);
      case 56: return launch_segprod<56, 1, 7, 2>(p, L, S, seg_out, st);
      case 64: return launch_segprod<64, 2, 4, 2>(p, L, S, seg_out, st);
    }
    return cudaErrorInvalidValue;
  }
  if (p_is_f32) {                                     // tcgen05 path: n <= 32
    switch (NP) {
      case 8: return launch_segprod<8, 1, 1, 1>(p, L, S, seg_out, st);
      case 16: return launch_segprod<16, 2, 2, 1>(p, L, S, seg_out, st);
      case 24: return launch_segprod<24, 1, 3, 1>(p, L, S, seg_out, st);
      case 32: return launch_segprod<32, 2, 4, 1>(p, L, S, seg_out, st);
    }
    return cudaErrorInvalidValue;
  }
  switch (NP) {
    case 8: return launch_segprod<8, 1, 1>(p, L, S, seg_out, st);
    case 16: return launch_segprod<16, 2, 2>(p, L, S, seg_out, st);
    case 24: return launch_segprod<24, 1, 3>(p, L, S, seg_out, st);
    case 32: return launch_segprod<32, 2, 4>(p, L, S, seg_out, st);
    case 40: return launch_segprod<40, 1, 5>(p, L, S, seg_out, st);
    case 48: return launch_segprod<48, 2, 3>(p, L, S, seg_out, st);
    case 56: return launch_segprod<56, 1, 7>(p, L, S, seg_out, st);
    case 64: return launch_segprod<64, 2, 4>(p, L, S, seg_out, st);
  }
  return cudaErrorInvalidValue;
}

cudaError_t qoc_launch_chain_f64(const QocParams& p, int NP, int p_is_f32, cudaStream_t st, int64_t* launches) {
  ++*launches;
  if (p_is_f32) {                                     // tcgen05 path: n <= 32
    switch (NP) {
      case 8: return launch_chain<8, 1, 1, 3, 2, true>(p, st);
      case 16: return launch_chain<16, 1, 1, 3, 2, true>(p, st);
      case 24: return launch_chain<24, 1, 1, 3, 2, true>(p, st);
      case 32: return launch_chain<32, 1, 2, 3, 2, true>(p, st);
    }
    return cudaErrorInvalidValue;
  }
  if (p.chain_no_psi) {                               // runs beside k_vec_sweep: two propagator buffers leave room for it
    switch (NP) {
      case 32: return launch_chain<32, 1, 2, 2, 2>(p, st);
      case 40: return launch_chain<40, 1, 5, 2, 2>(p, st);
      case 48: return launch_chain<48, 1, 3, 2, 2>(p, st);
    }
  }
  switch (NP) {
    case 8: return launch_chain<8, 1, 1, 3, 2>(p, st);
    case 16: return launch_chain<16, 1, 1, 3, 2>(p, st);
    case 24: return launch_chain<24, 1, 1, 3, 2>(p, st);
    case 32: return launch_chain<32, 1, 2, 3, 2>(p, st);
    case 40: return launch_chain<40, 1, 5, 3, 2>(p, st);
    case 48: return launch_chain<48, 1, 3, 3, 2>(p, st);
    case 56: return launch_chain<56, 1, 7, 2, 2>(p, st);      // 4 x 50 KB
    case 64: return launch_chain<64, 2, 4, 2, 1>(p, st);      // 3 x 64 KB: X updated in place
  }
  return cudaErrorInvalidValue;
}

// dense-m (m >= NP/2) costate / gradient on the DMMA path; returns cudaErrorNotSupported when not applicable
cudaError_t qoc_launch_costate_mma(const QocParams& p, int NP, cudaStream_t st, int64_t* launches) {
  if (2 * p.m < NP || p.m > NP) return cudaErrorNotSupported;
  ++*launches;
  switch (NP) {
    case 8: return launch_costate_mma<8, 1, 1, 3, 2>(p, st);
    case 16: return launch_costate_mma<16, 1, 1, 3, 2>(p, st);
    case 24: return launch_costate_mma<24, 1, 1, 3, 2>(p, st);
    case 32: return launch_costate_mma<32, 1, 2, 3, 2>(p, st);
    case 40: return launch_costate_mma<40, 1, 5, 3, 2>(p, st);
    case 48: return launch_costate_mma<48, 1, 3, 3, 2>(p, st);
    case 56: return launch_costate_mma<56, 1, 7, 2, 2>(p, st);
    case 64: return launch_costate_mma<64, 2, 4, 2, 1>(p, st);
  }
  --*launches;
  return cudaErrorNotSupported;
}

cudaError_t qoc_launch_grad_mma(const QocParams& p, int NP, int sm_count, cudaStream_t st, int64_t* launches) {
  if (2 * p.m < NP || p.m > NP) return cudaErrorNotSupported;
  ++*launches;
  switch (NP) {
    case 8: return launch_grad_mma<8, 1, 1>(p, sm_count, st);
    case 16: return launch_grad_mma<16, 1, 1>(p, sm_count, st);
    case 24: return launch_grad_mma<24, 1, 1>(p, sm_count, st);
    case 32: return launch_grad_mma<32, 1, 2>(p, sm_count, st);
    case 40: return launch_grad_mma<40, 1, 5>(p, sm_count, st);
    case 48: return launch_grad_mma<48, 1, 3>(p, sm_count, st);
    case 56: return launch_grad_mma<56, 1, 7>(p, sm_count, st);
    case 64: return launch_grad_mma<64, 2, 4>(p, sm_count, st);
  }
  --*launches;
  return cudaErrorNotSupported;
}
