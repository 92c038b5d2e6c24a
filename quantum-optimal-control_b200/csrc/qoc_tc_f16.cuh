// Declarations of the tcgen05 / TMEM / TMA complex-GEMM "program" engine (csrc/qoc_tc_f16.cu), the
// fp32-class arithmetic of the propagator and chain stages for 32 < n <= 256 (QOC_F16X2).
//
// Split-plane matrix ("plane set").  An n x n complex matrix X with a power-of-two scale 2^e is stored
// as four row-major fp16 planes [4][n][ld], ld = n rounded up to 16 (32-byte rows: every TMA box row and every
// epilogue store covers whole L2 sectors):
//     plane 0 = h0(Re), 1 = h1(Re), 2 = h0(Im), 3 = h1(Im),   2^e x = h0 + h1,
// h0 = fp16(2^e x), h1 = fp16(2^e x - h0): 22 significant bits, 4 bytes per real number -- the size of
// the fp32 the reference stores (core/tensorflow_state.py:49,70,205) -- and directly consumable by
// tcgen05.mma kind::f16 through TMA.  A product is three MMAs  A0 B0 + A0 B1 + A1 B0  with the fp32
// accumulator in TMEM (the dropped A1 B1 term is 2^-22 relative).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <vector>

#define TC_EU 13                 // scale exponent of unitary-like matrices (entries <= ~1 -> <= 8192 stored)
#define TC_NSLOT 5               // per-item scratch matrices: X (even rounds), Y, Z ping-pong, X (odd rounds)
#define TC_SLOT_OUT 15           // "the program's output matrix" as an expm-op destination
#define TC_MAX_N 256
#define TC_MAX_OPS 96           // products of one propagator program (Paterson-Stockmeyer steps + squarings)

enum { TC_CLS_SCR = 0, TC_CLS_P = 1, TC_CLS_SEG = 2, TC_CLS_CONST = 3, TC_NCLS = 4 };
// programs: an item is a sequence of dependent complex products run by one CTA
//   EXPM : item (b,t)  -> P[b][t] = (sum_{j<=p} X^j/j!)^(2^s), X = (A_0 + sum_k u_k A_k)/2^s     (get_matexp, tensorflow_state.py:25-46)
//   SEG  : item (b,sg) -> seg[b][sg] = P[b][t1-1] ... P[b][t0]                                    (re-associated chain, :214-220)
//   CHAIN: item b      -> U_final[b] = M[b][len-1] ... M[b][0] U0 (M = seg or P), unitary_scale   (:214-227)
//   GEMM : item i      -> seg[i] = P[2i] P[2i+1]                                                  (unit test of one product)
enum { TC_PROG_EXPM = 0, TC_PROG_SEG = 1, TC_PROG_CHAIN = 2, TC_PROG_GEMM = 3 };

// one product of the expm program: D = slot[sa] * slot[sb];  dst = c[0] D + c[1] X + c[2] I  (stored units)
struct TcExpmOp {
  int8_t sa, sb, d1, d2;         // scratch slots; d1/d2 = -1: none, TC_SLOT_OUT: P[b][t]
  int8_t se;                     // slot of the elementwise source "X" of c[1]
  float c1[3], c2[3];
};

struct TcMaps {                  // TMA descriptors per buffer class
  CUtensorMap a[TC_NCLS];        // A form: box {32 k, 128 rows, 4 planes, 1}, SWIZZLE_64B  (K-major operand tiles)
  CUtensorMap b[TC_NCLS];        // B form: box {64 n, 32 k, 4 planes, 1},     SWIZZLE_128B (MN-major operand tiles)
};

struct TcStoreMaps {             // epilogue stores of the pair kernel: box {16 columns, 32 rows, 4 planes, 1}, SWIZZLE_32B
  CUtensorMap st[TC_NCLS];
};

struct TcParams {
  int prog;
  int n, ld, N16, NT0, NH, NGT, DIOFF, NBUF, RB, KBLK, stages, tmem_cols;
  long long items;
  int ilv;                       // items interleaved per CTA (EXPM: 2 -- consecutive products of the stream are independent)
  __half* base[TC_NCLS];         // plane-set arrays; matrix i of a class at base + i * 4 n ld
  // EXPM
  int nops; const TcExpmOp* ops;
  int K, T; const double* ctrl; const double* maxA; const float2* A_f; float xscale;   // X' = xscale * (A_0 + sum u_k A_k)
  // union sparsity pattern of A_0..A_K (sparse Hamiltonians: only these entries of X are ever written; the X slots / images
  // are zero elsewhere): pat_rc[e] = row << 16 | column, pat_coef_f[e][K + 1]; pat_n = 0: dense assembly from A_f
  int pat_n; const int* pat_rc; const float2* pat_coef_f;
  // SEG / CHAIN
  int L, S, chain_cls, chain_len;
  double2* Ufin; double* scal;
  int* err_flag;
  int tma_store;                 // pair kernel: results leave through shared-memory staging + TMA stores
  unsigned long long* prof;      // optional [grid][8] cycle counters (tools/tc_prog_test.cu)
  // shared-memory matrix descriptor fields (>> 4), overridable by the probe tool
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
};

// N16 = n rounded up to 16 (MMA N granularity at M = 128).  Output tiles are 128 rows x NT columns with NT = N16 when
// N16 <= 128, else two column halves [0,128) and [128,N16) (aligned to the 64-column TMA boxes); NGT = 64-column groups
// per half.  Accumulator buffer: Dr at +0, Di at +DIOFF, NBUF buffers (two when a product has more than one tile).
struct TcGeom { int n, ld, N16, NT0, NH, NGT, DIOFF, NBUF, RB, KBLK, stages, tmem_cols, ctas_per_sm; size_t smem, mat_halfs; };
static __host__ __device__ inline int tc_ld(int n) { return (n + 15) / 16 * 16; }

// host helpers (qoc_tc_f16.cu)
bool tc_geometry(int n, TcGeom* g);
const char* tc_make_map(CUtensorMap* map, const void* base, int n, int ld, unsigned long long n_mats, bool b_form);
// Paterson-Stockmeyer program for (p, s) with operand scales 2^eX (X) and 2^eY (X^2)
void tc_build_expm_ops(int p, int s, int eX, int eY, std::vector<TcExpmOp>& ops);
// scale exponents from the entrywise bound |X| <= xmax (max entry) and ||X||_2 <= theta
void tc_pick_scales(double xmax, double theta, int* eX, int* eY);
cudaError_t tc_launch(const TcParams& q, const TcMaps& maps, const TcGeom& g, int grid, cudaStream_t st);
// small Hilbert dimensions (n <= 64): the propagator program with shared-memory-resident operands (qoc_tc_small.cu)
bool tc_small_supported(int n);
cudaError_t tc_small_launch_expm(const TcParams& q, int n, int sm_count, cudaStream_t st);
// 128 < n <= 256: the propagator program on CTA pairs (tcgen05 cta_group::2) in clusters of cs = 2 or 4 CTAs (qoc_tc_pair.cu)
bool tc_pair_supported(int n);
int tc_pair_max_clusters(int cs);
cudaError_t tc_pair_launch_expm(const TcParams& q, const TcMaps& maps, const TcStoreMaps& smaps, const TcGeom& g, int cs, cudaStream_t st);
const char* tc_make_store_map(CUtensorMap* map, const void* base, int n, int ld, unsigned long long n_mats);
// plane-set <-> complex double converters (host side of tests, constants upload)
void tc_pack_host(const double* z /* [n][n] (re,im) */, int n, int ld, int e, __half* out /* [4][n][ld] */);
