// Internal declarations shared by the kernel translation units and the C-ABI front end.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "qoc_b200.h"
#include "qoc_tc_f16.cuh"

typedef double2 cplx;

#define QOC_SEG_LEN 16   // propagators per segment product (k_segprod)

// Everything the kernels need, passed by value.
struct QocParams {
  int n, K, T, m, B, p, s;
  int has_cidx;
  int state_transfer;
  int herm;             // every generator A_k is anti-Hermitian (Hermitian Hamiltonians)
  int chain_no_psi;     // k_chain_mma only propagates X (U_final, unitary_scale); psi comes from k_vec_sweep
  double dt, inv2s;
  double invfact[32];   // 1/j!
  // constants
  const cplx* A;        // [K+1][n][n]
  const cplx* U0;       // [n][n]
  const cplx* phi;      // [m][n]
  const cplx* V;        // [m][n]
  const int* cidx;      // [m]
  const double* maxA;   // [K]
  const double* env;    // [K][T] or null
  const double* fw;     // [n] or null   (forbidden weights, not yet /T)
  const cplx* dressW;   // [n][n] or null: forbidden populations taken on W psi (forbid_dressed)
  cplx* psid;           // [B][T+1][m][n]: W psi, then (in place) the dressed costate sources
  const int* coo_off;   // [K+1]
  const int* coo_r;     // [nnz]
  const int* coo_c;     // [nnz]
  const cplx* coo_v;    // [nnz]
  int pat_n;            // union sparsity pattern of A_0..A_K
  const int* pat_rc;    // [pat_n]  (row << 16) | col
  const cplx* pat_coef; // [K+1][pat_n] (k-major)
  const float2* pat_coef_f;  // same in fp32 (tcgen05 path)
  qoc_reg_t reg;
  // per-call
  const double* base;   // [B][K][T]
  // workspace
  void* P;              // [B][T][n][n] complex (double2 or float2)
  cplx* psi;            // [B][T+1][m][n]
  cplx* lam;            // [B][T+1][m][n]
  double* gctrl;        // [B][K][T]
  cplx* ot;             // [B][T+1]
  double* scal;         // [B][8]: o.re o.im loss statereg spd unitary_scale - -
  cplx* Ufin;           // [B][n][n]
  // outputs
  double* loss; double* reg_loss; double* grad; double* unitary_scale; double* grad_squared;
};

struct qoc_handle_s {
  qoc_dims_t d;
  int NP;
  int Bc;                                 // instances per pass (batch chunk), <= d.B
  std::string err;
  int sm_count;
  bool problem_set, ws_set;
  // device constants (cudaMalloc'd by the handle; a few hundred KB)
  cplx *A, *U0, *phi, *V, *coo_v, *pat_coef;
  float2* pat_coef_f;
  int* err_flag;
  int *cidx, *coo_off, *coo_r, *coo_c, *pat_rc;
  int pat_n;
  double *maxA, *env, *fw;
  cplx *dressW, *psid;
  int has_cidx, nnz, herm;
  double dt;
  qoc_reg_t reg;
  // workspace carve-up
  char* ws; size_t ws_bytes;
  void* P; cplx *psi, *lam, *ot, *Ufin; double *gctrl, *scal;
  double *st_base, *st_grad, *st_out;     // staging for the *_host entry points
  void* scratch;                          // n > 64: per-CTA global intermediates
  int64_t launches;
  bool profiling; int ev_recorded;
  cudaStream_t hi;                        // high-priority stream of the loss / gradient critical path (vec-sweep mode)
  cudaEvent_t ev_fork, ev_join;
  bool hi_pending;
  cudaStream_t work;                      // where the current pass's sweep / gradient kernels go (hi or the caller's stream)
  cplx* seg;                              // [Bc][ceil(T/QOC_SEG_LEN)][n][n] segment products
  // QOC_F16X2: tcgen05 / TMA program engine (qoc_tc_f16.cu)
  bool tc, tc_ready;
  TcGeom tg; TcMaps tmaps;
  TcStoreMaps tsmaps;          // store-side descriptors of the pair kernel's epilogue
  __half *tc_seg, *tc_scr, *tc_const;     // plane-set arrays inside the workspace (P is h->P)
  TcExpmOp* tc_ops; int tc_nops; float tc_xscale;
  float2* A_f;                            // dense fp32 copy of A_0..A_K for the generator assembly
  double* U0_host;                        // kept for the constant plane sets
  int tc_grid, tc_S;
  cudaEvent_t ev[QOC_NUM_KERNELS + 1];
};

// kernel launchers (qoc_mma_f64.cu, qoc_sweeps.cu); return cudaError_t, bump *launches
cudaError_t qoc_launch_expm_f64(const QocParams& p, int NP, int sm_count, cudaStream_t st, int64_t* launches);
cudaError_t qoc_launch_chain_f64(const QocParams& p, int NP, int p_is_f32, cudaStream_t st, int64_t* launches);
cudaError_t qoc_launch_segprod_f64(const QocParams& p, int NP, int p_is_f32, int L, int S, cplx* seg_out, cudaStream_t st,
                                   int64_t* launches);
bool qoc_vec_sweep_supported(const QocParams& p);
cudaError_t qoc_launch_vec_sweep(const QocParams& p, int reverse, int p_is_f32, cudaStream_t st, int64_t* launches);
cudaError_t qoc_launch_expm_tc32(const QocParams& p, int sm_count, int* err_flag, cudaStream_t st, int64_t* launches);
size_t qoc_large_scratch_elems(int n, int B, int sm_count);
cudaError_t qoc_launch_expm_large(const QocParams& p, int sm_count, void* scratch, cudaStream_t st, int64_t* launches);
cudaError_t qoc_launch_chain_large(const QocParams& p, void* scratch, cudaStream_t st, int64_t* launches);
cudaError_t qoc_launch_costate_large(const QocParams& p, cudaStream_t st, int64_t* launches);
bool qoc_grad_large_supported(const QocParams& p);
cudaError_t qoc_launch_grad_large(const QocParams& p, int sm_count, cudaStream_t st, int64_t* launches);
cudaError_t qoc_launch_dress(const QocParams& p, int phase, cudaStream_t st, int64_t* launches);
cudaError_t qoc_launch_costate_mma(const QocParams& p, int NP, cudaStream_t st, int64_t* launches);
cudaError_t qoc_launch_grad_mma(const QocParams& p, int NP, int sm_count, cudaStream_t st, int64_t* launches);
cudaError_t qoc_launch_fwd_reduce(const QocParams& p, cudaStream_t st, int64_t* launches);
cudaError_t qoc_launch_costate(const QocParams& p, int p_is_f32, cudaStream_t st, int64_t* launches);
cudaError_t qoc_launch_grad(const QocParams& p, int sm_count, cudaStream_t st, int64_t* launches);
cudaError_t qoc_launch_finalize(const QocParams& p, cudaStream_t st, int64_t* launches);
bool qoc_plane_sweep_supported(int n, int m);
cudaError_t qoc_launch_plane_sweep(const QocParams& p, const void* planes, int reverse, cudaStream_t st, int64_t* launches);
