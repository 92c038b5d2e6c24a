// C-ABI front end of libqoc_b200.so (declared in include/qoc_b200.h).
#include "qoc_internal.cuh"
#include <cstring>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <vector>
#include <new>

#define QOC_CHECK_H(h) do { if (!(h)) return QOC_EINVAL; } while (0)
#define CUDA_TRY(h, expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { \
    (h)->err = std::string(#expr) + ": " + cudaGetErrorString(_e); return QOC_ECUDA; } } while (0)

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

static int pick_np(int n) { return n <= 64 ? (n + 7) / 8 * 8 : -1; }

struct WsLayout { size_t P, psi, lam, gctrl, ot, scal, Ufin, st_base, st_grad, st_out, seg, scratch, tc_seg, tc_scr, tc_const, total; };

#define QOC_TC_SEG_LEN 16   // propagators per segment product of the QOC_F16X2 U_final branch
#define QOC_TC_ILV 2        // (b,t) items interleaved per CTA in the propagator program: hides the product-to-product dependency

static int tc_grid_for(const qoc_dims_t& d, int sm_count) {
  TcGeom g;
  if (!tc_geometry(d.n, &g)) return 0;
  return sm_count * g.ctas_per_sm;
}

static WsLayout ws_layout(const qoc_dims_t& d, int sm_count, int Bc) {
  // Bc = instances processed per pass (batch chunk); P / psi / lam / gctrl / ot / scratch are reused by every pass
  WsLayout L;
  const size_t nn = (size_t)d.n * d.n, mn = (size_t)d.m * d.n;
  // propagators: fp64 interleaved [n][n] complex, or (QOC_TF32X3) fp32 planar padded [2][32][32]
  const bool tc = d.dtype == QOC_F16X2;
  const size_t tc_mat = tc ? (size_t)4 * d.n * tc_ld(d.n) * sizeof(__half) : 0;      // one split-plane matrix
  const size_t p_item = d.dtype == QOC_F64 ? nn * sizeof(cplx) : tc ? tc_mat : (size_t)2 * 32 * 32 * sizeof(float);
  size_t off = 0;
  L.P = off; off += align_up((size_t)Bc * d.T * p_item);
  L.psi = off; off += align_up((size_t)Bc * (d.T + 1) * mn * sizeof(cplx));
  L.lam = off; off += align_up((size_t)Bc * (d.T + 1) * mn * sizeof(cplx));
  L.gctrl = off; off += align_up((size_t)Bc * d.K * d.T * sizeof(double));
  L.ot = off; off += align_up((size_t)Bc * (d.T + 1) * sizeof(cplx));
  L.scal = off; off += align_up((size_t)d.B * 8 * sizeof(double));
  L.Ufin = off; off += align_up((size_t)d.B * nn * sizeof(cplx));
  L.st_base = off; off += align_up((size_t)d.B * d.K * d.T * sizeof(double));
  L.st_grad = off; off += align_up((size_t)d.B * d.K * d.T * sizeof(double));
  L.st_out = off; off += align_up((size_t)d.B * 4 * sizeof(double));
  L.seg = off;                                   // segment products of the re-associated U_final chain (n <= 64, fp64)
  if (d.n <= 64)                                 // (also QOC_F16X2: its U_final branch for n <= 64 runs on the fp64 segment kernels)
    off += align_up((size_t)Bc * ((d.T + QOC_SEG_LEN - 1) / QOC_SEG_LEN) * nn * sizeof(cplx));
  L.scratch = off;
  if (d.n > 64 && !tc) off += align_up(qoc_large_scratch_elems(d.n, Bc, sm_count) * sizeof(cplx));
  L.tc_seg = L.tc_scr = L.tc_const = off;
  if (tc) {
    off = align_up(off, 1024);
    L.tc_seg = off; off += align_up((size_t)Bc * ((d.T + QOC_TC_SEG_LEN - 1) / QOC_TC_SEG_LEN) * tc_mat, 1024);
    L.tc_scr = off; off += align_up((size_t)tc_grid_for(d, sm_count) * QOC_TC_ILV * TC_NSLOT * tc_mat, 1024);
    L.tc_const = off; off += align_up(2 * tc_mat, 1024);
  }
  L.total = off;
  return L;
}

extern "C" {

int qoc_abi_version(void) { return QOC_ABI_VERSION; }

int qoc_create(qoc_handle_t* out, const qoc_dims_t* dims) {
  if (!out || !dims) return QOC_EINVAL;
  *out = nullptr;
  const qoc_dims_t& d = *dims;
  if (d.n < 1 || d.K < 0 || d.K > 31 || d.T < 1 || d.m < 1 || d.B < 1 || d.exp_terms < 1 || d.exp_terms > 31 || ((d.flags & QOC_FLAG_STATE_TRANSFER) && d.exp_terms < 2) || d.scaling < 0 ||
      d.scaling > 60)
    return QOC_EINVAL;
  if (d.dtype != QOC_F64 && d.dtype != QOC_TF32X3 && d.dtype != QOC_F16X2) return QOC_EINVAL;
  qoc_handle_s* h = new (std::nothrow) qoc_handle_s();
  if (!h) return QOC_ENOMEM;
  h->d = d;
  h->NP = pick_np(d.n);
  h->problem_set = h->ws_set = false;
  h->A = h->U0 = h->phi = h->V = h->coo_v = h->pat_coef = nullptr;
  h->pat_coef_f = nullptr; h->err_flag = nullptr;
  h->cidx = h->coo_off = h->coo_r = h->coo_c = h->pat_rc = nullptr;
  h->pat_n = 0;
  h->maxA = h->env = h->fw = nullptr;
  h->dressW = h->psid = nullptr;
  h->has_cidx = 0; h->nnz = 0; h->dt = 0.0;
  std::memset(&h->reg, 0, sizeof(h->reg));
  h->ws = nullptr; h->ws_bytes = 0;
  h->launches = 0;
  h->sm_count = 0;
  h->profiling = false; h->ev_recorded = 0;
  for (int i = 0; i <= QOC_NUM_KERNELS; ++i) h->ev[i] = nullptr;
  h->hi = nullptr; h->ev_fork = h->ev_join = nullptr; h->hi_pending = false; h->work = nullptr; h->seg = nullptr;
  h->tc = d.dtype == QOC_F16X2; h->tc_ready = false;
  h->tc_seg = h->tc_scr = h->tc_const = nullptr; h->tc_ops = nullptr; h->tc_nops = 0; h->tc_xscale = 0.f;
  h->A_f = nullptr; h->U0_host = nullptr; h->tc_grid = 0; h->tc_S = 0;
  *out = h;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) {
    h->err = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)";
    return QOC_ECUDA;
  }
  if (d.n > 1024) { h->err = "n too large"; return QOC_EINVAL; }
  {   // batch chunk: the largest Bc <= B whose workspace fits the memory budget
    size_t free_b = 0, total_b = 0;
    double budget = 0.0;
    const char* env = getenv("QOC_B200_MAX_WS_GB");
    if (env && atof(env) > 0.0) budget = atof(env) * 1e9;
    else if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) budget = 0.8 * (double)free_b;
    h->Bc = d.B;
    if (budget > 0.0) {
      const double w1 = (double)ws_layout(d, h->sm_count, 1).total, wB = (double)ws_layout(d, h->sm_count, d.B).total;
      if (wB > budget) {
        const double per = (wB - w1) / (d.B > 1 ? d.B - 1 : 1);
        long long bc = (long long)((budget - w1) / (per > 0 ? per : 1.0)) + 1;
        if (bc < 1) bc = 1;
        if (bc > d.B) bc = d.B;
        while (bc > 1 && (double)ws_layout(d, h->sm_count, (int)bc).total > budget) --bc;
        h->Bc = (int)bc;
      }
    }
  }
  if (h->tc) {
    if (d.n > TC_MAX_N || d.K > 31 || (d.flags & QOC_FLAG_STATE_TRANSFER) || !qoc_plane_sweep_supported(d.n, d.m) ||
        !tc_geometry(d.n, &h->tg)) {
      h->err = "QOC_F16X2 (tcgen05 / TMA path) supports n <= 256, m <= 8, unitary mode in this build";
      return QOC_EINVAL;
    }
    h->tc_grid = tc_grid_for(d, h->sm_count);
    h->tc_S = (d.T + QOC_TC_SEG_LEN - 1) / QOC_TC_SEG_LEN;
  }
  if (d.dtype == QOC_TF32X3 && (d.n > 32 || d.K > 15)) {
    h->err = "QOC_TF32X3 (tcgen05 path) supports n <= 32, K <= 15 in this build";
    return QOC_EINVAL;
  }
  if (cudaMalloc((void**)&h->err_flag, sizeof(int)) != cudaSuccess || cudaMemset(h->err_flag, 0, sizeof(int)) != cudaSuccess) {
    h->err = "cudaMalloc failed"; return QOC_ECUDA;
  }
  {   // high-priority stream for the loss / gradient critical path; the U_final branch stays on the caller's stream
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&h->hi, cudaStreamNonBlocking, hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) {
      h->err = "cannot create the high-priority stream"; return QOC_ECUDA;
    }
  }
  return QOC_OK;
}

int qoc_destroy(qoc_handle_t h) {
  QOC_CHECK_H(h);
  cudaFree(h->A); cudaFree(h->U0); cudaFree(h->phi); cudaFree(h->V); cudaFree(h->coo_v);
  cudaFree(h->cidx); cudaFree(h->coo_off); cudaFree(h->coo_r); cudaFree(h->coo_c);
  cudaFree(h->tc_ops); cudaFree(h->A_f); free(h->U0_host);
  cudaFree(h->maxA); cudaFree(h->env); cudaFree(h->fw); cudaFree(h->dressW); cudaFree(h->psid); cudaFree(h->pat_rc); cudaFree(h->pat_coef); cudaFree(h->pat_coef_f); cudaFree(h->err_flag);
  for (int i = 0; i <= QOC_NUM_KERNELS; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  if (h->hi) { cudaStreamSynchronize(h->hi); cudaStreamDestroy(h->hi); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  delete h;
  return QOC_OK;
}

const char* qoc_last_error(qoc_handle_t h) { return h ? h->err.c_str() : "null handle"; }

int qoc_workspace_bytes(qoc_handle_t h, size_t* bytes) {
  QOC_CHECK_H(h);
  if (!bytes) return QOC_EINVAL;
  *bytes = ws_layout(h->d, h->sm_count, h->Bc).total;
  return QOC_OK;
}

int qoc_set_workspace(qoc_handle_t h, void* dev_ptr, size_t bytes) {
  QOC_CHECK_H(h);
  const WsLayout L = ws_layout(h->d, h->sm_count, h->Bc);
  if (!dev_ptr || ((uintptr_t)dev_ptr & 255)) { h->err = "workspace must be 256-byte aligned"; return QOC_EINVAL; }
  if (bytes < L.total) { h->err = "workspace too small"; return QOC_ENOMEM; }
  char* w = (char*)dev_ptr;
  h->ws = w; h->ws_bytes = bytes;
  h->P = w + L.P;
  h->psi = (cplx*)(w + L.psi); h->lam = (cplx*)(w + L.lam);
  h->gctrl = (double*)(w + L.gctrl); h->ot = (cplx*)(w + L.ot); h->scal = (double*)(w + L.scal);
  h->Ufin = (cplx*)(w + L.Ufin);
  h->st_base = (double*)(w + L.st_base); h->st_grad = (double*)(w + L.st_grad); h->st_out = (double*)(w + L.st_out);
  h->seg = (cplx*)(w + L.seg);
  h->scratch = w + L.scratch;
  h->tc_seg = (__half*)(w + L.tc_seg); h->tc_scr = (__half*)(w + L.tc_scr); h->tc_const = (__half*)(w + L.tc_const);
  h->tc_ready = false;
  h->ws_set = true;
  return QOC_OK;
}

}  // extern "C"

template <typename T>
static cudaError_t upload(T** dst, const void* src, size_t count, cudaStream_t st) {
  cudaError_t e = cudaSuccess;
  if (*dst) { cudaFree(*dst); *dst = nullptr; }
  if (count == 0) count = 1;
  e = cudaMalloc((void**)dst, count * sizeof(T));
  if (e != cudaSuccess) return e;
  if (src) e = cudaMemcpyAsync(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice, st);
  else e = cudaMemsetAsync(*dst, 0, count * sizeof(T), st);
  return e;
}

extern "C" {

int qoc_set_problem(qoc_handle_t h, const double* A_host, const double* U0_host, const double* phi_host,
                    const double* V_host, const int32_t* concerned_idx, const double* maxA_host, double dt,
                    void* stream) {
  QOC_CHECK_H(h);
  if (!A_host || !U0_host || !phi_host || !V_host || (h->d.K > 0 && !maxA_host)) { h->err = "null problem array"; return QOC_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  const qoc_dims_t& d = h->d;
  const size_t nn = (size_t)d.n * d.n, mn = (size_t)d.m * d.n;
  if (concerned_idx)
    for (int j = 0; j < d.m; ++j)
      if (concerned_idx[j] < 0 || concerned_idx[j] >= d.n) { h->err = "concerned_idx out of range"; return QOC_EINVAL; }
  // sparse (COO) form of the control operators A_1..A_K for the gradient kernel
  std::vector<int> off(d.K + 1, 0), rr, cc;
  std::vector<double> vv;
  for (int k = 0; k < d.K; ++k) {
    const double* Ak = A_host + (size_t)(k + 1) * nn * 2;
    for (int r = 0; r < d.n; ++r)
      for (int c = 0; c < d.n; ++c) {
        const double re = Ak[((size_t)r * d.n + c) * 2], im = Ak[((size_t)r * d.n + c) * 2 + 1];
        if (re != 0.0 || im != 0.0) { rr.push_back(r); cc.push_back(c); vv.push_back(re); vv.push_back(im); }
      }
    off[k + 1] = (int)rr.size();
  }
  h->nnz = (int)rr.size();
  // anti-Hermitian generators (Hermitian Hamiltonians) enable the triangle-only Taylor evaluation
  h->herm = 1;
  for (int k = 0; k <= d.K && h->herm; ++k) {
    const double* Ak = A_host + (size_t)k * nn * 2;
    double amax = 0.0;
    for (size_t i = 0; i < nn * 2; ++i) amax = fabs(Ak[i]) > amax ? fabs(Ak[i]) : amax;
    const double tol = 1e-13 * (amax > 1.0 ? amax : 1.0);
    for (int r = 0; r < d.n && h->herm; ++r)
      for (int c = r; c < d.n; ++c) {
        const double* x = Ak + ((size_t)r * d.n + c) * 2;
        const double* y = Ak + ((size_t)c * d.n + r) * 2;
        if (fabs(x[0] + y[0]) > tol || fabs(x[1] - y[1]) > tol) { h->herm = 0; break; }
      }
  }
  // union sparsity pattern of A_0..A_K with per-entry coefficient vectors, for the H assembly
  std::vector<int> prc;
  std::vector<double> pcf;
  for (int r = 0; r < d.n; ++r)
    for (int c = 0; c < d.n; ++c) {
      bool any = false;
      for (int k = 0; k <= d.K && !any; ++k) {
        const double* e = A_host + ((size_t)k * nn + (size_t)r * d.n + c) * 2;
        any = e[0] != 0.0 || e[1] != 0.0;
      }
      if (!any) continue;
      prc.push_back((r << 16) | c);
      for (int k = 0; k <= d.K; ++k) {
        const double* e = A_host + ((size_t)k * nn + (size_t)r * d.n + c) * 2;
        pcf.push_back(e[0]); pcf.push_back(e[1]);
      }
    }
  h->pat_n = (int)prc.size();
  CUDA_TRY(h, upload(&h->pat_rc, prc.data(), prc.size(), st));
  {   // fp64 table k-major ([K+1][pat_n]) so that the H assembly of k_expm_mma reads it coalesced
    const size_t pn = prc.size();
    std::vector<double> pkm(pcf.size());
    for (size_t e = 0; e < pn; ++e)
      for (int k = 0; k <= d.K; ++k) {
        pkm[((size_t)k * pn + e) * 2] = pcf[(e * (d.K + 1) + k) * 2];
        pkm[((size_t)k * pn + e) * 2 + 1] = pcf[(e * (d.K + 1) + k) * 2 + 1];
      }
    CUDA_TRY(h, upload(&h->pat_coef, pkm.data(), pkm.size() / 2, st));
  }
  std::vector<float> pcf32(pcf.begin(), pcf.end());
  CUDA_TRY(h, upload(&h->pat_coef_f, pcf32.data(), pcf32.size() / 2, st));
  CUDA_TRY(h, upload(&h->A, A_host, (size_t)(d.K + 1) * nn, st));
  CUDA_TRY(h, upload(&h->U0, U0_host, nn, st));
  CUDA_TRY(h, upload(&h->phi, phi_host, mn, st));
  CUDA_TRY(h, upload(&h->V, V_host, mn, st));
  CUDA_TRY(h, upload(&h->cidx, concerned_idx, (size_t)d.m, st));
  CUDA_TRY(h, upload(&h->maxA, maxA_host, (size_t)d.K, st));
  CUDA_TRY(h, upload(&h->coo_off, off.data(), off.size(), st));
  CUDA_TRY(h, upload(&h->coo_r, rr.data(), rr.size(), st));
  CUDA_TRY(h, upload(&h->coo_c, cc.data(), cc.size(), st));
  CUDA_TRY(h, upload(&h->coo_v, vv.data(), vv.size() / 2, st));
  if (h->tc) {
    if (!h->herm) { h->err = "QOC_F16X2 needs Hermitian H0 / Hops (anti-Hermitian generators)"; return QOC_EINVAL; }
    // entrywise bound |X| <= (|A_0| + sum_k maxA_k |A_k|) / 2^s and a bound of its 2-norm pick the power-of-two scales
    std::vector<double> bnd(nn, 0.0);
    for (int k = 0; k <= d.K; ++k) {
      const double w = k == 0 ? 1.0 : fabs(maxA_host[k - 1]);
      for (size_t i = 0; i < nn; ++i) bnd[i] += w * hypot(A_host[((size_t)k * nn + i) * 2], A_host[((size_t)k * nn + i) * 2 + 1]);
    }
    double xmax = 0.0, fro = 0.0, n1 = 0.0, ninf = 0.0;
    std::vector<double> colsum(d.n, 0.0);
    for (int r = 0; r < d.n; ++r) {
      double rs = 0.0;
      for (int c = 0; c < d.n; ++c) { const double v = bnd[(size_t)r * d.n + c]; xmax = v > xmax ? v : xmax; fro += v * v; rs += v; colsum[c] += v; }
      ninf = rs > ninf ? rs : ninf;
    }
    for (int c = 0; c < d.n; ++c) n1 = colsum[c] > n1 ? colsum[c] : n1;
    const double inv2s = ldexp(1.0, -d.scaling);
    double theta = sqrt(n1 * ninf);
    if (sqrt(fro) < theta) theta = sqrt(fro);
    int eX, eY;
    tc_pick_scales(xmax * inv2s, theta * inv2s, &eX, &eY);
    std::vector<TcExpmOp> ops;
    tc_build_expm_ops(d.exp_terms, d.scaling, eX, eY, ops);
    h->tc_nops = (int)ops.size();
    h->tc_xscale = (float)ldexp(1.0, eX - d.scaling);
    CUDA_TRY(h, upload(&h->tc_ops, ops.data(), ops.size(), st));
    std::vector<float> af((size_t)(d.K + 1) * nn * 2);
    for (size_t i = 0; i < af.size(); ++i) af[i] = (float)A_host[i];
    CUDA_TRY(h, upload(&h->A_f, af.data(), af.size() / 2, st));
    free(h->U0_host);
    h->U0_host = (double*)malloc(nn * 2 * sizeof(double));
    if (!h->U0_host) return QOC_ENOMEM;
    std::memcpy(h->U0_host, U0_host, nn * 2 * sizeof(double));
    h->tc_ready = false;
  }
  CUDA_TRY(h, cudaStreamSynchronize(st));       // host vectors above go out of scope
  h->has_cidx = concerned_idx ? 1 : 0;
  h->dt = dt;
  h->problem_set = true;
  return QOC_OK;
}

int qoc_set_regularizers(qoc_handle_t h, const qoc_reg_t* reg, const double* envelope_host,
                         const double* forbid_weight_host, void* stream) {
  QOC_CHECK_H(h);
  cudaStream_t st = (cudaStream_t)stream;
  if (!reg) { std::memset(&h->reg, 0, sizeof(h->reg)); return QOC_OK; }
  if (reg->has_envelope && !envelope_host) { h->err = "envelope array required"; return QOC_EINVAL; }
  if (reg->has_forbidden && !forbid_weight_host) { h->err = "forbid_weight array required"; return QOC_EINVAL; }
  if (reg->has_d2wdt2 && !reg->has_dwdt) {
    // the reference raises NameError here (regularization_functions.py:30 vs :41); the Python layer
    // reproduces that, the C layer just refuses.
    h->err = "'d2wdt2' requires 'dwdt'"; return QOC_EINVAL;
  }
  h->reg = *reg;
  if (reg->has_envelope) CUDA_TRY(h, upload(&h->env, envelope_host, (size_t)h->d.K * h->d.T, st));
  if (reg->has_forbidden) CUDA_TRY(h, upload(&h->fw, forbid_weight_host, (size_t)h->d.n, st));
  CUDA_TRY(h, cudaStreamSynchronize(st));
  return QOC_OK;
}

extern "C" int qoc_set_forbid_basis(qoc_handle_t h, const double* W_host, void* stream) {
  QOC_CHECK_H(h);
  cudaStream_t st = (cudaStream_t)stream;
  if (!W_host) { cudaFree(h->dressW); h->dressW = nullptr; return QOC_OK; }
  const qoc_dims_t& d = h->d;
  CUDA_TRY(h, upload(&h->dressW, W_host, (size_t)d.n * d.n, st));
  if (!h->psid) CUDA_TRY(h, cudaMalloc((void**)&h->psid, (size_t)h->Bc * (d.T + 1) * d.m * d.n * sizeof(cplx)));
  CUDA_TRY(h, cudaStreamSynchronize(st));
  return QOC_OK;
}

static int fill_params(qoc_handle_t h, QocParams& p, const double* base) {
  if (!h->ws_set) { h->err = "qoc_set_workspace not called"; return QOC_ESTATE; }
  if (!h->problem_set) { h->err = "qoc_set_problem not called"; return QOC_ESTATE; }
  if (!base) { h->err = "base is null"; return QOC_EINVAL; }
  const qoc_dims_t& d = h->d;
  std::memset(&p, 0, sizeof(p));
  p.n = d.n; p.K = d.K; p.T = d.T; p.m = d.m; p.B = d.B; p.p = d.exp_terms; p.s = d.scaling;
  p.state_transfer = (d.flags & QOC_FLAG_STATE_TRANSFER) ? 1 : 0;
  if (p.state_transfer) { p.p = d.exp_terms - 1; p.s = 0; }      // order p-1, no squaring (tensorflow_state.py:92)
  p.has_cidx = h->has_cidx;
  p.herm = h->herm;
  p.dt = h->dt; p.inv2s = 1.0 / (double)(1ull << p.s);
  { double f = 1.0; p.invfact[0] = 1.0; for (int j = 1; j < 32; ++j) { f *= (double)j; p.invfact[j] = 1.0 / f; } }
  p.A = h->A; p.U0 = h->U0; p.phi = h->phi; p.V = h->V; p.cidx = h->cidx; p.maxA = h->maxA;
  p.env = h->env; p.fw = h->fw;
  p.dressW = (h->reg.has_forbidden && h->fw) ? h->dressW : nullptr; p.psid = h->psid;
  p.coo_off = h->coo_off; p.coo_r = h->coo_r; p.coo_c = h->coo_c; p.coo_v = h->coo_v;
  p.pat_n = h->pat_n; p.pat_rc = h->pat_rc; p.pat_coef = h->pat_coef; p.pat_coef_f = h->pat_coef_f;
  p.reg = h->reg;
  p.base = base;
  p.P = h->P; p.psi = h->psi; p.lam = h->lam; p.gctrl = h->gctrl; p.ot = h->ot; p.scal = h->scal; p.Ufin = h->Ufin;
  return QOC_OK;
}

// per-kernel CUDA-event timing (qoc_set_profiling): event i is recorded before kernel i, event
// QOC_NUM_KERNELS after the last one
static int prof_mark(qoc_handle_t h, int i, cudaStream_t st) {
  if (!h->profiling) return QOC_OK;
  CUDA_TRY(h, cudaEventRecord(h->ev[i], st));
  if (i + 1 > h->ev_recorded) h->ev_recorded = i + 1;
  return QOC_OK;
}

// restrict the parameter block to instances [b0, b0 + bc): per-instance user arrays and the persistent
// scal / Ufin arrays are offset, the chunk-sized workspace arrays are reused from their start
static QocParams chunk_params(const QocParams& p, const qoc_dims_t& d, int b0, int bc) {
  QocParams q = p;
  q.B = bc;
  q.base = p.base + (size_t)b0 * d.K * d.T;
  q.scal = p.scal + (size_t)b0 * 8;
  q.Ufin = p.Ufin + (size_t)b0 * d.n * d.n;
  if (p.loss) q.loss = p.loss + b0;
  if (p.reg_loss) q.reg_loss = p.reg_loss + b0;
  if (p.grad) q.grad = p.grad + (size_t)b0 * d.K * d.T;
  if (p.unitary_scale) q.unitary_scale = p.unitary_scale + b0;
  if (p.grad_squared) q.grad_squared = p.grad_squared + b0;
  return q;
}

// Few concerned states (m < NP/2) on the fp64 shared-memory path: the loss and the gradient only need
// the m state columns, which k_vec_sweep propagates (HBM-bound) on the handle's high-priority stream;
// the full n x n product -- U_final / unitary_scale, wanted by the caller but by nothing downstream --
// is re-associated into segment products (k_segprod) + a short chain and runs on the caller's stream
// beside it.  QOC_B200_NO_VEC_SWEEP=1 restores the single-stream chain.
static bool use_vec_sweeps(qoc_handle_t h, const QocParams& p) {
  const bool off = getenv("QOC_B200_NO_VEC_SWEEP") != nullptr;     // read per call: tests flip it
  if (off || 2 * h->d.m >= h->NP || !qoc_vec_sweep_supported(p)) return false;
  return h->d.dtype == QOC_F64 ? h->d.n <= 64 : h->d.n <= 32;      // QOC_TF32X3: fp32 32 x 32 propagator tiles
}

// the caller's stream waits for the critical-path branch
static int join_hi(qoc_handle_t h, cudaStream_t st) {
  if (!h->hi_pending) return QOC_OK;
  h->hi_pending = false;
  CUDA_TRY(h, cudaEventRecord(h->ev_join, h->hi));
  CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_join, 0));
  h->work = st;
  return QOC_OK;
}

// U_final / unitary_scale without the states: segment products + chain over the segments (T >= 4 segments),
// else the plain chain
static int launch_xchain(qoc_handle_t h, const QocParams& p, cudaStream_t st) {
  QocParams q = p;
  q.chain_no_psi = 1;
  int L = QOC_SEG_LEN;
  if (getenv("QOC_B200_SEG_LEN")) { L = atoi(getenv("QOC_B200_SEG_LEN")); if (L < QOC_SEG_LEN) L = QOC_SEG_LEN; }   // experiments: longer segments only
  const int S = (p.T + L - 1) / L;
  bool pf32 = h->d.dtype != QOC_F64;                // fp32 propagator tiles (tcgen05 path); the segment matrices are fp64
  if (S >= 4) {
    CUDA_TRY(h, qoc_launch_segprod_f64(p, h->NP, pf32 ? 1 : 0, L, S, h->seg, st, &h->launches));
    q.P = h->seg; q.T = S;
    pf32 = false;
  }
  CUDA_TRY(h, qoc_launch_chain_f64(q, h->NP, pf32 ? 1 : 0, st, &h->launches));
  return QOC_OK;
}

// QOC_F16X2: constant plane sets (U0, I) and the TMA descriptors need both the problem and the workspace
static int tc_prepare(qoc_handle_t h, cudaStream_t st) {
  if (h->tc_ready) return QOC_OK;
  const qoc_dims_t& d = h->d;
  const TcGeom& g = h->tg;
  std::vector<__half> hbuf(2 * g.mat_halfs);
  std::vector<double> I((size_t)d.n * d.n * 2, 0.0);
  for (int i = 0; i < d.n; ++i) I[((size_t)i * d.n + i) * 2] = 1.0;
  tc_pack_host(h->U0_host, d.n, g.ld, TC_EU, hbuf.data());
  tc_pack_host(I.data(), d.n, g.ld, TC_EU, hbuf.data() + g.mat_halfs);
  CUDA_TRY(h, cudaMemcpyAsync(h->tc_const, hbuf.data(), hbuf.size() * sizeof(__half), cudaMemcpyHostToDevice, st));
  // the generator slots of the scratch are only ever written on the union sparsity pattern of the A_k: zero everything once
  CUDA_TRY(h, cudaMemsetAsync(h->tc_scr, 0, (size_t)h->tc_grid * QOC_TC_ILV * TC_NSLOT * g.mat_halfs * sizeof(__half), st));
  CUDA_TRY(h, cudaStreamSynchronize(st));
  const void* base[TC_NCLS] = {h->tc_scr, h->P, h->tc_seg, h->tc_const};
  const unsigned long long cnt[TC_NCLS] = {(unsigned long long)h->tc_grid * QOC_TC_ILV * TC_NSLOT, (unsigned long long)h->Bc * d.T,
                                           (unsigned long long)h->Bc * h->tc_S, 2ull};
  for (int c = 0; c < TC_NCLS; ++c) {
    const char* e = tc_make_map(&h->tmaps.a[c], base[c], d.n, g.ld, cnt[c], false);
    if (!e) e = tc_make_map(&h->tmaps.b[c], base[c], d.n, g.ld, cnt[c], true);
    if (!e) e = tc_make_store_map(&h->tsmaps.st[c], base[c], d.n, g.ld, cnt[c]);
    if (e) { h->err = e; return QOC_ECUDA; }
  }
  h->tc_ready = true;
  return QOC_OK;
}

static void tc_base_params(qoc_handle_t h, TcParams& q) {
  std::memset(&q, 0, sizeof(q));
  q.base[TC_CLS_SCR] = h->tc_scr; q.base[TC_CLS_P] = (__half*)h->P; q.base[TC_CLS_SEG] = h->tc_seg; q.base[TC_CLS_CONST] = h->tc_const;
  q.err_flag = h->err_flag;
}

static int tc_launch_expm(qoc_handle_t h, const QocParams& p, cudaStream_t st) {
  TcParams q;
  tc_base_params(h, q);
  q.prog = TC_PROG_EXPM; q.items = (long long)p.B * p.T;
  q.ilv = p.n > 128 ? 1 : QOC_TC_ILV;               // n > 128: four tiles per product already decouple MMA and epilogue, and one
                                                    // item per CTA keeps the scratch working set (148 x 1.5 MB) near the L2 size
  q.nops = h->tc_nops; q.ops = h->tc_ops; q.K = p.K; q.T = p.T; q.ctrl = p.base; q.maxA = p.maxA; q.A_f = h->A_f; q.xscale = h->tc_xscale;
  if (!getenv("QOC_B200_TC_DENSE_X")) { q.pat_n = h->pat_n; q.pat_rc = h->pat_rc; q.pat_coef_f = h->pat_coef_f; }   // sparse generators: scatter the pattern
  ++h->launches;
  const int small_on = getenv("QOC_B200_TC_SMALL") ? atoi(getenv("QOC_B200_TC_SMALL")) : 1;     // read per call: tests flip it
  if (small_on && tc_small_supported(p.n)) {       // n <= 64: operands resident in shared memory (qoc_tc_small.cu)
    CUDA_TRY(h, tc_small_launch_expm(q, p.n, h->sm_count, st));
    return QOC_OK;
  }
  const int pair_on = getenv("QOC_B200_TC_PAIR") ? atoi(getenv("QOC_B200_TC_PAIR")) : 1;
  if (pair_on && tc_pair_supported(p.n) && tc_pair_max_clusters(2) > 0) {
    // 128 < n <= 256: CTA pairs (tcgen05 cta_group::2, M = 256), two items interleaved per pair, TMA-store epilogue (qoc_tc_pair.cu)
    q.ilv = QOC_TC_ILV;
    q.tma_store = 1;
    CUDA_TRY(h, tc_pair_launch_expm(q, h->tmaps, h->tsmaps, h->tg, 2, st));
    return QOC_OK;
  }
  {
    const long long rounds = (q.items + q.ilv - 1) / q.ilv;
    CUDA_TRY(h, tc_launch(q, h->tmaps, h->tg, (int)(rounds < h->tc_grid ? rounds : h->tc_grid), st));
  }
  return QOC_OK;
}

// U_final / unitary_scale: segment products (16 propagators each) + the chain over the segments, all on tcgen05
static int tc_launch_xchain(qoc_handle_t h, const QocParams& p, cudaStream_t st) {
  const int seg64 = getenv("QOC_B200_TC_SEG_F64") ? atoi(getenv("QOC_B200_TC_SEG_F64")) : 1;
  const int S64 = (p.T + QOC_SEG_LEN - 1) / QOC_SEG_LEN;
  if (seg64 && p.n <= 64 && S64 >= 4) {
    // n <= 64: a product of this size leaves the streamed-operand engine idle most of the time (one 128-row MMA tile per
    // 36 rows, a TMA round trip per product); the DMMA segment kernel reads the plane sets straight from the cache
    // (widening h0 + h1 is exact) and runs beside the state sweeps; the segment matrices and the chain are fp64
    QocParams q = p;
    q.chain_no_psi = 1;
    CUDA_TRY(h, qoc_launch_segprod_f64(p, h->NP, 2, QOC_SEG_LEN, S64, h->seg, st, &h->launches));
    q.P = h->seg; q.T = S64;
    CUDA_TRY(h, qoc_launch_chain_f64(q, h->NP, 0, st, &h->launches));
    return QOC_OK;
  }
  TcParams q;
  tc_base_params(h, q);
  const int L = QOC_TC_SEG_LEN, S = (p.T + L - 1) / L;
  if (S >= 2) {
    q.prog = TC_PROG_SEG; q.items = (long long)p.B * S; q.T = p.T; q.L = L; q.S = S;
    ++h->launches;
    const int pair_on = getenv("QOC_B200_TC_PAIR") ? atoi(getenv("QOC_B200_TC_PAIR")) : 1;
    if (pair_on && tc_pair_supported(p.n) && p.T % L == 0 && tc_pair_max_clusters(2) > 0) {
      q.ilv = QOC_TC_ILV; q.tma_store = 1;          // equal-length segments: the CTA-pair kernel, two segments interleaved
      CUDA_TRY(h, tc_pair_launch_expm(q, h->tmaps, h->tsmaps, h->tg, 2, st));
    } else
    CUDA_TRY(h, tc_launch(q, h->tmaps, h->tg, (int)(q.items < h->tc_grid ? q.items : h->tc_grid), st));
  }
  tc_base_params(h, q);
  q.prog = TC_PROG_CHAIN; q.items = p.B;
  q.chain_cls = S >= 2 ? TC_CLS_SEG : TC_CLS_P; q.chain_len = S >= 2 ? S : p.T;
  q.Ufin = p.Ufin; q.scal = p.scal;
  ++h->launches;
  CUDA_TRY(h, tc_launch(q, h->tmaps, h->tg, (int)(q.items < h->tc_grid ? q.items : h->tc_grid), st));
  return QOC_OK;
}

// Forward pass.  On return h->work is the stream the caller must use for everything that consumes psi
// (fwd_reduce has run on it); join_hi() brings the caller's stream back in.
static int run_forward(qoc_handle_t h, const QocParams& p, cudaStream_t st) {
  int rc;
  h->ev_recorded = 0;
  if ((rc = join_hi(h, st))) return rc;
  h->work = st;
  if ((rc = prof_mark(h, 0, st))) return rc;
  if (h->tc) {
    // propagators on tcgen05 (TMA-fed fp16-pair tiles); then the two-branch tail of the few-state problems: the state
    // sweep over the split-plane propagators on the high-priority stream, the U_final products on tcgen05 beside it
    if ((rc = tc_prepare(h, st))) return rc;
    if ((rc = tc_launch_expm(h, p, st))) return rc;
    if ((rc = prof_mark(h, 1, st))) return rc;
    static const int xmode = getenv("QOC_B200_XCHAIN") ? atoi(getenv("QOC_B200_XCHAIN")) : 2;
    if (xmode == 2) {
      CUDA_TRY(h, cudaEventRecord(h->ev_fork, st));
      CUDA_TRY(h, cudaStreamWaitEvent(h->hi, h->ev_fork, 0));
      h->work = h->hi;
      h->hi_pending = true;
    }
    CUDA_TRY(h, qoc_launch_plane_sweep(p, h->P, 0, h->work, &h->launches));
    if (xmode != 0 && (rc = tc_launch_xchain(h, p, st))) return rc;
    cudaStream_t ws = h->work;
    if ((rc = prof_mark(h, 2, ws))) return rc;
    if (p.dressW) CUDA_TRY(h, qoc_launch_dress(p, 0, ws, &h->launches));
    CUDA_TRY(h, qoc_launch_fwd_reduce(p, ws, &h->launches));
    if ((rc = prof_mark(h, 3, ws))) return rc;
    return QOC_OK;
  }
  const bool large = h->d.n > 64;                   // matrices do not fit in shared memory: tiled global-operand path
  if (h->d.dtype == QOC_TF32X3) CUDA_TRY(h, qoc_launch_expm_tc32(p, h->sm_count, h->err_flag, st, &h->launches));
  else if (large) CUDA_TRY(h, qoc_launch_expm_large(p, h->sm_count, h->scratch, st, &h->launches));
  else CUDA_TRY(h, qoc_launch_expm_f64(p, h->NP, h->sm_count, st, &h->launches));
  if ((rc = prof_mark(h, 1, st))) return rc;
  if (large) CUDA_TRY(h, qoc_launch_chain_large(p, h->scratch, st, &h->launches));
  else if (use_vec_sweeps(h, p)) {
    static const int xmode = getenv("QOC_B200_XCHAIN") ? atoi(getenv("QOC_B200_XCHAIN")) : 2;   // debug: 0 skip, 1 serial
    if (xmode == 2) {
      CUDA_TRY(h, cudaEventRecord(h->ev_fork, st));                    // both branches start when the propagators are done
      CUDA_TRY(h, cudaStreamWaitEvent(h->hi, h->ev_fork, 0));
      h->work = h->hi;
      h->hi_pending = true;
    }
    CUDA_TRY(h, qoc_launch_vec_sweep(p, 0, h->d.dtype != QOC_F64, h->work, &h->launches));   // critical path first: psi(t)
    if (!p.state_transfer && xmode != 0 && (rc = launch_xchain(h, p, st))) return rc;
  } else CUDA_TRY(h, qoc_launch_chain_f64(p, h->NP, h->d.dtype != QOC_F64, st, &h->launches));
  cudaStream_t ws = h->work;
  if ((rc = prof_mark(h, 2, ws))) return rc;
  if (p.dressW) CUDA_TRY(h, qoc_launch_dress(p, 0, ws, &h->launches));
  CUDA_TRY(h, qoc_launch_fwd_reduce(p, ws, &h->launches));
  if ((rc = prof_mark(h, 3, ws))) return rc;
  return QOC_OK;
}

int qoc_value_and_grad(qoc_handle_t h, const double* base_dev, double* loss_dev, double* reg_loss_dev,
                       double* grad_dev, double* unitary_scale_dev, double* grad_squared_dev, void* stream) {
  QOC_CHECK_H(h);
  QocParams p;
  int rc = fill_params(h, p, base_dev);
  if (rc) return rc;
  if (!grad_dev) { h->err = "grad is null"; return QOC_EINVAL; }
  p.loss = loss_dev; p.reg_loss = reg_loss_dev; p.grad = grad_dev;
  p.unitary_scale = unitary_scale_dev; p.grad_squared = grad_squared_dev;
  cudaStream_t st = (cudaStream_t)stream;
  const QocParams full = p;
  for (int b0 = 0; b0 < h->d.B; b0 += h->Bc) {           // one pass per batch chunk (a single pass when everything fits)
    p = chunk_params(full, h->d, b0, h->d.B - b0 < h->Bc ? h->d.B - b0 : h->Bc);
    rc = run_forward(h, p, st);
    if (rc) return rc;
    cudaStream_t ws = h->work;
    if (p.dressW) CUDA_TRY(h, qoc_launch_dress(p, 1, ws, &h->launches));
    // dense-m problems (m >= NP/2, dense controls) take the DMMA costate / gradient kernels
    const bool dense_m = h->d.n <= 64 && h->d.dtype == QOC_F64 && 2 * h->d.m >= h->NP && h->d.m <= h->NP;
    const bool dense_A = (double)h->nnz >= 0.25 * (double)h->d.K * h->d.n * h->d.n;
    if (h->tc) CUDA_TRY(h, qoc_launch_plane_sweep(p, h->P, 1, ws, &h->launches));
    else if (h->d.n > 64) CUDA_TRY(h, qoc_launch_costate_large(p, ws, &h->launches));
    else if (dense_m) CUDA_TRY(h, qoc_launch_costate_mma(p, h->NP, ws, &h->launches));
    else if (use_vec_sweeps(h, p)) CUDA_TRY(h, qoc_launch_vec_sweep(p, 1, h->d.dtype != QOC_F64, ws, &h->launches));
    else CUDA_TRY(h, qoc_launch_costate(p, h->d.dtype != QOC_F64, ws, &h->launches));
    if ((rc = prof_mark(h, 4, ws))) return rc;
    // n > 64: the GEMM form pays when the K operators together hold more entries than one n x n matrix per m / 8 states
    const bool gemm_grad_large = !h->tc && h->d.n > 64 && h->d.dtype == QOC_F64 && qoc_grad_large_supported(p) && !getenv("QOC_B200_NO_GRAD_GEMM") &&
                                 (double)h->nnz * h->d.m >= 8.0 * (double)h->d.n * h->d.n * 1.5;
    if (dense_m && dense_A) CUDA_TRY(h, qoc_launch_grad_mma(p, h->NP, h->sm_count, ws, &h->launches));
    else if (gemm_grad_large) CUDA_TRY(h, qoc_launch_grad_large(p, h->sm_count, ws, &h->launches));
    else CUDA_TRY(h, qoc_launch_grad(p, h->sm_count, ws, &h->launches));
    if ((rc = prof_mark(h, 5, ws))) return rc;
    if ((rc = join_hi(h, st))) return rc;                  // finalize reads both branches; the next pass rewrites P
    CUDA_TRY(h, qoc_launch_finalize(p, st, &h->launches));
    if ((rc = prof_mark(h, 6, st))) return rc;
  }
  return QOC_OK;
}

int qoc_evolve(qoc_handle_t h, const double* base_dev, double* U_final_dev, double* inter_vecs_dev,
               double* loss_dev, double* unitary_scale_dev, void* stream) {
  QOC_CHECK_H(h);
  QocParams p;
  int rc = fill_params(h, p, base_dev);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const qoc_dims_t& d = h->d;
  const QocParams full = p;
  for (int b0 = 0; b0 < d.B; b0 += h->Bc) {
    const int bc = d.B - b0 < h->Bc ? d.B - b0 : h->Bc;
    p = chunk_params(full, d, b0, bc);
    rc = run_forward(h, p, st);
    if (rc) return rc;
    if ((rc = join_hi(h, st))) return rc;
    if (inter_vecs_dev) {
      const size_t per = (size_t)(d.T + 1) * d.m * d.n * 2;      // doubles per instance
      CUDA_TRY(h, cudaMemcpyAsync(inter_vecs_dev + (size_t)b0 * per, h->psi, (size_t)bc * per * sizeof(double),
                                  cudaMemcpyDeviceToDevice, st));
    }
  }
  if (U_final_dev)
    CUDA_TRY(h, cudaMemcpyAsync(U_final_dev, h->Ufin, (size_t)d.B * d.n * d.n * sizeof(cplx), cudaMemcpyDeviceToDevice, st));
  if (loss_dev)
    CUDA_TRY(h, cudaMemcpy2DAsync(loss_dev, sizeof(double), h->scal + 2, 8 * sizeof(double), sizeof(double), d.B,
                                  cudaMemcpyDeviceToDevice, st));
  if (unitary_scale_dev)
    CUDA_TRY(h, cudaMemcpy2DAsync(unitary_scale_dev, sizeof(double), h->scal + 5, 8 * sizeof(double), sizeof(double),
                                  d.B, cudaMemcpyDeviceToDevice, st));
  return QOC_OK;
}

int qoc_value_and_grad_host(qoc_handle_t h, const double* base_host, double* loss_host, double* reg_loss_host,
                            double* grad_host, double* unitary_scale_host, double* grad_squared_host,
                            void* stream) {
  QOC_CHECK_H(h);
  if (!h->ws_set) { h->err = "qoc_set_workspace not called"; return QOC_ESTATE; }
  if (!base_host || !grad_host) { h->err = "null host buffer"; return QOC_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  const qoc_dims_t& d = h->d;
  const size_t nb = (size_t)d.B * d.K * d.T * sizeof(double), sb = (size_t)d.B * sizeof(double);
  CUDA_TRY(h, cudaMemcpyAsync(h->st_base, base_host, nb, cudaMemcpyHostToDevice, st));
  double* o = h->st_out;
  int rc = qoc_value_and_grad(h, h->st_base, o, o + d.B, h->st_grad, o + 2 * d.B, o + 3 * d.B, stream);
  if (rc) return rc;
  CUDA_TRY(h, cudaMemcpyAsync(grad_host, h->st_grad, nb, cudaMemcpyDeviceToHost, st));
  if (loss_host) CUDA_TRY(h, cudaMemcpyAsync(loss_host, o, sb, cudaMemcpyDeviceToHost, st));
  if (reg_loss_host) CUDA_TRY(h, cudaMemcpyAsync(reg_loss_host, o + d.B, sb, cudaMemcpyDeviceToHost, st));
  if (unitary_scale_host) CUDA_TRY(h, cudaMemcpyAsync(unitary_scale_host, o + 2 * d.B, sb, cudaMemcpyDeviceToHost, st));
  if (grad_squared_host) CUDA_TRY(h, cudaMemcpyAsync(grad_squared_host, o + 3 * d.B, sb, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(h, cudaStreamSynchronize(st));
  return QOC_OK;
}

int qoc_evolve_host(qoc_handle_t h, const double* base_host, double* U_final_host, double* inter_vecs_host,
                    double* loss_host, double* unitary_scale_host, void* stream) {
  QOC_CHECK_H(h);
  if (!h->ws_set) { h->err = "qoc_set_workspace not called"; return QOC_ESTATE; }
  if (!base_host) { h->err = "null host buffer"; return QOC_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  const qoc_dims_t& d = h->d;
  const size_t nb = (size_t)d.B * d.K * d.T * sizeof(double);
  CUDA_TRY(h, cudaMemcpyAsync(h->st_base, base_host, nb, cudaMemcpyHostToDevice, st));
  double* o = h->st_out;
  int rc = QOC_OK;
  if (h->Bc >= d.B || !inter_vecs_host) {
    rc = qoc_evolve(h, h->st_base, nullptr, nullptr, o, o + d.B, stream);
    if (rc) return rc;
    if (inter_vecs_host)
      CUDA_TRY(h, cudaMemcpyAsync(inter_vecs_host, h->psi, (size_t)d.B * (d.T + 1) * d.m * d.n * sizeof(cplx),
                                  cudaMemcpyDeviceToHost, st));
  } else {                                               // chunked: stream each pass's states to the host
    QocParams p;
    rc = fill_params(h, p, h->st_base);
    if (rc) return rc;
    const QocParams full = p;
    const size_t per = (size_t)(d.T + 1) * d.m * d.n * 2;
    for (int b0 = 0; b0 < d.B; b0 += h->Bc) {
      const int bc = d.B - b0 < h->Bc ? d.B - b0 : h->Bc;
      p = chunk_params(full, d, b0, bc);
      rc = run_forward(h, p, st);
      if (rc) return rc;
      if ((rc = join_hi(h, st))) return rc;
      CUDA_TRY(h, cudaMemcpyAsync(inter_vecs_host + (size_t)b0 * per, h->psi, (size_t)bc * per * sizeof(double),
                                  cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(h, cudaMemcpy2DAsync(o, sizeof(double), h->scal + 2, 8 * sizeof(double), sizeof(double), d.B, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(h, cudaMemcpy2DAsync(o + d.B, sizeof(double), h->scal + 5, 8 * sizeof(double), sizeof(double), d.B, cudaMemcpyDeviceToDevice, st));
  }
  if (U_final_host)
    CUDA_TRY(h, cudaMemcpyAsync(U_final_host, h->Ufin, (size_t)d.B * d.n * d.n * sizeof(cplx), cudaMemcpyDeviceToHost, st));
  if (loss_host) CUDA_TRY(h, cudaMemcpyAsync(loss_host, o, d.B * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (unitary_scale_host)
    CUDA_TRY(h, cudaMemcpyAsync(unitary_scale_host, o + d.B, d.B * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(h, cudaStreamSynchronize(st));
  return QOC_OK;
}

int qoc_debug_propagators(qoc_handle_t h, void** P_dev, int* elem_bytes) {
  QOC_CHECK_H(h);
  if (!h->ws_set) return QOC_ESTATE;
  if (P_dev) *P_dev = h->P;
  if (elem_bytes) *elem_bytes = h->d.dtype == QOC_F64 ? (int)sizeof(cplx) : h->tc ? (int)sizeof(__half) : (int)sizeof(float2);
  return QOC_OK;
}

int64_t qoc_launch_count(qoc_handle_t h) { return h ? h->launches : -1; }

int qoc_batch_chunk(qoc_handle_t h) { return h ? h->Bc : -1; }

int qoc_poll_error(qoc_handle_t h, void* stream) {
  QOC_CHECK_H(h);
  int flag = 0;
  CUDA_TRY(h, cudaStreamSynchronize((cudaStream_t)stream));
  CUDA_TRY(h, cudaMemcpy(&flag, h->err_flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag) { h->err = "tcgen05 pipeline timed out waiting for a TMA / MMA completion barrier"; return QOC_ECUDA; }
  return QOC_OK;
}

int qoc_set_profiling(qoc_handle_t h, int enable) {
  QOC_CHECK_H(h);
  if (enable && !h->ev[0])
    for (int i = 0; i <= QOC_NUM_KERNELS; ++i) CUDA_TRY(h, cudaEventCreate(&h->ev[i]));
  h->profiling = enable != 0;
  h->ev_recorded = 0;
  return QOC_OK;
}

int qoc_kernel_times_ms(qoc_handle_t h, float* ms_out) {
  QOC_CHECK_H(h);
  if (!ms_out) return QOC_EINVAL;
  for (int i = 0; i < QOC_NUM_KERNELS; ++i) ms_out[i] = 0.f;
  if (!h->profiling || h->ev_recorded < 2) return QOC_ESTATE;
  CUDA_TRY(h, cudaEventSynchronize(h->ev[h->ev_recorded - 1]));
  for (int i = 0; i + 1 < h->ev_recorded; ++i) CUDA_TRY(h, cudaEventElapsedTime(&ms_out[i], h->ev[i], h->ev[i + 1]));
  return QOC_OK;
}

// TF-1 AdamOptimizer apply step on HOST arrays in one fused, multi-threaded pass (the optimiser of callers that keep the
// weights on the host and use the *_host entry points; init_optimizer / apply_gradients, core/tensorflow_state.py:342-356):
//   m <- b1 m + (1-b1) g;  v <- b2 v + (1-b2) g^2;  theta <- theta - lr_t m / (sqrt(v) + eps)
// lr_t = lr sqrt(1-b2^t)/(1-b1^t) is formed by the caller.  No FMA contraction, so the result equals the five-pass
// NumPy / torch formulation to the last bit or two.
int qoc_adam_host(double* theta, const double* grad, double* m, double* v, size_t count, double lr_t, double beta1,
                  double beta2, double eps, int threads) {
  if (!theta || !grad || !m || !v) return QOC_EINVAL;
  if (threads < 1) threads = 1;
  const double c1 = 1.0 - beta1, c2 = 1.0 - beta2;
  const long long n = (long long)count;
#pragma omp parallel for num_threads(threads) schedule(static)
  for (long long i = 0; i < n; ++i) {
    const double g = grad[i];
    const double a = beta1 * m[i], bb = c1 * g;
    const double mi = a + bb;
    const double c = beta2 * v[i], gg = g * g, dd = c2 * gg;
    const double vi = c + dd;
    m[i] = mi; v[i] = vi;
    const double den = sqrt(vi) + eps;
    const double q = mi / den, u = lr_t * q;
    theta[i] = theta[i] - u;
  }
  return QOC_OK;
}

}  // extern "C"
