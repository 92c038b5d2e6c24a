// Hardware check + timing of the tcgen05/TMA complex-GEMM program engine (csrc/qoc_tc_f16.cu) against a CPU
// double-precision evaluation of the same products.  Build: tools/build_tc_prog_test.sh;  run on a B200:
//   tools/tc_prog_test.bin [gemm|expm|chain|time|all]
// Prints one JSON line per case.  Exit code 0 even on numeric failure (the lines carry "ok").
#include "qoc_tc_f16.cuh"
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

typedef std::complex<double> cd;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"cuda_error\": \"%s\", \"at\": \"%s\"}\n", cudaGetErrorString(e_), #x); fflush(stdout); exit(0); } } while (0)

static void matmul(const std::vector<cd>& A, const std::vector<cd>& B, std::vector<cd>& C, int n) {
  C.assign((size_t)n * n, cd(0, 0));
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < n; ++k) {
      const cd a = A[(size_t)i * n + k];
      for (int j = 0; j < n; ++j) C[(size_t)i * n + j] += a * B[(size_t)k * n + j];
    }
}

static void unpack(const __half* h, int n, int ld, int e, std::vector<cd>& out) {
  const size_t plane = (size_t)n * ld;
  const double sc = ldexp(1.0, -e);
  out.resize((size_t)n * n);
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) {
      const size_t o = (size_t)r * ld + c;
      const double re = (double)__half2float(h[o]) + (double)__half2float(h[plane + o]);
      const double im = (double)__half2float(h[2 * plane + o]) + (double)__half2float(h[3 * plane + o]);
      out[(size_t)r * n + c] = cd(re * sc, im * sc);
    }
}

static double max_abs_diff(const std::vector<cd>& a, const std::vector<cd>& b) {
  double m = 0;
  for (size_t i = 0; i < a.size(); ++i) m = std::max(m, std::abs(a[i] - b[i]));
  return m;
}

struct Dev {
  TcGeom g;
  TcMaps maps;
  TcStoreMaps smaps;
  __half* base[TC_NCLS];
  size_t count[TC_NCLS];
  int* err;
  int sm;
};

static void dev_setup(Dev& d, int n, size_t nP, size_t nSeg, int grid) {
  if (!tc_geometry(n, &d.g)) { printf("{\"error\": \"geometry\"}\n"); exit(0); }
  d.count[TC_CLS_SCR] = (size_t)grid * 2 * TC_NSLOT;          // two interleaved items per CTA in the expm program
  d.count[TC_CLS_P] = nP;
  d.count[TC_CLS_SEG] = nSeg;
  d.count[TC_CLS_CONST] = 2;
  for (int c = 0; c < TC_NCLS; ++c) {
    const size_t bytes = (d.count[c] ? d.count[c] : 1) * d.g.mat_halfs * sizeof(__half);
    CK(cudaMalloc((void**)&d.base[c], bytes));
    CK(cudaMemset(d.base[c], 0xff, bytes));        // NaN patterns: any read of unwritten data shows up
    const char* ea = tc_make_map(&d.maps.a[c], d.base[c], n, d.g.ld, d.count[c] ? d.count[c] : 1, false);
    const char* eb = tc_make_map(&d.maps.b[c], d.base[c], n, d.g.ld, d.count[c] ? d.count[c] : 1, true);
    const char* es = tc_make_store_map(&d.smaps.st[c], d.base[c], n, d.g.ld, d.count[c] ? d.count[c] : 1);
    if (ea || eb || es) { printf("{\"error\": \"%s\"}\n", ea ? ea : (eb ? eb : es)); exit(0); }
  }
  CK(cudaMalloc((void**)&d.err, sizeof(int)));
  CK(cudaMemset(d.err, 0, sizeof(int)));
  // constants: U0 = I, identity
  std::vector<double> I((size_t)n * n * 2, 0.0);
  for (int i = 0; i < n; ++i) I[((size_t)i * n + i) * 2] = 1.0;
  std::vector<__half> h(d.g.mat_halfs);
  tc_pack_host(I.data(), n, d.g.ld, TC_EU, h.data());
  for (int k = 0; k < 2; ++k) CK(cudaMemcpy(d.base[TC_CLS_CONST] + k * d.g.mat_halfs, h.data(), h.size() * sizeof(__half), cudaMemcpyHostToDevice));
}
static void dev_free(Dev& d) {
  for (int c = 0; c < TC_NCLS; ++c) cudaFree(d.base[c]);
  cudaFree(d.err);
}
static int dev_err(Dev& d) {
  int f = 0;
  CK(cudaMemcpy(&f, d.err, sizeof(int), cudaMemcpyDeviceToHost));
  return f;
}
static void upload_mat(Dev& d, int cls, size_t idx, const std::vector<cd>& M, int e) {
  const int n = d.g.n;
  std::vector<double> z((size_t)n * n * 2);
  for (size_t i = 0; i < (size_t)n * n; ++i) { z[2 * i] = M[i].real(); z[2 * i + 1] = M[i].imag(); }
  std::vector<__half> h(d.g.mat_halfs);
  tc_pack_host(z.data(), n, d.g.ld, e, h.data());
  CK(cudaMemcpy(d.base[cls] + idx * d.g.mat_halfs, h.data(), h.size() * sizeof(__half), cudaMemcpyHostToDevice));
}
static void download_mat(Dev& d, int cls, size_t idx, int e, std::vector<cd>& M) {
  std::vector<__half> h(d.g.mat_halfs);
  CK(cudaMemcpy(h.data(), d.base[cls] + idx * d.g.mat_halfs, h.size() * sizeof(__half), cudaMemcpyDeviceToHost));
  unpack(h.data(), d.g.n, d.g.ld, e, M);
}
static void fill_params(Dev& d, TcParams& q) {
  memset(&q, 0, sizeof(q));
  for (int c = 0; c < TC_NCLS; ++c) q.base[c] = d.base[c];
  q.err_flag = d.err;
}

// one product per item, optional descriptor override
static void test_gemm(int n, int items, uint32_t b_lbo, uint32_t b_sbo, uint32_t a_sbo, const char* tag) {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  const int grid = std::min(items, pr.multiProcessorCount);
  Dev d; dev_setup(d, n, 2 * (size_t)items, items, grid);
  std::mt19937_64 rng(1234 + n);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  std::vector<std::vector<cd>> A(items), B(items);
  const double amp = 1.0 / sqrt((double)n);
  for (int i = 0; i < items; ++i) {
    A[i].resize((size_t)n * n); B[i].resize((size_t)n * n);
    for (auto& x : A[i]) x = cd(U(rng), U(rng)) * amp;
    for (auto& x : B[i]) x = cd(U(rng), U(rng)) * amp;
    upload_mat(d, TC_CLS_P, 2 * i, A[i], TC_EU);
    upload_mat(d, TC_CLS_P, 2 * i + 1, B[i], TC_EU);
  }
  TcParams q; fill_params(d, q);
  q.prog = TC_PROG_GEMM; q.items = items;
  q.a_lbo = 1; q.a_sbo = a_sbo; q.b_lbo = b_lbo; q.b_sbo = b_sbo;      // zeros: the engine's defaults
  CK(tc_launch(q, d.maps, d.g, grid, 0));
  CK(cudaDeviceSynchronize());
  double err = 0, ref = 0, bias_num = 0, bias_den = 0, rms = 0;
  int nan = 0;
  for (int i = 0; i < items; ++i) {
    std::vector<cd> C, G, As, Bs;
    // reference on the operands as stored (split-rounded), so only the product's own error is measured
    download_mat(d, TC_CLS_P, 2 * i, TC_EU, As);
    download_mat(d, TC_CLS_P, 2 * i + 1, TC_EU, Bs);
    matmul(As, Bs, C, n);
    download_mat(d, TC_CLS_SEG, i, TC_EU, G);
    for (auto& x : G) if (!(std::abs(x) < 1e30)) { ++nan; x = cd(0, 0); }
    err = std::max(err, max_abs_diff(C, G));
    for (size_t e = 0; e < C.size(); ++e) {
      ref = std::max(ref, std::abs(C[e]));
      // signed error along the true value: < 0 means the result is shrunk towards zero
      bias_num += (G[e].real() - C[e].real()) * C[e].real() + (G[e].imag() - C[e].imag()) * C[e].imag();
      bias_den += std::norm(C[e]);
      rms += std::norm(G[e] - C[e]);
    }
  }
  printf("{\"test\": \"gemm_bias\", \"n\": %d, \"relative_bias\": %.3e, \"rel_rms_err\": %.3e}\n", n, bias_num / bias_den, sqrt(rms / bias_den));
  printf("{\"test\": \"gemm\", \"tag\": \"%s\", \"n\": %d, \"items\": %d, \"b_lbo\": %u, \"b_sbo\": %u, \"a_sbo\": %u, \"max_abs_err\": %.3e, \"max_ref\": %.3e, \"nan\": %d, \"timeout\": %d, \"ok\": %s}\n",
         tag, n, items, b_lbo, b_sbo, a_sbo, err, ref, nan, dev_err(d), (err < 2e-5 * ref && !nan) ? "true" : "false");
  fflush(stdout);
  dev_free(d);
}

// random anti-Hermitian generators; returns per-item X (double) for the CPU reference
struct ExpmProblem {
  int n, K, T, B, p, s;
  std::vector<std::vector<cd>> A;      // [K+1] each n*n = -i dt H_k
  std::vector<double> ctrl, maxA;
  double xmax, theta;
};
static void make_problem(ExpmProblem& P, int n, int K, int T, int B, int p, int s, double norm) {
  P.n = n; P.K = K; P.T = T; P.B = B; P.p = p; P.s = s;
  std::mt19937_64 rng(99 + n);
  std::normal_distribution<double> N(0.0, 1.0);
  P.A.assign(K + 1, std::vector<cd>((size_t)n * n));
  for (int k = 0; k <= K; ++k) {
    std::vector<cd> H((size_t)n * n);
    for (int i = 0; i < n; ++i)
      for (int j = i; j < n; ++j) {
        cd v(N(rng), i == j ? 0.0 : N(rng));
        H[(size_t)i * n + j] = v; H[(size_t)j * n + i] = std::conj(v);
      }
    const double sc = norm / sqrt((double)n) / (K + 1);
    for (size_t i = 0; i < H.size(); ++i) P.A[k][i] = cd(0, -1) * H[i] * sc;
  }
  P.maxA.assign(K, 1.0);
  P.ctrl.resize((size_t)B * K * T);
  for (auto& x : P.ctrl) x = 0.7 * N(rng);
  // bounds: entrywise |A_0| + sum maxA |A_k|, over 2^s
  std::vector<double> bnd((size_t)n * n, 0.0);
  for (int k = 0; k <= K; ++k)
    for (size_t i = 0; i < bnd.size(); ++i) bnd[i] += std::abs(P.A[k][i]);
  double xmax = 0, fro = 0;
  for (double v : bnd) { xmax = std::max(xmax, v); fro += v * v; }
  const double inv = ldexp(1.0, -s);
  P.xmax = xmax * inv; P.theta = sqrt(fro) * inv;
}
static void cpu_expm(const ExpmProblem& P, int b, int t, std::vector<cd>& out) {
  const int n = P.n;
  std::vector<cd> X((size_t)n * n);
  const double inv = ldexp(1.0, -P.s);
  for (size_t i = 0; i < X.size(); ++i) {
    cd v = P.A[0][i];
    for (int k = 1; k <= P.K; ++k) v += P.maxA[k - 1] * sin(P.ctrl[((size_t)b * P.K + k - 1) * P.T + t]) * P.A[k][i];
    X[i] = v * inv;
  }
  // plain Horner of the same polynomial: S = I + X(I + X/2 (I + X/3 ...))
  std::vector<cd> S((size_t)n * n, cd(0, 0)), Tm;
  for (int i = 0; i < n; ++i) S[(size_t)i * n + i] = 1.0;
  for (int j = P.p; j >= 1; --j) {
    matmul(X, S, Tm, n);
    for (size_t i = 0; i < Tm.size(); ++i) Tm[i] /= (double)j;
    for (int i = 0; i < n; ++i) Tm[(size_t)i * n + i] += 1.0;
    S.swap(Tm);
  }
  for (int i = 0; i < P.s; ++i) { matmul(S, S, Tm, n); S.swap(Tm); }
  out = S;
}

struct ExpmDev { TcExpmOp* ops; double* ctrl; double* maxA; float2* A_f; int nops; float xscale; };
static void expm_upload(const ExpmProblem& P, ExpmDev& e, TcParams& q) {
  int eX, eY;
  tc_pick_scales(P.xmax, P.theta, &eX, &eY);
  std::vector<TcExpmOp> ops;
  tc_build_expm_ops(P.p, P.s, eX, eY, ops);
  e.nops = (int)ops.size();
  CK(cudaMalloc((void**)&e.ops, ops.size() * sizeof(TcExpmOp)));
  CK(cudaMemcpy(e.ops, ops.data(), ops.size() * sizeof(TcExpmOp), cudaMemcpyHostToDevice));
  CK(cudaMalloc((void**)&e.ctrl, P.ctrl.size() * sizeof(double)));
  CK(cudaMemcpy(e.ctrl, P.ctrl.data(), P.ctrl.size() * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMalloc((void**)&e.maxA, P.K * sizeof(double)));
  CK(cudaMemcpy(e.maxA, P.maxA.data(), P.K * sizeof(double), cudaMemcpyHostToDevice));
  std::vector<float2> Af((size_t)(P.K + 1) * P.n * P.n);
  for (int k = 0; k <= P.K; ++k)
    for (size_t i = 0; i < (size_t)P.n * P.n; ++i) Af[(size_t)k * P.n * P.n + i] = make_float2((float)P.A[k][i].real(), (float)P.A[k][i].imag());
  CK(cudaMalloc((void**)&e.A_f, Af.size() * sizeof(float2)));
  CK(cudaMemcpy(e.A_f, Af.data(), Af.size() * sizeof(float2), cudaMemcpyHostToDevice));
  e.xscale = (float)ldexp(1.0, eX - P.s);
  q.prog = TC_PROG_EXPM; q.items = (long long)P.B * P.T;
  q.ilv = getenv("TC_ILV") ? atoi(getenv("TC_ILV")) : 2;
  q.nops = e.nops; q.ops = e.ops; q.K = P.K; q.T = P.T; q.ctrl = e.ctrl; q.maxA = e.maxA; q.A_f = e.A_f; q.xscale = e.xscale;
}

static bool g_small = false;      // run the propagator program on the shared-memory-resident kernel (n <= 64)
static int g_pair = 0;            // CTAs per cluster (2 / 4) of the cta_group::2 kernel for n > 128; 0 = single-CTA engine
static cudaError_t launch_expm(const TcParams& q, const Dev& d, int n, int sms, int grid) {
  if (g_small && tc_small_supported(n)) return tc_small_launch_expm(q, n, sms, 0);
  if (g_pair && tc_pair_supported(n)) {
    TcParams q2 = q;
    q2.tma_store = getenv("TC_TST") ? atoi(getenv("TC_TST")) : 1;
    return tc_pair_launch_expm(q2, d.maps, d.smaps, d.g, g_pair, 0);
  }
  return tc_launch(q, d.maps, d.g, grid, 0);
}
static void test_expm_chain(int n, int K, int T, int B, int p, int s, int L, double norm) {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  ExpmProblem P; make_problem(P, n, K, T, B, p, s, norm);
  const int S = (T + L - 1) / L;
  TcGeom g; tc_geometry(n, &g);
  const int grid = pr.multiProcessorCount * g.ctas_per_sm;
  Dev d; dev_setup(d, n, (size_t)B * T, (size_t)B * S, grid);
  TcParams q; fill_params(d, q);
  ExpmDev e; expm_upload(P, e, q);
  CK(launch_expm(q, d, n, pr.multiProcessorCount, (int)std::min<long long>(grid, q.items)));
  CK(cudaDeviceSynchronize());
  double err = 0; int nan = 0;
  std::vector<std::vector<cd>> Pref((size_t)B * T);
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < T; ++t) {
      std::vector<cd> G;
      cpu_expm(P, b, t, Pref[(size_t)b * T + t]);
      download_mat(d, TC_CLS_P, (size_t)b * T + t, TC_EU, G);
      for (auto& x : G) if (!(std::abs(x) < 1e30)) { ++nan; x = cd(0, 0); }
      err = std::max(err, max_abs_diff(Pref[(size_t)b * T + t], G));
    }
  printf("{\"test\": \"expm%s\", \"n\": %d, \"K\": %d, \"T\": %d, \"B\": %d, \"p\": %d, \"s\": %d, \"nops\": %d, \"max_abs_err\": %.3e, \"nan\": %d, \"timeout\": %d, \"ok\": %s}\n",
         (g_small && tc_small_supported(n)) ? "_small" : "", n, K, T, B, p, s, e.nops, err, nan, dev_err(d), (err < 2e-5 && !nan) ? "true" : "false");
  fflush(stdout);
  // segment products + chain
  double2* Ufin; double* scal;
  CK(cudaMalloc((void**)&Ufin, (size_t)B * n * n * sizeof(double2)));
  CK(cudaMalloc((void**)&scal, (size_t)B * 8 * sizeof(double)));
  TcParams qs; fill_params(d, qs);
  qs.prog = TC_PROG_SEG; qs.items = (long long)B * S; qs.T = T; qs.L = L; qs.S = S;
  CK(tc_launch(qs, d.maps, d.g, (int)std::min<long long>(grid, qs.items), 0));
  TcParams qc; fill_params(d, qc);
  qc.prog = TC_PROG_CHAIN; qc.items = B; qc.chain_cls = TC_CLS_SEG; qc.chain_len = S; qc.Ufin = Ufin; qc.scal = scal;
  CK(tc_launch(qc, d.maps, d.g, (int)std::min<long long>(grid, qc.items), 0));
  CK(cudaDeviceSynchronize());
  std::vector<double> U((size_t)B * n * n * 2), sc((size_t)B * 8);
  CK(cudaMemcpy(U.data(), Ufin, U.size() * sizeof(double), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(sc.data(), scal, sc.size() * sizeof(double), cudaMemcpyDeviceToHost));
  double uerr = 0, serr = 0;
  for (int b = 0; b < B; ++b) {
    std::vector<cd> X((size_t)n * n, cd(0, 0)), Tm;
    for (int i = 0; i < n; ++i) X[(size_t)i * n + i] = 1.0;
    for (int t = 0; t < T; ++t) { matmul(Pref[(size_t)b * T + t], X, Tm, n); X.swap(Tm); }
    double us = 0;
    for (int r = 0; r < n; ++r) {
      cd rs(0, 0);
      for (int c = 0; c < n; ++c) {
        rs += X[(size_t)r * n + c];
        const cd gv(U[((size_t)b * n * n + (size_t)r * n + c) * 2], U[((size_t)b * n * n + (size_t)r * n + c) * 2 + 1]);
        uerr = std::max(uerr, std::abs(gv - X[(size_t)r * n + c]));
      }
      us += std::norm(rs);
    }
    serr = std::max(serr, std::abs(us / n - sc[(size_t)b * 8 + 5]));
  }
  printf("{\"test\": \"chain\", \"n\": %d, \"T\": %d, \"B\": %d, \"L\": %d, \"S\": %d, \"U_final_max_abs_err\": %.3e, \"unitary_scale_err\": %.3e, \"timeout\": %d, \"ok\": %s}\n",
         n, T, B, L, S, uerr, serr, dev_err(d), (uerr < 1e-4 && uerr == uerr) ? "true" : "false");
  fflush(stdout);
  cudaFree(Ufin); cudaFree(scal);
  cudaFree(e.ops); cudaFree(e.ctrl); cudaFree(e.maxA); cudaFree(e.A_f);
  dev_free(d);
}

static void time_expm(int n, int K, int T, int B, int p, int s, int reps) {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  ExpmProblem P; make_problem(P, n, K, T, B, p, s, 0.5);
  TcGeom g; tc_geometry(n, &g);
  const int grid = pr.multiProcessorCount * g.ctas_per_sm;
  Dev d; dev_setup(d, n, (size_t)B * T, 1, grid);
  TcParams q; fill_params(d, q);
  ExpmDev e; expm_upload(P, e, q);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int g2 = (int)std::min<long long>(grid, q.items);
  const bool small = g_small && tc_small_supported(n);
  unsigned long long* prof; CK(cudaMalloc((void**)&prof, (size_t)g2 * 8 * sizeof(unsigned long long)));
  CK(cudaMemset(prof, 0, (size_t)g2 * 8 * sizeof(unsigned long long)));
  q.prof = prof;
  CK(launch_expm(q, d, n, pr.multiProcessorCount, g2));
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; ++r) CK(launch_expm(q, d, n, pr.multiProcessorCount, g2));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= reps;
  if (small) {
    std::vector<unsigned long long> h((size_t)g2 * 8);
    CK(cudaMemcpy(h.data(), prof, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    printf("{\"test\": \"small_epilogue_cycles_cta0\", \"n\": %d, \"x_assembly\": %llu, \"wait_acc\": %llu, \"tmem_ld\": %llu, \"math_stores\": %llu, \"fence_arrive\": %llu, \"kernel_cycles\": %.0f}\n",
           n, h[0], h[1], h[2], h[3], h[4], ms * 1.965e6);
  }
  if (!small) {   // per-role cycle attribution of the last launch (averages over CTAs)
    std::vector<unsigned long long> h((size_t)g2 * 8);
    CK(cudaMemcpy(h.data(), prof, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    double a[8] = {0};
    for (int c = 0; c < g2; ++c) for (int i = 0; i < 8; ++i) a[i] += (double)h[(size_t)c * 8 + i] / g2;
    printf("{\"test\": \"roles\", \"n\": %d, \"producer_wait_opdone_Mcyc\": %.2f, \"producer_wait_empty_Mcyc\": %.2f, \"mma_wait_tmem_empty_Mcyc\": %.2f, \"mma_wait_full_Mcyc\": %.2f, \"epi_wait_tmem_full_Mcyc\": %.2f, \"epi_work_Mcyc\": %.2f, \"producer_wait_prologue_Mcyc\": %.2f, \"producer_wait_prev_product_Mcyc\": %.2f}\n",
           n, a[0] * 1e-6, a[1] * 1e-6, a[2] * 1e-6, a[3] * 1e-6, a[4] * 1e-6, a[5] * 1e-6, a[6] * 1e-6, a[7] * 1e-6);
  }
  const double alg = 8.0 * n * n * n * (double)(p - 1 + s) * (double)B * T;       // SURVEY 8d count
  const double issued = 8.0 * n * n * n * (double)e.nops * (double)B * T;
  printf("{\"test\": \"time_expm%s\", \"pair\": %d, \"n\": %d, \"T\": %d, \"B\": %d, \"p\": %d, \"s\": %d, \"nops\": %d, \"grid\": %d, \"stages\": %d, \"ms\": %.4f, \"us_per_item\": %.3f, \"alg_tflops\": %.2f, \"issued_complex_tflops\": %.2f, \"timeout\": %d}\n",
         small ? "_small" : "", g_pair, n, T, B, p, s, e.nops, g2, g.stages, ms, 1e3 * ms / ((double)B * T) , alg / ms * 1e-9, issued / ms * 1e-9, dev_err(d));
  fflush(stdout);
  cudaFree(e.ops); cudaFree(e.ctrl); cudaFree(e.maxA); cudaFree(e.A_f);
  dev_free(d);
}

// streaming throughput of INDEPENDENT products (TC_PROG_GEMM): no dependency chain, no elementwise source
static void time_gemm(int n, int items, int reps) {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  const int grid = std::min(items, pr.multiProcessorCount);
  Dev d; dev_setup(d, n, 2 * (size_t)items, items, grid);
  CK(cudaMemset(d.base[TC_CLS_P], 0, 2 * (size_t)items * d.g.mat_halfs * sizeof(__half)));
  TcParams q; fill_params(d, q);
  q.prog = TC_PROG_GEMM; q.items = items;
  unsigned long long* prof; CK(cudaMalloc((void**)&prof, (size_t)grid * 8 * sizeof(unsigned long long)));
  CK(cudaMemset(prof, 0, (size_t)grid * 8 * sizeof(unsigned long long)));
  q.prof = prof;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(tc_launch(q, d.maps, d.g, grid, 0));
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int r = 0; r < reps; ++r) CK(tc_launch(q, d.maps, d.g, grid, 0));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= reps;
  std::vector<unsigned long long> h((size_t)grid * 8);
  CK(cudaMemcpy(h.data(), prof, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  double a[8] = {0};
  for (int c = 0; c < grid; ++c) for (int i = 0; i < 8; ++i) a[i] += (double)h[(size_t)c * 8 + i] / grid;
  printf("{\"test\": \"time_gemm\", \"n\": %d, \"items\": %d, \"ms\": %.4f, \"complex_tflops\": %.2f, \"producer_wait_done_Mcyc\": %.2f, \"producer_wait_empty_Mcyc\": %.2f, \"mma_wait_tmem_empty_Mcyc\": %.2f, \"mma_wait_full_Mcyc\": %.2f, \"epi_wait_tmem_full_Mcyc\": %.2f, \"epi_work_Mcyc\": %.2f, \"timeout\": %d}\n",
         n, items, ms, 8.0 * n * n * n * items / ms * 1e-9, a[0] * 1e-6, a[1] * 1e-6, a[2] * 1e-6, a[3] * 1e-6, a[4] * 1e-6, a[5] * 1e-6, dev_err(d));
  fflush(stdout);
  dev_free(d);
}

int main(int argc, char** argv) {
  const std::string what = argc > 1 ? argv[1] : "all";
  if (what == "gemm" || what == "all") {
    const int ns[] = {16, 36, 64, 100, 128, 216, 256};
    for (int n : ns) test_gemm(n, 3, 0, 0, 0, "default");
  }
  if (what == "small") {
    g_small = true;
    const int ns[] = {8, 16, 24, 30, 32, 36, 40, 48, 64};
    for (int n : ns) test_expm_chain(n, 2, 6, 3, 6, 3, 4, 0.8);
    test_expm_chain(36, 4, 40, 7, 8, 2, 16, 0.8);
    test_expm_chain(30, 4, 25, 5, 7, 3, 16, 0.8);
    test_expm_chain(12, 2, 9, 3, 1, 0, 4, 0.01);
    test_expm_chain(20, 2, 9, 3, 2, 0, 4, 0.1);
    test_expm_chain(20, 2, 9, 3, 5, 0, 4, 0.3);
    time_expm(36, 4, 512, 148, 8, 2, 3);
    time_expm(64, 2, 128, 148, 8, 2, 3);
    time_expm(30, 4, 512, 148, 7, 3, 3);
    time_expm(16, 2, 1024, 148, 8, 2, 3);
  }
  if (what == "small1" && argc >= 9) {   // small1 n K T B p s reps
    g_small = true;
    time_expm(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]), atoi(argv[8]));
  }
  if (what == "pair" && argc >= 3) {      // pair <cs>: correctness of the cta_group::2 propagator kernel, then timing at the C4 item size
    g_pair = atoi(argv[2]);
    test_expm_chain(216, 3, 9, 2, 5, 8, 4, 3.0);
    test_expm_chain(256, 2, 5, 2, 6, 3, 4, 0.8);
    test_expm_chain(130, 2, 5, 2, 6, 3, 4, 0.8);
    test_expm_chain(176, 2, 7, 3, 4, 2, 4, 0.8);
    time_expm(216, 3, 16, 148, 5, 8, 3);
  }
  if (what == "pair1" && argc >= 10) {    // pair1 cs n K T B p s reps
    g_pair = atoi(argv[2]);
    time_expm(atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]), atoi(argv[8]), atoi(argv[9]));
  }
  if (what == "timegemm") {
    time_gemm(216, 148 * 24, 2);
    time_gemm(128, 148 * 64, 2);
    time_gemm(256, 148 * 16, 2);
    time_gemm(64, 148 * 128, 2);
  }
  if (what == "time1" && argc >= 9)   // time1 n K T B p s reps
    time_expm(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]), atoi(argv[8]));
  if (what == "expm" || what == "all") {
    test_expm_chain(16, 2, 6, 2, 6, 3, 4, 0.8);
    test_expm_chain(24, 2, 6, 2, 6, 3, 4, 0.8);
    test_expm_chain(32, 2, 6, 2, 6, 3, 4, 0.8);
    test_expm_chain(36, 2, 6, 2, 6, 0, 4, 0.1);
    test_expm_chain(36, 2, 6, 2, 1, 0, 4, 0.01);
    test_expm_chain(48, 2, 6, 2, 6, 3, 4, 0.8);
    test_expm_chain(36, 4, 40, 2, 8, 2, 16, 0.8);
    test_expm_chain(64, 2, 20, 2, 7, 3, 8, 0.8);
    test_expm_chain(128, 2, 10, 2, 8, 2, 4, 0.8);
    test_expm_chain(216, 3, 9, 2, 5, 8, 4, 3.0);
  }
  if (what == "time" || what == "all") {
    time_expm(216, 3, 8, 148, 5, 8, 3);
    time_expm(128, 2, 16, 148, 8, 2, 3);
    time_expm(64, 2, 64, 148, 8, 2, 3);
    time_expm(36, 4, 128, 148, 8, 2, 3);
  }
  return 0;
}
