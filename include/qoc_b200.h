/*
 * qoc_b200.h -- C ABI of the B200-native GRAPE engine (libqoc_b200.so).
 *
 * The reference (SchusterLab/quantum-optimal-control) has no FFI of its own: its hot path is a
 * TensorFlow-1 graph that the Python driver evaluates through four `session.run` fetch sites.
 * Each entry point below replaces one of those seams (reference paths are relative to
 * /root/reference/quantum_optimal_control/):
 *
 *   qoc_set_problem      <- constants baked into the graph by TensorflowState.init_variables /
 *                           init_tf_vectors / init_tf_propagators / init_tf_propagator
 *                           (core/tensorflow_state.py:146-165,205) from SystemParameters
 *                           (core/system_parameters.py:194-266)
 *   qoc_set_regularizers <- get_reg_loss(tfs)            (core/regularization_functions.py:7-97)
 *   qoc_value_and_grad   <- run_session.get_error        (core/run_session.py:119-127) and the
 *                           Adam-loop fetch [grad_squared, loss, reg_loss, unitary_scale]
 *                           (core/run_session.py:53-54) + compute_gradients
 *                           (core/tensorflow_state.py:348-353)
 *   qoc_evolve           <- Analysis.get_final_state / get_inter_vecs fetches
 *                           (core/analysis.py:26-35,44-65)
 *   qoc_*_host           <- same calls with HOST buffers (copies inside), the form a ctypes/cffi
 *                           binding in the reference's run_session.py would use
 *
 * Conventions
 *   - plain C, no torch types; complex numbers are interleaved (re,im) double pairs, i.e. the
 *     memory layout of numpy complex128 / torch complex128;
 *   - all matrices row-major; B = number of independent problem instances (seeds) that share the
 *     problem constants and differ only in their control amplitudes;
 *   - "dev" pointers are device pointers owned by the caller (e.g. torch tensors), "host"
 *     pointers are host memory; `stream` is a cudaStream_t passed as void*; device entry points
 *     are asynchronous on that stream, *_host entry points synchronise it before returning; a handle
 *     also owns one high-priority stream on which the loss / gradient kernels of few-state problems
 *     run beside the U_final kernels -- it forks from and joins back into `stream` inside every call,
 *     so the caller only ever orders against `stream`;
 *   - return value 0 = ok, negative = QOC_E* below; qoc_last_error(h) gives the message;
 *   - a handle is not thread-safe; distinct handles are independent;
 *   - there is NO CPU fallback: every compute entry point fails with QOC_ECUDA when no sm_100
 *     device is present.
 */
#ifndef QOC_B200_H
#define QOC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QOC_ABI_VERSION 2

enum {
  QOC_OK = 0,
  QOC_EINVAL = -1,      /* bad argument / unsupported size */
  QOC_ESTATE = -2,      /* call order (workspace or problem not set) */
  QOC_ECUDA = -3,       /* CUDA runtime error, see qoc_last_error */
  QOC_ENOMEM = -4       /* workspace too small */
};

/* arithmetic of the propagator (expm + chain) stage */
enum {
  QOC_F64 = 0,          /* IEEE double on the FP64 pipe (parity / config C2) */
  QOC_TF32X3 = 1,       /* fp32-class, n <= 32: tcgen05 kind::tf32 with 3-way operand split, fp32 accumulate in TMEM */
  QOC_F16X2 = 2         /* fp32-class, n <= 256, m <= 8: tcgen05 kind::f16 on TMA-fed tiles, every real number held as a pair
                         * of fp16 halves (22+ bits, the size of the reference's float32), three MMAs per real product,
                         * fp32 accumulate in TMEM; propagators cached as split fp16 planes; states / costates stay fp64 */
};

/* qoc_dims_t.flags */
/* state-transfer mode (core/tensorflow_state.py:77-142,244-261,331-335): per-step propagator is the
 * Taylor sum of order exp_terms-1 WITHOUT scaling/squaring, unitary_scale is
 * (sum_j |psi_j(T)|^2)^2 / m^2, V / phi are arbitrary state vectors (pass concerned_idx = NULL).
 * The reverse sweep uses Q_t^dagger, which equals the reference's sum_j (-H)^j/j! for Hermitian
 * Hamiltonians (the host layer checks that). */
#define QOC_FLAG_STATE_TRANSFER 1u

typedef struct qoc_handle_s* qoc_handle_t;

typedef struct {
  int32_t n;            /* Hilbert-space dimension (state_num, core/system_parameters.py:165) */
  int32_t K;            /* number of control operators (ops_len, :202) */
  int32_t T;            /* time steps (steps) */
  int32_t m;            /* number of concerned states (len(states_concerned_list)) */
  int32_t B;            /* batch of independent instances (ours; the reference has B = 1) */
  int32_t exp_terms;    /* Taylor order p, INCLUSIVE (core/tensorflow_state.py:37) */
  int32_t scaling;      /* squarings s (:43-44) */
  int32_t dtype;        /* QOC_F64 | QOC_TF32X3 */
  uint32_t flags;
} qoc_dims_t;

/* Regulariser description; a term is enabled by its `has_*` flag (the reference enables by key
 * PRESENCE in reg_coeffs, even with coefficient 0: core/regularization_functions.py:15,21,28,38,71,88). */
typedef struct {
  int32_t has_amplitude;  double amplitude;   /* :15-18 */
  int32_t has_envelope;   double envelope;    /* :21-25, uses the envelope[K][T] array */
  int32_t has_dwdt;       double dwdt;        /* :28-35 */
  int32_t has_d2wdt2;     double d2wdt2;      /* :38-45 */
  int32_t has_forbidden;                      /* :71-85, uses forbid_weight[n] */
  int32_t has_speed_up;   double speed_up;    /* :88-95 */
} qoc_reg_t;

int qoc_abi_version(void);
int qoc_create(qoc_handle_t* out, const qoc_dims_t* dims);
int qoc_destroy(qoc_handle_t h);
const char* qoc_last_error(qoc_handle_t h);

/* Scratch (propagators P[B][T][n][n], states psi / costates lambda [B][T+1][m][n], ...) lives in
 * ONE caller-owned device buffer so the caller's allocator (torch) stays in charge of HBM. */
int qoc_workspace_bytes(qoc_handle_t h, size_t* bytes);
int qoc_set_workspace(qoc_handle_t h, void* dev_ptr, size_t bytes);

/* Problem constants, all HOST pointers, copied once.
 *   A_host        [K+1][n][n] complex : A_0 = -i*dt*H0, A_k = -i*dt*Hops[k-1]  (system_parameters.py:199,204)
 *   U0_host       [n][n] complex      : initial unitary (:57)
 *   phi_host      [m][n] complex      : target vectors U_target * V_j (tensorflow_state.py:165)
 *   V_host        [m][n] complex      : initial vectors (system_parameters.py:168-187)
 *   concerned_idx [m] or NULL         : if non-NULL, V_j is the basis vector e_{idx[j]} (bare states)
 *   maxA_host     [K]                 : ops_max_amp (tensorflow_state.py:178)
 *   dt                                 : total_time / steps */
int qoc_set_problem(qoc_handle_t h, const double* A_host, const double* U0_host, const double* phi_host,
                    const double* V_host, const int32_t* concerned_idx, const double* maxA_host, double dt,
                    void* stream);

/* envelope_host [K][T] (one_minus_gauss, system_parameters.py:253-266) may be NULL unless has_envelope;
 * forbid_weight_host [n] = sum over forbidden entries f with state s_f == i of coeff_f (NOT yet divided
 * by T) may be NULL unless has_forbidden. reg == NULL clears all terms. */
int qoc_set_regularizers(qoc_handle_t h, const qoc_reg_t* reg, const double* envelope_host,
                         const double* forbid_weight_host, void* stream);

/* forbid_dressed (core/regularization_functions.py:73-80): evaluate the forbidden-state populations on
 * psi' = W psi with W_host [n][n] complex = v_sorted^dagger (host pointer, copied); NULL switches it off.
 * Allocates one internal device buffer of B*(T+1)*m*n complex for the transformed states. */
int qoc_set_forbid_basis(qoc_handle_t h, const double* W_host, void* stream);

/* One fwd+bwd evaluation for all B instances (== B calls of run_session.get_error).
 *   base_dev          [B][K][T]  ops_weight_base (controls are maxA_k * sin(base))
 *   loss_dev          [B]        1 - |sum_j <phi_j|psi_j(T)>|^2 / m^2
 *   reg_loss_dev      [B]        loss + regularisers
 *   grad_dev          [B][K][T]  d reg_loss / d base, the reference's first-order GRAPE gradient
 *   unitary_scale_dev [B]        (tensorflow_state.py:225)
 *   grad_squared_dev  [B]        sum g^2 / 2 (:352-353)
 * Any output pointer except grad_dev may be NULL. */
int qoc_value_and_grad(qoc_handle_t h, const double* base_dev, double* loss_dev, double* reg_loss_dev,
                       double* grad_dev, double* unitary_scale_dev, double* grad_squared_dev, void* stream);

/* Forward only.  U_final_dev [B][n][n] complex (may be NULL); inter_vecs_dev [B][T+1][m][n] complex
 * (may be NULL), entry t=0 is V, entry t>=1 is X_{t-1} V (tensorflow_state.py:233-238);
 * loss_dev / unitary_scale_dev [B] may be NULL. */
int qoc_evolve(qoc_handle_t h, const double* base_dev, double* U_final_dev, double* inter_vecs_dev,
               double* loss_dev, double* unitary_scale_dev, void* stream);

/* Same two calls with HOST buffers: H2D of base, compute, D2H of the results, stream synchronised. */
int qoc_value_and_grad_host(qoc_handle_t h, const double* base_host, double* loss_host, double* reg_loss_host,
                            double* grad_host, double* unitary_scale_host, double* grad_squared_host,
                            void* stream);
int qoc_evolve_host(qoc_handle_t h, const double* base_host, double* U_final_host, double* inter_vecs_host,
                    double* loss_host, double* unitary_scale_host, void* stream);

/* Read-only views into the workspace after a value_and_grad / evolve call (device pointers):
 * propagators P[B][T][n][n] (complex double for QOC_F64, complex float tiles for QOC_TF32X3; for QOC_F16X2
 * elem_bytes = 2 and the layout is [B][T][4][n][ld] fp16 planes (Re h0, Re h1, Im h0, Im h1), ld = n rounded up
 * to 16, value = (h0 + h1) / 8192). */
int qoc_debug_propagators(qoc_handle_t h, void** P_dev, int* elem_bytes);

/* Instances processed per pass.  When the workspace for all B instances would exceed the memory budget
 * (80 % of free device memory at qoc_create, or QOC_B200_MAX_WS_GB), the batch is processed in chunks that
 * reuse the same workspace; results are identical, qoc_workspace_bytes reports the reduced size. */
int qoc_batch_chunk(qoc_handle_t h);

/* Number of kernel launches issued by this handle since creation (bench.py's gpu_launches). */
int64_t qoc_launch_count(qoc_handle_t h);

/* Synchronises `stream` and reports device-side failures that cannot surface through a launch status
 * (the tcgen05 pipeline bails out instead of hanging if an MMA completion barrier never fires). */
int qoc_poll_error(qoc_handle_t h, void* stream);

/* Optional per-stage timing with CUDA events recorded around the QOC_NUM_KERNELS stages of the last
 * qoc_value_and_grad on the stream each stage runs on (order: expm, chain, fwd_reduce, costate, grad,
 * finalize).  For few-state problems "chain" is the forward state sweep (the U_final kernels run
 * concurrently on the caller's stream) and "finalize" includes the join of the two branches.
 * qoc_kernel_times_ms synchronises on the last event. */
#define QOC_NUM_KERNELS 6
int qoc_set_profiling(qoc_handle_t h, int enable);
int qoc_kernel_times_ms(qoc_handle_t h, float* ms_out /* [QOC_NUM_KERNELS] */);

/* TF-1 AdamOptimizer apply step (tf.train.AdamOptimizer as constructed at core/tensorflow_state.py:345, applied by
 * run_session.py:66-67) on HOST arrays of `count` doubles, one fused multi-threaded pass:
 *   m <- beta1 m + (1-beta1) g;  v <- beta2 v + (1-beta2) g^2;  theta <- theta - lr_t m / (sqrt(v) + eps)
 * with lr_t = lr sqrt(1-beta2^t) / (1-beta1^t) formed by the caller.  Companion of qoc_value_and_grad_host for
 * callers that keep the weights on the host; needs no GPU. */
int qoc_adam_host(double* theta, const double* grad, double* m, double* v, size_t count, double lr_t, double beta1,
                  double beta2, double eps, int threads);

#ifdef __cplusplus
}
#endif
#endif /* QOC_B200_H */
