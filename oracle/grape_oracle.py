"""CPU oracle for the GRAPE hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module.  The shipped path
(``quantum-optimal-control_b200/``) never imports it and has no CPU fallback.

What this is
------------
A line-by-line restatement (Python 3 / torch-CPU) of the algorithm that the reference
``/root/reference/quantum_optimal_control`` builds as a TensorFlow-1 graph, in the SAME
real-embedded (2n x 2n) form, with torch autograd standing in for TF autodiff and a
``torch.autograd.Function`` standing in for the reference's ``function.Defun`` custom
gradient.  Every function cites the reference file:line it follows (paths relative to
``quantum_optimal_control/``).

A second, independent formulation (``costate_value_and_grad``: complex n x n costate
recursion, NumPy) is what the CUDA kernels implement; the test-suite cross-checks the two.

Parity pin status
-----------------
The reference ships NO tests, golden vectors or fixtures and its
arithmetic lives in an absent third-party dependency (``tensorflow>=1.0``, unpinned,
``setup.py:40``), so it cannot run as-is: **parity unpinned by the reference's own tests**.
Pins we supply instead (see tests/test_oracle.py, tests/golden/):
  * the reference's own UNMODIFIED graph-definition source, executed in this container by
    ``oracle/run_reference.py`` through a small eager stand-in for the TF-1 ops it calls
    (``oracle/tf1_shim.py``) -> committed fixtures ``tests/golden/ref_*.npz``;
  * analytic pi-pulse, ``scipy.linalg.expm`` product, unitarity, finite differences.

dtype: ``torch.float64`` for parity work, ``torch.float32`` for the reference-cost timing
mode (the reference is float32 throughout: ``core/tensorflow_state.py:49,70,147-174,205``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import torch


# ----------------------------------------------------------------------------------------
# helper_functions/grape_functions.py
# ----------------------------------------------------------------------------------------
def c_to_r_mat(M):
    """complex n x n -> real 2n x 2n [[Re,-Im],[Im,Re]]  (helper_functions/grape_functions.py:211-213)"""
    M = np.asarray(M)
    return np.block([[M.real, -M.imag], [M.imag, M.real]]).astype(np.float64)


def c_to_r_vec(v):
    """complex n -> real 2n [Re;Im]  (helper_functions/grape_functions.py:215-220)"""
    v = np.asarray(v)
    return np.concatenate([v.real, v.imag]).astype(np.float64)


def get_state_index(bareindex, dressed_id):
    """helper_functions/grape_functions.py:204-209"""
    if len(dressed_id) > 0:
        return list(dressed_id).index(bareindex)
    return bareindex


def sort_ev(v, dressed_id):
    """helper_functions/grape_functions.py:194-202 -- column ii = eigenvector dressed as bare ii"""
    cols = [v[:, get_state_index(ii, dressed_id)] for ii in range(len(dressed_id))]
    return np.transpose(np.reshape(cols, [len(dressed_id), len(dressed_id)]))


def r_to_c_mat(M, n):
    """core/analysis.py:18-24 -- left block column of the real embedding"""
    return M[:n, :n] + 1j * M[n:2 * n, :n]


# ----------------------------------------------------------------------------------------
# core/system_parameters.py  (problem setup)
# ----------------------------------------------------------------------------------------
def _approx_expm(M, exp_t, scaling_terms):
    """core/system_parameters.py:88-103 (sum over ii < exp_t, i.e. order exp_t-1)"""
    U = np.identity(len(M), dtype=M.dtype)
    Mt = np.identity(len(M), dtype=M.dtype)
    fact = 1.0
    for ii in range(1, exp_t):
        fact *= ii
        Mt = np.dot(Mt, M)
        U = U + Mt / ((2. ** float(ii * scaling_terms)) * fact)
    for _ in range(scaling_terms):
        U = np.dot(U, U)
    return U


def _approx_exp(M, exp_t, scaling_terms):
    """core/system_parameters.py:105-120 (scalar version)"""
    U = 1.0
    Mt = 1.0
    fact = 1.0
    for ii in range(1, exp_t):
        fact *= ii
        Mt = M * Mt
        U += Mt / ((2. ** float(ii * scaling_terms)) * fact)
    for _ in range(scaling_terms):
        U = U * U
    return U


class _TermChooser:
    """core/system_parameters.py:122-158 + :208-230.  ``scaling`` persists between calls
    (the cumulative ``+= d`` quirk) and U_f is not reset between exp_t tries."""

    def __init__(self, H0, Hops, maxA, U0, dt, steps, unitary_error, state_transfer, no_scaling):
        self.H0, self.Hops, self.maxA, self.U0 = H0, Hops, maxA, U0
        self.dt, self.steps, self.err = dt, steps, unitary_error
        self.fixed_zero = bool(state_transfer or no_scaling)
        self.scaling = None
        self.n = len(H0)

    def choose(self, d):
        exp_t = 20
        H = self.H0
        U_f = self.U0
        for ii in range(len(self.Hops)):
            H = H + self.maxA[ii] * self.Hops[ii]
        if d == 0:
            self.scaling = max(int(2 * np.log2(np.max(np.abs(-(0 + 1j) * self.dt * H)))), 0)
        else:
            self.scaling += d
        if self.fixed_zero:
            self.scaling = 0
        while True:
            if self.n < 10:
                for _ in range(self.steps):
                    U_f = np.dot(U_f, _approx_expm((0 - 1j) * self.dt * H, exp_t, self.scaling))
                metric = np.abs(np.trace(np.dot(np.conjugate(np.transpose(U_f)), U_f))) / self.n
            else:
                max_term = np.max(np.abs(-(0 + 1j) * self.dt * H))
                metric = 1 + self.steps * np.abs(
                    (_approx_exp(max_term, exp_t, self.scaling) - np.exp(max_term)) / np.exp(max_term))
            if exp_t == 3:
                break
            if np.abs(metric - 1.0) < self.err:
                exp_t -= 1
            else:
                break
        return exp_t

    def select(self):
        comparisons = 1 if self.fixed_zero else 6
        exps, scalings = [], []
        d = 0
        while comparisons > 0:
            exps.append(self.choose(d))
            scalings.append(self.scaling)
            comparisons -= 1
            d += 1
        a = int(np.argmin(np.add(exps, scalings)))
        return exps[a], scalings[a]


@dataclass
class OracleSetup:
    """What core/system_parameters.py:12-86 leaves on ``sys_para`` for the graph."""
    n: int
    K: int
    steps: int
    dt: float
    exp_terms: int
    scaling: int
    matrix_list: np.ndarray          # [K+2, 2n, 2n]  (:246-251)
    initial_unitary: np.ndarray      # [2n,2n]        (:57)
    target_unitary: Optional[np.ndarray]   # [2n,2n]  (:59)
    target_vectors: Optional[list]   # state transfer (:61-65)
    initial_vectors: list            # m x [2n]       (:168-187)
    initial_vectors_c: list
    one_minus_gauss: np.ndarray      # [K,T]          (:253-266)
    ops_weight_base: np.ndarray      # [K,T]          (:272-284)
    ops_max_amp: np.ndarray
    states_concerned_list: list
    reg_coeffs: dict
    state_transfer: bool
    use_inter_vecs: bool
    is_dressed: bool = False
    v_c: Optional[np.ndarray] = None
    dressed_id: Optional[list] = None
    total_time: float = 0.0


def make_setup(H0, Hops, U, total_time, steps, states_concerned_list, U0=None, reg_coeffs=None,
               dressed_info=None, maxA=None, initial_guess=None, unitary_error=1e-4,
               state_transfer=False, no_scaling=False, Taylor_terms=None, use_inter_vecs=True,
               rng=None) -> OracleSetup:
    """main_grape/grape.py:89-104 defaults + core/system_parameters.py:12-286."""
    H0 = np.asarray(H0)
    Hops = [np.asarray(h) for h in Hops]
    n = len(H0)
    K = len(Hops)
    if U0 is None:                                   # grape.py:89-90
        U0 = np.identity(n)
    if maxA is None:                                 # grape.py:95-101
        if initial_guess is None:
            maxA = 4 * np.ones(K)
        else:
            maxA = 1.5 * np.max(np.abs(initial_guess)) * np.ones(K)
    maxA = np.asarray(maxA, dtype=np.float64)
    reg_coeffs = {} if reg_coeffs is None else reg_coeffs

    u0_base = None
    if initial_guess is not None:                    # system_parameters.py:38-46
        u0 = np.asarray(initial_guess, dtype=np.float64)
        u0_base = np.zeros_like(u0)
        for ii in range(len(u0_base)):
            u0_base[ii] = u0[ii] / maxA[ii]
            if max(u0_base[ii]) > 1.0:
                raise ValueError('Initial guess has strength > max_amp for op %d' % (ii))
        u0_base = np.arcsin(u0_base)

    is_dressed, v_c, dressed_id = False, None, None
    if dressed_info is not None:                     # :75-80
        v_c = dressed_info['eigenvectors']
        dressed_id = dressed_info['dressed_id']
        is_dressed = dressed_info['is_dressed']

    dt = float(total_time) / steps                   # :163-165

    init_c, init_r = [], []                          # :168-187
    for state in states_concerned_list:
        if state_transfer:
            vc = np.array(state)
        elif is_dressed:
            vc = v_c[:, get_state_index(state, dressed_id)]
        else:
            vc = np.zeros(n)
            vc[state] = 1
        init_c.append(vc)
        init_r.append(c_to_r_vec(vc))

    ops = [c_to_r_mat(-1j * dt * h) for h in Hops]   # :197-204
    H0r = c_to_r_mat(-1j * dt * H0)
    if Taylor_terms is None:                         # :208-230
        exp_terms, scaling = _TermChooser(H0, Hops, maxA, np.asarray(U0), dt, steps, unitary_error,
                                          state_transfer, no_scaling).select()
    else:
        exp_terms, scaling = int(Taylor_terms[0]), int(Taylor_terms[1])
    matrix_list = np.array([H0r] + ops + [np.eye(2 * n)])     # :240-251

    x = np.linspace(-2, 2, steps)                    # :253-270
    shape = np.ones(steps) - np.exp(-np.power(x - 0., 2.) / (2 * np.power(1., 2.))) - 0.0
    shape = shape * (shape > 0) + 0.01 * np.ones(steps)
    one_minus_gauss = np.array([shape for _ in range(K)])

    if u0_base is not None:                          # :272-284
        base = np.reshape(u0_base, [K, steps])
    else:
        rng = np.random if rng is None else rng
        base = rng.normal(0, 1. / np.sqrt(steps), [K, steps])

    if state_transfer:
        target_unitary, target_vectors = None, [c_to_r_vec(np.asarray(v)) for v in U]
    else:
        target_unitary, target_vectors = c_to_r_mat(np.asarray(U)), None

    return OracleSetup(n=n, K=K, steps=steps, dt=dt, exp_terms=exp_terms, scaling=scaling,
                       matrix_list=matrix_list, initial_unitary=c_to_r_mat(np.asarray(U0)),
                       target_unitary=target_unitary, target_vectors=target_vectors,
                       initial_vectors=init_r, initial_vectors_c=init_c,
                       one_minus_gauss=one_minus_gauss, ops_weight_base=base, ops_max_amp=maxA,
                       states_concerned_list=list(states_concerned_list), reg_coeffs=reg_coeffs,
                       state_transfer=bool(state_transfer), use_inter_vecs=use_inter_vecs,
                       is_dressed=bool(is_dressed), v_c=v_c, dressed_id=dressed_id,
                       total_time=float(total_time))


# ----------------------------------------------------------------------------------------
# core/tensorflow_state.py  (the graph)
# ----------------------------------------------------------------------------------------
def _get_matexp(uks, H_all, input_num, taylor_terms, scaling):
    """core/tensorflow_state.py:25-46.  Order ``taylor_terms`` INCLUSIVE; weights (incl. the
    constant-1 drift weight) are divided by 2**scaling, not the assembled matrix."""
    matexp = H_all[input_num]
    H = None
    for ii in range(input_num):
        term = (uks[ii] / (2. ** scaling)) * H_all[ii]
        H = term if H is None else H + term
    H_n = H
    factorial = 1.
    for ii in range(1, taylor_terms + 1):
        factorial = factorial * ii
        matexp = matexp + H_n / factorial
        if not ii == taylor_terms:
            H_n = torch.matmul(H, H_n)
    for _ in range(scaling):
        matexp = torch.matmul(matexp, matexp)
    return matexp


class _MatExpOp(torch.autograd.Function):
    """core/tensorflow_state.py:49-75: Defun ``matexp_op`` with ``grad_func=matexp_op_grad``:
    d/d uks[k] = sum(grad * (H_all[k] @ matexp)) for k>=1, 0 for the drift weight, zero wrt
    H_all; the backward RE-COMPUTES matexp (:58)."""

    @staticmethod
    def forward(ctx, uks, H_all, input_num, taylor_terms, scaling):
        ctx.save_for_backward(uks, H_all)
        ctx.cfg = (input_num, taylor_terms, scaling)
        return _get_matexp(uks, H_all, input_num, taylor_terms, scaling)

    @staticmethod
    def backward(ctx, grad):
        uks, H_all = ctx.saved_tensors
        input_num, taylor_terms, scaling = ctx.cfg
        matexp = _get_matexp(uks, H_all, input_num, taylor_terms, scaling)
        coeff = [torch.zeros((), dtype=grad.dtype)]
        for ii in range(1, input_num):
            coeff.append(torch.sum(grad * torch.matmul(H_all[ii], matexp)))
        return torch.stack(coeff), None, None, None, None


def _get_matvecexp(uks, H_all, psi, input_num, taylor_terms):
    """core/tensorflow_state.py:77-97.  Order ``taylor_terms``-1, no scaling."""
    matvecexp = psi
    H = None
    for ii in range(input_num):
        term = uks[ii] * H_all[ii]
        H = term if H is None else H + term
    psi_n = psi
    factorial = 1.
    for ii in range(1, taylor_terms):
        factorial = factorial * ii
        psi_n = torch.matmul(H, psi_n)
        matvecexp = matvecexp + psi_n / factorial
    return matvecexp


class _MatVecExpOp(torch.autograd.Function):
    """core/tensorflow_state.py:100-142."""

    @staticmethod
    def forward(ctx, uks, H_all, psi, input_num, taylor_terms):
        ctx.save_for_backward(uks, H_all, psi)
        ctx.cfg = (input_num, taylor_terms)
        return _get_matvecexp(uks, H_all, psi, input_num, taylor_terms)

    @staticmethod
    def backward(ctx, grad):
        uks, H_all, psi = ctx.saved_tensors
        input_num, taylor_terms = ctx.cfg
        matvecexp = _get_matvecexp(uks, H_all, psi, input_num, taylor_terms)
        coeff = [torch.zeros((), dtype=grad.dtype)]
        for ii in range(1, input_num):
            coeff.append(torch.sum(grad * torch.matmul(H_all[ii], matvecexp)))
        vec_grad = grad                                   # :118-131
        H = None
        for ii in range(input_num):
            term = (-uks[ii]) * H_all[ii]
            H = term if H is None else H + term
        vec_grad_n = grad
        factorial = 1.
        for ii in range(1, taylor_terms):
            factorial = factorial * ii
            vec_grad_n = torch.matmul(H, vec_grad_n)
            vec_grad = vec_grad + vec_grad_n / factorial
        return torch.stack(coeff), None, vec_grad, None, None


def _l2_loss(x):
    """tf.nn.l2_loss = sum(x**2)/2"""
    return torch.sum(x * x) / 2


def _inner_product_2D(psi1, psi2, n, m):
    """core/tensorflow_state.py:282-300"""
    a, b = psi1[0:n, :], psi1[n:2 * n, :]
    c, d = psi2[0:n, :], psi2[n:2 * n, :]
    ac = torch.sum(a * c, 0)
    bd = torch.sum(b * d, 0)
    bc = torch.sum(b * c, 0)
    ad = torch.sum(a * d, 0)
    reals = torch.square(torch.sum(ac + bd))
    imags = torch.square(torch.sum(bc - ad))
    return (reals + imags) / (m ** 2)


def _inner_product_3D(psi1, psi2, n, m):
    """core/tensorflow_state.py:302-321  (psi: [2n, T+1, m])"""
    a, b = psi1[0:n, :], psi1[n:2 * n, :]
    c, d = psi2[0:n, :], psi2[n:2 * n, :]
    ac = torch.sum(a * c, 0)
    bd = torch.sum(b * d, 0)
    bc = torch.sum(b * c, 0)
    ad = torch.sum(a * d, 0)
    reals = torch.sum(torch.square(torch.sum(ac + bd, 1)))
    imags = torch.sum(torch.square(torch.sum(bc - ad, 1)))
    return (reals + imags) / (m ** 2)


@dataclass
class GraphOutputs:
    loss: float
    reg_loss: float
    grad: np.ndarray            # [K,T] d reg_loss / d ops_weight_base
    unitary_scale: float
    grad_squared: float         # sum l2_loss(g) = sum g^2 / 2   (:352-353)
    final_state: np.ndarray     # [2n,2n] (unitary mode) or [2n,m] (state transfer)
    inter_vecs: Optional[np.ndarray]   # [m, 2n, T+1]  (= tf.stack(tfs.inter_vecs))
    ops_weight: np.ndarray      # sin(base)


def graph_value_and_grad(setup: OracleSetup, base, dtype=torch.float64, want_grad=True) -> GraphOutputs:
    """One evaluation of the reference graph (what ``run_session.get_error`` fetches,
    core/run_session.py:119-127): build_graph order follows core/tensorflow_state.py:366-394."""
    n, K, T = setup.n, setup.K, setup.steps
    input_num = K + 1
    m = len(setup.states_concerned_list)
    td = dict(dtype=dtype)
    one_minus_gauss = torch.tensor(setup.one_minus_gauss, **td)                 # :146-147
    V = torch.tensor(np.array(setup.initial_vectors), **td).t()                # :150-156  [2n,m]
    if setup.state_transfer:                                                    # :158-165
        target_vecs = torch.tensor(np.array(setup.target_vectors), **td).t()
    else:
        U0 = torch.tensor(setup.initial_unitary, **td)
        target_vecs = torch.matmul(torch.tensor(setup.target_unitary, **td), V)
    H_all = torch.tensor(setup.matrix_list, **td)

    w_base = torch.tensor(np.asarray(base), **td).clone().requires_grad_(want_grad)   # :174
    ops_weight = torch.sin(w_base)                                              # :176
    rows = [torch.ones(T, **td)]                                                # :172-173
    for ii in range(K):
        rows.append(float(setup.ops_max_amp[ii]) * ops_weight[ii, :])           # :177-178
    H_weights = torch.stack(rows)                                               # :181

    inter_vecs_packed = None
    if not setup.state_transfer:
        inter_states = []                                                       # :204-223
        for t in range(T):
            P = _MatExpOp.apply(H_weights[:, t], H_all, input_num, setup.exp_terms, setup.scaling)
            inter_states.append(torch.matmul(P, U0 if t == 0 else inter_states[t - 1]))
        final_state = inter_states[T - 1]
        unitary_scale = (0.5 / n) * torch.sum(torch.matmul(final_state.t(), final_state))   # :225
        if setup.use_inter_vecs:                                                # :229-240
            lst = [V] + [torch.matmul(inter_states[t], V) for t in range(T)]
            inter_vecs_packed = torch.stack(lst, dim=1)                         # [2n, T+1, m]
        final_vecs = torch.matmul(final_state, V)                               # :327
        loss = 1 - _inner_product_2D(final_vecs, target_vecs, n, m)            # :329
    else:
        lst = [V]                                                               # :244-258
        vec = V
        for t in range(T):
            vec = _MatVecExpOp.apply(H_weights[:, t], H_all, vec, input_num, setup.exp_terms)
            lst.append(vec)
        inter_vecs_packed = torch.stack(lst, dim=1)
        final_state = inter_vecs_packed[:, T, :]                                # :333
        loss = 1 - _inner_product_2D(final_state, target_vecs, n, m)           # :334
        unitary_scale = _inner_product_2D(final_state, final_state, n, m)      # :335

    reg_loss = _reg_loss(setup, loss, ops_weight, one_minus_gauss, inter_vecs_packed, target_vecs, dtype)

    if want_grad:
        (g,) = torch.autograd.grad(reg_loss, w_base)                            # :348-350
        grad = g.detach().numpy().copy()
        grad_squared = float(torch.sum(g * g) / 2)                              # :352-353
    else:
        grad = np.zeros((K, T))
        grad_squared = 0.0
    iv = None
    if inter_vecs_packed is not None:
        iv = inter_vecs_packed.detach().permute(2, 0, 1).numpy().copy()         # unstack axis 2 -> m x [2n,T+1]
    return GraphOutputs(loss=float(loss.detach()), reg_loss=float(reg_loss.detach()), grad=grad,
                        unitary_scale=float(unitary_scale.detach()), grad_squared=grad_squared,
                        final_state=final_state.detach().numpy().copy(), inter_vecs=iv,
                        ops_weight=ops_weight.detach().numpy().copy())


def _reg_loss(setup, loss, ops_weight, one_minus_gauss, inter_vecs_packed, target_vecs, dtype):
    """core/regularization_functions.py:7-97.  Terms are enabled by key PRESENCE."""
    rc = setup.reg_coeffs
    n, K, T = setup.n, setup.K, setup.steps
    steps_f = float(T)
    reg_loss = loss
    if 'amplitude' in rc:                                                       # :15-18
        reg_loss = reg_loss + (rc['amplitude'] / steps_f) * _l2_loss(ops_weight)
    if 'envelope' in rc:                                                        # :21-25
        reg_loss = reg_loss + (rc['envelope'] / steps_f) * _l2_loss(one_minus_gauss * ops_weight)
    new_weights = None
    if 'dwdt' in rc:                                                            # :28-35
        z2 = torch.zeros([K, 2], dtype=dtype)
        new_weights = torch.cat([z2, torch.cat([ops_weight, z2], 1)], 1)
        reg_loss = reg_loss + (rc['dwdt'] / steps_f) * _l2_loss(
            (new_weights[:, 1:] - new_weights[:, :T + 3]) / setup.dt)
    if 'd2wdt2' in rc:                                                          # :38-45
        if new_weights is None:
            raise NameError("name 'new_weights' is not defined")   # reference quirk
        reg_loss = reg_loss + (rc['d2wdt2'] / steps_f) * _l2_loss(
            (new_weights[:, 2:] - 2 * new_weights[:, 1:T + 3] + new_weights[:, :T + 2]) / (setup.dt ** 2))
    if 'bandpass' in rc:                                                        # :47-67 (dead code: tf.complex_abs)
        raise ValueError('bandpass regulariser is out of scope (tf.complex_abs was removed in TF 1.0)')
    if 'forbidden_coeff_list' in rc:                                            # :71-85
        m = inter_vecs_packed.shape[2]
        v_sorted = None
        if setup.is_dressed:
            v_sorted = torch.tensor(c_to_r_mat(np.reshape(
                sort_ev(setup.v_c, setup.dressed_id), [len(setup.dressed_id), len(setup.dressed_id)])), dtype=dtype)
        for j in range(m):
            inter_vec = inter_vecs_packed[:, :, j]                              # [2n, T+1]
            if setup.is_dressed and ('forbid_dressed' in rc and rc['forbid_dressed']):
                inter_vec = torch.matmul(v_sorted.t(), inter_vec)
            for coeff, state in zip(rc['forbidden_coeff_list'], rc['states_forbidden_list']):
                alpha = coeff / steps_f
                pop = torch.square(inter_vec[state, :]) + torch.square(inter_vec[n + state, :])
                reg_loss = reg_loss + alpha * _l2_loss(pop)
    if 'speed_up' in rc:                                                        # :88-95
        m = inter_vecs_packed.shape[2]
        tgt = target_vecs.reshape(2 * n, 1, m).repeat(1, T + 1, 1)
        ip = _inner_product_3D(inter_vecs_packed, tgt, n, m)
        reg_loss = reg_loss + (rc['speed_up'] / steps_f) * _l2_loss(T + 1 - ip)
    return reg_loss


# ----------------------------------------------------------------------------------------
# core/run_session.py (driver) + core/convergence.py:16-49 (defaults) + core/analysis.py
# ----------------------------------------------------------------------------------------
CONVERGENCE_DEFAULTS = dict(rate=0.01, update_step=100, evol_save_step=100, conv_target=1e-8,
                            max_iterations=5000, learning_rate_decay=2500, min_grad=1e-25)


class TF1Adam:
    """tf.train.AdamOptimizer (TF 1.x; third-party, formula from its documentation):
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
    theta -= lr_t * m / (sqrt(v) + eps).  Defaults b1=.9, b2=.999, eps=1e-8
    (core/tensorflow_state.py:345 passes only learning_rate)."""

    def __init__(self, shape, dtype=np.float64, beta1=0.9, beta2=0.999, eps=1e-8):
        self.m = np.zeros(shape, dtype=dtype)
        self.v = np.zeros(shape, dtype=dtype)
        self.t = 0
        self.b1, self.b2, self.eps = beta1, beta2, eps
        self.dtype = dtype

    def step(self, theta, g, lr):
        dt = self.dtype
        self.t += 1
        b1, b2 = dt(self.b1), dt(self.b2)
        lr_t = dt(lr) * np.sqrt(dt(1) - b2 ** self.t) / (dt(1) - b1 ** self.t)
        g = g.astype(dt)
        self.m = b1 * self.m + (dt(1) - b1) * g
        self.v = b2 * self.v + (dt(1) - b2) * g * g
        return (theta - lr_t * self.m / (np.sqrt(self.v) + dt(self.eps))).astype(dt)


@dataclass
class GrapeResult:
    uks: np.ndarray
    U_final: object
    iterations: int
    loss: float
    reg_loss: float
    history: list = field(default_factory=list)
    base: Optional[np.ndarray] = None


def run_adam(setup: OracleSetup, convergence=None, dtype=torch.float64, reference_cost=False) -> GrapeResult:
    """core/run_session.py:47-69 (loop, stop rules, lr schedule with the already-incremented
    counter), :94-117 (end results, uks = maxA * sin(base)), core/analysis.py:18-35.
    ``reference_cost=True`` repeats the reference's second full fwd+bwd per iteration
    (:53-54 fetch, then :69 optimizer run) for honest CPU timing."""
    conv = dict(CONVERGENCE_DEFAULTS)
    conv.update(convergence or {})
    npdt = np.float32 if dtype == torch.float32 else np.float64
    base = np.asarray(setup.ops_weight_base, dtype=npdt)
    adam = TF1Adam(base.shape, dtype=npdt)
    iterations = 0
    hist = []
    while True:
        out = graph_value_and_grad(setup, base, dtype)                     # :53-54
        hist.append((out.loss, out.reg_loss, out.grad_squared, out.unitary_scale))
        end = (out.loss < conv['conv_target']) or (out.grad_squared < conv['min_grad']) \
            or (iterations >= conv['max_iterations'])                           # :56-58
        if not end:
            iterations += 1                                                     # :92
        if end:
            break                                                               # :62-64
        lr = float(conv['rate']) * np.exp(-float(iterations) / conv['learning_rate_decay'])   # :66
        if reference_cost:
            out = graph_value_and_grad(setup, base, dtype)                 # :69 re-runs the graph
        base = adam.step(base, out.grad, lr)
    uks = out.ops_weight.copy()                                                 # :112-117
    for ii in range(len(uks)):
        uks[ii] = setup.ops_max_amp[ii] * uks[ii]
    Uf = [] if setup.state_transfer else r_to_c_mat(out.final_state, setup.n)  # :106-110
    return GrapeResult(uks=uks, U_final=Uf, iterations=iterations, loss=out.loss, reg_loss=out.reg_loss,
                       history=hist, base=base)


def grape(H0, Hops, Hnames, U, total_time, steps, states_concerned_list, convergence=None, U0=None,
          reg_coeffs=None, dressed_info=None, maxA=None, initial_guess=None, unitary_error=1e-4,
          method='Adam', state_transfer=False, no_scaling=False, Taylor_terms=None, use_inter_vecs=True,
          dtype=torch.float64, **_ignored):
    """main_grape/grape.py:19-129 without the I/O: returns (uks, U_final)."""
    setup = make_setup(H0, Hops, U, total_time, steps, states_concerned_list, U0=U0, reg_coeffs=reg_coeffs,
                       dressed_info=dressed_info, maxA=maxA, initial_guess=initial_guess,
                       unitary_error=unitary_error, state_transfer=state_transfer, no_scaling=no_scaling,
                       Taylor_terms=Taylor_terms, use_inter_vecs=use_inter_vecs)
    if method.upper() == 'EVOLVE':                                              # run_session.py:33-37
        conv = dict(convergence or {})
        conv['max_iterations'] = 0
        res = run_adam(setup, conv, dtype)
    elif method.upper() == 'ADAM':
        res = run_adam(setup, convergence, dtype)
    else:
        raise NotImplementedError('oracle covers ADAM and EVOLVE')
    return res.uks, res.U_final


# ----------------------------------------------------------------------------------------
# Independent formulation: complex costate recursion (what the CUDA kernels implement)
# ----------------------------------------------------------------------------------------
def costate_value_and_grad(setup: OracleSetup, base):
    """Complex n x n restatement of SURVEY.md 3.4 (NumPy complex128), unitary mode, bare or
    dressed initial vectors, regularisers amplitude/envelope/dwdt/d2wdt2/forbidden/speed_up.
    Returns dict(loss, reg_loss, grad[K,T], unitary_scale, grad_squared, U_final[n,n],
    inter_vecs[m,T+1,n], P[T,n,n])."""
    assert not setup.state_transfer
    n, K, T = setup.n, setup.K, setup.steps
    p, s = setup.exp_terms, setup.scaling
    rc = setup.reg_coeffs
    A = np.array([r_to_c_mat(setup.matrix_list[k], n) for k in range(K + 1)])   # -i dt H_k
    U0 = r_to_c_mat(setup.initial_unitary, n)
    Ut = r_to_c_mat(setup.target_unitary, n)
    V = np.array([v[:n] + 1j * v[n:] for v in setup.initial_vectors])           # [m,n]
    m = len(V)
    Phi = (Ut @ V.T).T                                                          # [m,n]
    base = np.asarray(base, dtype=np.float64)
    w = np.sin(base)
    u = setup.ops_max_amp[:, None] * w

    P = np.empty((T, n, n), dtype=np.complex128)
    for t in range(T):
        H = (A[0] + np.tensordot(u[:, t], A[1:], axes=1)) / 2.0 ** s
        S = np.eye(n, dtype=np.complex128)
        Hn = np.eye(n, dtype=np.complex128)
        for j in range(1, p + 1):
            Hn = H @ Hn / j
            S = S + Hn
        for _ in range(s):
            S = S @ S
        P[t] = S
    X = U0.copy()
    psi = np.empty((T + 1, m, n), dtype=np.complex128)
    psi[0] = V
    chi = (U0 @ V.T).T
    for t in range(T):
        X = P[t] @ X
        chi = (P[t] @ chi.T).T
        psi[t + 1] = chi
    rows = X.sum(axis=1)
    unitary_scale = float(np.sum(np.abs(rows) ** 2) / n)
    o = np.sum(np.conj(Phi) * psi[T])
    loss = 1 - abs(o) ** 2 / m ** 2

    reg = loss
    src = np.zeros((T + 1, m, n), dtype=np.complex128)
    if 'forbidden_coeff_list' in rc:
        if setup.is_dressed and rc.get('forbid_dressed'):
            raise NotImplementedError('forbid_dressed is not covered by the costate form')
        for coeff, state in zip(rc['forbidden_coeff_list'], rc['states_forbidden_list']):
            alpha = coeff / float(T)
            amp = psi[:, :, state]
            pop = np.abs(amp) ** 2
            reg = reg + alpha * 0.5 * np.sum(pop ** 2)
            src[:, :, state] += alpha * 2 * pop * amp
    if 'speed_up' in rc:
        c = rc['speed_up'] / float(T)
        ot = np.einsum('jn,tjn->t', np.conj(Phi), psi)
        S_ = np.sum(np.abs(ot) ** 2) / m ** 2
        reg = reg + c * 0.5 * (T + 1 - S_) ** 2
        src += (-c * (T + 1 - S_) * (2.0 / m ** 2)) * ot[:, None, None] * Phi[None, :, :]

    lam = -(2.0 / m ** 2) * o * Phi + src[T]
    gu = np.zeros((K, T))
    for t in range(T - 1, -1, -1):
        for k in range(K):
            gu[k, t] = np.sum(np.real(np.conj(lam) * (A[k + 1] @ psi[t + 1].T).T))
        lam = (P[t].conj().T @ lam.T).T
        if t >= 1:
            lam = lam + src[t]
    gw = setup.ops_max_amp[:, None] * gu
    dt = setup.dt
    if 'amplitude' in rc:
        c = rc['amplitude'] / float(T)
        reg = reg + c * 0.5 * np.sum(w ** 2)
        gw = gw + c * w
    if 'envelope' in rc:
        c = rc['envelope'] / float(T)
        E = setup.one_minus_gauss
        reg = reg + c * 0.5 * np.sum((E * w) ** 2)
        gw = gw + c * E * E * w
    if 'dwdt' in rc:
        c = rc['dwdt'] / float(T)
        z = np.pad(w, ((0, 0), (2, 2)))
        d = (z[:, 1:] - z[:, :-1]) / dt
        reg = reg + c * 0.5 * np.sum(d ** 2)
        gz = np.zeros_like(z)
        gz[:, 1:] += d / dt
        gz[:, :-1] -= d / dt
        gw = gw + c * gz[:, 2:-2]
    if 'd2wdt2' in rc:
        c = rc['d2wdt2'] / float(T)
        z = np.pad(w, ((0, 0), (2, 2)))
        e = (z[:, 2:] - 2 * z[:, 1:-1] + z[:, :-2]) / dt ** 2
        reg = reg + c * 0.5 * np.sum(e ** 2)
        gz = np.zeros_like(z)
        gz[:, 2:] += e / dt ** 2
        gz[:, 1:-1] -= 2 * e / dt ** 2
        gz[:, :-2] += e / dt ** 2
        gw = gw + c * gz[:, 2:-2]
    grad = gw * np.cos(base)
    return dict(loss=float(loss), reg_loss=float(reg), grad=grad, unitary_scale=unitary_scale,
                grad_squared=float(np.sum(grad ** 2) / 2), U_final=X, inter_vecs=np.transpose(psi, (1, 0, 2)), P=P)
