"""Host-side problem setup: everything the reference's ``SystemParameters`` object
(core/system_parameters.py:10-286) computes before the graph is built, restated for Python 3 in
complex (n x n) form -- the layout the CUDA engine consumes -- instead of the reference's
real-embedded 2n x 2n float32 form.

The quirks that change numerical results are kept bug-for-bug: the Taylor-order
chooser (cumulative scaling candidates, un-reset test propagator, "first failing order"), the
one-sided initial-guess check, the envelope offset, N(0, 1/sqrt(T)) initial weights.
"""
import numpy as np

from ..helper_functions.grape_functions import c_to_r_mat, c_to_r_vec, get_state_index


def _series_matrix(M, order_excl, squarings):
    """sum_{j<order_excl} (M/2^squarings)^j / j!, squared ``squarings`` times
    (the estimator of core/system_parameters.py:88-103; one order LOWER than what the graph sums)."""
    eye = np.identity(len(M), dtype=M.dtype)
    acc, power, fact = eye.copy(), eye.copy(), 1.0
    for j in range(1, order_excl):
        fact *= j
        power = power @ M
        acc = acc + power / ((2.0 ** float(j * squarings)) * fact)
    for _ in range(squarings):
        acc = acc @ acc
    return acc


def _series_scalar(x, order_excl, squarings):
    """Scalar twin of :func:`_series_matrix` (core/system_parameters.py:105-120)."""
    acc, power, fact = 1.0, 1.0, 1.0
    for j in range(1, order_excl):
        fact *= j
        power = x * power
        acc += power / ((2.0 ** float(j * squarings)) * fact)
    for _ in range(squarings):
        acc = acc * acc
    return acc


def choose_taylor_terms(H0, Hops, maxA, U0, dt, steps, unitary_error, fixed_zero_scaling=False):
    """Pick (exp_terms, scaling) exactly as core/system_parameters.py:122-158 + :208-227 does.

    Candidates: s0 = max(int(2*log2(max|dt*H_max|)), 0) and then s0+1, +3, +6, +10, +15 (the
    reference's ``scaling += d`` is cumulative); for each, the Taylor order is lowered from 20
    while the unitarity metric stays within ``unitary_error`` and the FIRST FAILING order (floor 3)
    is returned; the candidate minimising order + squarings wins (first on ties)."""
    n = len(H0)
    Hmax = np.asarray(H0)
    for amp, op in zip(maxA, Hops):
        Hmax = Hmax + amp * np.asarray(op)
    gen = (0 - 1j) * dt * Hmax
    peak = np.max(np.abs(gen))
    candidates = 1 if fixed_zero_scaling else 6
    found = []
    scaling = None
    for d in range(candidates):
        scaling = max(int(2 * np.log2(peak)), 0) if d == 0 else scaling + d
        if fixed_zero_scaling:
            scaling = 0
        order = 20
        probe = np.asarray(U0)                 # keeps accumulating across orders, like the reference
        while True:
            if n < 10:
                step = _series_matrix(gen, order, scaling)
                for _ in range(steps):
                    probe = probe @ step
                metric = np.abs(np.trace(probe.conj().T @ probe)) / n
            else:
                metric = 1 + steps * np.abs((_series_scalar(peak, order, scaling) - np.exp(peak)) / np.exp(peak))
            if order == 3 or not (np.abs(metric - 1.0) < unitary_error):
                break
            order -= 1
        found.append((order, scaling))
    best = int(np.argmin([o + s for o, s in found]))
    return found[best]


def one_minus_gaussian_envelope(K, steps):
    """[K, T] envelope max(1 - exp(-x^2/2), 0) + 0.01, x = linspace(-2, 2, T)
    (core/system_parameters.py:253-270)."""
    x = np.linspace(-2, 2, steps)
    row = np.ones(steps) - np.exp(-np.power(x, 2.) / 2.0)
    row = row * (row > 0) + 0.01 * np.ones(steps)
    return np.tile(row, (K, 1))


class SystemParameters:
    """Same constructor signature and attribute names as the reference class
    (core/system_parameters.py:12-13); complex-form extras carry a ``_c`` suffix.

    Batched use (ours): ``initial_guess`` of shape [B, K, T] -> ``ops_weight_base`` [B, K, T];
    with no guess, ``batch`` independent N(0, 1/sqrt(T)) draws are taken from the global NumPy RNG
    (the reference draws one, :278-282)."""

    def __init__(self, H0, Hops, Hnames, U, U0, total_time, steps, states_concerned_list, dressed_info, maxA,
                 draw, initial_guess, show_plots, Unitary_error, state_transfer, no_scaling, reg_coeffs, save,
                 file_path, Taylor_terms, use_gpu, use_inter_vecs, sparse_H, sparse_U, sparse_K, batch=None):
        self.sparse_H, self.sparse_U, self.sparse_K = sparse_H, sparse_U, sparse_K
        self.use_inter_vecs, self.use_gpu = use_inter_vecs, use_gpu
        self.Taylor_terms = Taylor_terms
        self.dressed_info = dressed_info
        self.reg_coeffs = {} if reg_coeffs is None else reg_coeffs
        self.file_path, self.save = file_path, save
        self.state_transfer, self.no_scaling = state_transfer, no_scaling
        self.H0_c = np.asarray(H0)
        self.ops_c = [np.asarray(h) for h in Hops]
        self.ops_max_amp = np.asarray(maxA, dtype=np.float64)
        self.Hnames = self.Hnames_original = Hnames
        self.total_time, self.steps = total_time, steps
        self.show_plots = show_plots
        self.Unitary_error = Unitary_error
        self.states_concerned_list = states_concerned_list
        self.U0_c = np.asarray(U0)
        self.draw_list, self.draw_names = (draw[0], draw[1]) if draw is not None else ([], [])

        self.ops_len = len(self.ops_c)
        self.state_num = len(self.H0_c)
        self.dt = float(total_time) / steps                      # :163-165

        self._init_guess_base(initial_guess)
        self._init_dressed(dressed_info)
        self._init_vectors(U)
        self._init_operators()
        self.one_minus_gauss = one_minus_gaussian_envelope(self.ops_len, steps)
        self._init_weights(batch)

    # -- :38-46 -------------------------------------------------------------------------------
    def _init_guess_base(self, initial_guess):
        self.u0, self.u0_base = [], None
        if initial_guess is None:
            return
        g = np.asarray(initial_guess, dtype=np.float64)
        self.u0 = g
        ratio = g / self.ops_max_amp.reshape((1,) * (g.ndim - 2) + (-1, 1))
        rows = ratio.reshape(-1, ratio.shape[-2], ratio.shape[-1])
        for inst in rows:
            for k, row in enumerate(inst):
                if np.max(row) > 1.0:                            # one-sided, like the reference (:44)
                    raise ValueError('Initial guess has strength > max_amp for op %d' % (k))
        self.u0_base = np.arcsin(ratio)

    # -- :75-80 -------------------------------------------------------------------------------
    def _init_dressed(self, info):
        self.is_dressed = False
        if info is not None:
            self.v_c = info['eigenvectors']
            self.dressed_id = info['dressed_id']
            self.w_c = info['eigenvalues']
            self.is_dressed = info['is_dressed']
            self.H0_diag = np.diag(self.w_c)

    # -- :57-65, :168-187 -----------------------------------------------------------------------
    def _init_vectors(self, U):
        n = self.state_num
        vecs = []
        bare = True
        for state in self.states_concerned_list:
            if self.state_transfer:
                v = np.array(state)
                bare = False
            elif self.is_dressed:
                v = self.v_c[:, get_state_index(state, self.dressed_id)]
                bare = False
            else:
                v = np.zeros(n)
                v[state] = 1
            vecs.append(np.asarray(v))
        self.initial_vectors_c = vecs
        self.initial_vectors = [c_to_r_vec(v) for v in vecs]
        self.concerned_idx = np.asarray(self.states_concerned_list, dtype=np.int32) if bare else None
        V = np.array(vecs, dtype=np.complex128).reshape(len(vecs), n)
        self.V_c = V
        if self.state_transfer:
            self.target_vectors_c = np.array([np.asarray(v) for v in U], dtype=np.complex128)
            self.target_vectors = [c_to_r_vec(v) for v in self.target_vectors_c]
        else:
            self.U_c = np.asarray(U, dtype=np.complex128)
            self.target_unitary = c_to_r_mat(self.U_c)
            self.target_vectors_c = (self.U_c @ V.T).T            # tensorflow_state.py:165
        self.initial_unitary = c_to_r_mat(self.U0_c)

    # -- :194-251 -------------------------------------------------------------------------------
    def _init_operators(self):
        if self.Taylor_terms is None:
            self.exp_terms, self.scaling = choose_taylor_terms(
                self.H0_c, self.ops_c, self.ops_max_amp, self.U0_c, self.dt, self.steps, self.Unitary_error,
                fixed_zero_scaling=bool(self.state_transfer or self.no_scaling))
        else:
            self.exp_terms, self.scaling = int(self.Taylor_terms[0]), int(self.Taylor_terms[1])
        print("Using " + str(self.exp_terms) + " Taylor terms and " + str(self.scaling) + " Scaling & Squaring terms")
        # generators -i*dt*H_k, complex form (the reference stores c_to_r_mat of these, :199,:204)
        self.A_c = np.array([-1j * self.dt * self.H0_c] + [-1j * self.dt * op for op in self.ops_c],
                            dtype=np.complex128)

    @property
    def matrix_list(self):
        """The reference's [K+2, 2n, 2n] real stack incl. the trailing identity (:246-251)."""
        return np.array([c_to_r_mat(a) for a in self.A_c] + [np.eye(2 * self.state_num)])

    # -- :272-284 -------------------------------------------------------------------------------
    def _init_weights(self, batch):
        K, T = self.ops_len, self.steps
        if self.u0_base is not None:
            base = np.asarray(self.u0_base)
            self.batched = base.ndim == 3
            self.ops_weight_base = base.reshape((-1, K, T)) if self.batched else base.reshape(K, T)
        else:
            sd = 1. / np.sqrt(T)
            self.batched = batch is not None
            if self.batched:
                self.ops_weight_base = np.stack([np.random.normal(0, sd, [K, T]) for _ in range(int(batch))])
            else:
                self.ops_weight_base = np.random.normal(0, sd, [K, T])
        self.raw_shape = np.shape(self.ops_weight_base)

    @property
    def batch_size(self):
        return self.ops_weight_base.shape[0] if self.batched else 1
