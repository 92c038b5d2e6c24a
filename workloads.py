"""Synthetic GRAPE problem definitions for the BASELINE.json configs.

The reference's ``examples`` submodule is absent from the tree, so these stand in for it.
Everything is built from ``np.random.default_rng(seed)`` and closed-form operators; used by
``tests/``, ``bench.py`` and ``__graft_entry__.smoke()``.  Pure NumPy -- no oracle, no CUDA.
"""
import os
import sys

import numpy as np

_PKG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "quantum-optimal-control_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from quantum_optimal_control.helper_functions.grape_functions import (  # noqa: E402
    transmon_gate, concerned, nn_chain_kron)


def _destroy(d):
    return np.diag(np.sqrt(np.arange(1, d)), 1).astype(complex)


def random_guess(K, T, maxA, seed, B=None):
    """base ~ N(0, 1/sqrt(T)) exactly as the reference draws it (core/system_parameters.py:278-282),
    returned as the ``initial_guess`` (= maxA*sin(base)) the reference API takes.  B=None -> [K,T]."""
    maxA = np.asarray(maxA, dtype=np.float64)
    seeds = [seed] if B is None else [seed + b for b in range(B)]
    out = np.stack([maxA[:, None] * np.sin(np.random.default_rng(sd).normal(0, 1 / np.sqrt(T), (K, T)))
                    for sd in seeds])
    return out[0] if B is None else out


def c1_pi_pulse(T=100):
    """C1: single-qubit pi pulse, lab frame 3.9 GHz, n=2, K=2, T=100, 10 ns."""
    sx = np.array([[0, 1], [1, 0]], dtype=complex)
    sz = np.array([[1, 0], [0, -1]], dtype=complex)
    return dict(H0=2 * np.pi * 3.9 * np.diag([0, 1]).astype(complex), Hops=[sx, sz], Hnames=['x', 'z'],
                U=sx.copy(), total_time=10.0, steps=T, states_concerned_list=[0, 1],
                maxA=[2.0, 2.0], reg_coeffs={}, unitary_error=1e-4)


def c2_transmon_cavity(T=500, n_q=3, n_c=10):
    """C2: 3-level transmon (x) 10-level cavity in the rotating frame, n=30, K=4, T=500, 1000 ns,
    m=2 (cavity |0>,|1> with the transmon in g); units GHz/ns."""
    a = np.kron(np.eye(n_q), _destroy(n_c))
    b = np.kron(_destroy(n_q), np.eye(n_c))
    ad, bd = a.conj().T, b.conj().T
    chi, alpha = -2 * np.pi * 2.2e-3, -2 * np.pi * 0.2
    H0 = chi * (ad @ a) @ (bd @ b) + 0.5 * alpha * (bd @ bd @ b @ b)
    Hops = [a + ad, 1j * (a - ad), b + bd, 1j * (b - bd)]
    n = n_q * n_c
    U = np.eye(n, dtype=complex)
    U[0, 0] = U[1, 1] = 0
    U[0, 1] = U[1, 0] = 1                      # X on the cavity {|0>,|1>} manifold
    return dict(H0=H0, Hops=Hops, Hnames=['ax', 'ay', 'bx', 'by'], U=U, total_time=1000.0, steps=T,
                states_concerned_list=[0, 1], maxA=[2 * np.pi * 0.05] * 4, reg_coeffs={}, unitary_error=1e-4)


def _transmon(levels, f, anh):
    d = _destroy(levels)
    num = d.conj().T @ d
    return 2 * np.pi * (f * num + 0.5 * anh * (num @ num - num)), d + d.conj().T, num


def c3_two_transmon_cnot(T=1000, levels=6):
    """C3: two coupled 6-level transmons (lab frame), CNOT, n=36, K=4, T=1000, 10 ns, m=4,
    forbidden-state regulariser on every state with a level >= 3."""
    H1, x1, n1 = _transmon(levels, 3.9, -0.225)
    H2, x2, n2 = _transmon(levels, 3.5, -0.225)
    I = np.eye(levels)
    H0 = np.kron(H1, I) + np.kron(I, H2) + 2 * np.pi * 0.1 * np.kron(x1, x2)
    Hops = [np.kron(x1, I), np.kron(I, x2), np.kron(n1, I), np.kron(I, n2)]
    cnot = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=complex)
    forb = [i for i in range(levels ** 2) if max(divmod(i, levels)) >= 3]
    return dict(H0=H0, Hops=Hops, Hnames=['x1', 'x2', 'n1', 'n2'], U=transmon_gate(cnot, levels),
                total_time=10.0, steps=T, states_concerned_list=concerned(2, levels), maxA=[2.0] * 4,
                reg_coeffs={'forbidden_coeff_list': [10.0] * len(forb), 'states_forbidden_list': forb},
                unitary_error=1e-4)


def c4_three_transmon_toffoli(T=2000, levels=6):
    """C4: three 6-level transmons, nearest-neighbour x(x)x coupling, Toffoli, n=216, K=3, T=2000,
    100 ns, m=8, dwdt + d2wdt2 regularisers."""
    I = np.eye(levels)
    freqs = [3.9, 3.5, 4.2]
    parts = [_transmon(levels, f, -0.225) for f in freqs]

    def on(site, op):
        mats = [op if k == site else I for k in range(3)]
        return np.kron(np.kron(mats[0], mats[1]), mats[2])

    H0 = sum(on(k, parts[k][0]) for k in range(3)) + 2 * np.pi * 0.05 * nn_chain_kron(parts[0][1], I, 3, levels)
    Hops = [on(k, parts[k][1]) for k in range(3)]
    toff = np.eye(8, dtype=complex)
    toff[6:, 6:] = [[0, 1], [1, 0]]
    return dict(H0=H0, Hops=Hops, Hnames=['x1', 'x2', 'x3'], U=transmon_gate(toff, levels),
                total_time=100.0, steps=T, states_concerned_list=concerned(3, levels), maxA=[1.0] * 3,
                reg_coeffs={'dwdt': 1e-3, 'd2wdt2': 1e-6}, unitary_error=1e-4)


def c5_random(n, T=1000, K=2, seed=1234):
    """C5: GUE-like drift scaled so max|dt*H| ~ 0.3, K random Hermitian controls, m = n,
    fixed Taylor_terms=[8,2]."""
    rng = np.random.default_rng(seed + n)

    def herm():
        M = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
        return (M + M.conj().T) / 2

    total_time = float(T) * 0.1
    dt = total_time / T
    H0 = herm()
    Hops = [herm() for _ in range(K)]
    maxA = [1.0] * K
    scale = 0.3 / (dt * np.max(np.abs(H0 + sum(Hops))))
    H0 = H0 * scale
    Hops = [h * scale for h in Hops]
    Q, _ = np.linalg.qr(rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n)))
    return dict(H0=H0, Hops=Hops, Hnames=['c%d' % k for k in range(K)], U=Q, total_time=total_time, steps=T,
                states_concerned_list=list(range(n)), maxA=maxA, reg_coeffs={}, Taylor_terms=[8, 2],
                unitary_error=1e-4)


def state_transfer_random(n=8, T=20, m=2, seed=77):
    """State-transfer problem (reference: state_transfer=True): m random initial / target state vectors under
    the C5-style random Hermitian Hamiltonian; ``states_concerned_list`` holds the initial vectors, ``U`` the targets."""
    pb = c5_random(n, T=T)
    rng = np.random.default_rng(seed)

    def vec():
        v = rng.normal(size=n) + 1j * rng.normal(size=n)
        return v / np.linalg.norm(v)

    pb.pop('Taylor_terms')
    pb.update(states_concerned_list=[vec() for _ in range(m)], U=[vec() for _ in range(m)], state_transfer=True)
    return pb


WORKLOADS = {
    'C1': (c1_pi_pulse, dict(B=1)),
    'C2': (c2_transmon_cavity, dict(B=256)),
    'C3': (c3_two_transmon_cnot, dict(B=1024)),
    'C4': (c4_three_transmon_toffoli, dict(B=128)),
    # C5 is the 8-GPU sweep over n with 4096 instances in total = 512 per GPU
    'C5n8': (lambda T=1000: c5_random(8, T=T), dict(B=512)),
    'C5n16': (lambda T=1000: c5_random(16, T=T), dict(B=512)),
    'C5n32': (lambda T=1000: c5_random(32, T=T), dict(B=512)),
    'C5n64': (lambda T=1000: c5_random(64, T=T), dict(B=512)),
    'C5n128': (lambda T=1000: c5_random(128, T=T), dict(B=512)),
}


def grape_kwargs(problem):
    """Split a workload dict into the positional/keyword arguments of ``Grape(...)``."""
    kw = dict(problem)
    args = [kw.pop(k) for k in ('H0', 'Hops', 'Hnames', 'U', 'total_time', 'steps', 'states_concerned_list')]
    return args, kw
