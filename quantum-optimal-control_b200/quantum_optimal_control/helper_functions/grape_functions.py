"""Problem-construction helpers (host side, NumPy).

Same names, arguments and results as the reference's
``quantum_optimal_control/helper_functions/grape_functions.py`` (cited per function) so user
scripts keep working; written from scratch for Python 3.
"""
import numpy as np
import scipy.linalg as la

_DIGITS = "0123456789abcdefghijklmnopqrstuvwxyz"


# -- complex <-> real embedding (define the reference's float layout; grape_functions.py:211-220) --
def c_to_r_mat(M):
    """complex n x n -> real 2n x 2n ``[[Re, -Im], [Im, Re]]`` (grape_functions.py:211-213)."""
    M = np.asarray(M)
    out = np.empty((2 * M.shape[0], 2 * M.shape[1]), dtype=np.float64)
    r, c = M.shape
    out[:r, :c] = M.real
    out[:r, c:] = -M.imag
    out[r:, :c] = M.imag
    out[r:, c:] = M.real
    return out


def c_to_r_vec(V):
    """complex n -> real 2n ``[Re; Im]`` (grape_functions.py:215-220)."""
    V = np.asarray(V)
    return np.concatenate((V.real, V.imag)).astype(np.float64)


# -- dressed basis (grape_functions.py:4-24,194-209) --
def get_state_index(bareindex, dressed_id):
    """Index of the dressed state assigned to bare state ``bareindex`` (grape_functions.py:204-209)."""
    return dressed_id.index(bareindex) if len(dressed_id) > 0 else bareindex


def sort_ev(v, dressed_id):
    """Eigenvector matrix with column i = dressed partner of bare state i (grape_functions.py:194-202)."""
    order = [get_state_index(i, dressed_id) for i in range(len(dressed_id))]
    return np.asarray(v)[:, order]


def get_dressed_info(H0):
    """Eigen-decompose H0 and greedily label each eigenvector by its largest not-yet-taken bare
    component (grape_functions.py:9-24).  Returns ``(w_c, v_c, dressed_id)``."""
    w_c, v_c = la.eig(H0)
    dressed_id = []
    for col in range(len(v_c)):
        weight = np.abs(v_c[:, col]).tolist()
        idx = int(np.argmax(weight))
        while idx in dressed_id:
            weight[idx] = 0
            idx = int(np.argmax(weight))
        dressed_id.append(idx)
    return w_c, v_c, dressed_id


def dressed_unitary(U, v, dressed_id):
    """U expressed in the dressed basis (grape_functions.py:4-7)."""
    S = sort_ev(v, dressed_id)
    return S @ U @ S.conj().T


# -- gates (grape_functions.py:26-79) --
def qft(N):
    """Quantum Fourier transform on N qubits (grape_functions.py:26-32)."""
    dim = 2 ** N
    idx = np.arange(dim)
    return np.exp(2.0j * np.pi * np.outer(idx, idx) / dim) / np.sqrt(dim)


def hamming_distance(x):
    """Population count (grape_functions.py:34-39)."""
    return bin(x).count("1")


def Hadamard(N=1):
    """N-qubit Hadamard (grape_functions.py:41-46)."""
    dim = 2 ** N
    sign = np.array([[(-1) ** hamming_distance(i & j) for i in range(dim)] for j in range(dim)])
    return (2.0 ** (-N / 2.0)) * sign


def rz(theta):
    return [[np.exp(-1j * theta / 2), 0], [0, np.exp(1j * theta / 2)]]


def rx(theta):
    return [[np.cos(theta / 2), -1j * np.sin(theta / 2)],
            [-1j * np.sin(theta / 2), np.cos(theta / 2)]]


def Bin(a, N):
    """Binary string of ``a`` zero-padded to N digits (grape_functions.py:81-85)."""
    return np.binary_repr(a).rjust(N, '0')


def baseN(num, b, numerals=_DIGITS):
    """``num`` written in base ``b`` (grape_functions.py:87-88)."""
    if num == 0:
        return numerals[0]
    out = ""
    while num:
        num, d = divmod(num, b)
        out = numerals[d] + out
    return out


def Basis(a, N, r):
    """Base-r digit string of ``a``, zero-padded to N digits (grape_functions.py:90-94)."""
    return baseN(a, r).rjust(N, '0')


def is_binary(num):
    return all(c in '01' for c in num)


def concerned(N, levels):
    """Indices of the computational (all digits 0/1) states of N ``levels``-level systems
    (grape_functions.py:48-54)."""
    return [i for i in range(levels ** N) if is_binary(Basis(i, N, levels))]


def transmon_gate(gate, levels):
    """Embed a 2^N x 2^N qubit gate into N ``levels``-level transmons; identity elsewhere
    (grape_functions.py:64-74)."""
    gate = np.asarray(gate)
    N = int(np.log2(len(gate)))
    out = np.identity(levels ** N, dtype=complex)
    comp = concerned(N, levels)
    bits = [int(Basis(i, N, levels), 2) for i in comp]
    for i, bi in zip(comp, bits):
        for j, bj in zip(comp, bits):
            out[i, j] = gate[bi, bj]
    return out


# -- operator builders (grape_functions.py:97-191) --
def multi_kron(op, num):
    """op (x) op (x) ... num times (grape_functions.py:117-122)."""
    out = op
    for _ in range(num - 1):
        out = np.kron(out, op)
    return out


def kron_all(op, num, op_2):
    """Reference behaviour kept: builds the terms of ``op(x)I(x)I + I(x)op(x)I + ...`` but returns
    only the LAST term (grape_functions.py:97-115 returns ``a``, not ``total``)."""
    last = op
    for jj in range(num):
        last = op if jj == 0 else op_2
        for ii in range(num - 1):
            last = np.kron(last, op if (jj - ii) == 1 else op_2)
    return last


def append_separate_krons(op, name, num, state_num, Hops, Hnames, ops_max_amp, amp=4.0):
    """Append op on each of ``num`` sites separately: ``op i i``, ``i op i``, ...
    (grape_functions.py:124-163)."""
    eye = np.identity(state_num)
    for site in range(num):
        mats = [op if k == site else eye for k in range(num)]
        full = mats[0]
        for mtx in mats[1:]:
            full = np.kron(full, mtx)
        Hops.append(full)
        ops_max_amp.append(amp)
        Hnames.append(''.join(name if k == site else 'i' for k in range(num)))
    return Hops, Hnames, ops_max_amp


def nn_chain_kron(op, op_I, qubit_num, qubit_state_num):
    """Nearest-neighbour chain op(x)op(x)I.. + I(x)op(x)op.. + ... (grape_functions.py:165-191)."""
    dim = qubit_state_num ** qubit_num
    total = np.zeros([dim, dim])
    for first in range(qubit_num - 1):
        term = None
        for site in range(qubit_num):
            f = op if site in (first, first + 1) else op_I
            term = f if term is None else np.kron(term, f)
        total = total + term
    return total
