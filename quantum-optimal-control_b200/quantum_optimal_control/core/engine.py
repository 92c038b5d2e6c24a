"""ctypes binding of libqoc_b200.so (include/qoc_b200.h) + a thin torch-facing wrapper.

PyTorch is plumbing here: it owns device memory (workspace, weights, results) and the stream;
every number is produced by the hand-written sm_100a kernels behind the C ABI.  There is no CPU
or eager fallback -- a missing library or GPU raises.
"""
import ctypes as C
import os

import numpy as np

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB_PATH = os.path.join(_PKG_ROOT, "lib", "libqoc_b200.so")

QOC_F64, QOC_TF32X3, QOC_F16X2 = 0, 1, 2
QOC_FLAG_STATE_TRANSFER = 1
QOC_ABI_VERSION = 2          # include/qoc_b200.h
# 'f16x2': fp32-class tcgen05 / TMA path (n <= 256, m <= 8): the reference's own working precision is float32
_DTYPES = {'f64': QOC_F64, 'fp64': QOC_F64, 'float64': QOC_F64, 'tf32x3': QOC_TF32X3, 'f16x2': QOC_F16X2}

SYMBOLS = ["qoc_abi_version", "qoc_create", "qoc_destroy", "qoc_last_error", "qoc_workspace_bytes",
           "qoc_set_workspace", "qoc_set_problem", "qoc_set_regularizers", "qoc_value_and_grad", "qoc_evolve",
           "qoc_value_and_grad_host", "qoc_evolve_host", "qoc_debug_propagators", "qoc_launch_count", "qoc_set_profiling",
           "qoc_kernel_times_ms", "qoc_poll_error", "qoc_set_forbid_basis", "qoc_batch_chunk", "qoc_adam_host"]


class QocDims(C.Structure):
    _fields_ = [("n", C.c_int32), ("K", C.c_int32), ("T", C.c_int32), ("m", C.c_int32), ("B", C.c_int32),
                ("exp_terms", C.c_int32), ("scaling", C.c_int32), ("dtype", C.c_int32), ("flags", C.c_uint32)]


class QocReg(C.Structure):
    _fields_ = [("has_amplitude", C.c_int32), ("amplitude", C.c_double),
                ("has_envelope", C.c_int32), ("envelope", C.c_double),
                ("has_dwdt", C.c_int32), ("dwdt", C.c_double),
                ("has_d2wdt2", C.c_int32), ("d2wdt2", C.c_double),
                ("has_forbidden", C.c_int32),
                ("has_speed_up", C.c_int32), ("speed_up", C.c_double)]


class QocError(RuntimeError):
    pass


_lib = None


def load_library(path=None):
    """dlopen the in-tree library; fail loudly when it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("QOC_B200_LIB") or LIB_PATH        # QOC_B200_LIB: A/B runs against another build
    if not os.path.exists(path):
        raise QocError("%s not found: build it with `python __graft_entry__.py` (or "
                       "quantum-optimal-control_b200/build.py); there is no CPU fallback" % path)
    lib = C.CDLL(path)
    vp, dp, ip = C.c_void_p, C.c_void_p, C.c_void_p
    lib.qoc_abi_version.restype = C.c_int
    if lib.qoc_abi_version() != QOC_ABI_VERSION:
        raise QocError("%s has ABI version %d, this package needs %d: rebuild it (python __graft_entry__.py)"
                       % (path, lib.qoc_abi_version(), QOC_ABI_VERSION))
    lib.qoc_create.argtypes = [C.POINTER(vp), C.POINTER(QocDims)]
    lib.qoc_destroy.argtypes = [vp]
    lib.qoc_last_error.argtypes = [vp]
    lib.qoc_last_error.restype = C.c_char_p
    lib.qoc_workspace_bytes.argtypes = [vp, C.POINTER(C.c_size_t)]
    lib.qoc_set_workspace.argtypes = [vp, vp, C.c_size_t]
    lib.qoc_set_problem.argtypes = [vp, dp, dp, dp, dp, ip, dp, C.c_double, vp]
    lib.qoc_set_regularizers.argtypes = [vp, C.POINTER(QocReg), dp, dp, vp]
    lib.qoc_value_and_grad.argtypes = [vp, dp, dp, dp, dp, dp, dp, vp]
    lib.qoc_evolve.argtypes = [vp, dp, dp, dp, dp, dp, vp]
    lib.qoc_value_and_grad_host.argtypes = [vp, dp, dp, dp, dp, dp, dp, vp]
    lib.qoc_evolve_host.argtypes = [vp, dp, dp, dp, dp, dp, vp]
    lib.qoc_debug_propagators.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_int)]
    lib.qoc_launch_count.argtypes = [vp]
    lib.qoc_launch_count.restype = C.c_int64
    lib.qoc_poll_error.argtypes = [vp, vp]
    lib.qoc_batch_chunk.argtypes = [vp]
    lib.qoc_set_forbid_basis.argtypes = [vp, dp, vp]
    lib.qoc_set_profiling.argtypes = [vp, C.c_int]
    lib.qoc_kernel_times_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.qoc_adam_host.argtypes = [dp, dp, dp, dp, C.c_size_t, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
    for fn in SYMBOLS:
        if fn not in ("qoc_last_error", "qoc_launch_count", "qoc_abi_version"):
            getattr(lib, fn).restype = C.c_int
    if path == LIB_PATH:
        _lib = lib
    return lib


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def reg_struct(reg_coeffs, n, T):
    """reg_coeffs dict (core/regularization_functions.py keys) -> (QocReg, forbid_weight[n] or None).
    Terms are enabled by key presence, like the reference."""
    rc = reg_coeffs or {}
    r = QocReg()
    for key in ('amplitude', 'envelope', 'dwdt', 'd2wdt2', 'speed_up'):
        if key in rc:
            setattr(r, 'has_' + key, 1)
            setattr(r, key, float(rc[key]))
    if 'bandpass' in rc:
        raise ValueError('bandpass regulariser is not supported (dead code in the reference: tf.complex_abs)')
    if 'd2wdt2' in rc and 'dwdt' not in rc:
        raise NameError("name 'new_weights' is not defined")     # regularization_functions.py:30 vs :41
    fw = None
    if 'forbidden_coeff_list' in rc:
        r.has_forbidden = 1
        fw = np.zeros(n, dtype=np.float64)
        for coeff, state in zip(rc['forbidden_coeff_list'], rc['states_forbidden_list']):
            fw[int(state)] += float(coeff)
    return r, fw


class GrapeEngine:
    """One problem (H_k, U_target, states, regularisers) x B independent control sets on one GPU."""

    def __init__(self, n, K, T, m, B, exp_terms, scaling, dtype='f64', device=None, flags=0):
        import torch
        self.torch = torch
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise QocError("no CUDA device visible: the GRAPE engine runs on sm_100a only (no CPU fallback)")
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device) \
            if not isinstance(device, torch.device) else device
        self.dims = QocDims(n, K, T, m, B, exp_terms, scaling, _DTYPES[dtype] if isinstance(dtype, str) else dtype, flags)
        self.n, self.K, self.T, self.m, self.B = n, K, T, m, B
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.qoc_create(C.byref(self._h), C.byref(self.dims))
            if rc:
                msg = self.lib.qoc_last_error(self._h).decode() if self._h else "invalid dimensions"
                if self._h:
                    self.lib.qoc_destroy(self._h)
                    self._h = C.c_void_p()
                raise QocError("qoc_create failed (%d): %s" % (rc, msg))
            nbytes = C.c_size_t()
            self._check(self.lib.qoc_workspace_bytes(self._h, C.byref(nbytes)))
            self.workspace_bytes = nbytes.value
            self.batch_chunk = int(self.lib.qoc_batch_chunk(self._h))
            self._ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=self.device)
            ptr = (self._ws.data_ptr() + 255) // 256 * 256
            self._check(self.lib.qoc_set_workspace(self._h, C.c_void_p(ptr), C.c_size_t(nbytes.value)))
        self._pinned = {}

    # ------------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc:
            raise QocError("libqoc_b200 error %d: %s" % (rc, self.lib.qoc_last_error(self._h).decode()))

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _on_device(self):
        """Every C call runs with this engine's device current (streams and events belong to it)."""
        return self.torch.cuda.device(self.device)

    def close(self):
        if getattr(self, '_h', None):
            with self._on_device():
                self.lib.qoc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def set_problem(self, A, U0, phi, V, concerned_idx, maxA, dt):
        """A [K+1,n,n] = -i*dt*[H0, Hops...]; U0 [n,n]; phi, V [m,n]; concerned_idx [m] or None."""
        A = np.ascontiguousarray(A, dtype=np.complex128)
        U0 = np.ascontiguousarray(U0, dtype=np.complex128)
        phi = np.ascontiguousarray(phi, dtype=np.complex128)
        V = np.ascontiguousarray(V, dtype=np.complex128)
        maxA = np.ascontiguousarray(maxA, dtype=np.float64)
        assert A.shape == (self.K + 1, self.n, self.n) and U0.shape == (self.n, self.n)
        assert phi.shape == (self.m, self.n) and V.shape == (self.m, self.n) and maxA.shape == (self.K,)
        idx = None if concerned_idx is None else np.ascontiguousarray(concerned_idx, dtype=np.int32)
        with self.torch.cuda.device(self.device):
            self._check(self.lib.qoc_set_problem(self._h, _np_ptr(A), _np_ptr(U0), _np_ptr(phi), _np_ptr(V),
                                                 _np_ptr(idx), _np_ptr(maxA), float(dt), self._stream()))

    def set_regularizers(self, reg_coeffs, envelope=None):
        r, fw = reg_struct(reg_coeffs, self.n, self.T)
        env = None
        if r.has_envelope:
            env = np.ascontiguousarray(envelope, dtype=np.float64)
            assert env.shape == (self.K, self.T)
        with self.torch.cuda.device(self.device):
            self._check(self.lib.qoc_set_regularizers(self._h, C.byref(r), _np_ptr(env), _np_ptr(fw), self._stream()))

    def set_forbid_basis(self, W):
        """forbid_dressed: forbidden populations are taken on W @ psi (W = v_sorted^dagger); None switches it off."""
        Wc = None if W is None else np.ascontiguousarray(W, dtype=np.complex128)
        with self.torch.cuda.device(self.device):
            self._check(self.lib.qoc_set_forbid_basis(self._h, _np_ptr(Wc), self._stream()))

    # ------------------------------------------------------------------------------------------
    def _base(self, base):
        t = self.torch
        assert base.is_cuda and base.dtype == t.float64 and base.is_contiguous()
        assert tuple(base.shape) == (self.B, self.K, self.T), (tuple(base.shape), (self.B, self.K, self.T))
        return base

    def value_and_grad(self, base, out=None):
        """base: cuda float64 [B,K,T] -> dict(loss, reg_loss, grad, unitary_scale, grad_squared) of
        cuda tensors (asynchronous on the current stream).  ``out`` recycles a previous result dict."""
        t = self.torch
        base = self._base(base)
        if out is None:
            out = dict(loss=t.empty(self.B, dtype=t.float64, device=self.device),
                       reg_loss=t.empty(self.B, dtype=t.float64, device=self.device),
                       grad=t.empty_like(base),
                       unitary_scale=t.empty(self.B, dtype=t.float64, device=self.device),
                       grad_squared=t.empty(self.B, dtype=t.float64, device=self.device))
        p = lambda x: C.c_void_p(x.data_ptr())
        with self._on_device():
            self._check(self.lib.qoc_value_and_grad(self._h, p(base), p(out['loss']), p(out['reg_loss']), p(out['grad']),
                                                    p(out['unitary_scale']), p(out['grad_squared']), self._stream()))
        return out

    def evolve(self, base, want_inter_vecs=True):
        """Forward only -> dict(U_final [B,n,n] c128, inter_vecs [B,T+1,m,n] c128 | None, loss, unitary_scale)."""
        t = self.torch
        base = self._base(base)
        U = t.empty(self.B, self.n, self.n, dtype=t.complex128, device=self.device)
        iv = t.empty(self.B, self.T + 1, self.m, self.n, dtype=t.complex128, device=self.device) if want_inter_vecs else None
        loss = t.empty(self.B, dtype=t.float64, device=self.device)
        us = t.empty(self.B, dtype=t.float64, device=self.device)
        p = lambda x: None if x is None else C.c_void_p(x.data_ptr())
        with self._on_device():
            self._check(self.lib.qoc_evolve(self._h, p(base), p(U), p(iv), p(loss), p(us), self._stream()))
        return dict(U_final=U, inter_vecs=iv, loss=loss, unitary_scale=us)

    # host-buffer entry points (the reference-facing call: run_session.get_error semantics) -----
    def _pin(self, name, shape, dtype=np.float64):
        t = self.torch
        key = (name, tuple(shape), np.dtype(dtype).str)
        if key not in self._pinned:
            td = {np.dtype(np.float64).str: t.float64, np.dtype(np.complex128).str: t.complex128}[np.dtype(dtype).str]
            self._pinned[key] = t.empty(tuple(shape), dtype=td).pin_memory()
        return self._pinned[key]

    def host_buffers(self):
        """Pinned host arrays (NumPy views) a caller may fill / read directly to avoid extra copies:
        dict(base[B,K,T], grad[B,K,T])."""
        return dict(base=self._pin('base', (self.B, self.K, self.T)).numpy(),
                    grad=self._pin('grad', (self.B, self.K, self.T)).numpy())

    def value_and_grad_host(self, base_np, copy=True):
        """NumPy in, NumPy out; H2D + kernels + D2H inside the call (pinned staging buffers).
        ``base_np`` may be the pinned ``host_buffers()['base']`` itself; with ``copy=False`` the returned
        arrays are views of the pinned result buffers (valid until the next call)."""
        hb = self._pin('base', (self.B, self.K, self.T))
        if base_np is not hb.numpy():
            hb.numpy()[...] = np.asarray(base_np, dtype=np.float64).reshape(self.B, self.K, self.T)
        hg = self._pin('grad', (self.B, self.K, self.T))
        ho = self._pin('out', (4, self.B))
        p = lambda x: C.c_void_p(x.data_ptr())
        with self.torch.cuda.device(self.device):
            self._check(self.lib.qoc_value_and_grad_host(
                self._h, p(hb), p(ho[0]), p(ho[1]), p(hg), p(ho[2]), p(ho[3]), self._stream()))
        o = ho.numpy()
        if not copy:
            return dict(loss=o[0], reg_loss=o[1], grad=hg.numpy(), unitary_scale=o[2], grad_squared=o[3])
        return dict(loss=o[0].copy(), reg_loss=o[1].copy(), grad=hg.numpy().copy(), unitary_scale=o[2].copy(),
                    grad_squared=o[3].copy())

    def evolve_host(self, base_np, want_inter_vecs=True):
        hb = self._pin('base', (self.B, self.K, self.T))
        hb.numpy()[...] = np.asarray(base_np, dtype=np.float64).reshape(self.B, self.K, self.T)
        hU = self._pin('U', (self.B, self.n, self.n), np.complex128)
        hiv = self._pin('iv', (self.B, self.T + 1, self.m, self.n), np.complex128) if want_inter_vecs else None
        ho = self._pin('eout', (2, self.B))
        p = lambda x: None if x is None else C.c_void_p(x.data_ptr())
        with self.torch.cuda.device(self.device):
            self._check(self.lib.qoc_evolve_host(self._h, p(hb), p(hU), p(hiv), p(ho[0]), p(ho[1]), self._stream()))
        return dict(U_final=hU.numpy().copy(), inter_vecs=None if hiv is None else hiv.numpy().copy(),
                    loss=ho.numpy()[0].copy(), unitary_scale=ho.numpy()[1].copy())

    def poll_error(self):
        """Synchronise and raise if a device-side pipeline flagged a failure (tcgen05 path)."""
        with self._on_device():
            self._check(self.lib.qoc_poll_error(self._h, self._stream()))

    def propagators(self):
        """Debug view of the cached propagators of the last call as complex128 [B,T,n,n] (clone)."""
        t = self.torch
        if self.batch_chunk < self.B:
            raise QocError("propagators(): the batch is processed in chunks of %d < B = %d, the cache only holds the "
                           "last chunk" % (self.batch_chunk, self.B))
        ptr, eb = C.c_void_p(), C.c_int()
        self._check(self.lib.qoc_debug_propagators(self._h, C.byref(ptr), C.byref(eb)))
        off = ptr.value - self._ws.data_ptr()
        if self.dims.dtype == QOC_F16X2:        # [B][T][4][n][ld] fp16 planes, value = (h0 + h1) / 2^13
            ld = (self.n + 15) // 16 * 16
            nel = self.B * self.T * 4 * self.n * ld
            raw = self._ws[off:off + nel * 2].view(t.float16).reshape(self.B, self.T, 4, self.n, ld)[..., :self.n].double()
            return t.complex(raw[:, :, 0] + raw[:, :, 1], raw[:, :, 2] + raw[:, :, 3]) / 8192.0
        if self.dims.dtype == QOC_F64:
            nel = self.B * self.T * self.n * self.n
            return self._ws[off:off + nel * 16].view(t.complex128).reshape(self.B, self.T, self.n, self.n).clone()
        raw = self._ws[off:off + self.B * self.T * 2048 * 4].view(t.float32).reshape(self.B, self.T, 2, 32, 32)
        return t.complex(raw[:, :, 0, :self.n, :self.n].double(), raw[:, :, 1, :self.n, :self.n].double())

    KERNELS = ("expm", "chain", "fwd_reduce", "costate", "grad", "finalize")

    def set_profiling(self, enable=True):
        with self._on_device():
            self._check(self.lib.qoc_set_profiling(self._h, int(bool(enable))))

    def kernel_times_ms(self):
        """Per-kernel CUDA-event durations (ms) of the last value_and_grad; synchronises."""
        buf = (C.c_float * len(self.KERNELS))()
        with self._on_device():
            self._check(self.lib.qoc_kernel_times_ms(self._h, buf))
        return dict(zip(self.KERNELS, [float(x) for x in buf]))

    @property
    def launch_count(self):
        return int(self.lib.qoc_launch_count(self._h))

    @classmethod
    def from_sys_para(cls, sp, B=None, dtype='f64', device=None):
        """Build an engine from a ``SystemParameters`` (unitary mode)."""
        flags = 0
        if sp.state_transfer:
            # the reverse sweep uses Q_t^dagger; the reference uses sum_j (-H)^j/j! (tensorflow_state.py:118-131),
            # identical for Hermitian Hamiltonians only
            for Hm in [sp.H0_c] + list(sp.ops_c):
                if not np.allclose(Hm, np.conj(np.transpose(Hm)), rtol=0, atol=1e-12 * max(1.0, np.abs(Hm).max())):
                    raise NotImplementedError("state_transfer=True needs Hermitian H0 / Hops in the CUDA engine")
            flags |= QOC_FLAG_STATE_TRANSFER
        B = sp.batch_size if B is None else B
        eng = cls(sp.state_num, sp.ops_len, sp.steps, len(sp.states_concerned_list), B, sp.exp_terms, sp.scaling,
                  dtype=dtype, device=device, flags=flags)
        eng.set_problem(sp.A_c, sp.U0_c, sp.target_vectors_c, sp.V_c, sp.concerned_idx, sp.ops_max_amp, sp.dt)
        eng.set_regularizers(sp.reg_coeffs, sp.one_minus_gauss)
        if sp.is_dressed and sp.reg_coeffs.get('forbid_dressed') and 'forbidden_coeff_list' in sp.reg_coeffs:
            from ..helper_functions.grape_functions import sort_ev      # regularization_functions.py:73-80
            eng.set_forbid_basis(np.conj(np.transpose(sort_ev(sp.v_c, sp.dressed_id))))
        return eng
