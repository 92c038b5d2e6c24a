"""Optimiser driver: the role of the reference's ``run_session`` (core/run_session.py:11-199), with
the TF session replaced by calls through the C ABI and the Adam update done in PyTorch on the
device (TF-1 ``AdamOptimizer`` arithmetic, NOT ``torch.optim.Adam``: the epsilon placement differs).

One engine evaluation per iteration (the reference runs the whole graph twice, :53-54 and :69).
For a batch of B instances every instance follows the reference's stop rules independently: once
an instance meets a stop criterion its weights are frozen on the device, so what is returned for it
is exactly the iterate whose loss / gradient was tested (:56-64).
"""
import time

import numpy as np


class TF1AdamState:
    """m, v and step counter of tf.train.AdamOptimizer(beta1=.9, beta2=.999, epsilon=1e-8)
    (core/tensorflow_state.py:345 passes only the learning rate)."""

    def __init__(self, like, beta1=0.9, beta2=0.999, eps=1e-8):
        import torch
        self.torch = torch
        self.m = torch.zeros_like(like)
        self.v = torch.zeros_like(like)
        self.t = 0
        self.b1, self.b2, self.eps = beta1, beta2, eps

    def step(self, theta, grad, lr, frozen=None):
        """theta <- theta - lr_t * m / (sqrt(v) + eps), lr_t = lr*sqrt(1-b2^t)/(1-b1^t); instances with
        frozen[b] == True keep their weights."""
        torch = self.torch
        self.t += 1
        lr_t = lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        self.m.mul_(self.b1).add_(grad, alpha=1.0 - self.b1)
        self.v.mul_(self.b2).addcmul_(grad, grad, value=1.0 - self.b2)
        upd = self.m / (self.v.sqrt() + self.eps)
        if frozen is not None:
            upd = torch.where(frozen.view(-1, 1, 1), torch.zeros_like(upd), upd)
        theta.add_(upd, alpha=-lr_t)
        return theta


class TF1AdamHost:
    """Host twin of :class:`TF1AdamState` for callers that keep the weights on the host and go through the
    host-buffer C-ABI entry point (``GrapeEngine.value_and_grad_host``).  ``theta`` / ``grad`` are NumPy arrays
    (typically the pinned ``host_buffers()``); the update is ONE fused multi-threaded pass in the library
    (``qoc_adam_host``) -- five torch passes spent more time waking their thread team after the GPU wait than on
    the arithmetic."""

    def __init__(self, shape, beta1=0.9, beta2=0.999, eps=1e-8, threads=None):
        import os
        from .engine import load_library
        self.lib = load_library()
        self.m = np.zeros(tuple(shape), dtype=np.float64)
        self.v = np.zeros(tuple(shape), dtype=np.float64)
        self.t = 0
        self.b1, self.b2, self.eps = beta1, beta2, eps
        # 4 threads: the gradient was just written by DMA and is cold in every cache; more threads only add wake-up
        # and cross-socket traffic (tools/e2e_breakdown.py: 1/2/4/8/16 threads -> 8.83/8.15/7.99/8.10/8.31 ms per C2 step)
        self.threads = int(threads) if threads else max(1, min(4, (os.cpu_count() or 1)))

    def step(self, theta, grad, lr):
        import ctypes as C
        self.t += 1
        lr_t = lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        g = np.ascontiguousarray(grad, dtype=np.float64)
        assert theta.dtype == np.float64 and theta.flags.c_contiguous and theta.shape == self.m.shape == g.shape
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = self.lib.qoc_adam_host(p(theta), p(g), p(self.m), p(self.v), C.c_size_t(theta.size), float(lr_t),
                                    float(self.b1), float(self.b2), float(self.eps), int(self.threads))
        if rc:
            raise RuntimeError("qoc_adam_host failed (%d)" % rc)
        return theta


class run_session:
    """Same constructor shape and result attributes (``uks``, ``Uf``) as the reference class."""

    def __init__(self, engine, conv, sys_para, method, show_plots=True, single_simulation=False, use_gpu=True,
                 quiet=False, run_file=None, return_dtype=None):
        import torch
        self.torch = torch
        self.return_dtype = return_dtype
        self.engine = engine
        self.conv = conv
        self.sys_para = sys_para
        self.update_step = conv.update_step
        self.iterations = 0
        self.method = method.upper()
        self.show_plots = show_plots
        self.quiet = quiet
        self.run_file = run_file
        self.elapsed = 0.0
        self.B = engine.B
        sp = sys_para
        base0 = np.asarray(sp.ops_weight_base, dtype=np.float64).reshape(self.B, sp.ops_len, sp.steps)
        self.base = torch.from_numpy(np.ascontiguousarray(base0)).to(engine.device)
        self.history = []
        self.start_time = time.time()
        if self.method == 'EVOLVE':                                    # :33-37
            self.out = self.engine.value_and_grad(self.base)
            self._fetch_scalars()
            self.get_end_results()
        elif self.method == 'ADAM':
            self.start_adam_optimizer()
        else:
            self.bfgs_optimize(method=self.method)

    # ------------------------------------------------------------------------------------------
    def _fetch_scalars(self):
        o = self.out
        s = self.torch.stack([o['loss'], o['reg_loss'], o['grad_squared'], o['unitary_scale']]).cpu().numpy()
        self.l, self.rl, self.g_squared, self.metric = s[0], s[1], s[2], s[3]

    def start_adam_optimizer(self):
        """core/run_session.py:47-69."""
        torch = self.torch
        conv = self.conv
        adam = TF1AdamState(self.base)
        done = torch.zeros(self.B, dtype=torch.bool, device=self.base.device)
        self.end = False
        out = None
        while True:
            out = self.engine.value_and_grad(self.base, out=out)       # :53-54 (one evaluation, not two)
            self.out = out
            stop = (out['loss'] < conv.conv_target) | (out['grad_squared'] < conv.min_grad)
            done |= stop
            self._fetch_scalars()
            self.history.append((self.l.copy(), self.rl.copy(), self.g_squared.copy(), self.metric.copy()))
            if bool(done.all().item()) or self.iterations >= conv.max_iterations:     # :56-58
                self.end = True
            self.update_and_save()
            if self.end:
                self.get_end_results()
                break
            lr = float(conv.rate) * np.exp(-float(self.iterations) / conv.learning_rate_decay)   # :66
            adam.step(self.base, out['grad'], lr, frozen=done if self.B > 1 else None)        # :67-69

    def update_and_save(self):
        """core/run_session.py:75-92: every update_step a summary save + console line, every evol_save_step the
        final state and the state trajectories (plots are out of scope; show_plots only silences the console)."""
        if not self.end:
            if self.iterations % self.conv.update_step == 0:
                self.save_data()
                self.display()
            if self.iterations % self.conv.evol_save_step == 0:
                if not (self.sys_para.show_plots and self.iterations % self.conv.update_step == 0):
                    if self.iterations % self.conv.update_step != 0:
                        self.save_data()
                    self.save_evol()
            self.iterations += 1

    def _uks_now(self):
        w = self.torch.sin(self.base).cpu().numpy()
        uks = np.asarray(self.sys_para.ops_max_amp)[None, :, None] * w
        return uks if self.sys_para.batched else uks[0]

    def _pick(self, x):
        x = np.asarray(x)
        return x if self.sys_para.batched else x[0]

    def save_data(self):
        """core/run_session.py:129-138."""
        if self.run_file is None:
            return
        self.elapsed = time.time() - self.start_time
        rf = self.run_file
        rf.append('error', self._pick(self.l))
        rf.append('reg_error', self._pick(self.rl))
        rf.append('uks', self._uks_now())
        rf.append('iteration', np.array(self.iterations))
        rf.append('run_time', np.array(self.elapsed))
        rf.append('unitary_scale', self._pick(self.metric))

    def save_evol(self, ev=None):
        """Convergence.save_evol -> Analysis.get_final_state / get_inter_vecs (core/analysis.py:26-35,44-101)."""
        if self.run_file is None:
            return
        from ..helper_functions.grape_functions import c_to_r_mat, sort_ev
        sp, rf = self.sys_para, self.run_file
        if ev is None:
            ev = self.engine.evolve(self.base, want_inter_vecs=sp.use_inter_vecs)
        if not sp.state_transfer:
            U = ev['U_final'].cpu().numpy()
            rf.append('final_state', self._pick(np.stack([c_to_r_mat(u) for u in U])))
        if sp.use_inter_vecs and ev['inter_vecs'] is not None:
            iv = np.transpose(ev['inter_vecs'].cpu().numpy(), (0, 2, 3, 1))        # [B, m, n, T+1]
            rf.append('inter_vecs_raw_real', self._pick(iv.real))
            rf.append('inter_vecs_raw_imag', self._pick(iv.imag))
            if sp.is_dressed:                                                         # analysis.py:76-84 (plain transpose)
                vt = np.transpose(sort_ev(sp.v_c, sp.dressed_id))
                iv = np.einsum('ac,bjct->bjat', vt, iv)
            rf.append('inter_vecs_mag_squared', self._pick(np.abs(iv) ** 2))
            rf.append('inter_vecs_real', self._pick(iv.real))
            rf.append('inter_vecs_imag', self._pick(iv.imag))

    def display(self):
        if self.quiet:
            return
        b = int(np.argmin(self.l))
        if self.run_file is None:
            self.elapsed = time.time() - self.start_time        # the reference only sets it in save_data (run_session.py:131)
        print('Error = :%1.2e; Runtime: %.1fs; Iterations = %d, grads =  %10.3e, unitary_metric = %.5f' % (
            self.l[b], self.elapsed, self.iterations, self.g_squared[b], self.metric[b]))

    def get_end_results(self):
        """core/run_session.py:94-117 + core/analysis.py:18-41."""
        sp = self.sys_para
        self.save_data()
        self.display()
        ev = self.engine.evolve(self.base, want_inter_vecs=sp.use_inter_vecs)
        if not self.show_plots:
            self.save_evol(ev)                                          # run_session.py:103-104
        w = self.torch.sin(self.base).cpu().numpy()                     # Analysis.get_ops_weight
        uks = np.asarray(sp.ops_max_amp)[None, :, None] * w             # Get_uks
        Uf = ev['U_final'].cpu().numpy()
        self.inter_vecs = None if ev['inter_vecs'] is None else ev['inter_vecs'].cpu().numpy()
        self.engine.poll_error()
        if getattr(self, 'return_dtype', None) == 'reference':
            # the reference returns float32 pulses (run_session.py:112-117, float32 session) and a complex64 final
            # state (analysis.py:18-24 on a float32 fetch)
            uks, Uf = uks.astype(np.float32), Uf.astype(np.complex64)
        if sp.batched:
            self.uks, self.Uf = uks, Uf
        else:
            self.uks, self.Uf = uks[0], Uf[0]
        if self.run_file is not None and not sp.state_transfer:        # get_final_state() appends once more (:107)
            from ..helper_functions.grape_functions import c_to_r_mat
            self.run_file.append('final_state', self._pick(np.stack([c_to_r_mat(u) for u in Uf])))
        if sp.state_transfer:
            self.Uf = []

    # -- SciPy path (core/run_session.py:119-127,151-196) -----------------------------------------
    def get_error(self, uks):
        o = self.engine.value_and_grad_host(np.asarray(uks, dtype=np.float64))
        K, T = self.sys_para.ops_len, self.sys_para.steps
        return o['loss'][0], o['reg_loss'][0], o['grad'].reshape(K * T), o['unitary_scale'][0], o['grad_squared'][0]

    def minimize_opt_fun(self, x):
        K = self.sys_para.ops_len
        # the reference assigns ops_weight_base inside get_error (run_session.py:121): the per-iteration records
        # (uks, final_state, inter_vecs) written by update_and_save must describe the CURRENT iterate
        self.base.copy_(self.torch.from_numpy(np.ascontiguousarray(np.reshape(x, (1, K, len(x) // K)), dtype=np.float64)))
        l, rl, grads, metric, g2 = self.get_error(np.reshape(x, (K, len(x) // K)))
        self.l, self.rl, self.g_squared, self.metric = (np.array([v]) for v in (l, rl, g2, metric))
        if l < self.conv.conv_target:
            self.conv_time = time.time() - self.start_time
            self.conv_iter = self.iterations
            self.end = True
            if not self.quiet:
                print('Target fidelity reached')
            grads = 0 * grads                                           # zero gradient stops SciPy (:155-160)
        self.update_and_save()
        return np.float64(rl), np.float64(grads)

    def bfgs_optimize(self, method='L-BFGS-B', jac=True, options=None):
        from scipy.optimize import minimize
        if self.B != 1:
            raise ValueError('SciPy methods optimise one instance at a time (B = 1)')
        self.conv.reset_convergence()
        self.conv_time, self.conv_iter, self.end = 0., 0, False
        if not self.quiet:
            print("Starting " + self.method + " Optimization")
        self.start_time = time.time()
        x0 = np.asarray(self.sys_para.ops_weight_base, dtype=np.float64).reshape(-1)
        options = {'maxfun': self.conv.max_iterations, 'gtol': self.conv.min_grad, 'disp': False, 'maxls': 40}
        if method.upper() not in ('L-BFGS-B',):
            options.pop('maxfun'); options.pop('maxls')
            options['maxiter'] = self.conv.max_iterations
        res = minimize(self.minimize_opt_fun, x0, method=method, jac=jac, options=options)
        K = self.sys_para.ops_len
        x = np.reshape(res['x'], (1, K, len(res['x']) // K))
        self.base = self.torch.from_numpy(np.ascontiguousarray(x)).to(self.engine.device)
        self.scipy_result = res
        if not self.quiet:
            print(self.method + ' optimization done')
            print(res.message)
        self.out = self.engine.value_and_grad(self.base)
        self._fetch_scalars()
        self.end = True
        self.get_end_results()
