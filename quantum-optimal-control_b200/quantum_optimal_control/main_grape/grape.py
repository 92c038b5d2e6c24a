"""``Grape(...)``: the reference's public entry point (main_grape/grape.py:19), kept verbatim in
signature and return value ``(uks, U_final)``, running on the B200 engine.

Additive keywords (ours): ``batch`` (number of independent random initialisations when no
``initial_guess`` is given), ``dtype`` ('f64' | 'f16x2' | 'tf32x3': arithmetic of the propagator stage),
``return_dtype`` (None: float64 / complex128 results; 'reference': float32 ``uks`` and complex64 ``U_final`` as the
reference returns them, core/run_session.py:112-117, core/analysis.py:18-24), ``device``, ``quiet``; ``initial_guess`` may be
[B, K, T], in which case ``uks`` is [B, K, T] and ``U_final`` is [B, n, n]; ``return_losses=True`` appends the
per-instance final losses (used by ``core.population`` for multi-GPU sweeps).
"""
import os
import time

import numpy as np

from ..core.problem import SystemParameters
from ..core.defaults import Convergence
from ..core.engine import GrapeEngine
from ..core.optimizer import run_session
from ..core import storage


def Grape(H0, Hops, Hnames, U, total_time, steps, states_concerned_list, convergence=None, U0=None,
          reg_coeffs=None, dressed_info=None, maxA=None, use_gpu=True, sparse_H=True, sparse_U=False,
          sparse_K=False, draw=None, initial_guess=None, show_plots=True, unitary_error=1e-4, method='Adam',
          state_transfer=False, no_scaling=False, freq_unit='GHz', file_name=None, save=True, data_path=None,
          Taylor_terms=None, use_inter_vecs=True, batch=None, dtype='f64', device=None, quiet=False,
          return_losses=False, return_dtype=None):
    grape_start_time = time.time()
    time_unit = {"GHz": "ns", "MHz": "us", "KHz": "ms", "Hz": "s"}[freq_unit]       # grape.py:25-26

    if return_dtype not in (None, 'reference'):
        raise ValueError("return_dtype must be None or 'reference'")
    if not use_gpu:
        raise NotImplementedError("use_gpu=False: this build has no CPU path; the GRAPE hot path runs on "
                                  "sm_100a (B200) only")
    sparse_H = sparse_U = sparse_K = False                                          # grape.py:29-32

    if U0 is None:                                                                  # grape.py:89-90
        U0 = np.identity(len(H0))
    if convergence is None:                                                         # grape.py:91-92
        convergence = {'rate': 0.01, 'update_step': 100, 'max_iterations': 5000, 'conv_target': 1e-8,
                       'learning_rate_decay': 2500}
    file_path, run_file = None, None
    if save:                                                                        # grape.py:36-87
        if file_name is None:
            raise ValueError('Grape function input: file_name, is not specified.')
        if data_path is None:
            raise ValueError('Grape function input: data_path, is not specified.')
        file_path = storage.new_run_file(data_path, file_name)
        run_file = storage.RunFile(file_path)
        if not quiet:
            print("data saved at: " + str(file_path))
        storage.save_inputs(run_file, dict(H0=H0, Hops=Hops, Hnames=Hnames, U=U, total_time=total_time, steps=steps,
                                            states_concerned_list=states_concerned_list, use_gpu=use_gpu,
                                            sparse_H=sparse_H, sparse_U=sparse_U, sparse_K=sparse_K, maxA=maxA,
                                            initial_guess=initial_guess, method=method),
                            convergence, reg_coeffs, dressed_info)

    if maxA is None:                                                                # grape.py:95-101
        if initial_guess is None:
            maxAmp = 4 * np.ones(len(Hops))
        else:
            maxAmp = 1.5 * np.max(np.abs(initial_guess)) * np.ones(len(Hops))
    else:
        maxAmp = maxA

    sys_para = SystemParameters(H0, Hops, Hnames, U, U0, total_time, steps, states_concerned_list, dressed_info,
                                maxAmp, draw, initial_guess, show_plots, unitary_error, state_transfer, no_scaling,
                                reg_coeffs, save, file_path, Taylor_terms, use_gpu, use_inter_vecs, sparse_H,
                                sparse_U, sparse_K, batch=batch)
    if save:                                                                        # system_parameters.py:189-191,233-236
        run_file.add('initial_vectors_c', np.array(sys_para.initial_vectors_c))
        run_file.add('taylor_terms', sys_para.exp_terms)
        run_file.add('taylor_scaling', sys_para.scaling)
    engine = GrapeEngine.from_sys_para(sys_para, dtype=dtype, device=device)
    conv = Convergence(sys_para, time_unit, convergence)
    try:
        SS = run_session(engine, conv, sys_para, method, show_plots=sys_para.show_plots, use_gpu=use_gpu, quiet=quiet,
                         run_file=run_file, return_dtype=return_dtype)
        if save:
            run_file.add('wall_clock_time', np.array(time.time() - grape_start_time))     # grape.py:123-127
            if not quiet:
                print("data saved at: " + str(file_path))
        if return_losses:
            return SS.uks, SS.Uf, np.asarray(SS.l)
        return SS.uks, SS.Uf
    except KeyboardInterrupt:                                                       # grape.py:130-139
        if save:
            run_file.add('wall_clock_time', np.array(time.time() - grape_start_time))
        return None
    finally:
        engine.close()
