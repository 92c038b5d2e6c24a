"""Shared helpers for the parity tests: build the same problem for the oracle and for the engine."""
import numpy as np

import workloads as W
from oracle import grape_oracle as O


def make_case(problem, seed=0, B=1, **overrides):
    """-> (setup for the oracle of instance 0.., list of oracle setups per instance, guesses [B,K,T], args, kwargs)"""
    pb = dict(problem)
    pb.update(overrides)
    K, T = len(pb['Hops']), pb['steps']
    guess = W.random_guess(K, T, pb['maxA'], seed, B=B)
    args, kw = W.grape_kwargs(pb)
    H0, Hops, Hn, U, tt, steps, scl = args
    setups = [O.make_setup(H0, Hops, U, tt, steps, scl, initial_guess=guess[b], **kw) for b in range(B)]
    return setups, guess, args, kw


def engine_for(problem_args, kw, guess):
    """SystemParameters + GrapeEngine for a batched guess [B,K,T]."""
    from quantum_optimal_control.core.problem import SystemParameters
    from quantum_optimal_control.core.engine import GrapeEngine
    H0, Hops, Hn, U, tt, steps, scl = problem_args
    kw = dict(kw)
    sp = SystemParameters(H0, Hops, Hn, U, kw.get('U0', np.identity(len(H0))), tt, steps, scl, kw.get('dressed_info'),
                          kw['maxA'], None, guess, False, kw.get('unitary_error', 1e-4), kw.get('state_transfer', False), False,
                          kw.get('reg_coeffs'), False, None, kw.get('Taylor_terms'), True, True, False, False, False)
    return sp, GrapeEngine.from_sys_para(sp)
