"""Exploration: the fp32-class tcgen05 path (dtype 'f16x2') against the fp64 path on the same inputs (GPU only).
python tools/f16x2_check.py [case ...]  -> one JSON line per case with the deviations and the per-stage times."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "quantum-optimal-control_b200"))
import workloads as W  # noqa: E402
from quantum_optimal_control.core.problem import SystemParameters  # noqa: E402
from quantum_optimal_control.core.engine import GrapeEngine  # noqa: E402

CASES = {
    'c1': (lambda: W.c1_pi_pulse(), {}, 2),
    'c2_T40': (lambda: W.c2_transmon_cavity(T=40), dict(total_time=80.0), 2),
    'c3_T30': (lambda: W.c3_two_transmon_cnot(T=30), dict(total_time=0.3), 2),
    'c3_T1000': (lambda: W.c3_two_transmon_cnot(T=1000), {}, 4),
    'c5_n64_m4': (lambda: W.c5_random(64, T=50), dict(states_concerned_list=[0, 3, 17, 63]), 2),
    'c5_n100_regs': (lambda: W.c5_random(100, T=30), dict(states_concerned_list=[0, 5, 99], reg_coeffs={'dwdt': 0.1, 'forbidden_coeff_list': [2.0], 'states_forbidden_list': [7]}), 2),
    'c5_n128_m8': (lambda: W.c5_random(128, T=40), dict(states_concerned_list=list(range(8))), 2),
    'c4_T50': (lambda: W.c4_three_transmon_toffoli(T=50), dict(total_time=2.5), 2),
    'c4_T400': (lambda: W.c4_three_transmon_toffoli(T=400), dict(total_time=20.0), 2),
}


def engine(pb, guess, dtype):
    args, kw = W.grape_kwargs(pb)
    H0, Hops, Hn, U, tt, steps, scl = args
    sp = SystemParameters(H0, Hops, Hn, U, kw.get('U0', np.identity(len(H0))), tt, steps, scl, kw.get('dressed_info'),
                          kw['maxA'], None, guess, False, kw.get('unitary_error', 1e-4), False, False,
                          kw.get('reg_coeffs'), False, None, kw.get('Taylor_terms'), True, True, False, False, False)
    return sp, GrapeEngine.from_sys_para(sp, dtype=dtype)


def run(name):
    fn, over, B = CASES[name]
    pb = dict(fn()); pb.update(over)
    K, T = len(pb['Hops']), pb['steps']
    guess = W.random_guess(K, T, pb['maxA'], 11, B=B)
    res = {}
    for dt in ('f64', 'f16x2'):
        sp, eng = engine(pb, guess, dt)
        base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
        eng.set_profiling(True)
        out = eng.value_and_grad(base)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = eng.value_and_grad(base)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        times = eng.kernel_times_ms()
        eng.poll_error()
        ev = eng.evolve(base)
        torch.cuda.synchronize()
        res[dt] = dict(out={k: v.cpu().numpy() for k, v in out.items()}, U=ev['U_final'].cpu().numpy(),
                       iv=ev['inter_vecs'].cpu().numpy(), P=eng.propagators().cpu().numpy() if B * T * sp.state_num ** 2 < 3e7 else None,
                       ms=ms, times=times, ps=(sp.exp_terms, sp.scaling))
        eng.close()
        del eng
        torch.cuda.empty_cache()
    a, b = res['f64'], res['f16x2']
    g0, g1 = a['out']['grad'], b['out']['grad']
    line = dict(case=name, n=int(a['U'].shape[-1]), T=T, B=B, ps=a['ps'],
                loss=[float(x) for x in a['out']['loss'][:2]],
                d_loss=float(np.abs(a['out']['loss'] - b['out']['loss']).max()),
                d_reg_loss=float(np.abs(a['out']['reg_loss'] - b['out']['reg_loss']).max()),
                grad_rel=float(np.abs(g0 - g1).max() / max(np.abs(g0).max(), 1e-300)),
                dU_fro=float(max(np.linalg.norm(a['U'][i] - b['U'][i]) for i in range(B))),
                d_uscale=float(np.abs(a['out']['unitary_scale'] - b['out']['unitary_scale']).max()),
                d_psi=float(np.abs(a['iv'] - b['iv']).max()),
                dP_max=None if a['P'] is None else float(np.abs(a['P'] - b['P']).max()),
                ms_f64=round(a['ms'], 3), ms_f16x2=round(b['ms'], 3),
                times_f16x2={k: round(v, 3) for k, v in b['times'].items()})
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(CASES)):
        try:
            run(nm)
        except Exception as e:  # keep going: one JSON line per case
            print(json.dumps(dict(case=nm, error=repr(e)[:400])), flush=True)
