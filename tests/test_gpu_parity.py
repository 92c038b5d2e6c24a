"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Tolerances (fp64 path): north_star asks ||dU_final||_F < 1e-5; the fp64 kernels are expected ~1e-12,
so the tests assert 1e-9 absolute on U_final / psi and 1e-9 relative-to-scale on gradients.
"""
import numpy as np
import pytest
import torch

import workloads as W
from oracle import grape_oracle as O
from helpers import make_case, engine_for

pytestmark = pytest.mark.gpu

ATOL_U = 1e-9
RTOL_G = 1e-9

ALL_REGS = {'amplitude': 0.3, 'envelope': 0.7, 'dwdt': 0.02, 'd2wdt2': 0.0005, 'speed_up': 0.4,
            'forbidden_coeff_list': [3.0, 5.0, 2.0], 'states_forbidden_list': [2, 3, 2]}

CASES = {
    'c1': (lambda: W.c1_pi_pulse(), {}, 1),
    'c1_regs_B3': (lambda: W.c1_pi_pulse(T=37), dict(reg_coeffs={'amplitude': 0.3, 'envelope': 0.7, 'dwdt': 0.02,
                                                                'd2wdt2': 0.0005, 'speed_up': 0.4,
                                                                'forbidden_coeff_list': [3.0], 'states_forbidden_list': [1]}), 3),
    'c2_T40': (lambda: W.c2_transmon_cavity(T=40), dict(total_time=80.0), 2),
    'c2_regs': (lambda: W.c2_transmon_cavity(T=25), dict(total_time=50.0, reg_coeffs=ALL_REGS), 2),
    'c3_T30': (lambda: W.c3_two_transmon_cnot(T=30), dict(total_time=0.3), 2),
    'c5_n8': (lambda: W.c5_random(8, T=20), {}, 2),
    'c5_n16': (lambda: W.c5_random(16, T=20), {}, 2),
    'c5_n20': (lambda: W.c5_random(20, T=12), {}, 2),
    'c5_n48': (lambda: W.c5_random(48, T=6), {}, 1),
    'c5_n50': (lambda: W.c5_random(50, T=5), {}, 1),
    'c5_n64': (lambda: W.c5_random(64, T=5), {}, 1),
    'c5_n72': (lambda: W.c5_random(72, T=3), {}, 1),
    # dense m and dense controls above n = 64: GEMM-form gradient (k_grad_large) and costate chain (k_costate_large_mma) with sources
    'c5_n80_dense_regs': (lambda: W.c5_random(80, T=6), dict(reg_coeffs={'dwdt': 0.1, 'speed_up': 0.3, 'forbidden_coeff_list': [2.0, 1.0], 'states_forbidden_list': [7, 50]}), 2),
    'c5_n100_regs': (lambda: W.c5_random(100, T=3), dict(states_concerned_list=[0, 5, 99], reg_coeffs={'dwdt': 0.1, 'forbidden_coeff_list': [2.0], 'states_forbidden_list': [7]}), 2),
    'c5_n128': (lambda: W.c5_random(128, T=2), dict(states_concerned_list=list(range(8))), 1),
    'c4_T4': (lambda: W.c4_three_transmon_toffoli(T=4), dict(total_time=0.2), 1),
    'n5_U0': (lambda: W.c5_random(5, T=15), dict(U0=np.linalg.qr(np.random.default_rng(5).normal(size=(5, 5)) +
                                                                  1j * np.random.default_rng(6).normal(size=(5, 5)))[0],
                                                  states_concerned_list=[1, 3]), 2),
}


@pytest.mark.parametrize("name", list(CASES))
def test_value_and_grad_matches_oracle(name, built_lib):
    fn, over, B = CASES[name]
    setups, guess, args, kw = make_case(fn(), seed=11, B=B, **over)
    sp, eng = engine_for(args, kw, guess)
    assert (sp.exp_terms, sp.scaling) == (setups[0].exp_terms, setups[0].scaling)
    base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
    out = eng.value_and_grad(base)
    ev = eng.evolve(base)
    torch.cuda.synchronize()
    for b in range(B):
        ref = O.graph_value_and_grad(setups[b], setups[b].ops_weight_base)
        n = setups[b].n
        Uref = O.r_to_c_mat(ref.final_state, n)
        assert np.linalg.norm(ev['U_final'][b].cpu().numpy() - Uref) < ATOL_U
        iv_ref = np.transpose(ref.inter_vecs[:, :n, :] + 1j * ref.inter_vecs[:, n:, :], (2, 0, 1))   # [T+1,m,n]
        assert np.abs(ev['inter_vecs'][b].cpu().numpy() - iv_ref).max() < ATOL_U
        assert abs(out['loss'][b].item() - ref.loss) < 1e-10
        assert abs(out['reg_loss'][b].item() - ref.reg_loss) < 1e-10 * max(1.0, abs(ref.reg_loss))
        assert abs(out['unitary_scale'][b].item() - ref.unitary_scale) < 1e-10
        g = out['grad'][b].cpu().numpy()
        scale = max(np.abs(ref.grad).max(), 1e-300)
        assert np.abs(g - ref.grad).max() < RTOL_G * scale
        assert abs(out['grad_squared'][b].item() - ref.grad_squared) < 1e-9 * max(ref.grad_squared, 1e-300)
    eng.close()


def test_host_entry_points_match_device(built_lib):
    setups, guess, args, kw = make_case(W.c2_transmon_cavity(T=20), seed=3, B=2, total_time=40.0)
    sp, eng = engine_for(args, kw, guess)
    base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
    d = eng.value_and_grad(base)
    h = eng.value_and_grad_host(sp.ops_weight_base)
    for k in ('loss', 'reg_loss', 'grad', 'unitary_scale', 'grad_squared'):
        assert np.array_equal(d[k].cpu().numpy(), h[k]), k
    e = eng.evolve(base)
    eh = eng.evolve_host(sp.ops_weight_base)
    assert np.array_equal(e['U_final'].cpu().numpy(), eh['U_final'])
    assert np.array_equal(e['inter_vecs'].cpu().numpy(), eh['inter_vecs'])
    eng.close()


@pytest.mark.parametrize("name,iters", [('c1', 25), ('c2', 12)])
def test_grape_adam_trajectory_matches_oracle(name, iters, built_lib):
    """Grape(...) drop-in: N Adam iterations -> (uks, U_final) vs the oracle's restated run_session loop."""
    from quantum_optimal_control.main_grape.grape import Grape
    pb = W.c1_pi_pulse() if name == 'c1' else dict(W.c2_transmon_cavity(T=30), total_time=60.0)
    if name == 'c2':
        pb['reg_coeffs'] = {'dwdt': 0.1, 'd2wdt2': 0.01, 'forbidden_coeff_list': [1.0, 1.0], 'states_forbidden_list': [20, 21]}
    K, T = len(pb['Hops']), pb['steps']
    guess = W.random_guess(K, T, pb['maxA'], 5)
    args, kw = W.grape_kwargs(pb)
    conv = {'rate': 0.01, 'update_step': 10, 'max_iterations': iters, 'conv_target': 1e-12, 'learning_rate_decay': 100}
    uks, Uf = Grape(*args, convergence=conv, initial_guess=guess, save=False, show_plots=False, quiet=True, **kw)
    ruks, rUf = O.grape(*args, convergence=conv, initial_guess=guess, **kw)
    assert uks.shape == ruks.shape and Uf.shape == rUf.shape
    assert np.abs(uks - ruks).max() < 1e-8
    assert np.linalg.norm(Uf - rUf) < 1e-7


def test_grape_batched_equals_single(built_lib):
    from quantum_optimal_control.main_grape.grape import Grape
    pb = W.c1_pi_pulse(T=40)
    args, kw = W.grape_kwargs(pb)
    g = W.random_guess(2, 40, pb['maxA'], 21, B=3)
    conv = {'rate': 0.02, 'update_step': 50, 'max_iterations': 15, 'conv_target': 1e-12, 'learning_rate_decay': 100}
    ub, Ub = Grape(*args, convergence=conv, initial_guess=g, save=False, show_plots=False, quiet=True, **kw)
    assert ub.shape == (3, 2, 40) and Ub.shape == (3, 2, 2)
    for b in range(3):
        u1, U1 = Grape(*args, convergence=conv, initial_guess=g[b], save=False, show_plots=False, quiet=True, **kw)
        assert np.abs(u1 - ub[b]).max() < 1e-12 and np.abs(U1 - Ub[b]).max() < 1e-12


def test_full_size_c2_properties(built_lib):
    """BASELINE config C2 at full size (n=30, T=500, B=256): size-independent properties --
    unitarity of U_final, loss consistent with U_final, linearity of the costate in the loss source
    checked through gradient == finite difference of the FIRST-ORDER model on one weight (cheap),
    and batch-permutation equivariance."""
    pb = W.c2_transmon_cavity()
    B = 256
    setups, guess, args, kw = make_case(pb, seed=100, B=1)
    guess = W.random_guess(4, 500, pb['maxA'], 100, B=B)
    sp, eng = engine_for(args, kw, guess)
    base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
    out = {k: v.clone() for k, v in eng.value_and_grad(base).items()}
    ev = eng.evolve(base, want_inter_vecs=False)
    U = ev['U_final']
    eye = torch.eye(30, dtype=torch.complex128, device='cuda')
    dev = (U.conj().transpose(1, 2) @ U - eye).abs().amax()
    assert dev.item() < 5e-3          # Taylor/squaring truncation at unitary_error=1e-4 (p=7, s=3)
    phi = torch.from_numpy(sp.target_vectors_c).cuda()          # [m,n]
    idx = torch.as_tensor(sp.concerned_idx, device='cuda', dtype=torch.long)
    o = (phi.conj()[None] * U[:, :, idx].transpose(1, 2)).sum(dim=(1, 2))
    loss = 1 - (o.abs() ** 2) / 4
    assert (loss - out['loss']).abs().max().item() < 1e-10
    perm = torch.randperm(B, device='cuda')
    outp = eng.value_and_grad(base[perm].contiguous())
    assert torch.equal(outp['grad'], out['grad'][perm])
    assert torch.equal(outp['loss'], out['loss'][perm])
    # instance 0 against the oracle at full size
    ref = O.graph_value_and_grad(setups[0], setups[0].ops_weight_base)
    assert np.linalg.norm(U[0].cpu().numpy() - O.r_to_c_mat(ref.final_state, 30)) < 1e-8
    assert np.abs(out['grad'][0].cpu().numpy() - ref.grad).max() < 1e-9 * np.abs(ref.grad).max()
    eng.close()


# ---- fixtures produced by the reference's own source (oracle/run_reference.py) ---------------------
import os as _os
from oracle.run_reference import golden_cases as _golden_cases

_GOLD = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", list(_golden_cases()))
def test_cuda_matches_reference_goldens(name, built_lib):
    """CUDA path vs what the reference's own graph code produced (fp64 arithmetic): one evaluation
    (loss, reg_loss, grad, unitary_scale, U_final, inter_vecs) and a full Grape() Adam run."""
    from quantum_optimal_control.main_grape.grape import Grape
    g = np.load(_os.path.join(_GOLD, "ref_%s_float64.npz" % name))
    pb, seed, conv = _golden_cases()[name]
    args, kw = W.grape_kwargs(pb)
    sp, eng = engine_for(args, kw, g['guess'][None])
    assert (sp.exp_terms, sp.scaling) == (int(g['exp_terms']), int(g['scaling']))
    base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
    out = eng.value_and_grad(base)
    ev = eng.evolve(base)
    n = sp.state_num
    assert abs(out['loss'][0].item() - g['eval_loss']) < 1e-10
    assert abs(out['reg_loss'][0].item() - g['eval_reg_loss']) < 1e-10 * max(1, abs(g['eval_reg_loss']))
    assert abs(out['unitary_scale'][0].item() - g['eval_unitary_scale']) < 1e-10
    assert np.abs(out['grad'][0].cpu().numpy() - g['eval_grad']).max() < 1e-9 * max(1.0, np.abs(g['eval_grad']).max())
    fs = g['eval_final_state']
    st = bool(pb.get('state_transfer', False))
    if not st:
        assert np.linalg.norm(ev['U_final'][0].cpu().numpy() - (fs[:n, :n] + 1j * fs[n:, :n])) < 1e-9
    ivp = g['eval_inter_vecs_packed']                                   # [2n, T+1, m]
    iv = np.transpose(ivp[:n] + 1j * ivp[n:], (1, 2, 0))                 # [T+1, m, n]
    assert np.abs(ev['inter_vecs'][0].cpu().numpy() - iv).max() < 1e-9
    eng.close()
    uks, Uf = Grape(*args, convergence=conv, initial_guess=g['guess'], save=False, show_plots=False, quiet=True, **kw)
    assert np.abs(uks - g['uks']).max() < 1e-8
    if st:
        assert len(Uf) == 0
        return
    assert np.linalg.norm(Uf - g['U_final']) < 1e-5        # north_star bar; observed ~1e-12
    assert np.linalg.norm(Uf - g['U_final']) < 1e-8
    g32 = np.load(_os.path.join(_GOLD, "ref_%s_float32.npz" % name))    # the reference's real dtype
    assert np.linalg.norm(Uf - g32['U_final']) < 2e-3


# ---- fp32-class tcgen05 path (QOC_TF32X3) -------------------------------------------------------------
def _engine_tf32(args, kw, guess):
    from quantum_optimal_control.core.problem import SystemParameters
    from quantum_optimal_control.core.engine import GrapeEngine
    H0, Hops, Hn, U, tt, steps, scl = args
    sp = SystemParameters(H0, Hops, Hn, U, kw.get('U0', np.identity(len(H0))), tt, steps, scl, None, kw['maxA'], None, guess,
                          False, kw.get('unitary_error', 1e-4), False, False, kw.get('reg_coeffs'), False, None,
                          kw.get('Taylor_terms'), True, True, False, False, False)
    return sp, GrapeEngine.from_sys_para(sp, dtype='tf32x3')


@pytest.mark.parametrize("name", ['c1', 'c2_T40', 'c2_regs', 'c5_n16', 'c5_n8', 'c5_n20'])
def test_tf32x3_tcgen05_path_matches_oracle(name, built_lib):
    """Propagators from the tcgen05 kernel (3xTF32, fp32 accumulate in TMEM) vs the fp64 oracle.
    Tolerances are fp32-class (the s squarings double the fp32 rounding error each time):
    |dP| < 1e-5 per propagator, ||dU_final||_F < 2e-3 * sqrt(T/40), loss 3e-4, gradient 3e-3 relative to its
    scale -- the reference itself computes in float32 and its fp32 goldens differ from fp64 by as much."""
    fn, over, B = CASES[name]
    setups, guess, args, kw = make_case(fn(), seed=11, B=B, **over)
    sp, eng = _engine_tf32(args, kw, guess)
    base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
    out = eng.value_and_grad(base)
    ev = eng.evolve(base)
    eng.poll_error()
    P = eng.propagators().cpu().numpy()
    errs = []
    for b in range(B):
        ref = O.costate_value_and_grad(setups[b], setups[b].ops_weight_base)
        errs.append((np.abs(P[b] - ref['P']).max(), np.linalg.norm(ev['U_final'][b].cpu().numpy() - ref['U_final']),
                     abs(out['loss'][b].item() - ref['loss']), np.abs(out['grad'][b].cpu().numpy() - ref['grad']).max() / max(np.abs(ref['grad']).max(), 1e-30)))
        print('tf32x3 errors (dP, dU, dloss, dgrad_rel):', errs[-1])
        assert np.abs(P[b] - ref['P']).max() < 1e-5
        T = setups[b].steps
        assert np.linalg.norm(ev['U_final'][b].cpu().numpy() - ref['U_final']) < 2e-3 * max(1.0, np.sqrt(T / 40.0))
        assert abs(out['loss'][b].item() - ref['loss']) < 3e-4
        g = out['grad'][b].cpu().numpy()
        assert np.abs(g - ref['grad']).max() < 3e-3 * max(np.abs(ref['grad']).max(), 1e-30)
    eng.close()


# ---- degenerate shapes ---------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", ["T1", "K1_m1", "B5_odd"])
def test_edge_shapes(shape, built_lib):
    """T = 1 (no costate step), a single control / single concerned state, odd batch sizes."""
    if shape == "T1":
        pb, B = dict(W.c5_random(6, T=1), reg_coeffs={'dwdt': 0.5, 'd2wdt2': 0.1, 'amplitude': 0.2}), 2
    elif shape == "K1_m1":
        pb, B = dict(W.c5_random(7, T=9, K=1), states_concerned_list=[3]), 1
    else:
        pb, B = dict(W.c2_transmon_cavity(T=6), total_time=12.0), 5
    setups, guess, args, kw = make_case(pb, seed=31, B=B)
    for dtype in ('f64', 'tf32x3'):
        if dtype == 'f64':
            sp, eng = engine_for(args, kw, guess)
        else:
            sp, eng = _engine_tf32(args, kw, guess)
        base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
        out = eng.value_and_grad(base)
        ev = eng.evolve(base)
        eng.poll_error()
        tol = 1e-9 if dtype == 'f64' else 2e-3
        for b in range(B):
            ref = O.graph_value_and_grad(setups[b], setups[b].ops_weight_base)
            assert abs(out['loss'][b].item() - ref.loss) < tol
            assert abs(out['reg_loss'][b].item() - ref.reg_loss) < tol * max(1.0, abs(ref.reg_loss))
            assert np.abs(out['grad'][b].cpu().numpy() - ref.grad).max() < tol * max(np.abs(ref.grad).max(), 1e-30) + (0 if dtype == 'f64' else 1e-7)
            assert np.linalg.norm(ev['U_final'][b].cpu().numpy() - O.r_to_c_mat(ref.final_state, setups[b].n)) < tol
        eng.close()


# ---- SciPy driver and dressed basis -----------------------------------------------------------------------
def test_lbfgs_driver_matches_scipy_on_oracle(built_lib):
    """method='L-BFGS-B' (core/run_session.py:151-196): the same scipy.optimize.minimize call driven by the
    CPU oracle's get_error must land on the same pulse."""
    from scipy.optimize import minimize
    from quantum_optimal_control.main_grape.grape import Grape
    pb = W.c1_pi_pulse(T=30)
    args, kw = W.grape_kwargs(pb)
    guess = W.random_guess(2, 30, pb['maxA'], 17)
    conv = {'rate': 0.01, 'update_step': 10, 'max_iterations': 12, 'conv_target': 1e-12, 'learning_rate_decay': 100}
    uks, Uf = Grape(*args, convergence=conv, initial_guess=guess, method='L-BFGS-B', save=False, show_plots=False,
                    quiet=True, **kw)
    H0, Hops, Hn, U, tt, steps, scl = args
    st = O.make_setup(H0, Hops, U, tt, steps, scl, initial_guess=guess, **kw)

    def fun(x):
        o = O.graph_value_and_grad(st, x.reshape(2, 30))
        return np.float64(o.reg_loss), np.float64(o.grad.reshape(-1))

    res = minimize(fun, np.asarray(st.ops_weight_base).reshape(-1), method='L-BFGS-B', jac=True,
                   options={'maxfun': 12, 'gtol': 1e-25, 'disp': False, 'maxls': 40})
    ref_uks = np.asarray(pb['maxA'])[:, None] * np.sin(res['x'].reshape(2, 30))
    assert np.abs(uks - ref_uks).max() < 1e-6
    assert O.graph_value_and_grad(st, res['x'].reshape(2, 30)).loss < O.graph_value_and_grad(st, st.ops_weight_base).loss


@pytest.mark.parametrize("forbid_dressed", [False, True])
def test_dressed_initial_vectors(forbid_dressed, built_lib):
    """dressed_info: initial vectors are eigenvectors of H0 (core/system_parameters.py:178-179) -> general-V path."""
    from quantum_optimal_control.helper_functions.grape_functions import get_dressed_info
    pb = dict(W.c2_transmon_cavity(T=12), total_time=24.0)
    H0d = pb['H0'] + 0.05 * (pb['Hops'][0] + pb['Hops'][2])           # make the dressed basis non-trivial
    w_c, v_c, ids = get_dressed_info(H0d)
    pb['H0'] = H0d
    pb['dressed_info'] = {'eigenvectors': v_c, 'dressed_id': ids, 'eigenvalues': w_c, 'is_dressed': True}
    pb['reg_coeffs'] = {'forbidden_coeff_list': [2.0, 1.0], 'states_forbidden_list': [20, 11], 'forbid_dressed': forbid_dressed}
    setups, guess, args, kw = make_case(pb, seed=41, B=2)
    sp, eng = engine_for(args, kw, guess)
    assert sp.concerned_idx is None
    base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
    out = eng.value_and_grad(base)
    ev = eng.evolve(base)
    for b in range(2):
        ref = O.graph_value_and_grad(setups[b], setups[b].ops_weight_base)
        n = setups[b].n
        assert abs(out['loss'][b].item() - ref.loss) < 1e-10 and abs(out['reg_loss'][b].item() - ref.reg_loss) < 1e-10
        assert np.abs(out['grad'][b].cpu().numpy() - ref.grad).max() < 1e-9 * np.abs(ref.grad).max()
        iv_ref = np.transpose(ref.inter_vecs[:, :n, :] + 1j * ref.inter_vecs[:, n:, :], (2, 0, 1))
        assert np.abs(ev['inter_vecs'][b].cpu().numpy() - iv_ref).max() < 1e-9
    eng.close()


def test_full_size_c3_properties(built_lib):
    """BASELINE config C3 at full size (n=36, T=1000, B=1024, forbidden-state regulariser on 27 states):
    instance 0 and 1023 against the oracle, batch-permutation equivariance, loss consistent with U_final."""
    pb = W.c3_two_transmon_cnot()
    B = 1024
    setups, _, args, kw = make_case(pb, seed=500, B=1)
    guess = W.random_guess(4, 1000, pb['maxA'], 500, B=B)
    sp, eng = engine_for(args, kw, guess)
    assert (sp.exp_terms, sp.scaling) == (8, 2)
    base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
    out = {k: v.clone() for k, v in eng.value_and_grad(base).items()}
    U = eng.evolve(base, want_inter_vecs=False)['U_final']
    phi = torch.from_numpy(sp.target_vectors_c).cuda()
    idx = torch.as_tensor(sp.concerned_idx, device='cuda', dtype=torch.long)
    o = (phi.conj()[None] * U[:, :, idx].transpose(1, 2)).sum(dim=(1, 2))
    assert ((1 - (o.abs() ** 2) / 16) - out['loss']).abs().max().item() < 1e-10
    assert (out['reg_loss'] >= out['loss']).all()
    perm = torch.randperm(B, device='cuda')
    outp = eng.value_and_grad(base[perm].contiguous())
    assert torch.equal(outp['grad'], out['grad'][perm]) and torch.equal(outp['reg_loss'], out['reg_loss'][perm])
    for b in (0, B - 1):
        H0, Hops, Hn, Ut, tt, steps, scl = args
        st = O.make_setup(H0, Hops, Ut, tt, steps, scl, initial_guess=guess[b], **kw)
        ref = O.graph_value_and_grad(st, st.ops_weight_base)
        assert np.linalg.norm(U[b].cpu().numpy() - O.r_to_c_mat(ref.final_state, 36)) < 1e-8
        assert abs(out['reg_loss'][b].item() - ref.reg_loss) < 1e-9
        assert np.abs(out['grad'][b].cpu().numpy() - ref.grad).max() < 1e-9 * np.abs(ref.grad).max()
    eng.close()


# ---- few-state path: vector sweeps (TMA ring) + re-associated U_final (segment products) ---------------
SWEEP_CASES = {
    # T > 48 -> the U_final branch runs as 16-step segment products + a chain over the segments
    'c2_T100_regs': (lambda: W.c2_transmon_cavity(T=100), dict(total_time=200.0, reg_coeffs=ALL_REGS), 3),
    'c2_T67': (lambda: W.c2_transmon_cavity(T=67), dict(total_time=134.0), 2),           # last segment has 3 steps
    'c2_T65': (lambda: W.c2_transmon_cavity(T=65), dict(total_time=130.0), 1),           # last segment has 1 step
    # n = 36 (two rows per lane), m = 4, forbidden-state sources in the costate sweep
    'c3_T70_forbidden': (lambda: W.c3_two_transmon_cnot(T=70), dict(total_time=0.7), 2),
    # odd m (a padded column), U0 != identity, n not a multiple of 4
    'n13_m3_U0': (lambda: W.c5_random(13, T=60), dict(U0=np.linalg.qr(np.random.default_rng(8).normal(size=(13, 13)) +
                                                                       1j * np.random.default_rng(9).normal(size=(13, 13)))[0],
                                                       states_concerned_list=[0, 4, 12],
                                                       reg_coeffs={'speed_up': 0.3, 'forbidden_coeff_list': [2.0], 'states_forbidden_list': [7]}), 2),
    # many columns on one CTA (m = 11 < NP/2 = 12): column pairs are looped over inside a warp
    'n24_m11': (lambda: W.c5_random(24, T=52), dict(states_concerned_list=list(range(11))), 1),
}


@pytest.mark.parametrize("name", list(SWEEP_CASES))
def test_vec_sweep_path_matches_oracle_and_single_stream_path(name, built_lib, monkeypatch):
    """The few-state path (k_vec_sweep on the high-priority stream, k_segprod + chain for U_final) against the
    oracle AND against the single-stream chain/costate kernels it replaces (QOC_B200_NO_VEC_SWEEP=1)."""
    fn, over, B = SWEEP_CASES[name]
    setups, guess, args, kw = make_case(fn(), seed=23, B=B, **over)
    sp, eng = engine_for(args, kw, guess)
    base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
    n_launch0 = eng.launch_count
    out = {k: v.clone() for k, v in eng.value_and_grad(base).items()}
    n_launch = eng.launch_count - n_launch0
    ev = {k: (None if v is None else v.clone()) for k, v in eng.evolve(base).items()}
    torch.cuda.synchronize()
    monkeypatch.setenv("QOC_B200_NO_VEC_SWEEP", "1")
    n_launch0 = eng.launch_count
    old = {k: v.clone() for k, v in eng.value_and_grad(base).items()}
    assert eng.launch_count - n_launch0 < n_launch            # the switch really selected the other kernels
    ev_old = eng.evolve(base)
    torch.cuda.synchronize()
    monkeypatch.delenv("QOC_B200_NO_VEC_SWEEP")
    for k in ('loss', 'reg_loss', 'unitary_scale', 'grad_squared'):
        assert (out[k] - old[k]).abs().max().item() < 1e-11 * max(1.0, old[k].abs().max().item()), k
    assert (out['grad'] - old['grad']).abs().max().item() < 1e-11 * old['grad'].abs().max().item()
    assert (ev['U_final'] - ev_old['U_final']).abs().max().item() < 1e-11
    assert (ev['inter_vecs'] - ev_old['inter_vecs']).abs().max().item() < 1e-11
    for b in range(B):
        ref = O.graph_value_and_grad(setups[b], setups[b].ops_weight_base)
        n = setups[b].n
        assert np.linalg.norm(ev['U_final'][b].cpu().numpy() - O.r_to_c_mat(ref.final_state, n)) < ATOL_U
        iv_ref = np.transpose(ref.inter_vecs[:, :n, :] + 1j * ref.inter_vecs[:, n:, :], (2, 0, 1))
        assert np.abs(ev['inter_vecs'][b].cpu().numpy() - iv_ref).max() < ATOL_U
        assert abs(out['loss'][b].item() - ref.loss) < 1e-10
        assert abs(out['reg_loss'][b].item() - ref.reg_loss) < 1e-10 * max(1.0, abs(ref.reg_loss))
        assert abs(out['unitary_scale'][b].item() - ref.unitary_scale) < 1e-10
        assert np.abs(out['grad'][b].cpu().numpy() - ref.grad).max() < RTOL_G * max(np.abs(ref.grad).max(), 1e-300)
    eng.close()


@pytest.mark.parametrize("name", ['c2_T100_regs', 'c2_T67'])
def test_tf32x3_few_state_path_equals_single_stream_path(name, built_lib, monkeypatch):
    """tcgen05 path: both tails consume the SAME fp32 propagator tiles (widened to double), so the TMA-ring sweeps +
    segment products must agree with the single-stream chain / scalar costate to fp64 rounding."""
    fn, over, B = SWEEP_CASES[name]
    setups, guess, args, kw = make_case(fn(), seed=29, B=B, **over)
    sp, eng = _engine_tf32(args, kw, guess)
    base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
    n0 = eng.launch_count
    out = {k: v.clone() for k, v in eng.value_and_grad(base).items()}
    n_new = eng.launch_count - n0
    ev = {k: (None if v is None else v.clone()) for k, v in eng.evolve(base).items()}
    eng.poll_error()
    monkeypatch.setenv("QOC_B200_NO_VEC_SWEEP", "1")
    n0 = eng.launch_count
    old = {k: v.clone() for k, v in eng.value_and_grad(base).items()}
    assert eng.launch_count - n0 < n_new
    ev_old = eng.evolve(base)
    eng.poll_error()
    monkeypatch.delenv("QOC_B200_NO_VEC_SWEEP")
    for k in ('loss', 'reg_loss', 'unitary_scale', 'grad_squared'):
        assert (out[k] - old[k]).abs().max().item() < 1e-11 * max(1.0, old[k].abs().max().item()), k
    assert (out['grad'] - old['grad']).abs().max().item() < 1e-11 * old['grad'].abs().max().item()
    assert (ev['U_final'] - ev_old['U_final']).abs().max().item() < 1e-11
    assert (ev['inter_vecs'] - ev_old['inter_vecs']).abs().max().item() < 1e-11
    eng.close()


def test_value_and_grad_is_repeatable_across_streams(built_lib):
    """Back-to-back calls reuse the propagator workspace while the high-priority branch of the previous call
    may still be reading it: results must not depend on that (joins are in place)."""
    setups, guess, args, kw = make_case(W.c2_transmon_cavity(T=80), seed=41, B=16, total_time=160.0)
    sp, eng = engine_for(args, kw, guess)
    base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
    first = {k: v.clone() for k, v in eng.value_and_grad(base).items()}
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(5):
            out = eng.value_and_grad(base)
        again = {k: v.clone() for k, v in out.items()}
    s.synchronize()
    for k in first:
        assert torch.equal(first[k], again[k]), k
    eng.close()


def test_batch_chunking_gives_identical_results(built_lib, monkeypatch):
    """A workspace budget smaller than the full batch makes the library process the batch in chunks
    (qoc_batch_chunk < B); every output must be bit-identical to the single-pass run."""
    pb = dict(W.c2_transmon_cavity(T=16), total_time=32.0)
    pb['reg_coeffs'] = {'dwdt': 0.1, 'forbidden_coeff_list': [1.0], 'states_forbidden_list': [20]}
    setups, guess, args, kw = make_case(pb, seed=77, B=7)
    sp, eng = engine_for(args, kw, guess)
    assert eng.batch_chunk == 7
    base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
    ref = {k: v.clone() for k, v in eng.value_and_grad(base).items()}
    ev_ref = {k: (None if v is None else v.clone()) for k, v in eng.evolve(base).items()}
    evh_ref = eng.evolve_host(sp.ops_weight_base)
    need = eng.workspace_bytes
    eng.close()
    monkeypatch.setenv("QOC_B200_MAX_WS_GB", str(need * 0.45 / 1e9))
    sp2, eng2 = engine_for(args, kw, guess)
    assert 1 <= eng2.batch_chunk < 7 and eng2.workspace_bytes < need
    out = eng2.value_and_grad(base)
    ev = eng2.evolve(base)
    evh = eng2.evolve_host(sp.ops_weight_base)
    for k in ref:
        assert torch.equal(out[k], ref[k]), k
    for k in ('U_final', 'inter_vecs', 'loss', 'unitary_scale'):
        assert torch.equal(ev[k], ev_ref[k]), k
        assert np.array_equal(evh[k], evh_ref[k]), k
    eng2.close()


_NCCL = r'''
import os, sys
sys.path.insert(0, os.path.join(%(root)r, "quantum-optimal-control_b200")); sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
rank = int(sys.argv[1]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%(port)d", rank=rank, world_size=2, device_id=torch.device("cuda", rank))
import workloads as W
from quantum_optimal_control.main_grape.grape import Grape
from quantum_optimal_control.core.population import grape_population
pb = W.c1_pi_pulse(T=30); args, kw = W.grape_kwargs(pb)
g = W.random_guess(2, 30, pb["maxA"], 3, B=5)
conv = {"rate": 0.02, "update_step": 50, "max_iterations": 10, "conv_target": 1e-12, "learning_rate_decay": 100}
res = grape_population(Grape, args, g, convergence=conv, save=False, show_plots=False, quiet=True, **kw)
if rank == 0:
    np.savez(sys.argv[2], best=res["best"], loss=res["loss"], uks=res["uks"], U=res["U_final"])
dist.destroy_process_group()
'''


def test_population_sweep_two_gpus_nccl(built_lib, tmp_path):
    """Batch sharded over 2 GPUs, one NCCL all-gather of the losses + broadcast of the winner (needs 2 GPUs)."""
    import os, subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from quantum_optimal_control.main_grape.grape import Grape
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "nccl_pop.py"
    out = tmp_path / "res.npz"
    script.write_text(_NCCL % dict(root=root, port=29650 + os.getpid() % 200))
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(out)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    res = np.load(out)
    pb = W.c1_pi_pulse(T=30)
    args, kw = W.grape_kwargs(pb)
    g = W.random_guess(2, 30, pb['maxA'], 3, B=5)
    conv = {"rate": 0.02, "update_step": 50, "max_iterations": 10, "conv_target": 1e-12, "learning_rate_decay": 100}
    uks, Uf, losses = Grape(*args, initial_guess=g, convergence=conv, save=False, show_plots=False, quiet=True, return_losses=True, **kw)
    assert np.allclose(res['loss'], losses, atol=1e-12) and int(res['best']) == int(np.argmin(losses))
    assert np.allclose(res['uks'], uks[int(res['best'])], atol=1e-12)


def test_run_file_schema(built_lib, tmp_path):
    """save=True writes the reference's datasets (core/run_session.py:129-138, core/analysis.py:31-33,62-65,95-99)."""
    import os
    from quantum_optimal_control.main_grape.grape import Grape
    pb = W.c1_pi_pulse(T=20)
    args, kw = W.grape_kwargs(pb)
    guess = W.random_guess(2, 20, pb['maxA'], 2)
    conv = {'rate': 0.02, 'update_step': 4, 'evol_save_step': 4, 'max_iterations': 9, 'conv_target': 1e-12,
            'learning_rate_decay': 100}
    uks, Uf = Grape(*args, convergence=conv, initial_guess=guess, save=True, file_name="run", data_path=str(tmp_path),
                    show_plots=False, quiet=True, **kw)
    files = sorted(os.listdir(tmp_path))
    assert len(files) == 1 and files[0].startswith("00000_run")
    path = str(tmp_path / files[0])
    if path.endswith(".npz"):
        f = dict(np.load(path, allow_pickle=True))
    else:
        import h5py
        with h5py.File(path, 'r') as hf:
            f = {}
            hf.visititems(lambda k, v: f.__setitem__(k, v[()]) if hasattr(v, 'shape') else None)
    n_saves = 3 + 1                     # iterations 0, 4, 8 and the end result
    assert f['error'].shape == (n_saves,) and f['uks'].shape == (n_saves, 2, 20)
    assert list(f['iteration']) == [0, 4, 8, 9]
    assert np.allclose(f['uks'][-1], uks)
    assert f['final_state'].shape == (n_saves + 1, 4, 4)          # the end result appends it twice, like the reference
    fs = f['final_state'][-1]
    assert np.allclose(fs[:2, :2] + 1j * fs[2:, :2], Uf)
    assert f['inter_vecs_raw_real'].shape == (n_saves, 2, 2, 21) and f['inter_vecs_mag_squared'].shape == (n_saves, 2, 2, 21)
    assert int(f['taylor_terms']) >= 3 and int(f['taylor_scaling']) >= 0 and 'wall_clock_time' in f
    assert f['convergence/max_iterations'] == 9 and np.allclose(f['H0'], pb['H0'])
