"""CPU tests (no GPU): pin the oracle against (a) fixtures produced by the reference's own source
(oracle/run_reference.py -> tests/golden/ref_*.npz), (b) closed-form / SciPy known answers,
(c) its own independent complex-costate formulation."""
import glob
import os

import numpy as np
import pytest
import scipy.linalg as sla
import torch

import workloads as W
from oracle import grape_oracle as O
from oracle.run_reference import golden_cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _setup(name, guess):
    pb, seed, conv = golden_cases()[name]
    args, kw = W.grape_kwargs(pb)
    H0, Hops, Hn, U, tt, steps, scl = args
    return O.make_setup(H0, Hops, U, tt, steps, scl, initial_guess=guess, **kw), args, kw, conv


@pytest.mark.parametrize("name", list(golden_cases()))
def test_oracle_matches_reference_goldens_fp64(name):
    g = np.load(os.path.join(GOLD, "ref_%s_float64.npz" % name))
    setup, args, kw, conv = _setup(name, g['guess'])
    assert (setup.exp_terms, setup.scaling) == (int(g['exp_terms']), int(g['scaling']))
    out = O.graph_value_and_grad(setup, setup.ops_weight_base)
    assert abs(out.loss - g['eval_loss']) < 1e-12
    assert abs(out.reg_loss - g['eval_reg_loss']) < 1e-12 * max(1, abs(g['eval_reg_loss']))
    assert abs(out.unitary_scale - g['eval_unitary_scale']) < 1e-12
    assert np.abs(out.grad - g['eval_grad']).max() < 1e-12 * max(1.0, np.abs(g['eval_grad']).max())
    assert abs(out.grad_squared - g['eval_grad_squared']) < 1e-12 * max(1, g['eval_grad_squared'])
    assert np.abs(out.final_state - g['eval_final_state']).max() < 1e-12
    assert np.abs(np.transpose(out.inter_vecs, (1, 2, 0)) - g['eval_inter_vecs_packed']).max() < 1e-12
    uks, Uf = O.grape(*args, convergence=conv, initial_guess=g['guess'], **kw)
    assert np.abs(uks - g['uks']).max() < 1e-10
    if setup.state_transfer:
        assert len(Uf) == 0 and g['U_final'].size == 0           # run_session.py:109-110
    else:
        assert np.linalg.norm(Uf - g['U_final']) < 1e-10


@pytest.mark.parametrize("name", ['c1_pi_pulse', 'c2_small_allregs'])
def test_oracle_matches_reference_goldens_fp32(name):
    """float32 is the reference's real dtype; op ordering inside torch/BLAS differs from run to run,
    so this is a tolerance check (1e-4 relative), not bit equality."""
    g = np.load(os.path.join(GOLD, "ref_%s_float32.npz" % name))
    setup, args, kw, conv = _setup(name, g['guess'])
    out = O.graph_value_and_grad(setup, setup.ops_weight_base, dtype=torch.float32)
    assert abs(out.loss - g['eval_loss']) < 1e-4
    assert np.abs(out.grad - g['eval_grad']).max() < 1e-4 * max(1.0, np.abs(g['eval_grad']).max())
    g64 = np.load(os.path.join(GOLD, "ref_%s_float64.npz" % name))
    assert np.linalg.norm(g['U_final'] - g64['U_final']) < 5e-4      # fp32 reference vs fp64 semantics


def test_costate_form_equals_graph_form():
    for name in golden_cases():
        g = np.load(os.path.join(GOLD, "ref_%s_float64.npz" % name))
        setup, *_ = _setup(name, g['guess'])
        if setup.state_transfer:
            continue
        a = O.graph_value_and_grad(setup, setup.ops_weight_base)
        b = O.costate_value_and_grad(setup, setup.ops_weight_base)
        assert abs(a.loss - b['loss']) < 1e-12 and abs(a.reg_loss - b['reg_loss']) < 1e-11 * max(1, abs(a.reg_loss))
        assert np.abs(a.grad - b['grad']).max() < 1e-11 * max(1.0, np.abs(a.grad).max())
        assert abs(a.unitary_scale - b['unitary_scale']) < 1e-12
        assert np.abs(O.r_to_c_mat(a.final_state, setup.n) - b['U_final']).max() < 1e-12


def test_analytic_pi_pulse():
    """H0 = 0, H = u*sx/2 constant with u*total_time = pi  ->  U = exp(-i pi sx / 2) = -i sx; loss 0."""
    sx = np.array([[0, 1], [1, 0]], dtype=complex)
    T, total = 50, 5.0
    u = np.pi / total
    guess = np.full((1, T), u)
    st = O.make_setup(np.zeros((2, 2), dtype=complex), [sx / 2], -1j * sx, total, T, [0, 1], maxA=[2 * u],
                      initial_guess=guess, Taylor_terms=[14, 2])
    out = O.graph_value_and_grad(st, st.ops_weight_base)
    U = O.r_to_c_mat(out.final_state, 2)
    assert np.abs(U - (-1j * sx)).max() < 1e-12
    assert abs(out.loss) < 1e-12 and abs(out.unitary_scale - 1.0) < 1e-12
    assert np.abs(out.grad).max() < 1e-12           # stationary point of the fidelity


def test_forward_matches_scipy_expm():
    pb = W.c5_random(12, T=15)
    args, kw = W.grape_kwargs(pb)
    H0, Hops, Hn, U, tt, steps, scl = args
    guess = W.random_guess(2, 15, pb['maxA'], 3)
    kw['Taylor_terms'] = [18, 4]
    st = O.make_setup(H0, Hops, U, tt, steps, scl, initial_guess=guess, **kw)
    out = O.graph_value_and_grad(st, st.ops_weight_base)
    X = np.eye(12, dtype=complex)
    dt = tt / steps
    for t in range(steps):
        H = H0 + sum(guess[k, t] * Hops[k] for k in range(2))
        X = sla.expm(-1j * dt * H) @ X
    assert np.abs(O.r_to_c_mat(out.final_state, 12) - X).max() < 1e-11
    assert np.abs(X.conj().T @ X - np.eye(12)).max() < 1e-12


def test_first_order_gradient_approaches_exact_for_small_dt():
    """The reference's matexp gradient is the first-order GRAPE approximation; it converges to the
    finite-difference gradient as dt -> 0 (sanity of sign / scaling conventions)."""
    pb = W.c5_random(6, T=8)
    pb['total_time'] = 0.008
    args, kw = W.grape_kwargs(pb)
    H0, Hops, Hn, U, tt, steps, scl = args
    guess = W.random_guess(2, 8, pb['maxA'], 4)
    st = O.make_setup(H0, Hops, U, tt, steps, scl, initial_guess=guess, **kw)
    base = st.ops_weight_base
    out = O.graph_value_and_grad(st, base)
    eps = 1e-6
    for (k, t) in [(0, 0), (1, 5)]:
        bp, bm = base.copy(), base.copy()
        bp[k, t] += eps
        bm[k, t] -= eps
        fd = (O.graph_value_and_grad(st, bp, want_grad=False).loss - O.graph_value_and_grad(st, bm, want_grad=False).loss) / (2 * eps)
        assert abs(fd - out.grad[k, t]) < 2e-2 * np.abs(out.grad).max()


def test_taylor_term_chooser_known_values():
    """(p, s) picked by the restated chooser (core/system_parameters.py:122-231)."""
    for fn, want in [(W.c1_pi_pulse, (7, 2)), (W.c2_transmon_cavity, (7, 3)), (W.c3_two_transmon_cnot, (8, 2))]:
        pb = fn()
        args, kw = W.grape_kwargs(pb)
        H0, Hops, Hn, U, tt, steps, scl = args
        st = O.make_setup(H0, Hops, U, tt, steps, scl, initial_guess=W.random_guess(len(Hops), steps, pb['maxA'], 0), **kw)
        assert (st.exp_terms, st.scaling) == want


def test_tf1_adam_formula():
    a = O.TF1Adam((3,))
    th = np.array([1.0, -2.0, 0.5])
    g = np.array([0.1, -0.2, 0.3])
    th1 = a.step(th, g, 0.01)
    # first step: m = .1 g, v = .001 g^2, lr_t = lr*sqrt(.001)/.1  ->  theta - lr * g/(|g| + eps*sqrt(1000)..)
    lr_t = 0.01 * np.sqrt(1 - 0.999) / (1 - 0.9)
    want = th - lr_t * (0.1 * g) / (np.sqrt(0.001 * g * g) + 1e-8)
    assert np.allclose(th1, want, rtol=0, atol=1e-15)


def test_reference_quirks():
    pb = W.c1_pi_pulse(T=10)
    args, kw = W.grape_kwargs(pb)
    H0, Hops, Hn, U, tt, steps, scl = args
    with pytest.raises(ValueError):                       # guess > maxA (system_parameters.py:44-45)
        O.make_setup(H0, Hops, U, tt, steps, scl, initial_guess=np.full((2, 10), 3.0), maxA=[2.0, 2.0])
    O.make_setup(H0, Hops, U, tt, steps, scl, initial_guess=np.full((2, 10), -1.0), maxA=[2.0, 2.0])
    kw2 = dict(kw, reg_coeffs={'d2wdt2': 1.0})
    st = O.make_setup(H0, Hops, U, tt, steps, scl, initial_guess=W.random_guess(2, 10, pb['maxA'], 0), **kw2)
    with pytest.raises(NameError):                        # 'd2wdt2' without 'dwdt' (regularization_functions.py:41)
        O.graph_value_and_grad(st, st.ops_weight_base)


def test_golden_files_present():
    assert len(glob.glob(os.path.join(GOLD, "ref_*_float64.npz"))) == len(golden_cases())


# ---- round-2 fixtures (oracle/run_reference.py r2): early stop, dressed + forbid_dressed, U0 != I, n = 64 / 216 ------------
from oracle.run_reference import golden_cases_r2  # noqa: E402


@pytest.mark.parametrize("name", [k for k, c in golden_cases_r2().items() if c['method'] == 'Adam'])
def test_oracle_matches_r2_reference_goldens_fp64(name):
    c = golden_cases_r2()[name]
    g = np.load(os.path.join(GOLD, "ref2_%s_float64.npz" % name))
    args, kw = W.grape_kwargs(c['pb'])
    H0, Hops, Hn, U, tt, steps, scl = args
    for b in range(len(g['seeds'])):
        setup = O.make_setup(H0, Hops, U, tt, steps, scl, initial_guess=g['guess'][b], **kw)
        assert (setup.exp_terms, setup.scaling) == (int(g['exp_terms']), int(g['scaling']))
        out = O.graph_value_and_grad(setup, setup.ops_weight_base)
        assert abs(out.loss - g['eval_loss'][b]) < 1e-12
        assert abs(out.reg_loss - g['eval_reg_loss'][b]) < 1e-12 * max(1, abs(g['eval_reg_loss'][b]))
        assert np.abs(out.grad - g['eval_grad'][b]).max() < 1e-11 * max(1.0, np.abs(g['eval_grad'][b]).max())
        assert np.abs(out.final_state - g['eval_final_state'][b]).max() < 1e-11
        if name == 'c4_T50':
            break                                  # n = 216: one evaluation of one seed keeps the CPU suite short
        uks, Uf = O.grape(*args, convergence=c['conv'], initial_guess=g['guess'][b], **kw)
        assert np.abs(uks - g['uks'][b]).max() < 1e-10           # incl. the iterate at which the early stop fired
        assert np.linalg.norm(Uf - g['U_final'][b]) < 1e-10


def test_r2_fixture_early_stop_has_mixed_stop_times():
    g = np.load(os.path.join(GOLD, "ref2_c1_earlystop_float64.npz"))
    its = [int(i) for i in g['run_iterations']]
    assert len(set(its)) > 1 and max(its) < int(g['max_iterations'])
    assert all(float(l) < 0.3 for l in g['run_final_loss'])     # stopped by loss < conv_target (run_session.py:56-58)
