"""GPU parity of the fp32-class tcgen05 / TMA path (dtype 'f16x2', QOC_F16X2) and of the round-2 reference-run
fixtures (tests/golden/ref2_*.npz, written by oracle/run_reference.py r2 from the reference's own source).

Tolerances of the fp32-class path.  The reference itself computes in float32 (core/tensorflow_state.py:49,70,205);
its own float32 run differs from its float64 run by ``gap`` (both fixtures are committed).  Our path holds every real
number as a pair of fp16 halves (>= 22 bits), accumulates in fp32 in TMEM and keeps states / costates in fp64, so it
must sit within a small multiple of that gap from BOTH reference runs:

    |ours - ref_float64| <= 4 gap + floor       and      |ours - ref_float32| <= 5 gap + floor

with floors of a few fp32 ulps of the quantity's scale (the gap of a single short run can be accidentally tiny).
"""
import os

import numpy as np
import pytest
import torch

import workloads as W
from oracle import grape_oracle as O
from oracle.run_reference import golden_cases, golden_cases_r2
from helpers import make_case

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _engine(args, kw, guess, dtype):
    from quantum_optimal_control.core.problem import SystemParameters
    from quantum_optimal_control.core.engine import GrapeEngine
    H0, Hops, Hn, U, tt, steps, scl = args
    sp = SystemParameters(H0, Hops, Hn, U, kw.get('U0', np.identity(len(H0))), tt, steps, scl, kw.get('dressed_info'),
                          kw['maxA'], None, guess, False, kw.get('unitary_error', 1e-4), kw.get('state_transfer', False), False,
                          kw.get('reg_coeffs'), False, None, kw.get('Taylor_terms'), True, True, False, False, False)
    return sp, GrapeEngine.from_sys_para(sp, dtype=dtype)


def _load_case(name):
    """-> (pb, conv, method, g64, g32) with leading seed axis on every array."""
    if name in golden_cases():
        pb, seed, conv = golden_cases()[name]
        g64 = {k: v[None] if k not in ('exp_terms', 'scaling', 'seed', 'max_iterations') else v
               for k, v in np.load(os.path.join(GOLD, "ref_%s_float64.npz" % name)).items()}
        g32 = {k: v[None] if k not in ('exp_terms', 'scaling', 'seed', 'max_iterations') else v
               for k, v in np.load(os.path.join(GOLD, "ref_%s_float32.npz" % name)).items()}
        return pb, conv, 'Adam', g64, g32
    c = golden_cases_r2()[name]
    g64 = dict(np.load(os.path.join(GOLD, "ref2_%s_float64.npz" % name)))
    g32 = dict(np.load(os.path.join(GOLD, "ref2_%s_float32.npz" % name)))
    return c['pb'], c['conv'], c['method'], g64, g32


def _complex_U(fs, n):
    return fs[..., :n, :n] + 1j * fs[..., n:, :n]


# ------------------------------------------------------------------------------------------------------------------
# fp32-class path against the reference's float32 AND float64 runs: n = 36, 64, 216
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ['c3_small_forbidden', 'c5_n64_m4', 'c4_T50', 'c2_dressed_forbid', 'n5_U0', 'c2_small_allregs'])
def test_f16x2_matches_reference_float32_goldens(name, built_lib):
    pb, conv, method, g64, g32 = _load_case(name)
    args, kw = W.grape_kwargs(pb)
    B = g64['guess'].shape[0]
    sp, eng = _engine(args, kw, g64['guess'], 'f16x2')
    n = sp.state_num
    assert (sp.exp_terms, sp.scaling) == (int(g64['exp_terms']), int(g64['scaling']))
    base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
    out = eng.value_and_grad(base)
    ev = eng.evolve(base)
    eng.poll_error()
    U = ev['U_final'].cpu().numpy()
    iv = ev['inter_vecs'].cpu().numpy()                                    # [B, T+1, m, n]
    for b in range(B):
        U64, U32 = _complex_U(g64['eval_final_state'][b], n), _complex_U(g32['eval_final_state'][b], n)
        gapU = np.linalg.norm(U32 - U64)
        assert np.linalg.norm(U[b] - U64) <= 4 * gapU + 2e-5 * np.sqrt(n), (name, b, np.linalg.norm(U[b] - U64), gapU)
        assert np.linalg.norm(U[b] - U32) <= 5 * gapU + 2e-5 * np.sqrt(n)
        ivp64, ivp32 = g64['eval_inter_vecs_packed'][b], g32['eval_inter_vecs_packed'][b]        # [2n, T+1, m]
        iv64 = np.transpose(ivp64[:n] + 1j * ivp64[n:], (1, 2, 0))
        iv32 = np.transpose(ivp32[:n] + 1j * ivp32[n:], (1, 2, 0))
        gap_iv = np.abs(iv32 - iv64).max()
        assert np.abs(iv[b] - iv64).max() <= 4 * gap_iv + 2e-6
        gap_l = abs(float(g32['eval_loss'][b]) - float(g64['eval_loss'][b]))
        assert abs(out['loss'][b].item() - float(g64['eval_loss'][b])) <= 4 * gap_l + 2e-6
        assert abs(out['reg_loss'][b].item() - float(g64['eval_reg_loss'][b])) <= 4 * abs(
            float(g32['eval_reg_loss'][b]) - float(g64['eval_reg_loss'][b])) + 2e-6 * max(1.0, abs(float(g64['eval_reg_loss'][b])))
        gs = max(np.abs(g64['eval_grad'][b]).max(), 1e-30)
        gap_g = np.abs(g32['eval_grad'][b] - g64['eval_grad'][b]).max()
        assert np.abs(out['grad'][b].cpu().numpy() - g64['eval_grad'][b]).max() <= 4 * gap_g + 2e-5 * gs
        assert abs(out['unitary_scale'][b].item() - float(g64['eval_unitary_scale'][b])) <= 4 * abs(
            float(g32['eval_unitary_scale'][b]) - float(g64['eval_unitary_scale'][b])) + 1e-4
    eng.close()


@pytest.mark.parametrize("name", ['c3_small_forbidden', 'c5_n64_m4'])
def test_f16x2_grape_run_returns_reference_dtypes(name, built_lib):
    """Full Grape() Adam run on the fp32-class path with return_dtype='reference': float32 pulses / complex64 final
    state as the reference returns them (run_session.py:112-117, analysis.py:18-24), within the fp32 gap of its runs."""
    from quantum_optimal_control.main_grape.grape import Grape
    pb, conv, method, g64, g32 = _load_case(name)
    args, kw = W.grape_kwargs(pb)
    for b in range(g64['guess'].shape[0]):
        uks, Uf = Grape(*args, convergence=conv, initial_guess=g64['guess'][b], save=False, show_plots=False, quiet=True,
                        dtype='f16x2', return_dtype='reference', **kw)
        assert uks.dtype == np.float32 and Uf.dtype == np.complex64
        gap_u = np.abs(g32['uks'][b] - g64['uks'][b]).max()
        assert np.abs(uks - g64['uks'][b]).max() <= 4 * gap_u + 1e-5 * max(1.0, np.abs(g64['uks'][b]).max())
        gapU = np.linalg.norm(g32['U_final'][b] - g64['U_final'][b])
        assert np.linalg.norm(Uf - g64['U_final'][b]) <= 4 * gapU + 2e-5 * np.sqrt(Uf.shape[0])


F16_CASES = {
    'c1': (lambda: W.c1_pi_pulse(), {}, 2),
    'c2_T40': (lambda: W.c2_transmon_cavity(T=40), dict(total_time=80.0), 2),
    'c2_regs': (lambda: W.c2_transmon_cavity(T=25), dict(total_time=50.0, reg_coeffs={
        'amplitude': 0.3, 'envelope': 0.7, 'dwdt': 0.02, 'd2wdt2': 0.0005, 'speed_up': 0.4,
        'forbidden_coeff_list': [3.0, 5.0, 2.0], 'states_forbidden_list': [2, 3, 2]}), 2),
    'c3_T30': (lambda: W.c3_two_transmon_cnot(T=30), dict(total_time=0.3), 2),
    'c5_n48_m3': (lambda: W.c5_random(48, T=12), dict(states_concerned_list=[0, 7, 47]), 2),
    'c5_n100_regs': (lambda: W.c5_random(100, T=20), dict(states_concerned_list=[0, 5, 99], reg_coeffs={
        'dwdt': 0.1, 'forbidden_coeff_list': [2.0], 'states_forbidden_list': [7]}), 2),
    'c5_n128_m8': (lambda: W.c5_random(128, T=12), dict(states_concerned_list=list(range(8))), 2),
    'c5_n200_m5': (lambda: W.c5_random(200, T=6), dict(states_concerned_list=[0, 1, 64, 128, 199]), 1),
    # T a multiple of the segment length: the U_final branch's segment products run on the CTA-pair kernel as well
    'c5_n136_T32': (lambda: W.c5_random(136, T=32), dict(states_concerned_list=[0, 5, 135]), 2),
    'c5_n256_m2': (lambda: W.c5_random(256, T=4), dict(states_concerned_list=[3, 255]), 1),
    'n5_U0': (lambda: W.c5_random(5, T=15), dict(U0=np.linalg.qr(np.random.default_rng(5).normal(size=(5, 5)) +
                                                                  1j * np.random.default_rng(6).normal(size=(5, 5)))[0],
                                                  states_concerned_list=[1, 3]), 2),
}


@pytest.mark.parametrize("name", list(F16_CASES))
def test_f16x2_matches_fp64_oracle(name, built_lib):
    """Every stage of the fp32-class path against the fp64 CPU oracle: propagators |dP| < 2e-6 (+ 2^s amplification
    of the Taylor stage), states 2e-6 + 3e-7 T, loss 2e-6 + 2e-7 T, gradient 1e-4 of its scale, ||dU_final||_F < 1e-4 sqrt(n T/10)."""
    fn, over, B = F16_CASES[name]
    setups, guess, args, kw = make_case(fn(), seed=11, B=B, **over)
    sp, eng = _engine(args, kw, guess, 'f16x2')
    n, T = sp.state_num, sp.steps
    base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
    out = eng.value_and_grad(base)
    ev = eng.evolve(base)
    P = eng.propagators().cpu().numpy()
    eng.poll_error()
    for b in range(B):
        ref = O.costate_value_and_grad(setups[b], setups[b].ops_weight_base) if not setups[b].is_dressed else None
        assert np.abs(P[b] - ref['P']).max() < 2e-6 * max(1, 2 ** (sp.scaling - 4))
        assert abs(out['loss'][b].item() - ref['loss']) < 2e-6 + 2e-7 * T
        assert abs(out['reg_loss'][b].item() - ref['reg_loss']) < (2e-6 + 2e-7 * T) * max(1.0, abs(ref['reg_loss']))
        assert np.abs(out['grad'][b].cpu().numpy() - ref['grad']).max() < 1e-4 * np.abs(ref['grad']).max()
        assert np.abs(ev['inter_vecs'][b].cpu().numpy() - np.transpose(ref['inter_vecs'], (1, 0, 2))).max() < 2e-6 + 3e-7 * T
        assert np.linalg.norm(ev['U_final'][b].cpu().numpy() - ref['U_final']) < 1e-4 * np.sqrt(n * max(T, 10) / 10.0)
        assert abs(out['unitary_scale'][b].item() - ref['unitary_scale']) < 2e-4
    eng.close()


@pytest.mark.parametrize("name,knob", [('c5_n200_m5', 'QOC_B200_TC_PAIR'), ('c5_n256_m2', 'QOC_B200_TC_PAIR'), ('c5_n136_T32', 'QOC_B200_TC_PAIR'),
                                       ('c3_T30', 'QOC_B200_TC_SMALL'), ('c5_n48_m3', 'QOC_B200_TC_SMALL'),
                                       ('c2_T40', 'QOC_B200_TC_SMALL')])
def test_f16x2_engines_agree_bit_for_bit(name, knob, built_lib, monkeypatch):
    """The three tcgen05 engines evaluate the same products in the same order (k-blocks ascending, 12 MMAs per 16 columns,
    fp32 accumulators, the same epilogue arithmetic): the CTA-pair kernel (n > 128) and the shared-memory-resident kernel
    (n <= 64, with its narrow products: QOC_B200_SMALL_WIDE=0) must reproduce the streamed-operand engine's propagators
    BIT FOR BIT, hence identical losses and gradients."""
    fn, over, B = F16_CASES[name]
    setups, guess, args, kw = make_case(fn(), seed=11, B=B, **over)
    monkeypatch.setenv('QOC_B200_SMALL_WIDE', '0')
    res = {}
    for v in ('1', '0'):
        monkeypatch.setenv(knob, v)
        sp, eng = _engine(args, kw, guess, 'f16x2')
        base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
        out = eng.value_and_grad(base)
        res[v] = (eng.propagators().clone(), out['loss'].clone(), out['grad'].clone(), eng.evolve(base)['U_final'].clone())
        eng.poll_error()
        eng.close()
    assert torch.equal(res['1'][0], res['0'][0])
    assert torch.equal(res['1'][1], res['0'][1]) and torch.equal(res['1'][2], res['0'][2])
    assert torch.equal(res['1'][3], res['0'][3])


@pytest.mark.parametrize("name", ['c3_T30', 'c5_n48_m3'])
def test_f16x2_wide_products_match_narrow_ones(name, built_lib, monkeypatch):
    """32 < n <= 64: the N = 128 products (D1 = Ar [Br | Bi], D2 = Ai [Br | Bi], combined in the epilogue) against the
    12-MMA form: the same fp16-pair terms, two fp32 accumulators instead of one per component -> equal to fp32 rounding."""
    fn, over, B = F16_CASES[name]
    setups, guess, args, kw = make_case(fn(), seed=11, B=B, **over)
    res = {}
    for v in ('1', '0'):
        monkeypatch.setenv('QOC_B200_SMALL_WIDE', v)
        sp, eng = _engine(args, kw, guess, 'f16x2')
        base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
        out = eng.value_and_grad(base)
        res[v] = (eng.propagators().clone(), out['loss'].clone(), out['grad'].clone())
        eng.poll_error()
        eng.close()
    assert (res['1'][0] - res['0'][0]).abs().max().item() < 1e-6
    assert (res['1'][1] - res['0'][1]).abs().max().item() < 1e-6
    assert (res['1'][2] - res['0'][2]).abs().max().item() < 1e-5 * res['0'][2].abs().max().item()


SHAPES = [  # (n, K, T, m, B): odd sizes, every engine boundary (32/33, 64/65, 128/129), T around the segment length, m = 1 .. 8
    (3, 1, 7, 1, 2), (9, 2, 17, 3, 3), (31, 2, 16, 2, 2), (33, 3, 33, 7, 2), (47, 1, 48, 8, 2), (63, 2, 15, 5, 2), (65, 2, 32, 4, 2),
    (97, 2, 18, 1, 2), (127, 3, 16, 6, 1), (129, 2, 32, 2, 2), (161, 1, 17, 8, 1), (209, 2, 16, 3, 1), (255, 2, 5, 2, 1),
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "n%d_K%d_T%d_m%d" % s[:4])
def test_f16x2_matches_f64_engine_on_odd_shapes(shape, built_lib):
    """The fp32-class path against the fp64 path of the SAME library on shapes that sit on the engines' boundaries (tile
    padding, partial chunks of the sweeps, segment remainders): propagators, states, loss, gradient, U_final."""
    n, K, T, m, B = shape
    pb = W.c5_random(n, T=T, K=K, seed=4000 + n)
    rng = np.random.default_rng(n)
    if n % 2 == 1 and n > 8:                      # sparse Hamiltonians (band + a few couplings): the pattern-scatter generator assembly
        i, j = np.indices((n, n))
        mask = (np.abs(i - j) <= 1) | ((i + 2 * j) % 11 == 0) | ((j + 2 * i) % 11 == 0)
        pb['H0'] = pb['H0'] * mask * 3.0
        pb['Hops'] = [h * mask * 3.0 for h in pb['Hops']]
    pb['states_concerned_list'] = sorted(rng.choice(n, size=m, replace=False).tolist())
    pb['reg_coeffs'] = {'dwdt': 0.05, 'forbidden_coeff_list': [1.5], 'states_forbidden_list': [n - 1]}
    setups, guess, args, kw = make_case(pb, seed=17, B=B)
    res = {}
    for dt in ('f64', 'f16x2'):
        sp, eng = _engine(args, kw, guess, dt)
        base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
        out = eng.value_and_grad(base)
        ev = eng.evolve(base)
        res[dt] = dict(loss=out['loss'].clone(), reg=out['reg_loss'].clone(), grad=out['grad'].clone(), U=ev['U_final'].clone(),
                       iv=ev['inter_vecs'].clone(), us=out['unitary_scale'].clone())
        eng.poll_error()
        eng.close()
    a, b = res['f64'], res['f16x2']
    assert torch.isfinite(b['U']).all() and torch.isfinite(b['grad']).all()
    assert (a['iv'] - b['iv']).abs().max().item() < 2e-6 + 3e-7 * T
    assert (a['loss'] - b['loss']).abs().max().item() < 2e-6 + 2e-7 * T
    assert (a['reg'] - b['reg']).abs().max().item() < (2e-6 + 2e-7 * T) * max(1.0, a['reg'].abs().max().item())
    assert (a['grad'] - b['grad']).abs().max().item() < 1e-4 * a['grad'].abs().max().item()
    assert (a['U'] - b['U']).flatten(1).norm(dim=1).max().item() < 1e-4 * np.sqrt(n * max(T, 10) / 10.0)
    assert (a['us'] - b['us']).abs().max().item() < 2e-4


def test_f16x2_rejects_what_it_does_not_cover(built_lib):
    from quantum_optimal_control.core.engine import GrapeEngine, QocError
    with pytest.raises(QocError):
        GrapeEngine(40, 2, 10, 12, 1, 6, 2, dtype='f16x2')          # m > 8: dense-state problems stay on the fp64 path
    with pytest.raises(QocError):
        GrapeEngine(300, 2, 10, 2, 1, 6, 2, dtype='f16x2')          # n > 256


def test_full_size_c3_f16x2_properties(built_lib):
    """BASELINE config C3 at full size (n=36, T=1000, B=1024, forbidden-state regulariser) on the fp32-class path:
    unitarity of U_final, loss consistent with the propagated states, batch-permutation equivariance (bit-exact),
    instance 0 against the fp64 oracle at fp32-class tolerance."""
    pb = W.c3_two_transmon_cnot()
    B = 1024
    setups, guess, args, kw = make_case(pb, seed=300, B=1)
    guess = W.random_guess(4, 1000, pb['maxA'], 300, B=B)
    sp, eng = _engine(args, kw, guess, 'f16x2')
    base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
    out = {k: v.clone() for k, v in eng.value_and_grad(base).items()}
    ev = eng.evolve(base, want_inter_vecs=True)
    eng.poll_error()
    U = ev['U_final']
    eye = torch.eye(36, dtype=torch.complex128, device='cuda')
    assert (U.conj().transpose(1, 2) @ U - eye).abs().amax().item() < 5e-3
    phi = torch.from_numpy(sp.target_vectors_c).cuda()
    o = (phi.conj()[None] * ev['inter_vecs'][:, -1]).sum(dim=(1, 2))
    assert ((1 - o.abs() ** 2 / 16) - out['loss']).abs().max().item() < 1e-10          # fp64 reduction of the fp64 states
    perm = torch.randperm(B, device='cuda')
    outp = eng.value_and_grad(base[perm].contiguous())
    assert torch.equal(outp['grad'], out['grad'][perm]) and torch.equal(outp['loss'], out['loss'][perm])
    ref = O.costate_value_and_grad(setups[0], setups[0].ops_weight_base)
    assert abs(out['loss'][0].item() - ref['loss']) < 2e-5
    assert np.abs(out['grad'][0].cpu().numpy() - ref['grad']).max() < 5e-4 * np.abs(ref['grad']).max()
    assert np.abs(ev['inter_vecs'][0].cpu().numpy() - np.transpose(ref['inter_vecs'], (1, 0, 2))).max() < 2e-4
    assert np.linalg.norm(U[0].cpu().numpy() - ref['U_final']) < 5e-3
    eng.close()


# ------------------------------------------------------------------------------------------------------------------
# round-2 reference-run fixtures on the fp64 path: early stop with mixed stop times, L-BFGS-B, dressed + forbid_dressed,
# U0 != I, n = 64 and n = 216 with two seeds
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(golden_cases_r2()))
def test_fp64_path_matches_r2_reference_goldens(name, built_lib):
    from quantum_optimal_control.main_grape.grape import Grape
    pb, conv, method, g64, g32 = _load_case(name)
    args, kw = W.grape_kwargs(pb)
    B = g64['guess'].shape[0]
    sp, eng = _engine(args, kw, g64['guess'], 'f64')
    n = sp.state_num
    assert (sp.exp_terms, sp.scaling) == (int(g64['exp_terms']), int(g64['scaling']))
    base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
    out = eng.value_and_grad(base)
    ev = eng.evolve(base)
    for b in range(B):
        assert abs(out['loss'][b].item() - float(g64['eval_loss'][b])) < 1e-10
        assert abs(out['reg_loss'][b].item() - float(g64['eval_reg_loss'][b])) < 1e-10 * max(1, abs(float(g64['eval_reg_loss'][b])))
        assert np.abs(out['grad'][b].cpu().numpy() - g64['eval_grad'][b]).max() < 1e-9 * max(1.0, np.abs(g64['eval_grad'][b]).max())
        assert np.linalg.norm(ev['U_final'][b].cpu().numpy() - _complex_U(g64['eval_final_state'][b], n)) < 1e-9
        ivp = g64['eval_inter_vecs_packed'][b]
        assert np.abs(ev['inter_vecs'][b].cpu().numpy() - np.transpose(ivp[:n] + 1j * ivp[n:], (1, 2, 0))).max() < 1e-9
    eng.close()
    if name == 'c4_T50':
        return                                   # the optimiser run at n = 216 is covered by the evaluation above
    if method == 'Adam':
        # ONE batched run: every instance follows the reference's stop rules on its own (run_session.py:56-64)
        uks, Uf = Grape(*args, convergence=conv, initial_guess=g64['guess'], save=False, show_plots=False, quiet=True, **kw)
        for b in range(B):
            assert np.abs(uks[b] - g64['uks'][b]).max() < 1e-8, (name, b, int(g64['run_iterations'][b]))
            assert np.linalg.norm(Uf[b] - g64['U_final'][b]) < 1e-8
        if name == 'c1_earlystop':
            assert len(set(int(i) for i in g64['run_iterations'])) > 1          # the fixture really has mixed stop times
            assert all(int(i) < conv['max_iterations'] for i in g64['run_iterations'])
    else:
        uks, Uf = Grape(*args, convergence=conv, initial_guess=g64['guess'][0], save=False, show_plots=False, quiet=True,
                        method=method, **kw)
        assert np.abs(uks - g64['uks'][0]).max() < 1e-6
        assert np.linalg.norm(Uf - g64['U_final'][0]) < 1e-6


def test_engine_on_a_non_current_device_or_same_device_guard(built_lib):
    """GrapeEngine(device=X) must work whatever the current device is: with one GPU this exercises the
    guard path (device 0 explicitly), with several it runs on the last one while device 0 is current."""
    from quantum_optimal_control.core.problem import SystemParameters
    from quantum_optimal_control.core.engine import GrapeEngine
    ndev = torch.cuda.device_count()
    target = ndev - 1
    torch.cuda.set_device(0)
    pb = W.c1_pi_pulse()
    args, kw = W.grape_kwargs(pb)
    H0, Hops, Hn, U, tt, steps, scl = args
    guess = W.random_guess(2, 100, pb['maxA'], 3, B=2)
    sp = SystemParameters(H0, Hops, Hn, U, np.identity(2), tt, steps, scl, None, kw['maxA'], None, guess, False, 1e-4, False,
                          False, {}, False, None, None, True, True, False, False, False)
    eng = GrapeEngine.from_sys_para(sp, device=target)
    base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).to('cuda:%d' % target)
    with torch.cuda.device(target):
        pass
    out = eng.value_and_grad(base)
    eng.set_profiling(True)
    out = eng.value_and_grad(base)
    eng.kernel_times_ms()
    eng.poll_error()
    torch.cuda.synchronize(target)
    ref = O.graph_value_and_grad(O.make_setup(H0, Hops, U, tt, steps, scl, initial_guess=guess[0], **kw),
                                 sp.ops_weight_base[0])
    assert abs(out['loss'][0].item() - ref.loss) < 1e-10
    eng.close()
    assert torch.cuda.current_device() == 0
