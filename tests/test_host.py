"""CPU tests of the host side: problem setup mirror vs the oracle's restatement, helper functions,
C-ABI library export table, error behaviour of the drop-in entry point, run files, and the
world_size-2 (gloo) sharding / all-gather logic of population sweeps.  No GPU compute."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import workloads as W
from oracle import grape_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sp(pb, guess, **over):
    from quantum_optimal_control.core.problem import SystemParameters
    args, kw = W.grape_kwargs(dict(pb, **over))
    H0, Hops, Hn, U, tt, steps, scl = args
    return SystemParameters(H0, Hops, Hn, U, kw.get('U0', np.identity(len(H0))), tt, steps, scl, kw.get('dressed_info'),
                            kw['maxA'], None, guess, False, kw.get('unitary_error', 1e-4), kw.get('state_transfer', False), False,
                            kw.get('reg_coeffs'), False, None, kw.get('Taylor_terms'), True, True, False, False, False), args, kw


@pytest.mark.parametrize("fn", [W.c1_pi_pulse, W.c2_transmon_cavity, W.c3_two_transmon_cnot, lambda: W.c5_random(16, T=50)])
def test_system_parameters_mirror_matches_oracle_setup(fn):
    pb = fn()
    K, T = len(pb['Hops']), pb['steps']
    guess = W.random_guess(K, T, pb['maxA'], 3)
    sp, args, kw = _sp(pb, guess)
    H0, Hops, Hn, U, tt, steps, scl = args
    st = O.make_setup(H0, Hops, U, tt, steps, scl, initial_guess=guess, **kw)
    assert (sp.exp_terms, sp.scaling) == (st.exp_terms, st.scaling)
    assert sp.dt == st.dt and sp.state_num == st.n
    assert np.array_equal(sp.matrix_list, st.matrix_list)
    assert np.array_equal(sp.one_minus_gauss, st.one_minus_gauss)
    assert np.array_equal(sp.ops_weight_base, st.ops_weight_base)
    assert np.array_equal(np.array(sp.initial_vectors), np.array(st.initial_vectors))
    assert np.array_equal(sp.initial_unitary, st.initial_unitary)
    n = sp.state_num
    V = np.array(st.initial_vectors)
    phi_ref = (st.target_unitary @ V.T)                       # tensorflow_state.py:165, real-embedded [2n,m]
    assert np.allclose(sp.target_vectors_c.T, phi_ref[:n] + 1j * phi_ref[n:], atol=0, rtol=0)


def test_batched_guess_and_errors():
    pb = W.c1_pi_pulse(T=12)
    g = W.random_guess(2, 12, pb['maxA'], 0, B=3)
    sp, *_ = _sp(pb, g)
    assert sp.batched and sp.batch_size == 3 and sp.ops_weight_base.shape == (3, 2, 12)
    assert np.allclose(np.sin(sp.ops_weight_base) * 2.0, g)
    with pytest.raises(ValueError):
        _sp(pb, np.full((2, 12), 2.5))
    _sp(pb, np.full((2, 12), -2.0))                             # one-sided check, like the reference


def test_random_initial_weights_follow_reference_distribution():
    pb = W.c1_pi_pulse(T=400)
    np.random.seed(7)
    sp, *_ = _sp(pb, None)
    np.random.seed(7)
    want = np.random.normal(0, 1 / np.sqrt(400), [2, 400])
    assert np.array_equal(sp.ops_weight_base, want)


def test_helper_functions():
    from quantum_optimal_control.helper_functions import grape_functions as G
    M = np.array([[1 + 2j, 3 - 1j], [0.5j, -2]])
    assert np.array_equal(G.c_to_r_mat(M), O.c_to_r_mat(M))
    assert np.array_equal(G.c_to_r_vec(M[0]), O.c_to_r_vec(M[0]))
    assert G.concerned(2, 6) == [0, 1, 6, 7]
    cnot = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]])
    Ug = G.transmon_gate(cnot, 3)
    assert Ug.shape == (9, 9) and Ug[3, 4] == 1 and Ug[4, 3] == 1 and Ug[3, 3] == 0 and Ug[8, 8] == 1
    assert np.allclose(G.qft(2) @ G.qft(2).conj().T, np.eye(4))
    assert np.allclose(G.Hadamard(1), np.array([[1, 1], [1, -1]]) / np.sqrt(2))
    assert G.Basis(7, 3, 6) == '011' and G.Bin(5, 4) == '0101'
    x = np.array([[0, 1], [1, 0]])
    assert np.array_equal(G.nn_chain_kron(x, np.eye(2), 3, 2), np.kron(np.kron(x, x), np.eye(2)) + np.kron(np.eye(2), np.kron(x, x)))
    Hops, names, amps = G.append_separate_krons(x, 'x', 2, 2, [], [], [])
    assert names == ['xi', 'ix'] and np.array_equal(Hops[1], np.kron(np.eye(2), x)) and amps == [4.0, 4.0]
    H0 = np.diag([0.0, 1.0, 2.0]) + 0.01 * (np.eye(3, k=1) + np.eye(3, k=-1))
    w, v, ids = G.get_dressed_info(H0)
    assert sorted(ids) == [0, 1, 2]
    assert np.array_equal(G.sort_ev(v, ids), O.sort_ev(v, ids))


def test_library_exports_every_declared_symbol(built_lib):
    import ctypes
    header = open(os.path.join(ROOT, "include", "qoc_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(qoc_[a-z_0-9]+)\s*\(", header)))
    lib = ctypes.CDLL(built_lib)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    from quantum_optimal_control.core.engine import SYMBOLS, load_library
    assert sorted(SYMBOLS) == declared
    from quantum_optimal_control.core.engine import QOC_ABI_VERSION
    assert load_library().qoc_abi_version() == QOC_ABI_VERSION == int(re.search(r"#define QOC_ABI_VERSION (\d+)", header).group(1))


def test_no_cpu_fallback(built_lib):
    """Without a GPU the product path must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from quantum_optimal_control.main_grape.grape import Grape
    from quantum_optimal_control.core.engine import QocError
    pb = W.c1_pi_pulse(T=10)
    args, kw = W.grape_kwargs(pb)
    with pytest.raises(QocError):
        Grape(*args, save=False, show_plots=False, quiet=True, **kw)
    with pytest.raises(NotImplementedError):
        Grape(*args, save=False, show_plots=False, use_gpu=False, **kw)


def test_grape_argument_errors():
    from quantum_optimal_control.main_grape.grape import Grape
    pb = W.c1_pi_pulse(T=10)
    args, kw = W.grape_kwargs(pb)
    with pytest.raises(ValueError, match="file_name"):
        Grape(*args, **kw)                                      # save=True is the default (grape.py:36-42)
    with pytest.raises(ValueError, match="data_path"):
        Grape(*args, file_name="x", **kw)
    with pytest.raises(KeyError):
        Grape(*args, freq_unit="kHz", save=False, **kw)         # grape.py:25-26 only knows "KHz"


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "quantum-optimal-control_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("oracle-free", ""), os.path.join(dp, f)


def test_reg_struct_and_storage(tmp_path):
    from quantum_optimal_control.core.engine import reg_struct
    from quantum_optimal_control.core import storage
    r, fw = reg_struct({'dwdt': 0.0, 'd2wdt2': 2.0, 'forbidden_coeff_list': [1.0, 2.0, 4.0], 'states_forbidden_list': [3, 1, 3]}, 5, 10)
    assert r.has_dwdt == 1 and r.dwdt == 0.0 and r.has_d2wdt2 == 1 and r.has_amplitude == 0 and r.has_forbidden == 1
    assert np.array_equal(fw, [0, 2, 0, 5, 0])
    with pytest.raises(NameError):
        reg_struct({'d2wdt2': 1.0}, 5, 10)
    with pytest.raises(ValueError):
        reg_struct({'bandpass': 1.0}, 5, 10)
    p1 = storage.new_run_file(str(tmp_path), "run")
    rf = storage.RunFile(p1)
    storage.save_inputs(rf, dict(H0=np.eye(2), steps=5, maxA=None), {'rate': 0.1}, {'dwdt': 1.0}, None)
    rf.add('wall_clock_time', 1.5)
    for it in range(3):                                         # H5File.append semantics: a new leading axis per save
        rf.append('error', np.array(0.5 / (it + 1)))
        rf.append('uks', np.full((2, 4), float(it)))
    p2 = storage.new_run_file(str(tmp_path), "run")
    assert os.path.basename(p1).startswith("00000_run") and os.path.basename(p2).startswith("00001_run")
    if p1.endswith(".npz"):
        with np.load(p1, allow_pickle=True) as f:
            assert f['error'].shape == (3,) and f['uks'].shape == (3, 2, 4) and f['uks'][2, 0, 0] == 2.0
            assert float(f['convergence/rate']) == 0.1 and float(f['reg_coeffs/dwdt']) == 1.0 and 'maxA' not in f.files


_GLOO = r'''
import os, sys
sys.path.insert(0, os.path.join(%(root)r, "quantum-optimal-control_b200")); sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from quantum_optimal_control.core import population as P
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
total = 5
lo, hi = P.shard_bounds(total, rank, 2)
losses_all = np.array([0.7, 0.3, 0.9, 0.05, 0.6])
K, T, n = 2, 4, 3
uks_all = np.arange(total * K * T, dtype=np.float64).reshape(total, K, T)
Uf_all = (np.arange(total * n * n).reshape(total, n, n) * (1 + 2j)).astype(np.complex128)
def fake_grape(*a, initial_guess=None, return_losses=False, **k):
    assert len(initial_guess) == hi - lo
    return uks_all[lo:hi], Uf_all[lo:hi], losses_all[lo:hi]
res = P.grape_population(fake_grape, (), np.zeros((total, K, T)))
assert res['best'] == 3, res['best']
assert np.array_equal(res['loss'], losses_all)
assert np.array_equal(res['uks'], uks_all[3]) and np.array_equal(res['U_final'], Uf_all[3])
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_population_sweep_world_size_2_gloo(tmp_path):
    from quantum_optimal_control.core.population import shard_bounds
    assert [shard_bounds(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    assert [shard_bounds(4096, r, 8) for r in (0, 7)] == [(0, 512), (3584, 4096)]
    script = tmp_path / "gloo_pop.py"
    script.write_text(_GLOO % dict(root=ROOT, port=29571 + os.getpid() % 200))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=120)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_host_adam_matches_tf1_formula():
    from quantum_optimal_control.core.optimizer import TF1AdamHost
    rng = np.random.default_rng(0)
    th = rng.normal(size=(3, 4, 5))
    th2 = th.copy()
    a, b = TF1AdamHost(th.shape), O.TF1Adam(th.shape)
    for i in range(6):
        g = rng.normal(size=th.shape)
        a.step(th, g, 0.01 * (i + 1))
        th2 = b.step(th2, g, 0.01 * (i + 1))
    assert np.abs(th - th2).max() < 1e-15


def test_qoc_adam_host_c_entry_point():
    """The fused C entry point itself: odd sizes, more threads than elements, zero elements, null pointers."""
    import ctypes as C
    from quantum_optimal_control.core.engine import load_library
    lib = load_library()
    rng = np.random.default_rng(3)
    for count, threads in ((1, 8), (7, 3), (1001, 4), (4096, 1)):
        th, g = rng.normal(size=count), rng.normal(size=count)
        m, v = rng.normal(size=count) * 0.1, np.abs(rng.normal(size=count)) * 0.01
        th0, m0, v0 = th.copy(), m.copy(), v.copy()
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        assert lib.qoc_adam_host(p(th), p(g), p(m), p(v), C.c_size_t(count), 0.02, 0.9, 0.999, 1e-8, threads) == 0
        m1 = 0.9 * m0 + (1 - 0.9) * g
        v1 = 0.999 * v0 + (1 - 0.999) * g * g
        assert np.abs(m - m1).max() < 1e-16 and np.abs(v - v1).max() < 1e-16
        assert np.abs(th - (th0 - 0.02 * m1 / (np.sqrt(v1) + 1e-8))).max() < 1e-15
    z = np.zeros(1)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.qoc_adam_host(p(z), p(z), p(z), p(z), C.c_size_t(0), 0.1, 0.9, 0.999, 1e-8, 4) == 0 and z[0] == 0.0
    assert lib.qoc_adam_host(None, p(z), p(z), p(z), C.c_size_t(1), 0.1, 0.9, 0.999, 1e-8, 1) == -1      # QOC_EINVAL


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU oracle in reference-cost mode; the only place outside tests/ and smoke()
    that executes oracle/) prints ONE JSON line with the keys the driver reads; shortened time grid so it takes seconds."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--steps-T", "20"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "instance-iterations/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["vs_baseline"] is None and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "T=20" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
