"""Run the UNMODIFIED reference (``/root/reference/quantum_optimal_control``) in this container and
write golden fixtures to tests/golden/ -- TEST INFRASTRUCTURE, never imported by the product.

    python oracle/run_reference.py            # regenerates tests/golden/ref_*.npz
    python oracle/run_reference.py r2 [name]  # regenerates tests/golden/ref2_*.npz (round-2 cases)

The reference is Python-2.7 / TensorFlow-1 source.  Nothing is copied or edited on disk: each
module is read from /root/reference at import time, passed through the purely syntactic py2->py3
fixes below, and executed with ``oracle/tf1_shim.py`` standing in for ``tensorflow`` (and empty
stubs for h5py / matplotlib / IPython, which the hot path never touches with save=False,
show_plots=False).  Every line of system_parameters.py, tensorflow_state.py,
regularization_functions.py, run_session.py, analysis.py and grape.py that executes is the
reference's own.

Syntactic fixes applied in memory: tabs->8 spaces; ``print x`` statements -> ``print(x)``;
``xrange`` -> ``range``; the three implicit relative imports -> absolute; ``len(a)/len(b)`` ->
``//`` inside np.reshape shapes (run_session.py:153,185); the 2-D ``x0`` handed to ``scipy.optimize.minimize``
(run_session.py:180) is flattened, which the SciPy of the reference's era did itself.

This script needs /root/reference and therefore only runs in the build container; the fixtures it
writes are committed and are what travels to the GPU box.
"""
import importlib.abc
import importlib.util
import io
import os
import re
import sys
import types
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("QOC_REFERENCE", "/root/reference")
PKG = "quantum_optimal_control"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import tf1_shim  # noqa: E402


def py2_to_py3(src):
    src = src.expandtabs(8)
    lines = src.split("\n")
    out = []
    i = 0
    while i < len(lines):
        line = lines[i]
        m = re.match(r"^(\s*)print\s+(?!\()(.*)$", line) or re.match(r"^(\s*)print\s+(\(.*\)\s*%.*)$", line) \
            or re.match(r"^(\s*)print\s+(\(.*)$", line)
        if m and not line.strip().startswith("print("):
            indent, rest = m.group(1), m.group(2)
            stmt = rest
            while stmt.count("(") > stmt.count(")") and i + 1 < len(lines):
                i += 1
                stmt += "\n" + lines[i]
            out.append("%sprint(%s)" % (indent, stmt))
        else:
            out.append(line)
        i += 1
    src = "\n".join(out)
    src = re.sub(r"\bxrange\(", "range(", src)
    src = src.replace("from analysis import Analysis", "from %s.core.analysis import Analysis" % PKG)
    src = src.replace("from regularization_functions import get_reg_loss",
                      "from %s.core.regularization_functions import get_reg_loss" % PKG)
    src = src.replace("len(x)/len(self.sys_para.ops_c)", "len(x)//len(self.sys_para.ops_c)")
    src = src.replace("len(res['x'])/len(self.sys_para.ops_c)", "len(res['x'])//len(self.sys_para.ops_c)")
    # the SciPy of the reference's era flattened a 2-D x0 itself (lbfgsb.py: x0 = asarray(x0).ravel()); today's raises
    head, sep, tail = src.partition("def bfgs_optimize")
    src = head + sep + tail.replace("x0 = self.sys_para.ops_weight_base\n", "x0 = np.reshape(self.sys_para.ops_weight_base, -1)\n", 1)
    return src


class _RefFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Imports quantum_optimal_control.* straight from the read-only reference tree; package
    __init__ files (py2 star-imports of everything incl. qutip) are replaced by empty packages."""

    def find_spec(self, fullname, path, target=None):
        if fullname != PKG and not fullname.startswith(PKG + "."):
            return None
        rel = fullname.split(".")[1:]
        d = os.path.join(REF, PKG, *rel)
        if os.path.isdir(d):
            return importlib.util.spec_from_loader(fullname, self, is_package=True)
        if os.path.isfile(d + ".py"):
            return importlib.util.spec_from_loader(fullname, self)
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        rel = module.__name__.split(".")[1:]
        d = os.path.join(REF, PKG, *rel)
        if os.path.isdir(d):
            module.__path__ = [d]
            return
        with open(d + ".py") as f:
            src = py2_to_py3(f.read())
        module.__file__ = d + ".py"
        exec(compile(src, d + ".py", "exec"), module.__dict__)


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def install_reference():
    """Make ``from quantum_optimal_control.main_grape.grape import Grape`` import the reference."""
    for k in [k for k in sys.modules if k == PKG or k.startswith(PKG + ".")]:
        del sys.modules[k]
    sys.path[:] = [p for p in sys.path if not p.rstrip("/").endswith("quantum-optimal-control_b200")]
    tf1_shim.install()
    _stub("h5py", File=type("File", (), {}))
    plt = _stub("matplotlib.pyplot")
    _stub("matplotlib", pyplot=plt)
    _stub("matplotlib.gridspec")
    disp = _stub("IPython.display")
    _stub("IPython", display=disp)
    if not any(isinstance(f, _RefFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _RefFinder())
    from quantum_optimal_control.main_grape.grape import Grape
    import quantum_optimal_control.core.run_session as rs
    # reference bug: display() reads self.elapsed, which only save_data() sets when save=True
    # (core/run_session.py:129-148); give the class a default so save=False, show_plots=False runs.
    rs.run_session.elapsed = 0.0
    return Grape


def run_reference_case(problem, guess, convergence, method='Adam', dtype='float32', capture_eval=True):
    """Run the reference's Grape() and also fetch one ``get_error`` evaluation (loss, reg_loss, grad,
    unitary_scale, grad_squared, final_state, inter_vecs) at the initial weights."""
    import torch
    import workloads as W
    Grape = install_reference()
    tf1_shim.set_float(torch.float32 if dtype == 'float32' else torch.float64)
    args, kw = W.grape_kwargs(problem)
    # list-of-rows form: modern NumPy raises on ``ndarray != []`` (core/system_parameters.py:274), which
    # the NumPy of the reference's era evaluated to True; a list of row arrays takes the same branch.
    guess = [np.array(row, dtype=np.float64) for row in np.asarray(guess)]
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        if capture_eval:
            # method='EVOLVE' = run_session.get_error at x0 + get_end_results (run_session.py:33-37)
            tf1_shim.reset()
            import quantum_optimal_control.core.run_session as rs
            captured = {}
            orig = rs.run_session.get_end_results

            def spy(self):
                captured['l'], captured['rl'], captured['g'] = self.l, self.rl, np.array(self.grads)
                captured['metric'], captured['g2'] = self.metric, self.g_squared
                captured['inter_vecs'] = np.array(self.session.run(self.tfs.inter_vecs_packed))
                captured['final_state'] = np.array(self.session.run(self.tfs.final_state))
                captured['exp_terms'], captured['scaling'] = self.sys_para.exp_terms, self.sys_para.scaling
                return orig(self)

            rs.run_session.get_end_results = spy
            try:
                Grape(*args, convergence=convergence, initial_guess=guess, method='EVOLVE', save=False,
                      show_plots=False, use_gpu=False, **kw)
            finally:
                rs.run_session.get_end_results = orig
            K, T = len(problem['Hops']), problem['steps']
            out.update(eval_loss=captured['l'], eval_reg_loss=captured['rl'], eval_grad=captured['g'].reshape(K, T),
                       eval_unitary_scale=captured['metric'], eval_grad_squared=captured['g2'],
                       eval_inter_vecs_packed=captured['inter_vecs'], eval_final_state=captured['final_state'],
                       exp_terms=captured['exp_terms'], scaling=captured['scaling'])
        tf1_shim.reset()
        import quantum_optimal_control.core.run_session as rs2
        fin = {}
        orig2 = rs2.run_session.get_end_results

        def spy2(self):
            fin['iterations'], fin['loss'] = int(self.iterations), float(self.l)
            return orig2(self)

        rs2.run_session.get_end_results = spy2
        try:
            uks, Uf = Grape(*args, convergence=convergence, initial_guess=guess, method=method, save=False,
                            show_plots=False, use_gpu=False, **kw)
        finally:
            rs2.run_session.get_end_results = orig2
    out.update(uks=np.array(uks), U_final=np.array(Uf), run_iterations=fin.get('iterations', -1),
               run_final_loss=fin.get('loss', np.nan))
    return out


def golden_cases():
    import workloads as W
    regs_all = {'amplitude': 0.3, 'envelope': 0.7, 'dwdt': 0.02, 'd2wdt2': 0.0005, 'speed_up': 0.4,
                'forbidden_coeff_list': [3.0, 5.0, 2.0], 'states_forbidden_list': [2, 3, 2]}
    conv = lambda it: {'rate': 0.01, 'update_step': 10, 'max_iterations': it, 'conv_target': 1e-12,
                       'learning_rate_decay': 100}
    c2s = dict(W.c2_transmon_cavity(T=30), total_time=60.0)
    c2r = dict(c2s, reg_coeffs=regs_all)
    c3s = dict(W.c3_two_transmon_cnot(T=20), total_time=0.2)
    return {
        'c1_pi_pulse': (W.c1_pi_pulse(), 5, conv(25)),
        'c2_small': (c2s, 7, conv(10)),
        'c2_small_allregs': (c2r, 8, conv(8)),
        'c3_small_forbidden': (c3s, 9, conv(6)),
        'c5_n16': (W.c5_random(16, T=20), 10, conv(5)),
        'state_transfer_n8': (W.state_transfer_random(8, T=20, m=2), 11, conv(6)),
    }


def golden_cases_r2():
    """Round-2 fixtures: name -> dict(pb, seeds, conv, method).  Early stop with mixed stop times, the SciPy driver,
    dressed basis + forbid_dressed, U0 != I, and the sizes the fp32-class tcgen05 path (dtype 'f16x2') is checked at
    (n = 36 is ``c3_small_forbidden`` above; n = 64 and n = 216 here, two seeds each)."""
    import workloads as W
    sys.path.insert(0, os.path.join(ROOT, "quantum-optimal-control_b200"))
    conv = lambda it, tgt=1e-12, rate=0.01: {'rate': rate, 'update_step': 10, 'max_iterations': it, 'conv_target': tgt,
                                              'learning_rate_decay': 100}
    # dressed basis of a perturbed C2 drift (helper_functions/grape_functions.py:194-209 semantics, computed by NumPy)
    c2d = dict(W.c2_transmon_cavity(T=12), total_time=24.0)
    H0d = c2d['H0'] + 0.05 * (c2d['Hops'][0] + c2d['Hops'][2])
    from quantum_optimal_control.helper_functions.grape_functions import get_dressed_info      # our NumPy helper
    w_c, v_c, ids = get_dressed_info(H0d)
    c2d['H0'] = H0d
    c2d['dressed_info'] = {'eigenvectors': v_c, 'dressed_id': ids, 'eigenvalues': w_c, 'is_dressed': True}
    c2d['reg_coeffs'] = {'forbidden_coeff_list': [2.0, 1.0], 'states_forbidden_list': [20, 11], 'forbid_dressed': True}
    rng5, rng6 = np.random.default_rng(5), np.random.default_rng(6)
    U0 = np.linalg.qr(rng5.normal(size=(5, 5)) + 1j * rng6.normal(size=(5, 5)))[0]
    n5 = dict(W.c5_random(5, T=15), U0=U0, states_concerned_list=[1, 3])
    c64 = dict(W.c5_random(64, T=10), states_concerned_list=[0, 3, 17, 63])
    c4 = dict(W.c4_three_transmon_toffoli(T=50), total_time=2.5)
    return {
        'c1_earlystop': dict(pb=W.c1_pi_pulse(), seeds=[21, 22, 24], conv=conv(40, 0.3, 0.012), method='Adam'),
        'c1_lbfgs': dict(pb=W.c1_pi_pulse(), seeds=[31], conv=conv(12), method='L-BFGS-B'),
        'c2_dressed_forbid': dict(pb=c2d, seeds=[41], conv=conv(5), method='Adam'),
        'n5_U0': dict(pb=n5, seeds=[51], conv=conv(6), method='Adam'),
        'c5_n64_m4': dict(pb=c64, seeds=[61, 62], conv=conv(3), method='Adam'),
        'c4_T50': dict(pb=c4, seeds=[71, 72], conv=conv(2), method='Adam'),
    }


def write_r2(names=None):
    import workloads as W
    outdir = os.path.join(ROOT, "tests", "golden")
    for name, c in golden_cases_r2().items():
        if names and name not in names:
            continue
        pb = c['pb']
        K, T = len(pb['Hops']), pb['steps']
        for dtype in ('float64', 'float32'):
            rows = []
            for seed in c['seeds']:
                guess = W.random_guess(K, T, pb['maxA'], seed)
                res = run_reference_case(pb, guess, c['conv'], method=c['method'], dtype=dtype)
                res['guess'] = guess
                rows.append(res)
            keys = [k for k in rows[0] if k not in ('exp_terms', 'scaling')]
            data = {k: np.stack([np.asarray(r[k]) for r in rows]) for k in keys}
            path = os.path.join(outdir, "ref2_%s_%s.npz" % (name, dtype))
            np.savez_compressed(path, seeds=np.array(c['seeds']), exp_terms=rows[0]['exp_terms'], scaling=rows[0]['scaling'],
                                max_iterations=c['conv']['max_iterations'], **data)
            print("%-20s %-8s (p,s)=(%d,%d) its=%s loss0=%s -> %s" % (
                name, dtype, rows[0]['exp_terms'], rows[0]['scaling'], [int(r['run_iterations']) for r in rows],
                ["%.5g" % r['eval_loss'] for r in rows], os.path.relpath(path, ROOT)), flush=True)


def main():
    import workloads as W
    if len(sys.argv) > 1 and sys.argv[1] == 'r2':
        write_r2(sys.argv[2:] or None)
        return
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    for name, (pb, seed, conv) in golden_cases().items():
        K, T = len(pb['Hops']), pb['steps']
        guess = W.random_guess(K, T, pb['maxA'], seed)
        for dtype in ('float64', 'float32'):
            res = run_reference_case(pb, guess, conv, dtype=dtype)
            path = os.path.join(outdir, "ref_%s_%s.npz" % (name, dtype))
            np.savez_compressed(path, guess=guess, seed=seed, max_iterations=conv['max_iterations'], **res)
            print("%-22s %-8s (p,s)=(%d,%d) loss0=%.6g  |U_final|_F=%.6f  -> %s" % (
                name, dtype, res['exp_terms'], res['scaling'], res['eval_loss'], np.linalg.norm(res['U_final']),
                os.path.relpath(path, ROOT)))


if __name__ == "__main__":
    main()
