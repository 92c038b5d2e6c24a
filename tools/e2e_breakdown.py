"""Where the host-buffer (e2e) step spends its time: C-ABI host call vs host Adam (C2)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "quantum-optimal-control_b200")]
import numpy as np, torch, contextlib, io
import workloads as W
from quantum_optimal_control.core.problem import SystemParameters
from quantum_optimal_control.core.engine import GrapeEngine
from quantum_optimal_control.core.optimizer import TF1AdamHost, TF1AdamState

pb = W.c2_transmon_cavity(); B = 256
n, K, T = len(pb['H0']), len(pb['Hops']), pb['steps']
guess = W.random_guess(K, T, pb['maxA'], 0, B=B)
pargs, kw = W.grape_kwargs(pb)
H0, Hops, Hn, U, tt, steps, scl = pargs
with contextlib.redirect_stdout(io.StringIO()):
    sp = SystemParameters(H0, Hops, Hn, U, np.identity(n), tt, steps, scl, None, kw['maxA'], None, guess, False,
                          1e-4, False, False, None, False, None, None, True, True, False, False, False)
eng = GrapeEngine.from_sys_para(sp)
hb = eng.host_buffers()['base']; hb[...] = sp.ops_weight_base
adam = TF1AdamHost(hb.shape)
base = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).cuda()
for _ in range(3):
    o = eng.value_and_grad_host(hb, copy=False); adam.step(hb, o['grad'], 0.01)
def t(fn, reps=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
out = None
def dev():
    global out
    out = eng.value_and_grad(base, out=out)
print("device call        %.3f ms" % t(dev))
print("host call          %.3f ms" % t(lambda: eng.value_and_grad_host(hb, copy=False)))
o = eng.value_and_grad_host(hb, copy=False)
print("host Adam          %.3f ms (threads %d)" % (t(lambda: adam.step(hb, o['grad'], 0.01)), torch.get_num_threads()))
def loop():
    o = eng.value_and_grad_host(hb, copy=False); adam.step(hb, o['grad'], 0.01)
print("host call + Adam   %.3f ms (alternating, as in bench.py's e2e leg)" % t(loop))
for th in (1, 2, 4, 8, 16):
    adam.threads = th
    print("  Adam threads %2d   %.3f ms alternating" % (th, t(loop)))
x = torch.empty(B * K * T, dtype=torch.float64, device='cuda'); h = torch.empty(B * K * T, dtype=torch.float64).pin_memory()
print("H2D 4 MB           %.3f ms" % t(lambda: x.copy_(h, non_blocking=True)))
print("D2H 4 MB           %.3f ms" % t(lambda: h.copy_(x, non_blocking=True)))
