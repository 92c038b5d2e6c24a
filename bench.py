#!/usr/bin/env python
"""bench.py -- GRAPE iterations/s on BASELINE.json's config (C2 by default) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C1|C2|C3|C4|C5n8..C5n128] [--dtype f64|f16x2|tf32x3]
                    [--batch B] [--impl ours|reference] [--no-secondary] [--no-cpu-baseline]

The default line is BASELINE config C2 (configs[1], fp64); at N = 1 it also carries ``config.secondary``: short runs of
every other BASELINE configuration at its stated size (C3 and C4 on the fp32-class tcgen05 path, the C5 sizes on fp64).

A "step" is one GRAPE optimiser iteration of the whole batch: one fwd+bwd (value_and_grad over all
T time steps for all B instances) plus the TF1-form Adam update.  ``value`` counts
instance-iterations/s (B * N iterations of ONE problem instance per step), which is the unit the
reference arm can be timed in too (it has no batch dimension: one instance per Grape() call).

Keys follow the driver contract: value = device-resident throughput (CUDA events, max over ranks);
e2e = the same step through the host-buffer C-ABI call (pinned H2D of the weights, D2H of gradient
and losses, Adam on the host); roofline = the dominant kernel (k_expm) against the FP64 peak;
cpu_baseline = the CPU oracle in reference-cost mode on a bounded sample.
"""
import argparse
import json
import os

# The e2e leg alternates a ~7 ms GPU wait with a short multi-threaded host update (qoc_adam_host): with libgomp's
# default passive wait policy the worker threads are asleep by then and waking them costs more than the update itself
# (tools/e2e_breakdown.py: 7.94 ms per C2 step passive / 4 threads vs 7.39 ms active / 16 threads).  Must be set before
# the OpenMP runtime is loaded; an explicit setting in the environment wins.  Single-process runs only: with one rank
# per GPU the ranks share the host's cores and spinning workers would fight each other (measured passive there).
if int(os.environ.get("WORLD_SIZE", "1")) == 1:
    os.environ.setdefault("OMP_WAIT_POLICY", "ACTIVE")
elif os.environ.get("QOC_BENCH_PIN", "1") != "0" and hasattr(os, "sched_setaffinity"):
    # one rank per GPU: give every rank its own slice of the host cores BEFORE any thread pool exists (the pools inherit
    # the mask), so the ranks' host legs (qoc_adam_host workers, the copy threads) stop migrating onto each other; with
    # dedicated cores the workers can spin through the GPU wait as in the single-process run
    try:
        _cores = sorted(os.sched_getaffinity(0))
        _lw = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))
        _lr = int(os.environ.get("LOCAL_RANK", "0")) % _lw
        _mine = None
        try:                                      # NUMA-aware: the cores next to this rank's GPU, shared out among the GPUs of that node
            import pynvml
            pynvml.nvmlInit()
            _words = (max(_cores) + 64) // 64
            _masks = []
            _vis = [int(x) for x in os.environ["CUDA_VISIBLE_DEVICES"].split(",")] if os.environ.get("CUDA_VISIBLE_DEVICES", "").replace(",", "").isdigit() else list(range(_lw))
            for _g in range(_lw):
                _m = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(_vis[_g]), _words)
                _masks.append(tuple(c for c in _cores if (_m[c // 64] >> (c % 64)) & 1))
            _same = [g for g in range(_lw) if _masks[g] == _masks[_lr]]
            _near = _masks[_lr]
            if len(_near) // len(_same) >= 2:
                _k, _p = _same.index(_lr), len(_near) // len(_same)
                _mine = list(_near[_k * _p:(_k + 1) * _p])
        except Exception:  # noqa: BLE001
            _mine = None
        if _mine is None:
            _per = len(_cores) // _lw
            if _per >= 2:
                _mine = _cores[_lr * _per:(_lr + 1) * _per]
        if _mine:
            os.sched_setaffinity(0, _mine)
            os.environ.setdefault("OMP_WAIT_POLICY", "ACTIVE")
            os.environ.setdefault("OMP_PROC_BIND", "false")
    except OSError:
        pass
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "quantum-optimal-control_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "GRAPE iterations/sec (fwd+bwd over T steps + Adam), summed over the batch of instances"
UNIT = "instance-iterations/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--batch", type=int, default=None, help="instances per GPU (default: the config's)")
    ap.add_argument("--steps-T", type=int, default=None, help="override the number of time steps (debug only)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "tf32x3", "f16x2"], help="arithmetic of the propagator stage")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short runs of the other BASELINE configs (config.secondary)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    return ap.parse_args()


def make_problem(name, T=None):
    import workloads as W
    fn, meta = W.WORKLOADS[name]
    pb = fn() if T is None else fn(T=T)
    if T is not None:
        full = fn()
        pb['total_time'] = full['total_time'] * T / full['steps']      # same dt
        if 'Taylor_terms' not in pb:
            pass
    return pb, dict(meta)


def flops_alg(n, T, m, K, p, s):
    """algorithmic real flops per instance per iteration."""
    return 8.0 * n ** 3 * T * (p + s) + 8.0 * n ** 2 * m * T * (K + 1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop, self._th = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([x.strip() for x in line.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:  # noqa: BLE001
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the CPU oracle in reference-cost mode (fp32 real-embedded graph,
# TWO fwd+bwd per optimiser iteration, backward re-computes every propagator) -- the only place
# bench.py executes oracle/.
# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(pb, steps, warmup, budget_s=None, T_sample=None):
    import torch
    import workloads as W
    from oracle import grape_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    K, T = len(pb['Hops']), pb['steps']
    Ts = T if T_sample is None else T_sample
    pbs = dict(pb)
    if Ts != T:
        pbs['steps'] = Ts
        pbs['total_time'] = pb['total_time'] * Ts / T
    args, kw = W.grape_kwargs(pbs)
    H0, Hops, Hn, U, tt, st, scl = args
    guess = W.random_guess(K, Ts, pb['maxA'], 0)
    setup = O.make_setup(H0, Hops, U, tt, st, scl, initial_guess=guess, **kw)
    base = np.asarray(setup.ops_weight_base, dtype=np.float32)
    adam = O.TF1Adam(base.shape, dtype=np.float32)

    def one_iteration(base):
        O.graph_value_and_grad(setup, base, torch.float32)            # run_session.py:53-54
        out = O.graph_value_and_grad(setup, base, torch.float32)      # :69 runs the graph again
        return adam.step(base, out.grad, 0.01)

    for _ in range(warmup):
        base = one_iteration(base)
    times = []
    t_start = time.perf_counter()
    for i in range(steps):
        t0 = time.perf_counter()
        base = one_iteration(base)
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_start > budget_s and i >= 1:
            break
    per_iter = float(np.mean(times)) * (T / Ts)                        # cost is linear in T
    return 1.0 / per_iter, cores, len(times), Ts, per_iter


def run_reference(args):
    """Reference arm: the CPU port of the reference graph on ALL host cores.  Two ways of using the cores are timed and the
    better one is the line's value: (a) one instance, every intra-op thread on its 2n x 2n matmuls (what a single
    Grape() call does under TensorFlow); (b) one instance per core, each on one thread (sequential Grape() calls spread
    over the cores) -- for the small matrices of C1-C3 (b) is ~10x higher."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pb, meta = make_problem(args.workload, args.steps_T)
    n, K, T, m = len(pb['H0']), len(pb['Hops']), pb['steps'], len(pb['states_concerned_list'])
    rate, cores, done, Ts, per_iter = cpu_reference_rate(pb, args.steps, min(args.warmup, 1), budget_s=90.0)
    sample = "1 instance x %d reference-style Adam iterations (2 fwd+bwd each, fp32 real-embedded 2n x 2n, T=%d), %d intra-op threads" % (done, Ts, cores)
    extra = cpu_extra_baselines(args.workload, pb, seconds=20.0)
    pc = extra.get("per_core", {})
    mode = "intra-op threads"
    if pc.get("value", 0.0) > rate:
        rate, per_iter, sample, mode = pc["value"], 1.0 / pc["value"], pc["sample"], "one instance per core"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": min(args.warmup, 1), "ms_per_step": per_iter * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: n=%d K=%d T=%d m=%d, one instance per Grape() call (the reference has no batch dim)" % (
            args.workload, n, K, T, m), "cores_used_as": mode,
            "note": "CPU oracle port in reference-cost mode; TF1/py2 reference cannot run here"},
        "cpu_baseline": dict({"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}, **extra),
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def measured_peak(torch, dev, dtype):
    """Tensor-pipe denominator of the roofline: MEASURED_PEAKS.json (driver-written) where it has the entry, else a
    cuBLAS GEMM timed here.  f64 -> DGEMM, tf32x3 -> TF32 GEMM, f16x2 -> the file's bf16 figure (kind::f16 and bf16 share
    the tensor pipe rate; sustained figure: the expm kernel runs for >= 0.1 s per launch on the big configs)."""
    if dtype == "f16x2":
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            return float(mp["bf16_tflops_sustained"]), "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)", float(mp["bf16_tflops"])
        except Exception:  # noqa: BLE001
            return 1400.0, "fallback 1.4 PFLOP/s sustained bf16 (B200_PROFILING.md; MEASURED_PEAKS.json absent: of fallback)", 1590.0
    tf32 = dtype == "tf32x3"
    N, tdt = (8192, torch.float32) if tf32 else (4096, torch.float64)
    old_flag = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(N, N, device=dev, dtype=tdt)
    bb = torch.randn(N, N, device=dev, dtype=tdt)
    for _ in range(2):
        a @ bb
    best = 1e9
    for _ in range(5):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); a @ bb; s1.record(); torch.cuda.synchronize()
        best = min(best, s0.elapsed_time(s1))
    torch.backends.cuda.matmul.allow_tf32 = old_flag
    del a, bb
    src = ("cuBLAS %s %d^3 via torch.matmul, best of 5, measured in this run (MEASURED_PEAKS.json has no %s entry)"
           % ("TF32 GEMM" if tf32 else "DGEMM", N, "tf32" if tf32 else "fp64"))
    return 2 * N ** 3 / best / 1e9, src, None


def measure(workload, dtype, B, steps, warmup, dev, local, world, rank, steps_T=None, e2e=True, want_peak=True):
    """One configuration on this rank's GPU: device-resident loop (value), host-buffer loop (e2e), per-stage CUDA events,
    roofline of the propagator kernel.  Returns a dict (rank-local times; the caller reduces over ranks)."""
    import torch
    import torch.distributed as dist
    import workloads as W
    from quantum_optimal_control.core.problem import SystemParameters
    from quantum_optimal_control.core.engine import GrapeEngine
    from quantum_optimal_control.core.optimizer import TF1AdamState, TF1AdamHost as HostAdam

    pb, meta = make_problem(workload, steps_T)
    B = B or meta['B']
    n, K, T, m = len(pb['H0']), len(pb['Hops']), pb['steps'], len(pb['states_concerned_list'])
    guess = W.random_guess(K, T, pb['maxA'], 1000 * rank, B=B)           # each rank: its own B seeds (weak scaling)
    pargs, kw = W.grape_kwargs(pb)
    H0, Hops, Hn, U, tt, nsteps, scl = pargs
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        sp = SystemParameters(H0, Hops, Hn, U, np.identity(n), tt, nsteps, scl, None, kw['maxA'], None, guess, False,
                              kw.get('unitary_error', 1e-4), False, False, kw.get('reg_coeffs'), False, None,
                              kw.get('Taylor_terms'), True, True, False, False, False)
    eng = GrapeEngine.from_sys_para(sp, device=dev, dtype=dtype)
    p, s = sp.exp_terms, sp.scaling
    base0 = torch.from_numpy(np.ascontiguousarray(sp.ops_weight_base)).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident loop: value --------------------------------------------------------
    base = base0.clone()
    adam = TF1AdamState(base)
    out = None
    lr = 0.01

    def step():
        nonlocal out
        out = eng.value_and_grad(base, out=out)
        adam.step(base, out['grad'], lr)

    for _ in range(warmup):
        step()
    eng.set_profiling(True)
    ktimes = {k: 0.0 for k in eng.KERNELS}
    launches0 = eng.launch_count
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        sampled = 0
        for i in range(steps):
            step()
            if i % 8 == 0 or i == steps - 1:                # per-kernel events are read (one sync) on a subset of steps
                for k, v in eng.kernel_times_ms().items():
                    ktimes[k] += v
                sampled += 1
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count - launches0
    eng.set_profiling(False)
    eng.poll_error()
    res = dict(workload=workload, dtype=dtype, B=B, n=n, K=K, T=T, m=m, p=p, s=s, ms=ms, steps=steps, launches=int(launches),
               clocks=clk.summary(), final_loss=float(out['loss'].min().item()), losses=out['loss'].clone(),
               batch_chunk=eng.batch_chunk)

    # ---- end to end through the host-buffer C-ABI call ------------------------------------------
    if e2e:
        ncores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else max(1, (os.cpu_count() or 1) // max(world, 1))
        torch.set_num_threads(max(1, ncores if world > 1 else (os.cpu_count() or 1)))      # host Adam: this rank's cores
        hbase = eng.host_buffers()['base']                      # pinned host weights, updated in place by the host Adam
        hbase[...] = np.asarray(sp.ops_weight_base, dtype=np.float64)
        hadam = HostAdam(hbase.shape, threads=max(1, min(16 if world == 1 else 4, ncores if world > 1 else (os.cpu_count() or 1))))
        for _ in range(max(1, min(warmup, 5))):
            o = eng.value_and_grad_host(hbase, copy=False)
            hadam.step(hbase, o['grad'], lr)
        barrier()
        e2e_steps = max(2, min(steps, 50))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            o = eng.value_and_grad_host(hbase, copy=False)      # pinned H2D of the weights, kernels, D2H of grad + losses
            hadam.step(hbase, o['grad'], lr)
            _ = float(o['loss'][0])
        torch.cuda.synchronize()
        res.update(e2e_s=time.perf_counter() - t0, e2e_steps=e2e_steps, h2d=int(hbase.nbytes), d2h=int(hbase.nbytes + 4 * B * 8))

    # ---- roofline of the dominant kernel (the propagator stage) ----------------------------------
    ktimes = {k: v * steps / max(sampled, 1) for k, v in ktimes.items()}      # scale the sampled sums to all steps
    if eng.batch_chunk and eng.batch_chunk < B:      # the per-kernel events of a step cover its LAST workspace-sized pass only
        last = B % eng.batch_chunk or eng.batch_chunk                        # passes: chunk, chunk, ..., remainder
        ktimes = {k: v * (B / float(last)) for k, v in ktimes.items()}
    expm_ms = ktimes['expm'] / steps
    expm_flops = 8.0 * n ** 3 * (p - 1 + s) * T * B                    # (p-1) Taylor products + s squarings per (b,t)
    achieved = expm_flops / (expm_ms * 1e-3) / 1e12 if expm_ms > 0 else None
    step_flops = flops_alg(n, T, m, K, p, s) * B
    ps_products = (1 + p // 2 - (1 if p % 2 == 0 else 0)) if p >= 2 else 0          # Paterson-Stockmeyer product count
    hermitian = all(np.allclose(h, np.conj(np.transpose(h))) for h in [H0] + list(Hops))
    peak = peak_src = peak_burst = None
    if want_peak and rank == 0:
        peak, peak_src, peak_burst = measured_peak(torch, dev, dtype)
    if dtype == "f16x2":
        n16 = (n + 15) // 16 * 16
        rows = 128 * (2 if n > 128 else 1)
        kpad = n16 if n <= 64 else (n + 31) // 32 * 32      # k-steps of 16 (shared-memory-resident operands) / stages of 32
        nprod = max(1, ps_products) + s                                 # products issued per (b,t)
        executed = 2.0 * rows * n16 * kpad * 12 * nprod * T * B         # 12 real MMAs (4 real products x 3 half-pairs) per k-step
        if n <= 64:
            ek = "k_tc_small_expm (tcgen05 kind::f16, shared-memory-resident fp16-pair operands)"
            xk = "k_segprod<plane sets> + k_chain_mma (DMMA, fp64 segment matrices)"
        elif n > 128:
            ek = "k_tc_pair_expm (tcgen05 cta_group::2 kind::f16, TMA-fed fp16-pair operands, TMA-store epilogue)"
            xk = "k_tc_prog<SEG> + k_tc_prog<CHAIN>"
        else:
            ek = "k_tc_prog<EXPM> (tcgen05 kind::f16, TMA-fed, fp16-pair operands)"
            xk = "k_tc_prog<SEG> + k_tc_prog<CHAIN>"
        kname = ek
        pipe = "fp16 tensor pipe via tcgen05 (3 MMAs per real product: h0 h0 + h0 h1 + h1 h0), fp32 accumulators in TMEM"
        kernels = ("f16x2: expm = " + ek.split(" ")[0] + "; chain = k_plane_sweep<fwd> (states, fp64) on the handle's high-priority "
                   "stream while " + xk + " (U_final, unitary_scale) run on the caller's stream; costate = "
                   "k_plane_sweep<rev>; grad / fwd_reduce / finalize as on the fp64 path")
    else:
        np_pad = 32 if dtype == "tf32x3" else (n + 7) // 8 * 8
        if dtype == "f64" and hermitian and np_pad in (16, 32) and p >= 2:
            nblk = np_pad // 8                   # Hermitian-structure path: p//2 + 1 products on the upper block triangle
            taylor_equiv = (p // 2 + 1) * (nblk * (nblk + 1) / 2.0) / (nblk * nblk)
        else:
            taylor_equiv = ps_products
        # fp64: complex products are evaluated with 3 real DMMA products (Gauss / 3M) instead of 4 -> 6 n^3 issued flops
        executed = (8.0 if dtype == "tf32x3" else 6.0) * np_pad ** 3 * (taylor_equiv + s) * T * B * (3 if dtype == "tf32x3" else 1)
        kname = "k_expm_tc32 (tcgen05)" if dtype == "tf32x3" else ("k_expm_mma (DMMA)" if n <= 64 else "k_expm_large (DMMA, tiled)")
        pipe = ("tf32 tensor pipe via tcgen05 (3 MMAs per product: 3xTF32 operand split)" if dtype == "tf32x3"
                else "fp64 (tcgen05 has no f64 kind; bound is the FP64 FMA/DMMA pipe)")
        kernels = ("f64, few concerned states (m < NP/2): expm = k_expm_mma; chain = k_vec_sweep<fwd> (states, TMA ring) on the "
                   "handle's high-priority stream while k_segprod + k_chain_mma (U_final, unitary_scale) run on the caller's "
                   "stream; costate = k_vec_sweep<rev>; finalize includes the join of the two branches")
    traffic = tr_src = None              # dram bytes per launch from the committed ncu capture of the SAME workload, else null
    try:
        trj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        key = "%s/%s" % (workload, dtype)
        if key in trj and steps_T is None and B == meta['B']:
            traffic = trj[key]["dram_bytes_read"] + trj[key]["dram_bytes_write"]
            tr_src = "profiles/r02_traffic.json[%s] (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, per launch)" % key
    except Exception:  # noqa: BLE001
        pass
    res['roofline'] = {
        "kernel": kname, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": (achieved / peak) if (achieved and peak) else None, "traffic": traffic, "traffic_source": tr_src,
        "peak_source": peak_src, "peak_burst": peak_burst, "pipe": pipe,
        "executed_flops_per_launch": executed,
        "executed_frac_of_peak": (executed / (expm_ms * 1e-3) / 1e12 / peak) if (expm_ms > 0 and peak) else None,
        "note": "achieved uses SURVEY 8(d)'s algorithmic count (p-1+s products of 8n^3 per time step); the kernels evaluate the "
                "SAME polynomial with fewer products (Paterson-Stockmeyer); executed_* counts the flops actually issued on the "
                "padded tiles (fp64: 3M products and Hermitian triangle; tf32x3 / f16x2: the 3x operand split) and is the "
                "hardware-utilisation figure",
        "kernels": kernels, "alg_flops_per_launch": expm_flops, "avg_launch_ms": expm_ms,
        "kernel_ms_per_step": {k: v / steps for k, v in ktimes.items()},
        "whole_step_alg_tflops": step_flops / (ms / steps * 1e-3) / 1e12}
    res['pb'] = pb
    eng.close()
    del eng
    torch.cuda.empty_cache()
    return res


def _percore_worker(a):
    """One reference-style iteration stream on ONE thread (the per-core CPU baseline: one instance per core)."""
    wl, Ts, seconds = a
    import torch
    torch.set_num_threads(1)
    pb, _ = make_problem(wl)
    import workloads as W
    from oracle import grape_oracle as O
    K, T = len(pb['Hops']), pb['steps']
    pbs = dict(pb, steps=Ts, total_time=pb['total_time'] * Ts / T)
    args, kw = W.grape_kwargs(pbs)
    H0, Hops, Hn, U, tt, st, scl = args
    setup = O.make_setup(H0, Hops, U, tt, st, scl, initial_guess=W.random_guess(K, Ts, pb['maxA'], 0), **kw)
    base = np.asarray(setup.ops_weight_base, dtype=np.float32)
    O.graph_value_and_grad(setup, base, torch.float32)
    t0, it = time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds:
        O.graph_value_and_grad(setup, base, torch.float32)
        O.graph_value_and_grad(setup, base, torch.float32)
        it += 1
    return it, time.perf_counter() - t0


def cpu_extra_baselines(workload, pb, seconds=8.0):
    """(a) per-core -- one instance per core, each on one thread, all cores busy;
    (b) generous -- ONE fwd+bwd per iteration in the complex n x n costate form (NumPy complex128, all cores)."""
    import multiprocessing as mp
    from oracle import grape_oracle as O
    import workloads as W
    cores = os.cpu_count() or 1
    K, T = len(pb['Hops']), pb['steps']
    Ts = max(4, min(T, 20))
    out = {}
    try:
        with mp.get_context("spawn").Pool(cores) as pool:
            rs = pool.map(_percore_worker, [(workload, Ts, seconds)] * cores)
        rate = sum(it / dt for it, dt in rs) * Ts / T
        out["per_core"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                           "sample": "%d processes x 1 thread, reference-style iterations (2 fwd+bwd, fp32 real-embedded) at T=%d "
                                     "for %.0f s, scaled linearly to T=%d" % (cores, Ts, seconds, T)}
    except Exception as e:  # noqa: BLE001
        out["per_core"] = {"error": repr(e)[:200]}
    try:
        pbs = dict(pb, steps=Ts, total_time=pb['total_time'] * Ts / T)
        args, kw = W.grape_kwargs(pbs)
        H0, Hops, Hn, U, tt, st, scl = args
        setup = O.make_setup(H0, Hops, U, tt, st, scl, initial_guess=W.random_guess(K, Ts, pb['maxA'], 0), **kw)
        O.costate_value_and_grad(setup, setup.ops_weight_base)
        t0, it = time.perf_counter(), 0
        while time.perf_counter() - t0 < seconds:
            O.costate_value_and_grad(setup, setup.ops_weight_base)
            it += 1
        dt = (time.perf_counter() - t0) / it * (T / Ts)
        out["generous"] = {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": "port",
                           "sample": "1 instance, ONE fwd+bwd per iteration, complex128 n x n costate form (NumPy/BLAS, all "
                                     "cores), T=%d scaled to T=%d; %.3f s per iteration" % (Ts, T, dt)}
    except Exception as e:  # noqa: BLE001
        out["generous"] = {"error": repr(e)[:200]}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    r = measure(args.workload, args.dtype, args.batch, args.steps, args.warmup, dev, local, world, rank, steps_T=args.steps_T)
    B, n, K, T, m, p, s = r['B'], r['n'], r['K'], r['T'], r['m'], r['p'], r['s']
    t_ms = torch.tensor([r['ms'], r['e2e_s']], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max, e2e_max = float(t_ms[0].item()), float(t_ms[1].item())
    ms_per_step = ms_max / args.steps
    value = B * world * args.steps / (ms_max * 1e-3)
    e2e_value = B * world * r['e2e_steps'] / e2e_max

    # the ONE collective of a population sweep (core/population.py): all-gather of the per-instance losses, verified
    allgather = None
    if world > 1:
        from quantum_optimal_control.core.population import gather_losses
        allv = gather_losses(r['losses'], B * world)
        mine = allv[rank * B:(rank + 1) * B]
        ok = torch.tensor([1.0 if (allv.numel() == B * world and torch.equal(mine, r['losses'])) else 0.0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        chk = torch.tensor([float(r['losses'].sum().item())], dtype=torch.float64, device=dev)
        dist.all_reduce(chk, op=dist.ReduceOp.SUM)
        allgather = {"backend": dist.get_backend(), "elements": int(allv.numel()), "verified": bool(ok.item() == 1.0) and
                     abs(float(chk.item()) - float(allv.sum().item())) < 1e-9 * max(1.0, abs(float(chk.item()))),
                     "best_instance": int(torch.argmin(allv).item())}

    secondary = None
    if world > 1 and not args.no_secondary and args.workload == "C2" and args.steps_T is None and args.batch is None:
        # BASELINE configs[4] through the multi-GPU launch: the Hilbert-dimension sweep, 512 instances per GPU (4096 over 8 GPUs),
        # every rank runs its shard, the time is the max over ranks (no collective in the iteration)
        secondary = {}
        for wl, dt, st, wu in (("C5n8", "f64", 10, 2), ("C5n16", "f64", 10, 2), ("C5n32", "f64", 5, 1), ("C5n64", "f64", 3, 1),
                               ("C5n128", "f64", 1, 1)):
            key = "%s/%s" % (wl, dt)
            try:
                q = measure(wl, dt, None, st, wu, dev, local, world, rank, e2e=False, want_peak=False)
                t = torch.tensor([q['ms']], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms_max = float(t.item())
                secondary[key] = {
                    "workload": "%s: n=%d K=%d T=%d m=%d, B=%d per GPU x %d GPUs, (p,s)=(%d,%d), %s" % (wl, q['n'], q['K'], q['T'], q['m'], q['B'], world, q['p'], q['s'], dt),
                    "ms_per_step": ms_max / st, "value": world * q['B'] * st / (ms_max * 1e-3), "unit": UNIT, "steps": st, "warmup": wu,
                    "n_gpus": world, "batch_chunk": q['batch_chunk'], "clocks": q['clocks'], "gpu_launches": q['launches'],
                    "kernel_ms_per_step": q['roofline']["kernel_ms_per_step"]}
            except Exception as e:  # noqa: BLE001
                secondary[key] = {"error": repr(e)[:300]}
        if rank != 0:
            secondary = None
    if rank == 0 and world == 1 and not args.no_secondary and args.workload == "C2" and args.steps_T is None and args.batch is None:
        # every other BASELINE configuration at its stated size, short runs (device-resident, CUDA events, clocks sampled)
        secondary = {}
        for wl, dt, st, wu in (("C3", "f16x2", 3, 1), ("C3", "f64", 3, 1), ("C4", "f16x2", 2, 1), ("C5n64", "f64", 3, 1),
                               ("C5n32", "f64", 5, 1), ("C5n16", "f64", 10, 2), ("C5n8", "f64", 10, 2), ("C5n128", "f64", 1, 1)):
            try:
                q = measure(wl, dt, None, st, wu, dev, local, 1, 0, e2e=False)
                rf = q['roofline']
                secondary["%s/%s" % (wl, dt)] = {
                    "workload": "%s: n=%d K=%d T=%d m=%d B=%d, (p,s)=(%d,%d), %s" % (wl, q['n'], q['K'], q['T'], q['m'], q['B'], q['p'], q['s'], dt),
                    "ms_per_step": q['ms'] / st, "value": q['B'] * st / (q['ms'] * 1e-3), "unit": UNIT, "steps": st, "warmup": wu,
                    "batch_chunk": q['batch_chunk'], "clocks": q['clocks'], "gpu_launches": q['launches'],
                    "roofline": {k: rf[k] for k in ("kernel", "achieved", "peak", "frac", "executed_frac_of_peak", "peak_source",
                                                    "avg_launch_ms", "kernel_ms_per_step", "whole_step_alg_tflops")}}
            except Exception as e:  # noqa: BLE001
                secondary["%s/%s" % (wl, dt)] = {"error": repr(e)[:300]}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            rate, cores, done, Ts, per_iter = cpu_reference_rate(r['pb'], 50, 1, budget_s=args.cpu_seconds)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "1 instance x %d reference-style Adam iterations (2 fwd+bwd each, fp32 real-embedded, T=%d), "
                             "%d intra-op threads; %.3f s per iteration" % (done, Ts, cores, per_iter)}
            cpu.update(cpu_extra_baselines(args.workload, r['pb'], seconds=min(8.0, args.cpu_seconds)))
        dname = {"f64": "f64", "tf32x3": "tf32x3 (fp32 accumulate)", "f16x2": "f16x2 (fp16-pair operands, fp32 accumulate, fp64 states)"}[args.dtype]
        p_bytes = 16 if args.dtype == "f64" else 8
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dname, "data": "synthetic",
            "config": {"workload": "%s: n=%d K=%d T=%d m=%d B=%d per GPU, (p,s)=(%d,%d), %s" % (
                args.workload, n, K, T, m, B, p, s, args.dtype), "batch_iterations_per_s": args.steps / (ms_max * 1e-3),
                "l2": "inputs exceed L2: %.2f GB of propagators are rewritten and re-read every step" % (
                    B * T * n * n * p_bytes / 1e9), "final_loss_min": r['final_loss'], "secondary": secondary,
                "population_allgather": allgather},
            "clocks": r['clocks'],
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": r['h2d'], "d2h_bytes_per_step": r['d2h'],
                    "steps": r['e2e_steps']},
            "gpu_launches": r['launches'],
            "roofline": r['roofline'],
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
        try:                                         # the peaks measured in this run, next to the other profiles
            json.dump({"workload": args.workload, "dtype": args.dtype, "peak_tflops": r['roofline']['peak'],
                       "peak_source": r['roofline']['peak_source'], "n_gpus": world},
                      open(os.path.join(ROOT, "profiles", "bench_peaks_last_run.json"), "w"), indent=1)
        except Exception:  # noqa: BLE001
            pass
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
