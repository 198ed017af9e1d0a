#!/usr/bin/env python
"""Benchmark of the log-posterior hot path (driver contract: one JSON line on stdout).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json config 3, the BSM fit the north star's roofline target names): the batched
BSM dim-6 / texture-OET log-posterior -- 6 SM parameters + log10(Lambda) per point, 20 energy bins,
Gaussian flavor-ratio likelihood + (truncated-)Gaussian priors -- on a synthetic batch of
4096 walkers x 1024 independent chains = 2^22 parameter points per step and per GPU (theta is
235 MB, larger than the 126 MB L2, so every step streams its input from HBM).  One step = one pass of
the hot path over that batch = ONE kernel launch.

  value    : log-posterior evaluations / s over all ranks, theta resident in HBM (CUDA events on the
             launching stream, barrier + synchronize on both sides, max over ranks)
  e2e      : same metric through the public host API (`llh.LnProb.evaluate_host` -> C ABI
             `gf_lnprob_host`): pinned host theta -> chunked H2D -> kernel -> D2H of the results
  roofline : algorithmic fp64 FLOPs (9252 per evaluation) / kernel time, against the
             fp64 FMA peak MEASURED in the same run by a DFMA microbenchmark (MEASURED_PEAKS.json
             carries no fp64 figure); HBM traffic is reported alongside
  scan     : secondary section -- the sharded Monte-Carlo scan (config 4), whole-job samples / s
             including the NCCL all-reduce of the histograms
  cpu_baseline : the oracle's scalar float128 ln_prob (port of the reference path) on 1 host core

`--impl reference` times the reference's CPU algorithm (the oracle port; the reference is pure
Python and cannot travel to the GPU box) on all host cores for the same metric and config.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

FLOP_PER_EVAL = 9252          # F_BSM for D=7, G=5, 20 bins
BYTES_PER_EVAL = 64           # 7 doubles in, 1 out
WALKERS, CHAINS = 4096, 1024
METRIC = 'log-posterior evals/sec'
WORKLOAD = ('C3: BSM dim-6 operator, texture OET, 6 SM params + logLambda, 20 energy bins, Gaussian flavor-ratio LLH '
            '+ priors; 4096 walkers x 1024 chains = 4194304 points per step per GPU')


def build_problem():
    import models
    from golemflavor_b200.enums import Texture
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_llh.npz'))
    return models.bsm_model_c3(g['asimov_angles'], dim=6, texture=Texture.OET), models


def synth_theta(pset, n, seed, models):
    """Uniform inside Param.ranges for the SM block (SURVEY 8d 'throughput batches'), logLam across
    the dim-6 scale boundaries."""
    rng = np.random.default_rng(seed)
    return models.draw_in_ranges(pset, n, rng)


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """Samples SM clock, power and throttle reasons of one GPU through NVML every 5 ms from a
    background thread while the timed region runs (the main thread sits in a CUDA synchronize,
    which releases the GIL)."""
    BAD = {'hw_slowdown': 0x8, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20, 'sw_power_cap': 0x4,
           'hw_power_brake_slowdown': 0x80}

    def __init__(self, index):
        self.index, self.rows, self._stop, self.thread, self.err = index, [], threading.Event(), None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(visible.split(',')[index]) if visible and visible.split(',')[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception as exc:  # noqa: BLE001
            self.nv, self.err = None, repr(exc)

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.rows.append((sm, pw, rs))
            except Exception as exc:  # noqa: BLE001
                self.err = repr(exc)
                break
            time.sleep(0.005)

    def start(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        if self.thread is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['clock sampling unavailable: %s' % self.err]}
        self._stop.set()
        self.thread.join()
        sm = [r[0] for r in self.rows]
        power = [r[1] for r in self.rows]
        reasons = sorted(name for name, bit in self.BAD.items() if any(r[2] & bit for r in self.rows))
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_min_mhz': min(sm) if sm else None, 'sm_max_mhz': self.max_sm,
                'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': reasons}


# ---------------------------------------------------------------------------------------------- CPU arm
def _cpu_eval_chunk(job):
    """Scalar float128 ln_prob of the oracle (restatement of llh.py:121-130 + fr.py) on a chunk."""
    seed, count = job
    import models  # noqa: F401  (sys.path set at import)
    from oracle import golem_oracle as go
    (args, asimov, pset), _ = build_problem()
    theta = synth_theta(pset, count, seed, sys.modules['models'])
    t0 = time.perf_counter()
    out = []
    for t in theta:
        try:
            out.append(go.ln_prob(list(t), args, asimov, pset))
        except AssertionError:   # the reference's unitarity assertion (fr.py:493-498): same work was done
            out.append(np.nan)
    return time.perf_counter() - t0, count, float(np.sum(np.isfinite(out)))


def cpu_baseline(per_core, cores):
    jobs = [(1000 + c, per_core) for c in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_eval_chunk(jobs[0])]
    else:
        import multiprocessing as mp
        with mp.get_context('fork').Pool(cores) as pool:
            res = pool.map(_cpu_eval_chunk, jobs)
    wall = time.perf_counter() - t0
    total = sum(r[1] for r in res)
    compute = max(r[0] for r in res)
    return total / compute, total, wall


def run_reference(opts):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_core = 150
    for _ in range(max(0, min(opts.warmup, 1))):
        cpu_baseline(8, cores)
    vals = []
    t0 = time.perf_counter()
    steps = max(1, min(opts.steps, 3))
    for _ in range(steps):
        v, total, _ = cpu_baseline(per_core, cores)
        vals.append(v)
    elapsed = time.perf_counter() - t0
    value = float(np.mean(vals))
    sample = '{0} scalar float128 ln_prob evaluations per step ({1} per process x {2} processes), same model and synthetic theta ' \
             'distribution as the GPU arm'.format(per_core * cores, per_core, cores)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'evals/s', 'n_gpus': opts.gpus, 'steps': steps,
        'warmup': min(opts.warmup, 1), 'ms_per_step': 1e3 * elapsed / steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f80 (x87 long double, as the reference)', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'sample': sample},
        'cpu_baseline': {'value': value, 'unit': 'evals/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'evals/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------- GPU arm
def run_gpu(opts):
    import torch
    import torch.distributed as dist
    from golemflavor_b200 import _lib, llh, scan

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the GPU arm has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    try:  # keep this rank (and the pinned buffers it first-touches) on the CPUs / NUMA node next to its GPU
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get('CUDA_VISIBLE_DEVICES')
        phys = int(visible.split(',')[local]) if visible and visible.split(',')[local].isdigit() else local
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
    except Exception:  # noqa: BLE001  (affinity is an optimisation, never a requirement)
        pass
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    lib = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    (args, asimov, pset), models = build_problem()
    fn = llh.LnProb(args, asimov, pset)
    n = WALKERS * CHAINS
    theta_host = torch.as_tensor(synth_theta(pset, n, 25 + rank, models)).pin_memory()
    theta = theta_host.cuda()
    out = torch.empty(n, dtype=torch.float64, device='cuda')
    stream = _lib.stream_ptr(torch)

    def step():
        _lib.check(lib.gf_lnprob(fn.model.ref, _lib.ptr(theta), n, fn.ndim, 1, _lib.ptr(out), None, None, stream))

    # -- fp64 peak probe (same run, same clocks)
    sink = torch.zeros(8, dtype=torch.float64, device='cuda')
    import ctypes as C
    flops = C.c_double()
    peak = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.gf_fp64_peak_probe(0, 200000, _lib.ptr(sink), C.byref(flops), stream))
        e1.record()
        torch.cuda.synchronize()
        peak = max(peak, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)

    # -- device-resident throughput
    for _ in range(opts.warmup):
        step()
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    launches0 = lib.gf_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(opts.steps):
        step()
    e1.record()
    barrier()
    launches = lib.gf_launch_count() - launches0
    ms = max_over_ranks(e0.elapsed_time(e1))
    clock_info = clocks.stop() if rank == 0 else None
    value = world * n * opts.steps / (ms * 1e-3)
    kernel_ms = ms / opts.steps

    # -- end to end through the host API
    out_host = torch.empty(n, dtype=torch.float64).pin_memory()
    th_np, out_np = theta_host.numpy(), out_host.numpy()
    e2e_steps = max(3, min(opts.steps, 10))
    for _ in range(2):
        fn.evaluate_host(th_np, out=out_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fn.evaluate_host(th_np, out=out_np)     # returns after the results are in host memory
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * n * e2e_steps / e2e_s
    assert np.array_equal(out_np, out.cpu().numpy()), 'host pipeline and device path disagree'
    # context for e2e: the plain pinned-host -> device copy rate of the same theta buffer (the link ceiling)
    link = 0.0
    for _ in range(3):
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        theta.copy_(theta_host, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        link = max(link, theta_host.numel() * 8 / (c0.elapsed_time(c1) * 1e-3) / 1e9)
    finite_frac = float(np.isfinite(out_np).mean())

    # -- secondary: sharded Monte-Carlo scan with the histogram all-reduce (config 4)
    def _scan_section():
        scan_info = None
        if opts.scan_samples > 0:
            fm = scan.scan_model(opts.scan_mode, dimension=6)
            scan.scan_histogram(fm, 10 ** 7, nb=25, seed=26)      # warm-up (also NCCL channel set-up)
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            hist, kept = scan.scan_histogram(fm, opts.scan_samples, nb=25, seed=26, return_tensor=True)
            s1.record()
            barrier()
            scan_ms = max_over_ranks(s0.elapsed_time(s1))
            scan_info = {'mode': opts.scan_mode, 'samples': opts.scan_samples, 'nb': 25, 'seconds': scan_ms * 1e-3,
                         'samples_per_s': opts.scan_samples / (scan_ms * 1e-3), 'scaling': 'strong',
                         'kept': int(kept.item()), 'hist_checksum': int((hist.flatten() * torch.arange(hist.numel(), device='cuda') % 1000003).sum().item()),
                         'collective': 'one NCCL all-reduce (sum, int64) of the 26^3 histogram' if world > 1 else 'none (1 GPU)'}
        return scan_info

    try:
        scan_info = _scan_section()
    except Exception as exc:  # noqa: BLE001  (a secondary section must never cost the headline line)
        scan_info = {'error': repr(exc)}

    # -- secondary: the sampler-shaped configs of BASELINE.json (latency-bound by design: 512 / 2048 /
    #    18000 points per half-step), run on the device-resident ensemble sampler
    def _configs_section():
        cfg_info = None
        if opts.configs:
            import models as _m
            from golemflavor_b200 import mcmc, sens
            g = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_llh.npz'))
            cfg_info = {}

            def timed(fn_):
                barrier()
                t0_ = time.perf_counter()
                r_ = fn_()
                torch.cuda.synchronize()
                return max_over_ranks(time.perf_counter() - t0_), r_

            a2, as2, ps2 = _m.notebook_model(g['asimov_angles'])
            f2 = llh.LnProb(a2, as2, ps2)
            np.random.seed(25)
            p0 = mcmc.flat_seed(ps2, 1024)
            p0[:, 4], p0[:, 5] = np.random.uniform(.9, 1, 1024), np.random.uniform(.8, 1, 1024)
            # K1: the SM-only log-posterior (161 algorithmic FLOP / 56 B per point) is HBM-bound: report it against HBM
            n1 = 1 << 24
            th1 = torch.as_tensor(_m.draw_in_ranges(ps2, 1 << 20, np.random.default_rng(3))).cuda().repeat(16, 1)
            o1 = torch.empty(n1, dtype=torch.float64, device='cuda')
            k1 = lambda: _lib.check(lib.gf_lnprob(f2.model.ref, _lib.ptr(th1), n1, 6, 1, _lib.ptr(o1), None, None, stream))
            for _ in range(3):
                k1()
            torch.cuda.synchronize()
            k0e, k1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0e.record()
            for _ in range(20):
                k1()
            k1e.record()
            torch.cuda.synchronize()
            k1ms = k0e.elapsed_time(k1e) / 20
            cfg_info['K1_sm_lnprob'] = {'points': n1, 'ms': k1ms, 'evals_per_s': n1 / (k1ms * 1e-3), 'bound': 'hbm',
                                        'achieved_gbs': 56.0 * n1 / (k1ms * 1e-3) / 1e9,
                                        'note': '6-D notebook model (4 PMNS coords + 2 source angles), theta 805 MB > L2'}
            del th1, o1
            smp = mcmc.DeviceEnsembleSampler(1024, 6, f2, seed=25)
            smp.run_mcmc(p0, 200, store=False)
            sec, _ = timed(lambda: smp.run_mcmc(None, 10000, store=True, return_tensor=True))
            cfg_info['C2_emcee_sm_fit'] = {'walkers': 1024, 'steps': 10000, 'ndim': 6, 'seconds': sec, 'evals_per_s': 1024 * 1e4 / sec,
                                           'acceptance': float(np.mean(smp.acceptance_fraction)), 'launches': 1,
                                           'note': 'replicas only: every rank runs the same chain shape independently'}
            # C1: the reference's own CPU-sized case (3 raw source ratios, fixed NuFIT PMNS), here on the device sampler
            a1, as1, ps1 = _m.sm_fit_c1(g['asimov_angles'])
            f1 = llh.LnProb(a1, as1, ps1)
            p1 = mcmc.flat_seed(ps1, 100)
            smp1 = mcmc.DeviceEnsembleSampler(100, 3, f1, seed=25)
            smp1.run_mcmc(p1, 100, store=False)
            sec, _ = timed(lambda: smp1.run_mcmc(None, 1000, store=True, return_tensor=True))
            cfg_info['C1_sm_fit_fixed_pmns'] = {'walkers': 100, 'steps': 1000, 'ndim': 3, 'seconds': sec, 'evals_per_s': 100 * 1000 / sec,
                                                'acceptance': float(np.mean(smp1.acceptance_fraction)), 'launches': 1}
            p3 = mcmc.flat_seed(pset, 4096)
            smp3 = mcmc.DeviceEnsembleSampler(4096, fn.ndim, fn, seed=25)
            smp3.run_mcmc(p3, 100, store=False)
            sec, _ = timed(lambda: smp3.run_mcmc(None, 2000, store=False, return_tensor=True))
            cfg_info['C3_bsm_dim6_fit'] = {'walkers': 4096, 'steps': 2000, 'ndim': fn.ndim, 'seconds': sec, 'evals_per_s': 4096 * 2000 / sec,
                                           'acceptance': float(np.mean(smp3.acceptance_fraction))}
            sens.sweep(segments=100, nwalkers=60, burnin=5, nsteps=5)   # warm-up (first cooperative launches, NCCL float64 path)
            sec, sw = timed(lambda: sens.sweep(segments=100, nwalkers=60, burnin=200, nsteps=1000))
            cfg_info['C5_sens_sweep'] = {'grid_points': int(len(sw['scale'])), 'walkers': 60, 'steps': 1200, 'seconds': sec,
                                         'evals_per_s': len(sw['scale']) * 60 * 1200 / sec, 'acceptance': float(sw['acceptance'].mean()),
                                         'sharding': 'grid points split over %d rank(s), one all-reduce of the summaries' % world}
            sens.evidence_grid(dimensions=(6,), segments=4, samples=10000)                 # warm-up
            sec, ev = timed(lambda: sens.evidence_grid(segments=100, samples=1000000))
            cfg_info['C5_evidence_grid'] = {'grid_points': int(sum(len(v) for v in ev.values())), 'samples_per_point': 1000000, 'seconds': sec,
                                            'samples_per_s': 6e8 / sec,
                                            'note': 'Monte-Carlo evidence ln mean(L) per (dimension, scale), what scripts/sens.py gets from MultiNest; '
                                                    'samples sharded over the ranks, two all-reduces of the 600 (max, sum-exp) slots'}
        return cfg_info

    try:
        cfg_info = _configs_section()
    except Exception as exc:  # noqa: BLE001  (a secondary section must never cost the headline line)
        cfg_info = {'error': repr(exc)}

    if rank == 0:
        base = None
        if world == 1 and opts.cpu_evals > 0:
            v, total, wall = cpu_baseline(opts.cpu_evals, 1)
            base = {'value': v, 'unit': 'evals/s', 'cores': 1, 'kind': 'port',
                    'sample': '{0} scalar float128 ln_prob evaluations of the same model ({1:.1f} s)'.format(total, wall)}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except (OSError, ValueError):
            pass
        hbm_peak = peaks.get('hbm_gbs', 6650.0)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json'))).get('k_lnprob_bytes_per_launch')
        except (OSError, ValueError):
            pass
        achieved = FLOP_PER_EVAL * n / (kernel_ms * 1e-3) / 1e12
        line = {
            'metric': METRIC, 'value': value, 'unit': 'evals/s', 'n_gpus': world, 'steps': opts.steps, 'warmup': opts.warmup,
            'ms_per_step': kernel_ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'points_per_step_per_gpu': n, 'ndim': fn.ndim, 'nbins': 20, 'parallelism': 'dp%d (independent shards, no data-path collective)' % world,
                       'l2': 'inputs (235 MB theta per step) larger than the 126 MB L2', 'finite_fraction': finite_frac},
            'roofline': {'bound': 'fp64', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak if peak else None,
                         'traffic': traffic, 'kernel': 'k_lnprob<0>', 'flop_per_eval': FLOP_PER_EVAL,
                         'peak_source': 'DFMA microbenchmark (gf_fp64_peak_probe) measured in this run; nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2',
                         'hbm': {'achieved': BYTES_PER_EVAL * n / (kernel_ms * 1e-3) / 1e9, 'peak': hbm_peak, 'unit': 'GB/s',
                                 'frac': BYTES_PER_EVAL * n / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                                 'peak_source': 'MEASURED_PEAKS.json' if peaks else 'fallback'}},
            'cpu_baseline': base,
            'e2e': {'value': e2e_value, 'unit': 'evals/s', 'h2d_bytes_per_step': n * fn.ndim * 8, 'd2h_bytes_per_step': n * 8,
                    'steps': e2e_steps, 'api': 'golemflavor_b200.llh.LnProb.evaluate_host -> gf_lnprob_host (pinned host buffers)',
                    'h2d_gbs': e2e_value / world * fn.ndim * 8 / 1e9, 'h2d_link_gbs': link,
                    'note': 'bound by the host link: h2d_gbs is the input rate the pipeline sustains per GPU, h2d_link_gbs a plain pinned copy of the same buffer'},
            'gpu_launches': int(launches),
            'clocks': clock_info,
            'scan': scan_info,
            'configs': cfg_info,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _claim_stdout():
    """The driver parses ONE JSON line from stdout, but libraries write there too (NCCL prints its
    version banner to fd 1 on communicator creation).  Keep a private duplicate of the real stdout
    for the result line and point fd 1 at stderr for everybody else."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=400)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-evals', type=int, default=1500, help='size of the bounded CPU-baseline sample (0 = skip)')
    ap.add_argument('--scan-samples', type=int, default=10 ** 10, help='samples of the secondary scan section (0 = skip)')
    ap.add_argument('--configs', type=int, default=1, help='1: also time the sampler-shaped configs C2/C3/C5 (secondary section)')
    ap.add_argument('--scan-mode', default='anarchic', choices=['unitary', 'x', 'texture', 'anarchic'])
    opts = ap.parse_args()
    opts.warmup = max(opts.warmup, 3) if opts.impl == 'b200' else opts.warmup
    if opts.impl == 'reference':
        run_reference(opts)
    else:
        run_gpu(opts)


if __name__ == '__main__':
    main()
