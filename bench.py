#!/usr/bin/env python
"""Benchmark of the EINCM objective+gradient hot path (BASELINE.json metric: warped-event objective+grad evals,
Gevents/s/GPU at DSEC 640x480).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload dsec] [--theta 16]

A *step* is one objective+gradient evaluation (all R reference times, forward + analytic backward) of EACH of ``--windows``
independent synthetic DSEC-shaped event windows resident on the GPU (a batch of windows; one plan and one stream per window).
Their combined working set is larger than the 126 MB L2.  ``single_window`` reports the same evaluations one at a time.

Own arm  : ``value``  = device-resident evaluation (theta, loss and gradient stay in HBM), timed with CUDA events on the
           launching stream, max over ranks.  ``e2e`` = the same evaluations through the reference-facing host call
           (``eincm_value_and_grad_host_batch``: theta of every window from host memory, loss + gradient of every window
           read back to the host every step; per window this is what jaxopt's ``scipy_fun`` does per line-search step).
           ``e2e_stateless`` additionally re-stages the whole window (events + edge images) from pinned host memory every
           step (``set_window`` + evaluation).
Reference arm (``--impl reference``): the CPU restatement of the reference (``oracle/``; JAX is not installable in this
           image) on the host cores, each step a bounded sample of the same workload.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'objective+grad evaluation throughput (events x evals / s)'
UNIT = 'Gevents/s'


def _peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def measured_traffic():
    """DRAM bytes per launch from the committed ncu capture of this command (profiles/r2_traffic.json)."""
    p = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
    try:
        with open(p) as f:
            return json.load(f)['dram_bytes_per_launch']
    except Exception:
        return {}


def algorithmic_bytes(N, R, H, W):
    """SURVEY.md §8(d), reference dtypes: events (x:i16, y:i16, t:f64) read once per event pass (2 passes);
    per reference image: IWE written + read, edge image read, dL/dIWE written + read (f64); dense theta field
    written and dense gradient field read (f64 x 2 channels)."""
    ev = 2 * N * 12
    img = R * H * W * 8 * 5
    fld = H * W * 2 * 8 * 2
    return {'eval': ev + img + fld,
            'k_splat': N * 12 + R * H * W * 8 + H * W * 16,                      # events, IWE written, dense theta field
            'k_image_stats': R * H * W * 8 * 2,                                 # IWE read, edge image read
            'k_image_grad': R * H * W * 8,                                      # dL/dIWE written
            'k_backward_events': N * 12 + R * H * W * 8 + H * W * 16,           # events, dL/dIWE read, dense gradient field
            'k_theta_grad': H * W * 16}


class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index=0):
        self.samples = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.gpu_index}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        rows = [s for s in self.samples if t0 - 0.05 <= s[0] <= t1 + 0.15] or self.samples[-3:]
        for _, line in rows:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def workload_string(name, W, H, N, R, theta):
    """config.workload: the same string in both arms (the driver compares them)."""
    shaped = 'DSEC-shaped ' if name.startswith(('dsec', 'e00')) else ''
    return (f'{name}: {shaped}{W}x{H}, N={N} events/window, R={R} reference times, theta {theta}x{theta}x2 '
            f'(finest pyramid level)')


DTYPE = 'mixed: f64 event coordinates, warp, rint and image statistics; f32 tap values; 2^-21 fixed-point (u32/u64) votes'


def make_windows(args, rank):
    from eincm_b200 import synth
    wins = []
    for k in range(args.windows):
        wins.append(synth.make_workload(args.workload, seed=1000 * rank + k, n_events=args.events))
    return wins



# --------------------------------------------------------------------------------------------------------------
# batched evaluation: B windows per launch (BASELINE.json configs[2]: MVSEC-shaped windows, tile and dense theta)
# --------------------------------------------------------------------------------------------------------------
def batched_measure(workload, B, theta, dense, steps, warmup, min_time_s, cpu=True, n_distinct=16, n_check=2, cpu_budget=6.0):
    """One batch of B staged windows on the current GPU: device-resident step (CUDA events), the host-operand call, per-kernel
    spans, roofline of the dominant kernel, parity of the first windows against the CPU restatement and its timing."""
    import torch
    from eincm_b200 import plan as P, synth
    wins = [synth.make_workload(workload, seed=100 + k) for k in range(min(n_distinct, B))]
    H, W = wins[0].sensor_size
    N, R = len(wins[0].xs), len(wins[0].edge_ts)
    hpd = wins[0].hparams
    hp = P.make_hparams(hpd['alpha'], hpd['beta'], 0.0, 0.0, 1)
    shape = (H, W) if dense else (theta, theta)
    plans, th_h, th_d, lo_d, gr_d = [], [], [], [], []
    for k in range(B):
        w = wins[k % len(wins)]
        p = P.Plan((H, W), max_events=N, max_refs=max(R, 3))
        p.set_window(*w.args())
        plans.append(p)
        t = synth.theta_test_points(w, shape, seed=k)['perturbed']
        th_h.append(t)
        th_d.append(torch.from_numpy(t).cuda())
        lo_d.append(torch.zeros(1, dtype=torch.float64, device='cuda'))
        gr_d.append(torch.zeros(shape + (2,), dtype=torch.float64, device='cuda'))
    batch = P.Batch(plans)
    for _ in range(warmup):
        batch.value_and_grad_device(th_d, hp, lo_d, gr_d)
    torch.cuda.synchronize()
    block_ms = []
    l0 = batch.launch_count()
    while True:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            batch.value_and_grad_device(th_d, hp, lo_d, gr_d)
        e1.record()
        torch.cuda.synchronize()
        block_ms.append(e0.elapsed_time(e1))
        if sum(block_ms) >= min_time_s * 1e3 or len(block_ms) >= 200:
            break
    launches = (batch.launch_count() - l0) // len(block_ms)
    ms_step = float(np.median(block_ms)) / steps
    # per-kernel spans (separate pass)
    batch.set_timing(True)
    for _ in range(max(3, steps // 2)):
        batch.value_and_grad_device(th_d, hp, lo_d, gr_d)
    kt = batch.get_timing()
    batch.set_timing(False)
    kern_ms = {k: v[0] / max(v[1], 1) for k, v in kt.items()}
    # host operands: thetas of all windows in, losses + gradients out, one synchronous call per step
    for _ in range(2):
        batch.value_and_grad_host(th_h, hp)
    n_e2e = max(3, min(steps, 10))
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        losses_h, grads_h = batch.value_and_grad_host(th_h, hp)
    e2e_s = (time.perf_counter() - t0) / n_e2e
    theta_bytes = int(np.prod(shape)) * 2 * 8
    alg = algorithmic_bytes(N, R, H, W)
    peak, peak_src = _peaks()
    dom = max(kern_ms, key=lambda k: kern_ms[k])
    achieved = B * alg[dom] / (kern_ms[dom] * 1e-3) / 1e9
    res = {'workload': f'{workload}: {W}x{H}, N={N} events/window, R={R} reference times, theta ' +
                       (f'dense {H}x{W}x2' if dense else f'{theta}x{theta}x2') + f', batch of {B} windows per launch',
           'value': B * N / (ms_step * 1e-3) / 1e9, 'unit': UNIT, 'ms_per_step': ms_step, 'us_per_window': ms_step / B * 1e3,
           'evals_per_s': B / (ms_step * 1e-3), 'gpu_launches_per_step': launches // steps,
           'e2e': {'value': B * N / e2e_s / 1e9, 'unit': UNIT, 'ms_per_step': e2e_s * 1e3, 'h2d_bytes_per_step': B * theta_bytes,
                   'd2h_bytes_per_step': B * (theta_bytes + 8), 'call': 'Batch.value_and_grad_host -> eincm_batch_value_and_grad_host'},
           'kernels_ms_per_launch': {k: round(v, 5) for k, v in sorted(kern_ms.items(), key=lambda kv: -kv[1])},
           'roofline': {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                        'traffic': None, 'algorithmic_bytes_per_launch': B * alg[dom], 'kernel_ms': kern_ms[dom],
                        'kernel_share_of_step': kern_ms[dom] / sum(kern_ms.values()), 'peak_source': peak_src},
           'roofline_eval': {'bound': 'hbm', 'achieved': B * alg['eval'] / (ms_step * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                             'frac': B * alg['eval'] / (ms_step * 1e-3) / 1e9 / peak, 'algorithmic_bytes_per_eval': alg['eval']},
           'l2': f'{B} windows x ~{(N * 16 + (3 * R + 6) * H * W * 8) / 1e6:.0f} MB = {B * (N * 16 + (3 * R + 6) * H * W * 8) / 1e6:.0f} MB working set'}
    if cpu:
        fn, kind, cores, label = cpu_eval_factory()
        worst_l, worst_g = 0.0, 0.0
        for k in range(min(n_check, B)):
            l_ref, g_ref = fn(th_h[k], wins[k % len(wins)], dict(hpd, gamma=0.0), 1)
            worst_l = max(worst_l, abs(losses_h[k] - l_ref) / abs(l_ref))
            worst_g = max(worst_g, float(np.abs(grads_h[k] - g_ref).max() / np.abs(g_ref).max()))
        res['check'] = {'windows': min(n_check, B), 'loss_rel': worst_l, 'grad_rel_inf': worst_g, 'tolerance': {'loss_rel': 1e-5, 'grad_rel_inf': 1e-4},
                        'ok': bool(worst_l <= 1e-5 and worst_g <= 1e-4), 'checker': label}
        # CPU restatement on the same windows: whole windows, as many evaluations as fit the budget
        t0 = time.perf_counter()
        n_cpu = 0
        while n_cpu < 4 or (time.perf_counter() - t0 < cpu_budget and n_cpu < 4000):
            k = n_cpu % min(B, len(wins))
            fn(th_h[k], wins[k], dict(hpd, gamma=0.0), 1)
            n_cpu += 1
        dt = time.perf_counter() - t0
        res['cpu_baseline'] = {'value': n_cpu * N / dt / 1e9, 'unit': UNIT, 'cores': cores, 'kind': kind,
                               'sample': f'{n_cpu} objective+grad evals of complete windows ({N} events each), {label}'}
    batch.close()
    for p in plans:
        p.close()
    torch.cuda.empty_cache()
    return res


def run_batched(args):
    """bench.py --workload mvsec_dt1|mvsec_dt4|mvsec_raw_dt4 [--dense] [--batch B]: the batched evaluation as the main line."""
    import torch
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.25)
    t0 = time.time()
    r = batched_measure(args.workload, args.batch, args.theta, args.dense, args.steps, args.warmup, args.min_time_s, cpu=not args.no_cpu_baseline,
                        cpu_budget=args.cpu_budget)
    clocks = sampler.stop(t0, time.time())
    line = {'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': 1, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': DTYPE,
            'data': 'synthetic', 'config': {'workload': r['workload'], 'l2': r['l2'], 'step': 'one objective+gradient evaluation of every window of the batch: five launches (four for dense theta)'},
            'clocks': clocks, 'e2e': r['e2e'], 'gpu_launches': r['gpu_launches_per_step'] * args.steps, 'roofline': r['roofline'],
            'roofline_eval': r['roofline_eval'], 'kernels_ms_per_launch': r['kernels_ms_per_launch'], 'cpu_baseline': r.get('cpu_baseline'),
            'check': r.get('check'), 'us_per_window': r['us_per_window'], 'evals_per_s': r['evals_per_s']}
    print(json.dumps(line))
    if (r.get('check') or {}).get('ok') is False:
        sys.exit(3)


def event_split_measure(world, rank, local_rank, total_events, steps=5, warmup=3, theta=16):
    """BASELINE.json configs[4]: ONE very large window (1280x720) whose events are split over the ranks (strong scaling: the same
    total at every N).  Every rank synthesises ITS share of the events on the shared scene (same edge images, own events), so
    the union over the ranks is one window of `total_events` events.  Two forms (eincm_b200.parallel.EventSplitObjective): NCCL
    all-reduce of the float64 images, and the splat fused with the reduction over peer memory (NVLink reductions into every
    rank's fixed-point image)."""
    import torch
    import torch.distributed as dist
    from eincm_b200 import parallel as PAR, plan as P, synth
    n_local = total_events // world
    win = synth.make_workload('large', seed=500 + rank, n_events=n_local, scene_seed=4242)
    H, W = win.sensor_size
    R = len(win.edge_ts)
    hpd = win.hparams
    shape = (theta, theta)
    th = torch.from_numpy(synth.theta_test_points(win, shape, seed=0)['perturbed']).cuda()
    if world > 1:
        dist.broadcast(th, 0)                               # one theta for all ranks
    res = {'workload': f'large: ONE window of {W}x{H}, N={n_local * world} events split over {world} GPU(s), R={R}, theta {theta}x{theta}x2',
           'scaling': 'strong', 'unit': UNIT}
    for mode, p2p, fixed in (('nccl_allreduce_of_images', False, False), ('nccl_allreduce_of_fixed_point_images', False, True),
                             ('peer_fused_splat', True, False)):
        plan = P.Plan((H, W), max_events=n_local, max_refs=max(R, 3), flags=P.FLAG_EVENT_SPLIT)
        obj = PAR.EventSplitObjective(plan, lambda lvl: P.make_hparams(hpd['alpha'], hpd['beta'], hpd['gamma'], hpd['delta'], lvl), p2p=p2p,
                                      fixed_point=fixed)
        obj.set_datasample(win.xs, win.ys, win.ts, win.edges, win.edge_ts, need_mask=False)
        loss = torch.zeros(1, dtype=torch.float64, device='cuda')
        grad = torch.zeros_like(th)
        for _ in range(warmup):
            obj.value_and_grad(th, 0, loss, grad)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = obj.collective_bytes
        e0.record()
        for _ in range(steps):
            obj.value_and_grad(th, 0, loss, grad)
        e1.record()
        torch.cuda.synchronize(); dist.barrier()
        tms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item()) / steps
        # the objective must be identical on every rank (same complete images everywhere)
        lo = torch.stack([loss.detach().clone().reshape(()), -loss.detach().clone().reshape(())])
        dist.all_reduce(lo, op=dist.ReduceOp.MAX)
        res[mode] = {'value': n_local * world / (ms * 1e-3) / 1e9, 'ms_per_eval': ms, 'loss': float(loss.item()),
                     'loss_identical_on_all_ranks': bool(float(lo[0]) == float(loss.item()) and -float(lo[1]) == float(loss.item())),
                     'collective_bytes_per_eval_per_rank': (obj.collective_bytes - c0) // steps}
        plan.close()
        torch.cuda.empty_cache()
    return res

# --------------------------------------------------------------------------------------------------------------
# reference arm: CPU restatement of the reference on the host cores
# --------------------------------------------------------------------------------------------------------------
def cpu_eval_factory():
    """Returns (fn(theta, win, hp, lvl) -> (loss, grad), kind, cores, label).  Prefers the multi-threaded C restatement
    (oracle/_build) and falls back to the NumPy one; both are restatements ("port"), not the JAX reference itself."""
    try:
        from oracle import c_oracle
        if c_oracle.available():
            # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1: override it for the CPU arm)
            n_cpu = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
            c_oracle.set_num_threads(n_cpu)
            cores = c_oracle.num_threads()
            return c_oracle.value_and_grad, 'port', cores, f'oracle/eincm_oracle_c.c (OpenMP, {cores} threads)'
    except Exception:
        pass
    from oracle import eincm_oracle as O

    def fn(theta, win, hp, lvl):
        return O.value_and_grad(theta, *win.args(), hp['alpha'], hp['beta'], hp['gamma'], hp['delta'], lvl, 5, win.sensor_size)

    return fn, 'port', 1, 'oracle/eincm_oracle.py (NumPy float64, 1 thread)'


def cpu_sample_window(win, n_sample):
    """A bounded sample of the workload: the first n_sample events of a random permutation (same sensor, same edges)."""
    from eincm_b200.synth import Window
    N = len(win.xs)
    if n_sample >= N:
        return win
    idx = np.sort(np.random.default_rng(7).permutation(N)[:n_sample])
    return Window(xs=np.ascontiguousarray(win.xs[idx]), ys=np.ascontiguousarray(win.ys[idx]), ts=np.ascontiguousarray(win.ts[idx]),
                  edges=win.edges, edge_ts=win.edge_ts, sensor_size=win.sensor_size, truth_theta=win.truth_theta,
                  hparams=win.hparams)


def time_cpu(win, theta, steps, warmup, budget_s):
    fn, kind, cores, label = cpu_eval_factory()
    N = len(win.xs)
    # probe on a small sample to size the per-step sample so that (steps + warmup) steps fit the budget
    probe_n = min(N, 50_000)
    pw = cpu_sample_window(win, probe_n)
    t0 = time.perf_counter(); fn(theta, pw, win.hparams, 0); tp = time.perf_counter() - t0
    t0 = time.perf_counter(); fn(theta, pw, win.hparams, 0); tp = min(tp, time.perf_counter() - t0)
    per_event = max(tp, 1e-4) / probe_n          # upper bound: includes the image-space fixed cost
    n_sample = int(min(N, max(probe_n, budget_s / max(1, steps + warmup) / per_event)))
    sw = cpu_sample_window(win, n_sample)
    for _ in range(warmup):
        fn(theta, sw, win.hparams, 0)
    t0 = time.perf_counter()
    for _ in range(steps):
        fn(theta, sw, win.hparams, 0)
    dt = time.perf_counter() - t0
    return {'value': n_sample * steps / dt / 1e9, 'unit': UNIT, 'cores': cores, 'kind': kind,
            'sample': f'{n_sample} of {N} events of one window, {steps} objective+grad evals, {label}',
            'ms_per_step': dt / steps * 1e3, 'n_sample': n_sample}


def time_edge_maps(H, W, R, cpu=True, reps=20):
    """Edge-image stage of one window: R uint8 frames (pinned host memory) -> Canny -> Gaussian -> normalise, result on the device
    (eincm_b200.img_utils.edge_maps -> eincm_edge_maps).  CPU side: OpenCV itself (the reference's implementation of this stage,
    exp_mgr.py:343-350) when cv2 is importable on the box, else the NumPy restatement."""
    import torch
    from eincm_b200 import img_utils, synth
    frames = synth.make_frames(H, W, R, seed=0)
    pinned = torch.from_numpy(frames).pin_memory()
    for _ in range(3):
        img_utils.edge_maps(pinned.cuda(non_blocking=True), 30, 80)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = img_utils.edge_maps(pinned.cuda(non_blocking=True), 30, 80)
    torch.cuda.synchronize()
    gpu_ms = (time.perf_counter() - t0) / reps * 1e3
    d_frames = pinned.cuda()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        img_utils.edge_maps(d_frames, 30, 80)
    b.record()
    torch.cuda.synchronize()
    for _ in range(2):
        den = img_utils.fast_nl_means_denoising(d_frames, 4, 3, 11)
    a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a2.record()
    for _ in range(reps):
        den = img_utils.fast_nl_means_denoising(d_frames, 4, 3, 11)
    b2.record()
    torch.cuda.synchronize()
    nlm = {'ms_per_window_device': a2.elapsed_time(b2) / reps, 'params': 'h 4, template 3, search 11 (denoise/default.yaml)'}
    # CLAHE + Gaussian sharpen of preprocess_image (img_utils.py:159-178) chained on the denoised frames
    for _ in range(2):
        shp = img_utils.sharpen(img_utils.clahe_apply(den, 5, (10, 10)), 3, 1.5, -0.5)
    a3, b3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a3.record()
    for _ in range(reps):
        shp = img_utils.sharpen(img_utils.clahe_apply(den, 5, (10, 10)), 3, 1.5, -0.5)
    b3.record()
    torch.cuda.synchronize()
    cs = {'ms_per_window_device': a3.elapsed_time(b3) / reps, 'params': 'CLAHE clip 5, 10 x 10 tiles; GaussianBlur sigma 3 (19 taps), addWeighted 1.5 / -0.5'}
    res = {'ms_per_window_e2e': gpu_ms, 'ms_per_window_device': a.elapsed_time(b) / reps, 'frames': f'{R} x {W}x{H} uint8',
           'h2d_bytes': int(frames.nbytes), 'stages': 'Canny (3x3 Sobel, L2, 30/80) + Gaussian sigma 1 (9 taps) + normalise',
           'edge_pixels': int((out > 0.5).sum().item())}
    if cpu:
        try:
            import cv2 as cv
            eps = np.finfo(np.float64).eps

            def ref():
                es = []
                for f in frames:
                    s_ = cv.GaussianBlur(cv.Canny(f, 30, 80, None, 3, True).astype(np.float64), None, 1, 1, 0)
                    es.append((s_ - s_.min()) / (s_.max() - s_.min() + eps))
                return np.stack(es)
            kind = f'reference (OpenCV {cv.__version__}, {cv.getNumThreads()} threads)'
        except ImportError:
            from oracle import edge_oracle as EO

            def ref():
                return np.stack([EO.edge_map(f, 30, 80) for f in frames])
            kind = 'port (oracle/edge_oracle.py, NumPy)'
        ref()
        t0 = time.perf_counter()
        n = 0
        while n < 5 or (time.perf_counter() - t0 < 1.0 and n < 200):
            e_ref = ref()
            n += 1
        res['cpu_ms_per_window'] = (time.perf_counter() - t0) / n * 1e3
        res['cpu_kind'] = kind
        res['max_abs_diff_vs_cpu'] = float(np.abs(out.cpu().numpy() - e_ref).max())
        try:
            import cv2 as cv
            t0 = time.perf_counter()
            ref_den = np.stack([cv.fastNlMeansDenoising(f, None, 4, 3, 11) for f in frames])
            nlm['cpu_ms_per_window'] = (time.perf_counter() - t0) * 1e3
            nlm['cpu_kind'] = f'reference (OpenCV {cv.__version__}, {cv.getNumThreads()} threads)'
            nlm['identical_to_cpu'] = bool(np.array_equal(den.cpu().numpy(), ref_den))

            def ref_cs():
                out_ = []
                for f in ref_den:
                    c_ = cv.createCLAHE(clipLimit=5, tileGridSize=(10, 10)).apply(f)
                    out_.append(cv.addWeighted(c_, 1.5, cv.GaussianBlur(c_, None, 3, 2, 0), -0.5, 0))
                return np.stack(out_)
            ref_cs()
            t0 = time.perf_counter()
            for _ in range(5):
                ref_shp = ref_cs()
            cs['cpu_ms_per_window'] = (time.perf_counter() - t0) / 5 * 1e3
            cs['cpu_kind'] = nlm['cpu_kind']
            cs['identical_to_cpu'] = bool(np.array_equal(shp.cpu().numpy(), ref_shp))
        except ImportError:
            pass
    res['nlm_denoise'] = nlm
    res['clahe_sharpen'] = cs
    return res


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from eincm_b200 import synth
    win = synth.make_workload(args.workload, seed=0, n_events=args.events)
    theta = synth.theta_test_points(win, (args.theta, args.theta))['perturbed']
    r = time_cpu(win, theta, args.steps, args.warmup, budget_s=args.cpu_budget)
    H, W = win.sensor_size
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_string(args.workload, W, H, len(win.xs), len(win.edge_ts), args.theta),
                   'sample_events': r['n_sample']},
        'cpu_baseline': {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
        'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'note': 'CPU restatement of the reference (JAX/jaxlib/jaxopt are not installable in this image: no network)',
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------
# own arm
# --------------------------------------------------------------------------------------------------------------
def run_own(args):
    import torch
    import torch.distributed as dist
    from eincm_b200 import plan as P, synth

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at the first collective: keep stdout for the ONE JSON line
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
            t = torch.zeros(1, device='cuda')
            dist.all_reduce(t)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    wins = make_windows(args, rank)
    H, W = wins[0].sensor_size
    N = len(wins[0].xs)
    R = len(wins[0].edge_ts)
    hpd = wins[0].hparams
    hp = P.make_hparams(hpd['alpha'], hpd['beta'], hpd['gamma'], hpd['delta'], 0)
    shape = (args.theta, args.theta)
    plans, thetas_h, thetas_d, losses_d, grads_d = [], [], [], [], []
    for w in wins:
        p = P.Plan((H, W), max_events=N, max_refs=max(R, 3))
        p.set_window(*w.args())
        plans.append(p)
        th = synth.theta_test_points(w, shape)['perturbed']
        thetas_h.append(th)
        thetas_d.append(torch.from_numpy(th).cuda())
        losses_d.append(torch.zeros(1, dtype=torch.float64, device='cuda'))
        grads_d.append(torch.zeros(shape + (2,), dtype=torch.float64, device='cuda'))
    nw = len(plans)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # A step = one objective+gradient evaluation of EACH of the nw independent windows resident on this GPU (a batch of
    # windows, BASELINE.json configs[2]/[3]); every window has its own plan and stream, so kernels of different windows overlap.
    streams = [torch.cuda.Stream() for _ in plans]
    main = torch.cuda.current_stream()

    def dev_eval(k):
        plans[k].value_and_grad_device(thetas_d[k], hp, losses_d[k], grads_d[k])

    def dev_step():
        for k in range(nw):
            with torch.cuda.stream(streams[k]):
                dev_eval(k)

    def fork():
        for st in streams:
            st.wait_stream(main)

    def join():
        for st in streams:
            main.wait_stream(st)

    # ---- device-resident timing -------------------------------------------------------------------------------
    fork()
    for i in range(args.warmup):
        dev_step()
    join()
    launches0 = sum(p.launch_count() for p in plans)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    # The timed region: blocks of EXACTLY args.steps steps, each block bracketed by CUDA events on the launching stream, repeated until
    # >= args.min_time_s of device time has been measured (one 20-step block is ~10 ms: a single sample).  The line reports the MEDIAN block.
    t_wall0 = time.time()
    block_ms = []
    while True:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        fork()
        for i in range(args.steps):
            dev_step()
        join()
        ev1.record()
        barrier()
        block_ms.append(max_over_ranks(ev0.elapsed_time(ev1)))
        if sum(block_ms) >= args.min_time_s * 1e3 or len(block_ms) >= 500:
            break
    t_wall1 = time.time()
    ms_total = float(np.median(block_ms))
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    launches = (sum(p.launch_count() for p in plans) - launches0) // len(block_ms)
    ms_per_step = ms_total / args.steps
    value = world * nw * N * args.steps / (ms_total * 1e-3) / 1e9

    # the same evaluations one window at a time on one stream (no overlap between windows)
    for i in range(args.warmup):
        dev_eval(i % nw)
    barrier()
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    es0.record()
    for i in range(args.steps * nw):
        dev_eval(i % nw)
    es1.record()
    barrier()
    ms_single = max_over_ranks(es0.elapsed_time(es1)) / (args.steps * nw)

    # per-kernel durations for the roofline: a third pass (one stream) with CUDA events around every launch
    # (eincm_plan_set_timing, events on the launching stream); kept out of the timed regions above
    for p in plans:
        p.set_timing(True)
    for i in range(args.steps * nw):
        dev_eval(i % nw)
    torch.cuda.synchronize()
    kt = {}
    for p in plans:
        for name, (ms, n) in p.get_timing().items():
            a = kt.setdefault(name, [0.0, 0])
            a[0] += ms; a[1] += n
        p.set_timing(False)

    # ---- end to end through the host-facing call --------------------------------------------------------------
    # per step: theta of every window from host memory in, loss + gradient of every window out (one synchronous batched call)
    def host_step():
        return P.value_and_grad_host_batch(plans, thetas_h, hp)

    for i in range(args.warmup):
        host_step()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        losses_h, grads_h = host_step()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * nw * N * args.steps / e2e_s / 1e9
    loss_h, grad_h = float(losses_h[-1]), grads_h[-1]
    theta_bytes = int(np.prod(shape)) * 2 * 8

    # the reference's own calling pattern: one window per synchronous host call (jaxopt's scipy_fun per line-search step)
    for i in range(args.warmup):
        plans[i % nw].value_and_grad_host(thetas_h[i % nw], hp)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps * nw):
        plans[i % nw].value_and_grad_host(thetas_h[i % nw], hp)
    torch.cuda.synchronize()
    e2e1_s = max_over_ranks(time.perf_counter() - t0)
    e2e1_value = world * N * args.steps * nw / e2e1_s / 1e9

    # ---- windows/s: complete multi-level solves (BASELINE.json metric, second half) -------------------------------
    # The reference's run loop (exp_mgr.py:615-659): per window the 5-level coarse-to-fine solve of solver.py with the
    # main.yaml defaults (BFGS 40/28/19/11/8 iterations, retries at levels 0/1, handover solved by L-BFGS-B at levels 1/0),
    # consecutive windows of a rank chained through the handover prior.  scipy drives, every evaluation is a host call.
    solve = None
    if args.solve_windows > 0:
        from eincm_b200 import losses, solver as SV

        def run_solves(backend, n_threads, wait, seqs):
            """n_threads independent sequences per GPU, each: warm-up solve of its first window, then args.solve_windows chained
            windows (handover prior from the previous window).  wait = 'block': the host entry points sleep on a blocking CUDA
            event (EINCM_FLAG_BLOCKING_SYNC) instead of spinning, so the sequences per GPU do not depend on the host cores per rank."""
            flags = P.FLAG_BLOCKING_SYNC if wait == 'block' else 0
            objs = [losses.WindowObjective((H, W), hpd['alpha'], hpd['beta'], hpd['gamma'], hpd['delta'], max_events=N, max_refs=max(R, 3), flags=flags)
                    for _ in range(n_threads)]
            sols = [SV.MultipleLevelEINCMSolver(o, backend=backend, own_stream=(backend != 'scipy')) for o in objs]
            finals = [None] * n_threads

            def work(t, first, last):
                torch.cuda.set_device(local_rank)
                for k in range(first, last):
                    sols[t].set_datasample(*seqs[t][k].args())
                    finals[t] = sols[t].solve()

            n_rep = 3 if backend != 'scipy' else 1
            n_win = n_threads * args.solve_windows
            runs = []
            for rep in range(n_rep):
                # every repeat starts from scratch: first window of the sequence untimed (no prior), then the chained windows
                for t, sol in enumerate(sols):
                    sols[t] = SV.MultipleLevelEINCMSolver(objs[t], backend=backend, own_stream=(backend != 'scipy'))
                    work(t, 0, 1)
                barrier()
                n0 = sum(o.n_evals for o in objs)
                t0 = time.perf_counter()
                if n_threads == 1:
                    work(0, 1, 1 + args.solve_windows)
                else:
                    ths = [threading.Thread(target=work, args=(t, 1, 1 + args.solve_windows)) for t in range(n_threads)]
                    for th in ths:
                        th.start()
                    for th in ths:
                        th.join()
                torch.cuda.synchronize()
                runs.append((max_over_ranks(time.perf_counter() - t0), sum(o.n_evals for o in objs) - n0))
            dt, n_ev = sorted(runs)[len(runs) // 2]
            rates = [world * n_win / r[0] for r in runs]
            res = {'value': world * n_win / dt, 'unit': 'windows/s', 'repeats_windows_per_s': [round(x, 2) for x in rates],
                   'repeat_spread': (max(rates) - min(rates)) / float(np.median(rates)),
                   'repeats_evals_per_window': [round(r[1] / n_win, 1) for r in runs], 'sequences_per_gpu': n_threads,
                   'windows_per_sequence': args.solve_windows, 'ms_per_window': dt / args.solve_windows * 1e3, 'evals_per_window': n_ev / n_win,
                   'final_loss': finals[0]['theta_opt_state_pyr']['pyr_lvl_0'].fun_val,
                   'host_wait': 'sleeping on a blocking CUDA event (EINCM_FLAG_BLOCKING_SYNC)' if wait == 'block' else 'spinning on mapped pinned memory'}
            for o in objs:
                o.close()
            return res

        # synthetic SEQUENCES: same scene, slowly drifting flow, fresh events (synth.make_sequence) - the handover prior of window k
        # is the solution of window k - 1, as on consecutive DSEC windows
        n_seq = max(1, args.solve_seqs)
        seqs = [synth.make_sequence(args.workload, 1 + args.solve_windows, seed=100 * rank + t, n_events=args.events) for t in range(n_seq)]
        solve = run_solves('native', n_seq, args.solve_wait, seqs)
        solve['note'] = ('eincm_b200.solver.MultipleLevelEINCMSolver (mirror of reference src/eincm/solver.py, main.yaml defaults: 5 levels, '
                         'BFGS 40/28/19/11/8 iterations, retries, handover solved at levels 1/0); optimizers native '
                         '(eincm_minimize_bfgs_host / eincm_minimize_handover_host), one host thread and one CUDA stream per '
                         'sequence, windows of a sequence chained by handover, set_datasample (staging) inside the timed region; '
                         'sequences: same scene, truth flow drifting 8 % of the flow magnitude per window')
        # The loop on the device (eincm_minimize_bfgs_graph_host: unrolled CUDA graphs of 8 x { evaluation ; k_bfgs_step }, the host relaunches
        # ~7 times per pyramid level; the scalar handover solves stay host-driven): the same sequences, spinning and asleep between launches -
        # asleep, the rate does not depend on host cores (profiles/r2_graph_solve.txt) - and one sequence per GPU against the host loop.
        def brief(r):
            return {'value': r['value'], 'repeats_windows_per_s': r['repeats_windows_per_s'], 'evals_per_window': r['evals_per_window']}
        solve['device_graph_loop'] = {
            'call': 'eincm_minimize_bfgs_graph_host per pyramid level (solver backend "graph")', 'unit': 'windows/s', 'sequences_per_gpu': n_seq,
            'spinning_wait': brief(run_solves('graph', n_seq, 'spin', seqs)),
            'blocking_wait': brief(run_solves('graph', n_seq, 'block', seqs)),
            'one_sequence_per_gpu': {'device_graph_loop': brief(run_solves('graph', 1, args.solve_wait, seqs[:1])),
                                     'host_loop': brief(run_solves('native', 1, args.solve_wait, seqs[:1]))}}
        if args.solve_compare and world == 1:
            solve['spinning'] = run_solves('native', n_seq, 'spin' if args.solve_wait == 'block' else 'block', seqs)
            solve['scipy_single_sequence'] = run_solves('scipy', 1, args.solve_wait, seqs[:1])
        del seqs

    # ---- stateless: stage the whole window from pinned host memory every step ---------------------------------
    pin = []
    for w in wins:
        pin.append(tuple(torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (w.xs, w.ys, w.ts, w.edges)))
    n_sl = max(2, min(args.steps, 8))

    def stateless_step(i):
        k = i % nw
        xs, ys, ts, ed = (t.cuda(non_blocking=True) for t in pin[k])
        plans[k].set_window(xs, ys, ts, ed, wins[k].edge_ts)
        return plans[k].value_and_grad_host(thetas_h[k], hp)

    stateless_step(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(n_sl):
        stateless_step(i + 1)
    torch.cuda.synchronize()
    sl_s = max_over_ranks(time.perf_counter() - t0)
    sl_value = world * N * n_sl / sl_s / 1e9
    window_bytes = N * 12 + R * H * W * 8 + R * 8

    # ---- event split of ONE very large window over the ranks (BASELINE.json configs[4]); at N = 1 the same window on one GPU ------
    esplit = None
    if not args.no_event_split:
        if world == 1 and not dist.is_initialized():
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1'); os.environ.setdefault('MASTER_PORT', '29533')
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group('nccl', rank=0, world_size=1, device_id=torch.device('cuda', local_rank))
            finally:
                sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
        esplit = event_split_measure(world, rank, local_rank, args.split_events)

    if rank != 0:
        if dist.is_initialized():
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------------
    peak, peak_src = _peaks()
    alg = algorithmic_bytes(N, R, H, W)
    kern_ms = {k: v[0] / max(v[1], 1) for k, v in kt.items()}
    span_total = sum(v[0] for v in kt.values())
    dom = max(kt, key=lambda k: kt[k][0]) if kt else None
    roofline = None
    if dom is not None:
        dom_bytes = alg.get(dom, alg['eval'])
        achieved = dom_bytes / (kern_ms[dom] * 1e-3) / 1e9
        roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                    'traffic': measured_traffic().get(dom) if args.workload == 'dsec' and args.events is None else None, 'algorithmic_bytes_per_launch': dom_bytes, 'kernel_ms': kern_ms[dom],
                    'kernel_share_of_step': kt[dom][0] / span_total if span_total else None, 'peak_source': peak_src}
    eval_achieved = nw * alg['eval'] / (ms_per_step * 1e-3) / 1e9

    # ---- CPU baseline (bounded sample, rank 0, N = 1 only) ----------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = time_cpu(wins[0], thetas_h[0], steps=6, warmup=1, budget_s=args.cpu_budget)
        cpu = {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}

    # ---- parity of THIS run: window 0 of the timed batch, same theta, against the CPU restatement of the reference -----------
    # (BASELINE.json tolerances: objective 1e-5, gradient 1e-4 relative).  The checker runs on the complete window (not a sample).
    check = {'loss': float(loss_h), 'grad_inf': float(np.abs(grad_h).max())}
    if not args.no_cpu_baseline:
        fn, kind, cores, label = cpu_eval_factory()
        l0, g0 = plans[0].value_and_grad_host(thetas_h[0], hp)
        l_ref, g_ref = fn(thetas_h[0], wins[0], hpd, 0)
        check = {'window': 'window 0 of the timed batch (seed 0), complete', 'loss': float(l0), 'loss_ref': float(l_ref),
                 'loss_rel': abs(l0 - l_ref) / abs(l_ref), 'grad_rel_inf': float(np.abs(g0 - g_ref).max() / np.abs(g_ref).max()),
                 'tolerance': {'loss_rel': 1e-5, 'grad_rel_inf': 1e-4}, 'checker': label}
        check['ok'] = bool(check['loss_rel'] <= 1e-5 and check['grad_rel_inf'] <= 1e-4)
        # The same window and theta against the committed output of the reference's OWN source (its unmodified eincm.losses executed over
        # a float64 stand-in for the JAX primitives: tests/golden/make_golden_refsrc_fullsize.py).  Reported, never fatal: the vector
        # exists for the default workload only and the regenerated window must carry the stored checksum.
        try:
            z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tests', 'golden', 'refsrc_fullsize', 'dsec_2m_theta16.npz'))
            w0 = wins[0]
            cs = np.array([float(np.asarray(w0.xs, dtype=np.float64).sum()), float(np.asarray(w0.ys, dtype=np.float64).sum()),
                           float(np.asarray(w0.ts).sum()), float(np.asarray(w0.edges).sum()), float(len(w0.xs))])
            if (args.workload == 'dsec' and z['theta'].shape == np.shape(thetas_h[0]) and np.array_equal(z['theta'], thetas_h[0])
                    and np.allclose(cs, z['checksum'], rtol=1e-12, atol=0)):
                check['reference_source'] = {'loss_ref': float(z['loss']), 'loss_rel': abs(l0 - float(z['loss'])) / abs(float(z['loss'])),
                                             'grad_rel_inf': float(np.abs(g0 - z['grad']).max() / np.abs(z['grad']).max()),
                                             'vector': 'tests/golden/refsrc_fullsize/dsec_2m_theta16.npz'}
        except Exception as e:  # noqa: BLE001
            check['reference_source'] = {'error': repr(e)[:120]}

    # ---- the same device-resident step with EINCM_FLAG_EXACT_F64 plans (float64 taps and float64 scatter-adds: `dtype` f64 throughout)
    exact = None
    if world == 1 and not args.no_exact:
        xplans = []
        for w in wins:
            p = P.Plan((H, W), max_events=N, max_refs=max(R, 3), flags=P.FLAG_EXACT_F64)
            p.set_window(*w.args())
            xplans.append(p)
        for i in range(3):
            for k in range(nw):
                xplans[k].value_and_grad_device(thetas_d[k], hp, losses_d[k], grads_d[k])
        torch.cuda.synchronize()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_x = 5
        x0.record()
        for i in range(n_x):
            for k in range(nw):
                xplans[k].value_and_grad_device(thetas_d[k], hp, losses_d[k], grads_d[k])
        x1.record()
        torch.cuda.synchronize()
        xms = x0.elapsed_time(x1) / n_x
        exact = {'value': nw * N / (xms * 1e-3) / 1e9, 'unit': UNIT, 'ms_per_step': xms, 'dtype': 'f64',
                 'note': 'EINCM_FLAG_EXACT_F64: nine float64 atomic adds per event and image, float64 taps (what XLA emits for the reference)'}
        for p in xplans:
            p.close()

    # ---- edge images of a window: uint8 frames in host memory -> float64 edge images on the device ----
    edge = None
    if world == 1:
        edge = time_edge_maps(H, W, R, cpu=not args.no_cpu_baseline)

    # ---- MVSEC-shaped windows (BASELINE.json configs[2]): batches of windows per launch, tile and dense theta ----------------------
    mvsec = None
    if world == 1 and not args.no_mvsec:
        for p in plans:
            p.close()
        plans = []
        torch.cuda.empty_cache()
        mvsec = {}
        for name, wl, Bn, dn in (('dt4_tile16_b512', 'mvsec_dt4', 512, False), ('dt1_tile16_b512', 'mvsec_dt1', 512, False),
                                 ('dt4_dense_b64', 'mvsec_dt4', 64, True), ('dt1_dense_b64', 'mvsec_dt1', 64, True)):
            mvsec[name] = batched_measure(wl, Bn, 16, dn, steps=5, warmup=3, min_time_s=0.2, cpu=not args.no_cpu_baseline, cpu_budget=3.0)

        # complete 5-level solves of 64 MVSEC-shaped sequences in LOCKSTEP: per pyramid level one batched device-side BFGS solve of all
        # windows (eincm_batch_minimize_bfgs_graph_host), beside one device loop per sequence on 3 host threads
        try:
            mvsec['lockstep_solve'] = mvsec_lockstep_solve('mvsec_dt4', 64, 2, local_rank)
        except Exception as e:                                 # a sub-line never takes the headline down
            mvsec['lockstep_solve'] = {'error': f'{type(e).__name__}: {e}'}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': DTYPE, 'data': 'synthetic',
        'config': {'workload': workload_string(args.workload, W, H, N, R, args.theta),
                   'hparams': f'alpha={hpd["alpha"]}, beta={hpd["beta"]}, gamma={hpd["gamma"]}, delta={hpd["delta"]}',
                   'step': f'one objective+gradient evaluation of each of {nw} independent windows per GPU (one plan and one '
                           f'stream per window: kernels of different windows overlap)',
                   'l2': f'inputs larger than L2: {nw} distinct windows per GPU '
                         f'(~{nw * (N * 16 + (3 * R + 6) * H * W * 8) / 1e6:.0f} MB working set)',
                   'parallelism': f'windows sharded over {world} GPU(s), no data-path collective',
                   'windows_per_step_per_gpu': nw, 'events_per_step_per_gpu': nw * N},
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': nw * theta_bytes, 'd2h_bytes_per_step': nw * (theta_bytes + 8),
                'call': 'plan.value_and_grad_host_batch -> eincm_value_and_grad_host_batch: one synchronous call per step, theta of '
                        'every window from host memory in, loss + gradient of every window out (window operands are staged once '
                        'per window like the reference\'s jnp arrays)',
                'ms_per_step': e2e_s / args.steps * 1e3},
        'single_window': {'value': world * N / (ms_single * 1e-3) / 1e9, 'ms_per_eval': ms_single,
                          'e2e_value': e2e1_value, 'e2e_ms_per_eval': e2e1_s / (args.steps * nw) * 1e3, 'unit': UNIT,
                          'note': 'one window at a time on one stream; e2e = eincm_value_and_grad_host per evaluation (what '
                                  'jaxopt\'s scipy_fun does per line-search step)'},
        'e2e_stateless': {'value': sl_value, 'unit': UNIT, 'h2d_bytes_per_step': window_bytes + theta_bytes,
                          'd2h_bytes_per_step': theta_bytes + 8, 'steps': n_sl, 'ms_per_step': sl_s / n_sl * 1e3,
                          'call': 'set_window from pinned host memory + value_and_grad_host every step'},
        'windows_per_s': solve,
        'edge_maps': edge,
        'gpu_launches': int(launches),
        'roofline': roofline,
        'roofline_eval': {'bound': 'hbm', 'achieved': eval_achieved, 'peak': peak, 'unit': 'GB/s', 'frac': eval_achieved / peak,
                          'algorithmic_bytes_per_eval': alg['eval']},
        'kernels_ms_per_launch': {k: round(v, 5) for k, v in sorted(kern_ms.items(), key=lambda kv: -kv[1])},
        'cpu_baseline': cpu,
        'check': check,
        'mvsec_batched': mvsec,
        'event_split': esplit,
        'exact_f64': exact,
        'timed_blocks': {'blocks': len(block_ms), 'steps_per_block': args.steps, 'ms_median': ms_total, 'ms_min': float(min(block_ms)),
                         'ms_max': float(max(block_ms)), 'ms_total': float(sum(block_ms))},
    }
    print(json.dumps(line))
    if dist.is_initialized():
        dist.destroy_process_group()
    if check.get('ok') is False:
        sys.exit(3)


# --------------------------------------------------------------------------------------------------------------
# event split of ONE large window over the GPUs (BASELINE.json configs[4]); not the driver's default line
# --------------------------------------------------------------------------------------------------------------
def run_event_split(args):
    import torch
    import torch.distributed as dist
    from eincm_b200 import parallel as PAR, plan as P, synth

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local_rank)
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        if world > 1:
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        else:
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1'); os.environ.setdefault('MASTER_PORT', '29533')
            dist.init_process_group('nccl', rank=0, world_size=1, device_id=torch.device('cuda', local_rank))
        t = torch.zeros(1, device='cuda'); dist.all_reduce(t); torch.cuda.synchronize()
    finally:
        sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    win = synth.make_workload(args.workload, seed=0, n_events=args.events)      # the same window on every rank
    H, W = win.sensor_size
    N, R = len(win.xs), len(win.edge_ts)
    xs, ys, ts = PAR.split_events(win.xs, win.ys, win.ts, world, rank)
    hpd = win.hparams
    shape = (args.theta, args.theta)
    theta = torch.from_numpy(synth.theta_test_points(win, shape)['perturbed']).cuda()
    res = {}
    for mode, p2p in (('nccl_allreduce_of_images', False), ('peer_fused_splat', True)):
        plan = P.Plan((H, W), max_events=len(xs), max_refs=max(R, 3), flags=P.FLAG_EVENT_SPLIT)
        obj = PAR.EventSplitObjective(plan, lambda lvl: P.make_hparams(hpd['alpha'], hpd['beta'], hpd['gamma'], hpd['delta'], lvl), p2p=p2p)
        obj.set_datasample(xs, ys, ts, win.edges, win.edge_ts, need_mask=False)
        loss = torch.zeros(1, dtype=torch.float64, device='cuda'); grad = torch.zeros_like(theta)
        for _ in range(args.warmup):
            obj.value_and_grad(theta, 0, loss, grad)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = obj.collective_bytes
        e0.record()
        for _ in range(args.steps):
            obj.value_and_grad(theta, 0, loss, grad)
        e1.record()
        torch.cuda.synchronize(); dist.barrier()
        tms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item()) / args.steps
        res[mode] = {'value': N / (ms * 1e-3) / 1e9, 'unit': UNIT, 'ms_per_eval': ms,
                     'collective_bytes_per_eval_per_rank': (obj.collective_bytes - c0) // args.steps, 'loss': float(loss.item())}
        plan.close()
    if rank == 0:
        print(json.dumps({'mode': 'event_split', 'metric': METRIC, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                          'config': {'workload': f'{args.workload}: ONE window of {W}x{H}, N={N} events split over {world} GPU(s), R={R}, '
                                                 f'theta {shape[0]}x{shape[1]}x2'}, 'scaling': 'strong', **res}))
    dist.destroy_process_group()


def mvsec_lockstep_solve(workload, n_seq, n_windows, device):
    """windows/s of complete multi-level solves of n_seq MVSEC-shaped sequences solved in lockstep (solver.BatchedMultipleLevelEINCMSolver)
    and, for comparison, of 3 of them with one device-side loop per sequence on its own host thread."""
    import torch
    from eincm_b200 import losses, solver as SV, synth
    seqs = [synth.make_sequence(workload, 1 + n_windows, seed=1000 + t) for t in range(n_seq)]
    w0 = seqs[0][0]
    H, W = w0.sensor_size
    hpd = w0.hparams
    N, R = len(w0.xs), len(w0.edge_ts)

    def make_objs(n):
        return [losses.WindowObjective((H, W), hpd['alpha'], hpd['beta'], 0.0, hpd['delta'], max_events=N, max_refs=max(R, 3)) for _ in range(n)]

    objs = make_objs(n_seq)
    lock = SV.BatchedMultipleLevelEINCMSolver(objs)
    lock.set_datasamples([s[0].args() for s in seqs])
    lock.solve()                                              # first windows untimed: no prior, graphs built
    torch.cuda.synchronize()
    n0, l0 = sum(o.n_evals for o in objs), lock.graph_launches
    t0 = time.perf_counter()
    for k in range(1, 1 + n_windows):
        lock.set_datasamples([s[k].args() for s in seqs])
        lock.solve()
    dt = time.perf_counter() - t0
    n_win = n_seq * n_windows
    n_ev = sum(o.n_evals for o in objs) - n0
    out = {'value': n_win / dt, 'unit': 'windows/s', 'workload': workload_string(workload, W, H, N, R, 16), 'sequences': n_seq,
           'windows_per_sequence': n_windows, 'evals_per_window': n_ev / n_win, 'us_per_evaluation': dt / n_ev * 1e6,
           'graph_launches_per_batch': (lock.graph_launches - l0) / n_windows,
           'call': 'eincm_batch_minimize_bfgs_graph_host per pyramid level (solver.BatchedMultipleLevelEINCMSolver); the scalar handover '
                   'solves per window and set_datasample (staging) are inside the timed region'}
    lock.close()
    for o in objs:
        o.close()
    # the same sequences with one device-side loop per sequence, driven by T host threads (every thread owns n_seq / T sequences)
    T = 3
    objs = make_objs(n_seq)
    sols = [SV.MultipleLevelEINCMSolver(o, backend='graph', own_stream=True) for o in objs]
    for t, sol in enumerate(sols):
        sol.set_datasample(*seqs[t][0].args())
        sol.solve()

    def work(t):
        torch.cuda.set_device(device)
        for k in range(1, 1 + n_windows):
            for q in range(t, n_seq, T):
                sols[q].set_datasample(*seqs[q][k].args())
                sols[q].solve()

    torch.cuda.synchronize()
    n0 = sum(o.n_evals for o in objs)
    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(t,)) for t in range(T)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    n_ev = sum(o.n_evals for o in objs) - n0
    out['one_device_loop_per_sequence'] = {'value': n_win / dt, 'unit': 'windows/s', 'sequences': n_seq, 'host_threads': T,
                                           'evals_per_window': n_ev / n_win, 'us_per_evaluation': dt / n_ev * 1e6}
    for o in objs:
        o.close()
    return out


def main():
    # NCCL's version banner goes to (buffered) stdout; rank 0 must print ONE JSON line
    if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
        os.environ['NCCL_DEBUG'] = 'WARN'
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--warmup', type=int, default=8)
    ap.add_argument('--impl', default='own', choices=['own', 'reference'])
    ap.add_argument('--workload', default='dsec')
    ap.add_argument('--events', type=int, default=None, help='override the number of events per window')
    ap.add_argument('--theta', type=int, default=16, help='tile-flow resolution (theta is theta x theta x 2)')
    ap.add_argument('--windows', type=int, default=4, help='distinct windows per GPU cycled round-robin')
    ap.add_argument('--cpu-budget', type=float, default=20.0, help='seconds of CPU work for the CPU baseline / reference arm')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--batch', type=int, default=512, help='windows per launch for the batched (MVSEC) workloads')
    ap.add_argument('--dense', action='store_true', help='dense per-pixel theta (H x W x 2) instead of theta x theta tiles (batched workloads)')
    ap.add_argument('--no-event-split', action='store_true', help='skip the event-split measurement of one very large window')
    ap.add_argument('--split-events', type=int, default=50_000_000, help='events of the large window that is split over the GPUs')
    ap.add_argument('--no-mvsec', action='store_true', help='skip the MVSEC-shaped batched sub-lines of the default run')
    ap.add_argument('--no-exact', action='store_true', help='skip the EINCM_FLAG_EXACT_F64 sub-line')
    ap.add_argument('--min-time-s', type=float, default=0.5, help='device time to accumulate over repeated blocks of --steps steps')
    ap.add_argument('--solve-seqs', type=int, default=3, help='sequences solved concurrently per GPU for the windows/s figure (3: the spinning driver threads of 8 ranks fit a 32-core box)')
    ap.add_argument('--solve-wait', default='spin', choices=['block', 'spin'], help='how the host entry points wait for an evaluation')
    ap.add_argument('--solve-compare', action='store_true', help='also measure the other wait mode and the scipy-driven solve (N = 1)')
    ap.add_argument('--solve-windows', type=int, default=6, help='chained windows per sequence for the windows/s figure (0: skip); the evaluations per window vary (430 - 560), and the figure is the slowest rank\'s: more windows per sequence average that out')
    ap.add_argument('--event-split', action='store_true', help='ONE window split over the GPUs (configs[4]) instead of windows sharded')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.event_split:
        run_event_split(args)
        return
    if args.impl == 'reference':
        args.cpu_budget = max(args.cpu_budget, 60.0)
        run_reference(args)
    elif args.workload.startswith('mvsec') or args.workload in ('ecd', 'tiny'):
        run_batched(args)
    else:
        run_own(args)


if __name__ == '__main__':
    main()
