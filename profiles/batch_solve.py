"""windows/s of complete multi-level solves of B MVSEC-shaped sequences in LOCKSTEP (solver.BatchedMultipleLevelEINCMSolver: one batched
device-side level solve per pyramid level, eincm_batch_minimize_bfgs_graph_host) against one solver per sequence on its own host thread
(the device-side loop per window, bench.py's windows_per_s arrangement).
usage: python profiles/batch_solve.py [--workload mvsec_dt4] [--batch 64] [--windows 3] [--threads 3]"""
import argparse
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from eincm_b200 import losses, solver as SV, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='mvsec_dt4')
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--windows', type=int, default=3)
ap.add_argument('--threads', type=int, default=3)
ap.add_argument('--no-threads', action='store_true')
a = ap.parse_args()
torch.cuda.set_device(0)
t0 = time.perf_counter()
seqs = [synth.make_sequence(a.workload, 1 + a.windows, seed=t) for t in range(a.batch)]
w0 = seqs[0][0]
H, W = w0.sensor_size
hpd = w0.hparams
N, R = len(w0.xs), len(w0.edge_ts)
print(f'# {a.workload}: {W}x{H}, {N} events, R = {R}; {a.batch} sequences x (1 + {a.windows}) windows generated in {time.perf_counter() - t0:.1f} s', flush=True)


def make_objs(n):
    return [losses.WindowObjective((H, W), hpd['alpha'], hpd['beta'], 0.0, hpd['delta'], max_events=N, max_refs=max(R, 3)) for _ in range(n)]


def aee(res, win):
    th = res['final_theta_pyr']['pyr_lvl_0']
    tr = synth.theta_test_points(win, th.shape[:2])['truth']
    return float(np.mean(np.linalg.norm(th - tr, axis=-1)))


# ---- lockstep --------------------------------------------------------------------------------------------------------------------------
objs = make_objs(a.batch)
lock = SV.BatchedMultipleLevelEINCMSolver(objs)
acc = {}


def timed(obj, name, label=None):
    f = getattr(obj, name)

    def g(*args, **kw):
        t = time.perf_counter()
        try:
            return f(*args, **kw)
        finally:
            key = label(*args) if label else name
            c = acc.setdefault(key, [0.0, 0])
            c[0] += time.perf_counter() - t
            c[1] += 1
    setattr(obj, name, g)


_rts = lock._run_theta_solvers
lvl_stats = {}


def rts(lvl, thetas0, active=None):
    t = time.perf_counter()
    out = _rts(lvl, thetas0, active)
    c = lvl_stats.setdefault(lvl, [0.0, 0, 0, 0])
    c[0] += time.perf_counter() - t
    c[1] += lock.batch.solve_launches()
    c[2] += sum(st.n_evals for k, st in enumerate(out[1]) if active is None or active[k])
    c[3] += 1
    return out


lock._run_theta_solvers = rts
timed(lock, 'set_datasamples')
for s_ in lock.solvers:
    timed(s_, '_perform_handover_at_level', lambda lvl: f'level {lvl} handover')
lock.set_datasamples([s[0].args() for s in seqs])
lock.solve()                                                     # first windows: no handover; graphs are built here
torch.cuda.synchronize()
n0 = sum(o.n_evals for o in objs)
l0 = lock.graph_launches
acc.clear()
lvl_stats.clear()
t0 = time.perf_counter()
errs = []
for k in range(1, 1 + a.windows):
    lock.set_datasamples([s[k].args() for s in seqs])
    res = lock.solve()
    errs += [aee(r, s[k]) for r, s in zip(res, seqs)]
dt = time.perf_counter() - t0
nw = a.batch * a.windows
ne = sum(o.n_evals for o in objs) - n0
print(f'lockstep, {a.batch} sequences: {nw / dt:.2f} windows/s ({dt / nw * 1e3:.2f} ms per window, {ne / nw:.0f} evaluations per window, '
      f'{dt / ne * 1e6:.1f} us of wall time per evaluation, {(lock.graph_launches - l0) / a.windows:.0f} graph launches per batch of windows, '
      f'AEE {np.mean(errs):.2f} px)', flush=True)
for lvl in sorted(lvl_stats, reverse=True):
    c = lvl_stats[lvl]
    print(f'   level {lvl} solve: {c[0] * 1e3 / a.windows:.1f} ms per batch of windows ({c[3] / a.windows:.1f} calls, {c[1] / a.windows:.0f} graph launches, '
          f'{c[2] / a.windows / a.batch:.0f} evaluations per window: {c[0] / max(c[2], 1) * 1e6:.1f} us per evaluation)')
for key in sorted(acc):
    print(f'   {key}: {acc[key][0] * 1e3 / a.windows:.1f} ms per batch of windows ({acc[key][1] // a.windows} calls)')
lock.close()
for o in objs:
    o.close()

# ---- one solver per sequence, T host threads ---------------------------------------------------------------------------------------------
if not a.no_threads:
    T = a.threads
    objs = make_objs(T)
    sols = [SV.MultipleLevelEINCMSolver(o, backend='graph', own_stream=True) for o in objs]
    for t, sol in enumerate(sols):
        sol.set_datasample(*seqs[t][0].args())
        sol.solve()
    errs = [[] for _ in range(T)]

    def work(t):
        torch.cuda.set_device(0)
        for k in range(1, 1 + a.windows):
            sols[t].set_datasample(*seqs[t][k].args())
            errs[t].append(aee(sols[t].solve(), seqs[t][k]))

    n0 = sum(o.n_evals for o in objs)
    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(t,)) for t in range(T)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    nw = T * a.windows
    ne = sum(o.n_evals for o in objs) - n0
    print(f'one device loop per sequence, {T} threads: {nw / dt:.2f} windows/s ({ne / nw:.0f} evaluations per window, '
          f'{dt / ne * 1e6:.1f} us of wall time per evaluation, AEE {np.mean(sum(errs, [])):.2f} px)', flush=True)
