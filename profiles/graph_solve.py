"""Level solves and complete multi-level solves with the loop on the device (backend 'graph') against the host-driven native loop.
usage: python profiles/graph_solve.py [--threads 1]"""
import argparse, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from eincm_b200 import losses, plan as P, solver as SV, synth
ap = argparse.ArgumentParser()
ap.add_argument('--threads', type=int, default=1)
ap.add_argument('--windows', type=int, default=4)
ap.add_argument('--workload', default='dsec')
ap.add_argument('--blocking', action='store_true', help='EINCM_FLAG_BLOCKING_SYNC: the host waits asleep on an event')
ap.add_argument('--skip-levels', action='store_true')
a = ap.parse_args()
torch.cuda.set_device(0)
seq = synth.make_sequence(a.workload, 1 + a.windows, seed=0)
H, W = seq[0].sensor_size
hpd = seq[0].hparams
N = len(seq[0].xs)
# one level at a time
p = P.Plan((H, W), max_events=N, max_refs=max(3, len(seq[0].edge_ts)))
p.set_window(*seq[0].args())
for shape, lvl, maxiter in (() if a.skip_levels else (((1, 1), 4, 8), ((4, 4), 2, 19), ((16, 16), 0, 40))):
    hp = P.make_hparams(hpd['alpha'], hpd['beta'], 0.0, 0.0, lvl)
    th0 = 0.5 * synth.theta_test_points(seq[0], shape)['truth']
    for name, fn in (('host loop ', lambda: p.minimize_bfgs_host(th0, hp, maxiter, 1e-7, own_stream=True)), ('graph loop', lambda: p.minimize_bfgs_graph_host(th0, hp, maxiter, 1e-7))):
        fn()
        t0 = time.perf_counter(); th, r = fn(); dt = time.perf_counter() - t0
        print(f'level {lvl} theta {shape[0]}x{shape[1]} {name}: {dt * 1e3:7.2f} ms, status {r.status} nit {r.nit:3d} nfev {r.nfev:3d} -> {dt / r.nfev * 1e6:6.1f} us per evaluation, loss {r.fun:.6f}')
p.close()
# complete solves
for backend in ('native', 'graph'):
    seqs = [synth.make_sequence(a.workload, 1 + a.windows, seed=t) for t in range(a.threads)]
    objs = [losses.WindowObjective((H, W), hpd['alpha'], hpd['beta'], hpd['gamma'], hpd['delta'], max_events=N, max_refs=max(3, len(seq[0].edge_ts)), flags=P.FLAG_BLOCKING_SYNC if a.blocking else 0) for _ in range(a.threads)]
    sols = [SV.MultipleLevelEINCMSolver(o, backend=backend, own_stream=True) for o in objs]
    for t in range(a.threads):
        sols[t].set_datasample(*seqs[t][0].args()); sols[t].solve()
    def work(t):
        torch.cuda.set_device(0)
        for k in range(1, 1 + a.windows):
            sols[t].set_datasample(*seqs[t][k].args()); sols[t].solve()
    n0 = sum(o.n_evals for o in objs)
    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(t,)) for t in range(a.threads)]
    [th.start() for th in ths]; [th.join() for th in ths]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ne = sum(o.n_evals for o in objs) - n0
    print(f'backend {backend:6s} {"blocking wait" if a.blocking else "spinning wait"}: {a.threads} sequence(s) x {a.windows} windows: {a.threads * a.windows / dt:6.2f} windows/s, {ne / (a.threads * a.windows):5.0f} evaluations per window, {dt / ne * 1e6:6.1f} us of wall time per evaluation')
    for o in objs: o.close()
