"""Where a complete multi-level solve spends its time (one sequence, native optimizers): per pyramid level wall time, objective
evaluations, host launch / wait time inside the evaluations, and the staging of the window.
usage: python profiles/solve_breakdown.py [--windows 3]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from eincm_b200 import losses, solver as SV, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='dsec')
ap.add_argument('--windows', type=int, default=3)
a = ap.parse_args()
torch.cuda.set_device(0)
wins = [synth.make_workload(a.workload, seed=k) for k in range(4)]
H, W = wins[0].sensor_size
hpd = wins[0].hparams
N, R = len(wins[0].xs), len(wins[0].edge_ts)
obj = losses.WindowObjective((H, W), hpd['alpha'], hpd['beta'], hpd['gamma'], hpd['delta'], max_events=N, max_refs=max(R, 3))
sol = SV.MultipleLevelEINCMSolver(obj, backend='native', own_stream=True)
sol.set_datasample(*wins[0].args())
sol.solve()

acc = {}
orig_theta, orig_ho = sol._run_theta_solver, sol._run_handover_solver


def timed(kind, fn):
    def wrapper(pyr_lvl, *args, **kw):
        n0 = obj.n_evals
        ht0 = obj.plan.host_times(reset=False)
        t0 = time.perf_counter()
        out = fn(pyr_lvl, *args, **kw)
        dt = time.perf_counter() - t0
        ht1 = obj.plan.host_times(reset=False)
        e = acc.setdefault((kind, pyr_lvl), [0.0, 0, 0.0, 0.0, 0])
        e[0] += dt; e[1] += obj.n_evals - n0; e[2] += ht1[0] - ht0[0]; e[3] += ht1[1] - ht0[1]; e[4] += 1
        return out
    return wrapper


sol._run_theta_solver = timed('theta', orig_theta)
sol._run_handover_solver = timed('handover', orig_ho)
t_stage = t_solve = 0.0
for k in range(a.windows):
    t0 = time.perf_counter()
    sol.set_datasample(*wins[(k + 1) % 4].args())
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    sol.solve()
    t2 = time.perf_counter()
    t_stage += t1 - t0; t_solve += t2 - t1
nw = a.windows
print(f'{a.workload}: per window: set_datasample {t_stage / nw * 1e3:.2f} ms, solve {t_solve / nw * 1e3:.2f} ms')
tot = 0.0
for (kind, lvl), (dt, ne, tl, tw, calls) in sorted(acc.items(), key=lambda kv: (-kv[0][1], kv[0][0])):
    tot += dt
    print(f'  level {lvl} {kind:8s}: {dt / nw * 1e3:7.2f} ms/window, {calls / nw:4.1f} calls, {ne / nw:6.1f} evals, per eval: wall {dt / max(ne, 1) * 1e6:6.0f} us '
          f'(launch {tl / max(ne, 1) * 1e6:4.0f}, wait {tw / max(ne, 1) * 1e6:4.0f}, other host {(dt - tl - tw) / max(ne, 1) * 1e6:5.0f})')
print(f'  inside the level solvers {tot / nw * 1e3:.2f} ms/window; Python between them {(t_solve - tot) / nw * 1e3:.2f} ms/window')
