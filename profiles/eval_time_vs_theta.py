"""Per-evaluation wall time against the flow magnitude of the trial point, over complete multi-level solves driven by scipy (the
Python callback sees every evaluation): which evaluations of a BFGS line search are the slow ones.
usage: python profiles/eval_time_vs_theta.py [--windows 3]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from eincm_b200 import losses, solver as SV, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='dsec')
ap.add_argument('--windows', type=int, default=3)
a = ap.parse_args()
torch.cuda.set_device(0)
seq = synth.make_sequence(a.workload, 1 + a.windows, seed=0)
H, W = seq[0].sensor_size
hpd = seq[0].hparams
obj = losses.WindowObjective((H, W), hpd['alpha'], hpd['beta'], hpd['gamma'], hpd['delta'], max_events=len(seq[0].xs), max_refs=3)
rec = []
orig = obj.plan.value_and_grad_host


def timed(theta, hp, *args, **kw):
    t0 = time.perf_counter()
    out = orig(theta, hp, *args, **kw)
    rec.append((theta.shape[0], float(np.abs(theta).max()), (time.perf_counter() - t0) * 1e6))
    return out


obj.plan.value_and_grad_host = timed
sol = SV.MultipleLevelEINCMSolver(obj, backend='scipy')
sol.set_datasample(*seq[0].args()); sol.solve()
rec.clear()
for k in range(1, 1 + a.windows):
    sol.set_datasample(*seq[k].args()); sol.solve()
r = np.array(rec)
print(f'{len(r) / a.windows:.0f} evaluations per window, {r[:, 2].sum() / a.windows / 1e3:.1f} ms of evaluations per window')
edges = [0, 25, 50, 100, 200, 400, 1e3, 1e4, 1e9]
for lo, hi in zip(edges[:-1], edges[1:]):
    m = (r[:, 1] >= lo) & (r[:, 1] < hi)
    if m.any():
        print(f'  max|theta| in [{lo:g}, {hi:g}) px: {m.sum() / a.windows:6.1f} evals/window, mean {r[m, 2].mean():8.1f} us, max {r[m, 2].max():9.1f} us, '
              f'{r[m, 2].sum() / r[:, 2].sum() * 100:5.1f} % of the evaluation time')
