"""Per-kernel CUDA-event times of objective+gradient evaluations (eincm_plan_set_timing) on a few windows, one stream.
usage: python profiles/kernel_times.py [--workload dsec] [--theta 16] [--evals 40] [--windows 4] [--point perturbed]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from eincm_b200 import plan as P, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='dsec')
ap.add_argument('--events', type=int, default=None)
ap.add_argument('--theta', type=int, default=16)
ap.add_argument('--evals', type=int, default=40)
ap.add_argument('--windows', type=int, default=4)
ap.add_argument('--point', default='perturbed')
ap.add_argument('--const-flow', type=float, default=None, help='theta = constant flow (c, -c) px/window instead of a test point')
ap.add_argument('--lib', default=None, help='alternative build of the library (A/B comparisons)')
a = ap.parse_args()
if a.lib:
    P.LIB_PATH = os.path.abspath(a.lib)
torch.cuda.set_device(0)
wins = [synth.make_workload(a.workload, seed=k, n_events=a.events) for k in range(a.windows)]
hpd = wins[0].hparams
hp = P.make_hparams(hpd['alpha'], hpd['beta'], hpd['gamma'], hpd['delta'], 0)
plans, thetas = [], []
for w in wins:
    p = P.Plan(w.sensor_size, max_events=len(w.xs), max_refs=max(len(w.edge_ts), 3))
    p.set_window(*w.args())
    plans.append(p)
    th = synth.theta_test_points(w, (a.theta, a.theta))[a.point]
    if a.const_flow is not None:
        th = th * 0.0 + np.array([a.const_flow, -a.const_flow])
    thetas.append(torch.from_numpy(th).cuda())
loss = torch.zeros(1, dtype=torch.float64, device='cuda')
grad = torch.zeros((a.theta, a.theta, 2), dtype=torch.float64, device='cuda')
for i in range(2 * a.windows):
    plans[i % a.windows].value_and_grad_device(thetas[i % a.windows], hp, loss, grad)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.evals):
    plans[i % a.windows].value_and_grad_device(thetas[i % a.windows], hp, loss, grad)
e1.record()
torch.cuda.synchronize()
total = e0.elapsed_time(e1) / a.evals
for p in plans:
    p.set_timing(True)
for i in range(a.evals):
    plans[i % a.windows].value_and_grad_device(thetas[i % a.windows], hp, loss, grad)
torch.cuda.synchronize()
kt = {}
for p in plans:
    for name, (ms, n) in p.get_timing().items():
        t = kt.setdefault(name, [0.0, 0])
        t[0] += ms; t[1] += n
print(f'{a.workload} theta {a.theta} {a.point if a.const_flow is None else a.const_flow}: {total * 1e3:.1f} us/eval (one stream);',
      ', '.join(f'{k} {v[0] / v[1] * 1e3:.1f}' for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0])), f'; loss {float(loss.item()):.12g}')
