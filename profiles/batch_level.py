"""One batched device-side level solve (eincm_batch_minimize_bfgs_graph_host) of B MVSEC-shaped windows, timed, plus the per-kernel spans of
the batched evaluation at the same theta shape (eincm_batch_set_timing).  Under ncu (-k regex:'_b$|k_bfgs') the launch list of the solve graph.
usage: python profiles/batch_level.py [--batch 256] [--level 0] [--maxiter 10]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from eincm_b200 import plan as P, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='mvsec_dt4')
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--level', type=int, default=0)
ap.add_argument('--maxiter', type=int, default=10)
ap.add_argument('--repeat', type=int, default=2)
a = ap.parse_args()
torch.cuda.set_device(0)
wins = [synth.make_sequence(a.workload, 1, seed=t)[0] for t in range(a.batch)]
w0 = wins[0]
H, W = w0.sensor_size
N, R = len(w0.xs), len(w0.edge_ts)
plans = []
for w in wins:
    p = P.Plan((H, W), max_events=N, max_refs=max(R, 3))
    p.set_window(*w.args())
    plans.append(p)
b = P.Batch(plans)
side = 16 >> a.level
hp = P.make_hparams(w0.hparams['alpha'], w0.hparams['beta'], 0.0, 0.0, a.level)
th0 = np.stack([0.5 * synth.theta_test_points(w, (side, side))['truth'] for w in wins])
for rep in range(a.repeat):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th, res = b.minimize_bfgs_graph_host(th0, hp, a.maxiter, 1e-7)
    dt = time.perf_counter() - t0
    nf = np.array([r.nfev for r in res])
    steps = b.solve_launches() * int(os.environ.get('EINCM_GRAPH_UNROLL', '8'))
    print(f'level {a.level} theta {side}x{side}, {a.batch} windows, maxiter {a.maxiter}: {dt * 1e3:.1f} ms, {b.solve_launches()} graph launches (<= {steps} steps), '
          f'evaluations per window {nf.mean():.1f} (min {nf.min()}, max {nf.max()}): {dt / nf.sum() * 1e6:.1f} us per window evaluation, '
          f'{dt / steps * 1e3:.2f} ms per step', flush=True)
# the batched evaluation alone at this theta shape
ths = [torch.from_numpy(t).cuda() for t in th0]
losses = [torch.zeros(1, dtype=torch.float64, device='cuda') for _ in plans]
grads = [torch.zeros_like(t) for t in ths]
for _ in range(3):
    b.value_and_grad_device(ths, hp, losses, grads)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    b.value_and_grad_device(ths, hp, losses, grads)
e1.record()
torch.cuda.synchronize()
print(f'batched evaluation alone: {e0.elapsed_time(e1) / 5:.3f} ms per step ({e0.elapsed_time(e1) / 5 / a.batch * 1e3:.1f} us per window)')
b.set_timing(True)
for _ in range(3):
    b.value_and_grad_device(ths, hp, losses, grads)
print('   spans:', {k: round(v[0] / v[1] * 1e3, 1) for k, v in b.get_timing().items()}, 'us per launch')
b.close()
for p in plans:
    p.close()
