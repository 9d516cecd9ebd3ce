"""The reference objective on a synthetic window along rays in theta space: where it has its minima.  Prints the loss and its two
terms (mean relative contrast, mean relative "correlation" = MSE ratio - src/eincm/losses.py:171-193) at the zero flow, at the truth
and at multiples of the truth flow / of a wrong direction."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from eincm_b200 import plan as P, synth
torch.cuda.set_device(0)
name = sys.argv[1] if len(sys.argv) > 1 else 'dsec'
kw = dict(a.split('=') for a in sys.argv[2:])
kw = {k: float(v) for k, v in kw.items()}
if 'n_segments' in kw: kw['n_segments'] = int(kw['n_segments'])
rs = np.random.default_rng(10_000)
theta0 = rs.uniform(-20, 20, size=(2, 2, 2))
win = synth.make_window(480, 640, 2_000_000, (0.0, 0.5, 1.0), seed=1, scene_seed=20_000, truth_theta=theta0, hparams=dict(alpha=2000.0, beta=4000.0, gamma=0.0, delta=0.0), **kw)
H, W = win.sensor_size
hp = P.make_hparams(win.hparams['alpha'], win.hparams['beta'], 0.0, 0.0, 1)
p = P.Plan((H, W), max_events=len(win.xs), max_refs=3)
p.set_window(*win.args())
truth = win.truth_theta                      # (2, 2, 2)
rng = np.random.default_rng(1)
wrong = rng.normal(size=truth.shape); wrong *= np.abs(truth).max() / np.abs(wrong).max()
def show(tag, th):
    l, g = p.value_and_grad_host(np.ascontiguousarray(th), hp)
    s = p.scalars()
    print(f'{tag:28s} loss {l:12.3f}  rel contrast {s["mean_rel_contrast"] * 3:8.4f}  rel mse {s["mean_rel_corr"] * 3:8.4f}  |grad|inf {np.abs(g).max():10.3f}   (zero mse {-s["zero_correlations"][0]:.5f}, mse {-s["correlations"][0]:.5f})')
for k in (0.0, 0.5, 0.9, 1.0, 1.1, 1.5, 2.0, 4.0, 10.0, 30.0, 100.0):
    show(f'{k:g} x truth', k * truth)
for k in (0.25, 0.5, 1.0, 2.0, 4.0, 10.0, 30.0, 100.0):
    show(f'{k:g} x wrong direction', k * wrong)
