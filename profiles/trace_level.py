"""Trace of the scipy BFGS solve of one pyramid level: every evaluation (max|theta|, loss, max|grad|, step from the level's start)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, scipy.optimize
from eincm_b200 import losses, synth
torch.cuda.set_device(0)
win = synth.make_sequence('dsec', 1, seed=0)[0]
H, W = win.sensor_size
hpd = win.hparams
obj = losses.WindowObjective((H, W), hpd['alpha'], hpd['beta'], hpd['gamma'], hpd['delta'], max_events=len(win.xs), max_refs=3)
obj.set_datasample(*win.args())
print('truth theta', np.round(win.truth_theta.reshape(-1, 2), 2).tolist())
x = np.zeros((1, 1, 2))
for lvl, shape in ((4, (1, 1)), (3, (2, 2))):
    if x.shape[:2] != shape:
        x = np.repeat(np.repeat(x, 2, axis=0), 2, axis=1)
    fun = obj.scipy_fun(shape + (2,), lvl)
    log = []
    def f(v):
        val, g = fun(v)
        log.append((float(np.abs(v).max()), val, float(np.abs(g).max()), float(np.abs(v - x.ravel()).max())))
        return val, g
    res = scipy.optimize.minimize(f, x.ravel(), jac=True, method='BFGS', options={'maxiter': 11 if lvl == 3 else 8, 'gtol': 1e-7})
    print(f'level {lvl}: status {res.status} nit {res.nit} nfev {res.nfev} fun {res.fun:.4f} x {np.round(res.x, 2).tolist()}')
    for i, (m, val, g, d) in enumerate(log):
        print(f'   eval {i:3d}: max|theta| {m:10.3f}  loss {val:14.6f}  max|grad| {g:12.5f}  max|theta - start| {d:10.3f}')
    x = res.x.reshape(shape + (2,))
