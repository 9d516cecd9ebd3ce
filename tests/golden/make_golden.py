"""Generates the golden fixtures under tests/golden/ (committed together with this script).

    python tests/golden/make_golden.py

PARITY UNPINNED: the reference is pure Python on JAX / jaxopt, which cannot be imported in this image (no jax, jaxlib, jaxopt
wheels; no network), so these vectors are outputs of the float64 NumPy restatement (oracle/eincm_oracle.py), NOT of the reference
itself.  They pin the restatement against regressions and give the C oracle and the CUDA path fixed targets that do not depend
on the oracle code at test time.  Since round 2 they are cross-checked against the reference's own source executed over a float64 stand-in for the JAX primitives
(tests/golden/make_golden_refsrc.py -> tests/golden/refsrc/, tests/test_reference_source.py): same numbers to 1e-16.  That pins the
composition, not JAX's primitives.  If a machine with JAX becomes available, regenerate them from the real
``jit(value_and_grad(loss_func))`` with the same seeds (the inputs are stored in the files) and commit the result.

Each case stores its inputs (xs, ys int16; ts float64; edges; edge_ts; theta; hyper-parameters) and, from the oracle: loss, gradient,
the images of warped events, the rounded pixel index stream per reference time, d loss / d IWE, the handover value and
d/d alpha, and the evaluation metrics of src/evaluations/theta_eval.py with a synthetic ground-truth flow.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from eincm_b200 import synth  # noqa: E402
from oracle import eincm_oracle as O  # noqa: E402

CASES = [
    # name, (H, W, N, edge_ts), theta shape, theta point, hyper-parameters
    ('tiny_4x4_lvl1', dict(H=48, W=64, N=4000, edge_ts=(0.0, 0.5, 1.0)), (4, 4), 'perturbed', dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.0, cur_pyr_lvl=1)),
    ('tiny_1x1_lvl4', dict(H=48, W=64, N=4000, edge_ts=(0.0, 0.5, 1.0)), (1, 1), 'truth', dict(alpha=2000.0, beta=4000.0, gamma=0.0, delta=0.0, cur_pyr_lvl=4)),
    ('tiny_16x16_tv_div', dict(H=48, W=64, N=4000, edge_ts=(0.0, 0.25, 0.5, 0.75, 1.0)), (16, 16), 'perturbed', dict(alpha=20.0, beta=35.0, gamma=0.0025, delta=0.3, cur_pyr_lvl=0)),
    ('ragged_dense_r2', dict(H=37, W=53, N=2500, edge_ts=(0.0, 1.0)), (37, 53), 'perturbed', dict(alpha=60.0, beta=60.0, gamma=0.0025, delta=0.0, cur_pyr_lvl=0)),
]


def main():
    for k, (name, wk, shape, point, hp) in enumerate(CASES):
        win = synth.make_window(seed=100 + k, **wk)
        pts = synth.theta_test_points(win, shape, seed=k)
        theta = pts[point]
        kw = dict(n_pyr_lvls=5, sensor_size=win.sensor_size, scale_to_sensor_size_method='bilinear', **hp)
        loss, grad, inter = O.value_and_grad(theta, *win.args(), return_intermediates=True, **kw)
        prev = pts['truth'] if point != 'truth' else pts['zero']
        a_ho = 0.37
        ho_loss, ho_dalpha = O.handover_value_and_grad(a_ho, prev, theta, *win.args(), **kw)
        theta_full = O.scale_theta_to_sensor_size(theta, win.sensor_size)
        obj = inter['objectives']
        rounded = np.stack([np.stack(O.rounded_event_pixels(obj['warped_xs'][r], obj['warped_ys'][r])) for r in range(len(win.edge_ts))])
        gt_flow = O.scale_theta_to_sensor_size(pts['truth'], win.sensor_size).copy()
        rng = np.random.default_rng(7 + k)
        gt_flow[rng.random(gt_flow.shape[:2]) < 0.15] = 0.0           # invalid ground truth (flow_eval.py:42-46)
        gt_flow[0, 0] = np.inf
        err_mask = rng.random(gt_flow.shape[:2]) < 0.8
        ev = O.evaluate_theta_array(theta_full, *win.args(), gt_flow, hp['alpha'], hp['beta'], hp['gamma'], hp['delta'], win.sensor_size,
                                    err_eval_event_mask=err_mask)
        ev_keys = sorted(k2 for k2, v in ev.items() if np.isscalar(v))
        out = dict(xs=win.xs, ys=win.ys, ts=win.ts, edges=win.edges, edge_ts=np.asarray(win.edge_ts), theta=theta, prev_theta=prev,
                   alpha_handover=a_ho, hp_names=np.array(sorted(hp)), hp_values=np.array([float(hp[n]) for n in sorted(hp)]),
                   loss=loss, grad=grad, iwes=inter['iwes'], zero_iwe=inter['zero_iwe'], dLdI=inter['dLdI'], rounded=rounded.astype(np.int32),
                   handover_loss=ho_loss, handover_dalpha=ho_dalpha, gt_flow=gt_flow, err_mask=err_mask,
                   eval_names=np.array(ev_keys), eval_values=np.array([float(ev[k2]) for k2 in ev_keys]),
                   eval_flow_warp_losses=ev['flow_warp_losses'], eval_rel_contrasts=ev['rel_contrasts'],
                   eval_rel_correlations=ev['rel_correlations'], eval_rel_iwe_divergences=ev['rel_iwe_divergences'])
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **out)
        print(f'{name}: loss {loss:.15g}, |grad|inf {np.abs(grad).max():.6g}, {os.path.getsize(path) / 1024:.0f} KB')


if __name__ == '__main__':
    main()
