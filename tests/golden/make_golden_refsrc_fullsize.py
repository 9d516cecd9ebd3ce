"""Reference-source vectors AT THE BENCHMARKED CONFIGURATIONS (see make_golden_refsrc.py for how the reference's unmodified
eincm.losses is executed here): loss and gradient of

    dsec        640 x 480, N = 2 M events, R = 3, theta 16 x 16   (the BENCH line: bench.py make_windows, rank 0, window 0 = seed 0)
    mvsec_dt4   346 x 260, N = 30 k, R = 5, dense theta           (BASELINE.json configs[2])

at the 'perturbed' test point.  The windows are regenerated from their seed by eincm_b200.synth (too large to commit); a checksum of
the operands is stored so that a test can tell a regenerated window from a different one.

    python tests/golden/make_golden_refsrc_fullsize.py          (needs /root/reference; ~1 min, ~10 GB of host memory)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from eincm_b200 import synth  # noqa: E402
import make_golden_refsrc as M  # noqa: E402

OUT = os.path.join(HERE, 'refsrc_fullsize')
CASES = [  # file, workload, seed, theta shape (None: dense), pyramid level
    ('dsec_2m_theta16', 'dsec', 0, (16, 16), 0),
    ('mvsec_dt4_dense', 'mvsec_dt4', 0, None, 0),
]


def checksum(win):
    return np.array([float(np.asarray(win.xs, dtype=np.float64).sum()), float(np.asarray(win.ys, dtype=np.float64).sum()),
                     float(np.asarray(win.ts).sum()), float(np.asarray(win.edges).sum()), float(len(win.xs))])


def case_inputs(workload, seed, shape):
    win = synth.make_workload(workload, seed=seed)
    shape = tuple(win.sensor_size) if shape is None else shape
    return win, synth.theta_test_points(win, shape)['perturbed']


def main():
    jax, L = M.import_reference()
    import jax.numpy as jnp
    os.makedirs(OUT, exist_ok=True)
    for name, workload, seed, shape, lvl in CASES:
        win, theta = case_inputs(workload, seed, shape)
        hp = win.hparams
        xs, ys, ts, edges, edge_ts = (jnp.array(np.asarray(a)) for a in win.args())
        t0 = time.time()
        (loss, _), grad = jax.value_and_grad(
            lambda th: L.loss_func(th, xs, ys, ts, edges, edge_ts, alpha=hp['alpha'], beta=hp['beta'], gamma=hp['gamma'], delta=hp['delta'],
                                   cur_pyr_lvl=lvl, n_pyr_lvls=5, sensor_size=tuple(win.sensor_size), scale_to_sensor_size_method='bilinear'),
            has_aux=True)(jnp.array(theta))
        np.savez_compressed(os.path.join(OUT, name + '.npz'), loss=float(loss), grad=np.asarray(grad), theta=theta, checksum=checksum(win),
                            workload=workload, seed=seed, cur_pyr_lvl=lvl)
        print(f'{name}: N = {len(win.xs)}, loss {float(loss):.12g}, |grad|inf {np.abs(np.asarray(grad)).max():.6g}, {time.time() - t0:.1f} s')


if __name__ == '__main__':
    main()
