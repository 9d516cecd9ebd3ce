"""The C restatement is the checker of the full-size GPU parity tests (tests/test_gpu_fullsize.py) and of bench.py's
``check``; here it is pinned to the NumPy oracle (which carries the hand-derived pins) on the SHAPES those tests use -
DSEC 640x480 with tile theta, MVSEC 256x336 / 260x346 with dense theta, 1280x720 - at event counts the NumPy oracle
finishes in seconds (N >= 200 k where the shape's full size is larger)."""
import os
import subprocess

import numpy as np
import pytest

import eincm_b200.synth as S
from oracle import c_oracle as C
from oracle import eincm_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module', autouse=True)
def built():
    if not C.available():
        subprocess.run(['make', '-C', os.path.join(ROOT, 'oracle')], check=True, stdout=subprocess.DEVNULL)
    assert C.available()


def _check(win, th, lvl):
    hp = win.hparams
    l, g = O.value_and_grad(th, *win.args(), hp['alpha'], hp['beta'], hp['gamma'], hp['delta'], lvl, 5, win.sensor_size)
    lc, gc, _ = C.value_and_grad_raw(th, *win.args(), hp['alpha'], hp['beta'], hp['gamma'], hp['delta'], lvl, win.sensor_size)
    assert lc == pytest.approx(l, rel=1e-11)
    assert np.abs(gc - g).max() <= 1e-9 * np.abs(g).max()


@pytest.mark.parametrize('point', ['zero', 'perturbed'])
def test_dsec_shape_tile_theta_200k(point):
    win = S.make_workload('dsec', seed=0, n_events=200_000)
    _check(win, S.theta_test_points(win, (16, 16))[point], 0)


def test_large_shape_tile_theta_200k():
    win = S.make_workload('large', seed=0, n_events=200_000)
    _check(win, S.theta_test_points(win, (16, 16))['perturbed'], 0)


@pytest.mark.parametrize('name', ['mvsec_dt1', 'mvsec_dt4', 'mvsec_raw_dt4'])
def test_mvsec_shapes_dense_theta(name):
    win = S.make_workload(name, seed=0)                    # the shape's full size: 30 k events
    H, W = win.sensor_size
    _check(win, S.theta_test_points(win, (H, W))['perturbed'], 0)
    _check(win, S.theta_test_points(win, (16, 16))['perturbed'], 0)


def test_dsec_shape_large_flow_200k():
    """60 px / window: the GPU path slices its shared-memory windows here (tests/test_gpu_fullsize.py)."""
    win = S.make_workload('dsec', seed=1, n_events=200_000)
    th = np.zeros((16, 16, 2))
    th[..., 0] = 60.0
    th[..., 1] = -47.0
    _check(win, th, 0)
