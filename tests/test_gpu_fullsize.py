"""Parity of the CUDA path against the C restatement of the reference AT THE SIZES bench.py measures (BASELINE.json
configs): DSEC 640x480 with N = 2 M (the bench line) and 5 M, 1280x720 at 5 M events, MVSEC dt = 1 / dt = 4 with dense
theta, and a 60 px flow at full N (sliced shared-memory windows).  The NumPy oracle is too slow here; the C oracle is
pinned to it on the same shapes by tests/test_oracle_c_bench_shapes.py.

Tolerances are BASELINE.json's: objective <= 1e-5 relative, gradient <= 1e-4 relative (inf-norm / inf-norm)."""
import os
import subprocess

import numpy as np
import pytest

import eincm_b200.synth as S
from oracle import c_oracle as C

pytestmark = pytest.mark.gpu

OBJ_RTOL = 1e-5      # BASELINE.json north_star
GRAD_RTOL = 1e-4     # BASELINE.json north_star
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module', autouse=True)
def built():
    if not C.available():
        subprocess.run(['make', '-C', os.path.join(ROOT, 'oracle')], check=True, stdout=subprocess.DEVNULL)
    assert C.available()
    C.set_num_threads(len(os.sched_getaffinity(0)))


def _rel_inf(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def _compare(win, thetas, lvl=0):
    from eincm_b200 import plan as P
    hp = win.hparams
    R = len(win.edge_ts)
    p = P.Plan(win.sensor_size, max_events=len(win.xs), max_refs=max(R, 3))
    try:
        p.set_window(*win.args())
        hpc = P.make_hparams(hp['alpha'], hp['beta'], hp['gamma'], hp['delta'], lvl)
        for th in thetas:
            loss, grad = p.value_and_grad_host(th, hpc)
            l_ref, g_ref, _ = C.value_and_grad_raw(th, *win.args(), hp['alpha'], hp['beta'], hp['gamma'], hp['delta'], lvl, win.sensor_size)
            assert abs(loss - l_ref) <= OBJ_RTOL * abs(l_ref), (loss, l_ref)
            assert _rel_inf(grad, g_ref) <= GRAD_RTOL, _rel_inf(grad, g_ref)
    finally:
        p.close()


def test_dsec_bench_configuration_2m():
    """The configuration of the BENCH line: dsec, N = 2 M, R = 3, theta 16x16, seed 0 (bench.py make_windows, rank 0 window 0)."""
    win = S.make_workload('dsec', seed=0)
    pts = S.theta_test_points(win, (16, 16))
    _compare(win, [pts['zero'], pts['perturbed']])


@pytest.mark.parametrize('shape', [(1, 1), (4, 4)])
def test_dsec_2m_coarse_levels(shape):
    win = S.make_workload('dsec', seed=1)
    _compare(win, [S.theta_test_points(win, shape)['perturbed']], lvl=2)


def test_dsec_5m():
    win = S.make_workload('dsec_5m', seed=0)
    pts = S.theta_test_points(win, (16, 16))
    _compare(win, [pts['zero'], pts['perturbed']])


def test_large_1280x720_5m():
    win = S.make_workload('large', seed=0, n_events=5_000_000)
    pts = S.theta_test_points(win, (16, 16))
    _compare(win, [pts['zero'], pts['perturbed']])


@pytest.mark.parametrize('name', ['mvsec_dt1', 'mvsec_dt4', 'mvsec_raw_dt4'])
def test_mvsec_dense_theta(name):
    win = S.make_workload(name, seed=0)
    H, W = win.sensor_size
    dense = S.theta_test_points(win, (H, W))
    tile = S.theta_test_points(win, (16, 16))
    _compare(win, [dense['zero'], dense['perturbed'], tile['perturbed']])


def test_dsec_2m_large_flow_sliced_windows():
    """60 px / window at full N: destination rectangles exceed a shared-memory window and are processed in row slices."""
    win = S.make_workload('dsec', seed=1)
    th = np.zeros((16, 16, 2))
    th[..., 0] = 60.0
    th[..., 1] = -47.0
    rng = np.random.default_rng(5)
    _compare(win, [th, th + rng.normal(0.0, 3.0, size=th.shape)])
