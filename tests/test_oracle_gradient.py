"""The oracle's hand-written reverse mode vs (a) torch autograd of an op-by-op forward mirror and
(b) central finite differences."""
import numpy as np
import pytest

import eincm_b200.synth as S
from oracle import eincm_oracle as O
from tests._torch_autodiff import value_and_grad_torch


@pytest.fixture(scope='module')
def win():
    return S.make_workload('tiny', seed=0)


CASES = [  # gamma, delta, cur_pyr_lvl
    (0.0, 0.0, 1), (0.0025, 0.0, 0), (0.0, 0.3, 2), (0.0025, 0.3, 0), (0.0025, 0.0, 3)]


@pytest.mark.parametrize('shape', [(1, 1), (2, 2), (4, 4), (16, 16), (48, 64)])
@pytest.mark.parametrize('gamma,delta,lvl', CASES)
def test_backward_matches_torch_autograd(win, shape, gamma, delta, lvl):
    pts = S.theta_test_points(win, shape)
    for name in ('truth', 'perturbed'):      # theta = 0 is a kink of |.| (delta term): see test below
        th = pts[name]
        l, g = O.value_and_grad(th, *win.args(), 20.0, 35.0, gamma, delta, lvl, 5, win.sensor_size)
        lt, gt = value_and_grad_torch(th, *win.args(), 20.0, 35.0, gamma, delta, lvl, win.sensor_size)
        assert l == pytest.approx(lt, rel=1e-12)
        assert np.abs(g - gt).max() <= 1e-10 * np.abs(gt).max()


def test_backward_at_zero_theta(win):
    th = np.zeros((4, 4, 2))
    l, g = O.value_and_grad(th, *win.args(), 20.0, 35.0, 0.0, 0.0, 1, 5, win.sensor_size)
    lt, gt = value_and_grad_torch(th, *win.args(), 20.0, 35.0, 0.0, 0.0, 1, win.sensor_size)
    assert l == pytest.approx(lt, rel=1e-13)
    assert np.abs(g - gt).max() <= 1e-10 * np.abs(gt).max()


@pytest.mark.parametrize('shape', [(1, 1), (2, 2), (4, 4)])
def test_backward_matches_finite_differences(win, shape):
    th = S.theta_test_points(win, shape)['perturbed']
    kw = dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.0, cur_pyr_lvl=1, n_pyr_lvls=5, sensor_size=win.sensor_size)
    _, g = O.value_and_grad(th, *win.args(), **kw)
    fd = np.zeros_like(th)
    h = 1e-6    # the objective is only piecewise smooth (rint, min/max): keep the step tiny
    for idx in np.ndindex(th.shape):
        e = np.zeros_like(th); e[idx] = h
        fd[idx] = (O.loss_func(th + e, *win.args(), **kw)[0] - O.loss_func(th - e, *win.args(), **kw)[0]) / (2 * h)
    assert np.abs(g - fd).max() <= 2e-5 * np.abs(fd).max()


def test_handover_gradient_is_dot_product(win):
    pts = S.theta_test_points(win, (4, 4))
    prev, cur = pts['truth'], pts['perturbed']
    kw = dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.0, cur_pyr_lvl=1, n_pyr_lvls=5, sensor_size=win.sensor_size)
    a0 = 0.37
    l, da = O.handover_value_and_grad(a0, prev, cur, *win.args(), **kw)
    assert l == pytest.approx(O.handover_loss_func(a0, prev, cur, *win.args(), **kw), rel=1e-14)
    h = 1e-6
    fd = (O.handover_loss_func(a0 + h, prev, cur, *win.args(), **kw)
          - O.handover_loss_func(a0 - h, prev, cur, *win.args(), **kw)) / (2 * h)
    assert da == pytest.approx(fd, rel=2e-5)


def test_wrap_quirk_switch_changes_border_only(win):
    # events pushed past the left/top border: the two index rules differ, in-sensor warps agree
    th = np.zeros((1, 1, 2)); th[..., 0] = 80.0; th[..., 1] = 60.0
    kw = dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.0, cur_pyr_lvl=1, n_pyr_lvls=5, sensor_size=win.sensor_size)
    la, _ = O.value_and_grad(th, *win.args(), wrap_negative=True, **kw)
    lb, _ = O.value_and_grad(th, *win.args(), wrap_negative=False, **kw)
    assert la != lb
    small = np.zeros((1, 1, 2))
    la, ga = O.value_and_grad(small, *win.args(), wrap_negative=True, **kw)
    lt, gt = value_and_grad_torch(small, *win.args(), 20.0, 35.0, 0.0, 0.0, 1, win.sensor_size, wrap_negative=True)
    assert la == pytest.approx(lt, rel=1e-13)
