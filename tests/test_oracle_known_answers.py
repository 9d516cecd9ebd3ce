"""Pins for the CPU oracle (the reference ships no tests / golden vectors: SURVEY.md §4, §8c).

Known answers are derived by hand from the reference source; primitive restatements are checked
against scipy / torch implementations of the same published JAX semantics."""
import math

import numpy as np
import pytest
import torch
from scipy.signal import convolve2d, correlate2d
from scipy.stats import multivariate_normal, norm

import eincm_b200.synth as S
from oracle import eincm_oracle as O


def test_single_event_patch():
    # event_utils.py:55-56: exp(-r^2/2)/(2*pi), r^2 in {0,1,2}
    f = O.events_to_pdf_frame(np.array([5.0]), np.array([7.0]), (16, 12))
    patch = f[6:9, 4:7]
    assert patch[1, 1] == pytest.approx(0.15915494309189535, rel=1e-15)
    assert patch[0, 1] == pytest.approx(0.09653235263005391, rel=1e-15)
    assert patch[0, 0] == pytest.approx(0.05854983152431917, rel=1e-15)
    assert f.sum() == pytest.approx(0.7794836797093877, rel=1e-14)
    assert np.count_nonzero(f) == 9


def test_tap_values_match_scipy_multivariate_normal():
    rng = np.random.default_rng(0)
    x = rng.uniform(3, 9, 50); y = rng.uniform(3, 9, 50)
    for k in range(50):
        f = O.events_to_pdf_frame(x[k:k + 1], y[k:k + 1], (16, 16))
        xr, yr = int(np.rint(x[k])), int(np.rint(y[k]))
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                q = np.array([xr + dx - x[k], yr + dy - y[k]])
                assert f[yr + dy, xr + dx] == pytest.approx(multivariate_normal.pdf(q, mean=[0, 0], cov=np.eye(2)), rel=1e-13)


def test_round_half_even_and_wrap_quirk():
    # jnp.round is half-to-even: 2.5 -> 2, 3.5 -> 4
    f = O.events_to_pdf_frame(np.array([2.5]), np.array([3.5]), (10, 10))
    rows, cols = np.nonzero(f)                                           # support: rows 4+-1 (3.5 -> 4), cols 2+-1 (2.5 -> 2)
    assert sorted(set(rows)) == [3, 4, 5] and sorted(set(cols)) == [1, 2, 3]
    xr, yr = O.rounded_event_pixels(np.array([2.5, 3.5, -0.5, -1.5]), np.array([0.5, 1.5, 2.5, -2.5]))
    assert xr.tolist() == [2, 4, 0, -2] and yr.tolist() == [0, 2, 2, -2]
    # event at x' in [-0.5, 0.5) => tap dx=-1 lands in column W-1 (negative index wraps)
    H, W = 8, 10
    f = O.events_to_pdf_frame(np.array([0.2]), np.array([4.0]), (H, W))
    assert f[4, W - 1] > 0 and f[4, 0] > 0 and f[4, 1] > 0
    f2 = O.events_to_pdf_frame(np.array([0.2]), np.array([4.0]), (H, W), wrap_negative=False)
    assert f2[4, W - 1] == 0
    # beyond -N the update is dropped; >= N is dropped (no wrap on the high side)
    assert O.events_to_pdf_frame(np.array([W + 0.6]), np.array([4.0]), (H, W)).sum() == 0
    f3 = O.events_to_pdf_frame(np.array([W - 0.6]), np.array([4.0]), (H, W))     # rounds to W-1... taps W-2, W-1, (W dropped)
    assert np.count_nonzero(f3) == 6
    f4 = O.events_to_pdf_frame(np.array([-W - 5.0]), np.array([4.0]), (H, W))
    assert f4.sum() == 0


def test_multi_ref_weights():
    # losses.py:39-46
    np.testing.assert_allclose(O.compute_weights_for_multi_reference(1), [1.0])
    np.testing.assert_allclose(O.compute_weights_for_multi_reference(2), [0.5, 0.5])
    np.testing.assert_allclose(O.compute_weights_for_multi_reference(3), [0.19684199, 0.60631602, 0.19684199], atol=5e-9)
    np.testing.assert_allclose(O.compute_weights_for_multi_reference(5),
                               [0.10277116, 0.23895011, 0.31655746, 0.23895011, 0.10277116], atol=5e-9)
    x = np.linspace(-1.5, 1.5, 4)
    np.testing.assert_allclose(O.compute_weights_for_multi_reference(4), norm.pdf(x) / norm.pdf(x).sum(), rtol=1e-15)


@pytest.mark.parametrize('m,n', [(16, 480), (16, 640), (8, 256), (8, 336), (1, 480), (2, 260), (2, 346), (4, 176), (480, 480)])
def test_resize_weights_match_torch_bilinear(m, n):
    # upsampling weights == torch bilinear, align_corners=False (half-pixel centres)
    Wm = O.compute_weight_mat(m, n, n / m)
    eye = torch.eye(m, dtype=torch.float64)[None]                        # (1, m, m): channel k = basis vector e_k
    up = torch.nn.functional.interpolate(eye, size=n, mode='linear', align_corners=False)[0].numpy()
    np.testing.assert_allclose(Wm, up, atol=1e-14)
    np.testing.assert_allclose(Wm.sum(axis=0), 1.0, atol=1e-15)
    assert (np.count_nonzero(Wm, axis=0) <= 2).all()


def test_scale_theta_constant_and_identity():
    th = np.array([[[1.25, -3.5]]])
    full = O.scale_theta_to_sensor_size(th, (12, 20))
    np.testing.assert_allclose(full[..., 0], 1.25, rtol=1e-15); np.testing.assert_allclose(full[..., 1], -3.5, rtol=1e-15)
    rng = np.random.default_rng(1)
    dense = rng.normal(size=(12, 20, 2))
    np.testing.assert_allclose(O.scale_theta_to_sensor_size(dense, (12, 20)), dense, atol=1e-15)


def test_scharr_canonical_equals_literal_and_adjoint():
    rng = np.random.default_rng(2)
    I = rng.normal(size=(37, 53))
    np.testing.assert_allclose(O.sobel_scharr_optimized_image_grads(I), O.sobel_scharr_literal(I), atol=1e-12)
    gx = rng.normal(size=I.shape); gy = rng.normal(size=I.shape)
    ref = correlate2d(gx, O.SCHARR_GX, mode='same') + correlate2d(gy, O.SCHARR_GY, mode='same')
    np.testing.assert_allclose(O._scharr_adjoint(gx, gy), ref, atol=1e-12)
    np.testing.assert_allclose(O.div_kern_conv(I), convolve2d(I, O.DIV_KERN, mode='same'), atol=1e-14)
    # exact zeros on constant fields away from the border (the property the TV nz-count relies on)
    g = O.sobel_scharr_optimized_image_grads(np.full((9, 9), 0.1234567))
    assert (g[1:-1, 1:-1] == 0).all()


@pytest.mark.parametrize('name', ['tiny'])
@pytest.mark.parametrize('R', [1, 2, 3, 5])
def test_loss_at_zero_theta(name, R):
    # loss(theta=0) = -(alpha+beta)/R for gamma=delta=0 (all IWE_r == zero-IWE, sum w_r = 1)
    w = S.make_window(48, 64, 3000, edge_ts=np.linspace(0, 1, R) if R > 1 else (0.0,), seed=3)
    alpha, beta = 20.0, 35.0
    for shape in [(1, 1), (4, 4)]:
        loss, _ = O.loss_func(np.zeros(shape + (2,)), *w.args(), alpha, beta, 0.0, 0.0, 1, 5, w.sensor_size)
        assert loss == pytest.approx(-(alpha + beta) / R, rel=1e-12)


def test_integer_translation_permutes_iwe():
    # constant integer flow, t_ref chosen so dt == 1 for all events => IWE is the zero-IWE shifted
    rng = np.random.default_rng(5)
    H, W, N = 40, 50, 500
    xs = rng.integers(12, W - 12, N).astype(np.int16); ys = rng.integers(12, H - 12, N).astype(np.int16)
    ts = np.ones(N)
    th = np.zeros((H, W, 2)); th[..., 0] = 3.0; th[..., 1] = -2.0
    xw, yw = O.per_pix_warp(th, xs, ys, ts, 0.0)
    z = O.events_to_pdf_frame(xs, ys, (H, W))
    i = O.events_to_pdf_frame(xw, yw, (H, W))
    np.testing.assert_allclose(i, np.roll(z, shift=(2, -3), axis=(0, 1)), atol=1e-15)


def test_event_split_additivity():
    w = S.make_workload('tiny', seed=7)
    th = S.theta_test_points(w, (4, 4))['perturbed']
    full = O.scale_theta_to_sensor_size(th, w.sensor_size)
    xw, yw = O.per_pix_warp(full, w.xs, w.ys, w.ts, 0.5)
    a = O.events_to_pdf_frame(xw[::2], yw[::2], w.sensor_size)
    b = O.events_to_pdf_frame(xw[1::2], yw[1::2], w.sensor_size)
    np.testing.assert_allclose(a + b, O.events_to_pdf_frame(xw, yw, w.sensor_size), atol=1e-13)


def test_maxiter_schedule_known_answer():
    # exp_mgr.py:177-184 with main.yaml defaults; used by the solver mirror
    from eincm_b200.solver import growing_maxiters
    assert growing_maxiters(8, 40, 5, 1.413) == {0: 40, 1: 28, 2: 19, 3: 11, 4: 8}
    assert growing_maxiters(4, 20, 5, 1.413) == {0: 20, 1: 14, 2: 10, 3: 6, 4: 4}
