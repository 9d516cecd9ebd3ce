"""Randomised (hypothesis) property tests of the CPU oracles: the size-independent properties SURVEY.md 8c lists as pins for a
restatement whose reference cannot be executed here - integer translation permutes the image of warped events, theta = 0 gives the
zero-warp image at every reference time, the image is additive in the events, the handover gradient is the projection of the theta
gradient, the gradient matches central differences - plus invariants of the edge-image restatement."""
import numpy as np
from hypothesis import given, settings, strategies as st

import eincm_b200.synth as S
from oracle import edge_oracle as E
from oracle import eincm_oracle as O

FAST = settings(max_examples=12, deadline=None, derandomize=True)


@FAST
@given(dx=st.integers(-6, 6), dy=st.integers(-6, 6), seed=st.integers(0, 1000))
def test_integer_translation_permutes_the_image(dx, dy, seed):
    rng = np.random.default_rng(seed)
    H, W, N = 36, 44, 300
    xs = rng.integers(10, W - 10, N).astype(np.int16); ys = rng.integers(10, H - 10, N).astype(np.int16)
    th = np.zeros((H, W, 2)); th[..., 0] = dx; th[..., 1] = dy
    xw, yw = O.per_pix_warp(th, xs, ys, np.ones(N), 0.0)                      # dt = 1 for every event
    np.testing.assert_allclose(O.events_to_pdf_frame(xw, yw, (H, W)),
                               np.roll(O.events_to_pdf_frame(xs, ys, (H, W)), shift=(-dy, -dx), axis=(0, 1)), atol=1e-14)


@FAST
@given(R=st.integers(1, 5), shape=st.sampled_from([(1, 1), (2, 3), (8, 8)]), seed=st.integers(0, 1000))
def test_zero_theta_gives_the_zero_warp_image_and_the_known_loss(R, shape, seed):
    w = S.make_window(32, 40, 800, edge_ts=np.linspace(0, 1, R) if R > 1 else (0.0,), seed=seed)
    loss, aux = O.loss_func(np.zeros(shape + (2,)), *w.args(), 20.0, 35.0, 0.0, 0.0, 1, 5, w.sensor_size)
    assert abs(loss + 55.0 / R) <= 1e-11 * 55.0 / R                            # -(alpha + beta) / R, sum of the weights = 1


@FAST
@given(seed=st.integers(0, 1000), frac=st.floats(0.1, 0.9), t_ref=st.floats(0.0, 1.0))
def test_image_of_warped_events_is_additive_in_the_events(seed, frac, t_ref):
    w = S.make_window(32, 40, 1200, seed=seed)
    th = S.theta_test_points(w, (4, 4), seed=seed)['perturbed']
    full = O.scale_theta_to_sensor_size(th, w.sensor_size)
    xw, yw = O.per_pix_warp(full, w.xs, w.ys, w.ts, t_ref)
    k = int(frac * len(xw))
    a = O.events_to_pdf_frame(xw[:k], yw[:k], w.sensor_size)
    b = O.events_to_pdf_frame(xw[k:], yw[k:], w.sensor_size)
    np.testing.assert_allclose(a + b, O.events_to_pdf_frame(xw, yw, w.sensor_size), atol=1e-12)


@FAST
@given(seed=st.integers(0, 1000), a=st.floats(0.05, 0.95))
def test_handover_gradient_is_the_projection_of_the_theta_gradient(seed, a):
    w = S.make_window(32, 40, 1500, seed=seed)
    pts = S.theta_test_points(w, (4, 4), seed=seed)
    prev, th = pts['truth'], pts['perturbed']
    kw = dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.0, cur_pyr_lvl=1, n_pyr_lvls=5, sensor_size=w.sensor_size)
    l_ho, d_alpha = O.handover_value_and_grad(a, prev, th, *w.args(), **kw)
    th_ho = a * prev + (1.0 - a) * th                                         # losses.py:269
    l, g = O.value_and_grad(th_ho, *w.args(), **kw)
    assert l_ho == l
    assert abs(d_alpha - float(np.sum(g * (prev - th)))) <= 1e-10 * max(1.0, abs(d_alpha))


@FAST
@given(seed=st.integers(0, 1000), idx=st.integers(0, 31))
def test_gradient_matches_central_differences(seed, idx):
    w = S.make_window(32, 40, 1500, seed=seed)
    th = S.theta_test_points(w, (4, 4), seed=seed)['perturbed']
    kw = dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.0, cur_pyr_lvl=1, n_pyr_lvls=5, sensor_size=w.sensor_size)
    _, g = O.value_and_grad(th, *w.args(), **kw)
    e = np.zeros(th.size); e[idx] = 1e-6
    e = e.reshape(th.shape)
    lp, _ = O.value_and_grad(th + e, *w.args(), **kw)
    lm, _ = O.value_and_grad(th - e, *w.args(), **kw)
    fd = (lp - lm) / 2e-6
    # rint() makes the objective piecewise smooth: a pixel crossing inside the 2e-6 bracket is possible but rare; min / max ties as well
    assert abs(fd - g.reshape(-1)[idx]) <= 2e-4 * max(1.0, np.abs(g).max())


@FAST
@given(seed=st.integers(0, 1000), shift=st.integers(1, 60), th=st.sampled_from([(30, 80), (100, 200), (5, 400)]))
def test_canny_ignores_a_brightness_offset_and_is_monotone_in_the_thresholds(seed, shift, th):
    f = S.make_frames(40, 56, 1, seed=seed)[0]
    f = np.clip(f, 0, 255 - shift).astype(np.uint8)
    base = E.canny(f, *th)
    assert np.array_equal(E.canny((f + shift).astype(np.uint8), *th), base)    # Sobel sees differences only
    looser = E.canny(f, th[0] * 0.5, th[1] * 0.5)
    assert np.all(looser[base > 0] > 0)                                      # lower thresholds keep every edge pixel
    m = E.canny_candidates(f, *th)
    assert np.all(m[base > 0] != 1) and np.all(base[m == 2] > 0)             # edges are candidates; strong candidates are edges


@FAST
@given(seed=st.integers(0, 1000))
def test_edge_images_lie_in_the_unit_range_and_peak_on_edges(seed):
    f = S.make_frames(40, 56, 1, seed=seed)[0]
    c = E.canny(f, 30, 80)
    for mode in ('gaussian', 'iedt'):
        e = E.edge_map(f, 30, 80, mode)
        assert e.min() >= 0.0 and e.max() <= 1.0
        if c.any():
            assert e.max() > 0.99 and (e.argmax() in np.flatnonzero(c) or mode == 'gaussian')
            if mode == 'iedt':
                assert np.all(e[c > 0] == e.max())                              # distance 0 on every edge pixel
