"""Parity of the CUDA path (through the C-ABI) against the float64 CPU oracle on the same seeded inputs.

Tolerances are BASELINE.json's: objective <= 1e-5 relative, gradient <= 1e-4 relative (inf-norm / inf-norm), the
event->pixel index stream bit-exact."""
import functools

import numpy as np
import pytest

import eincm_b200.synth as S
from oracle import eincm_oracle as O

pytestmark = pytest.mark.gpu

OBJ_RTOL = 1e-5      # BASELINE.json north_star
GRAD_RTOL = 1e-4     # BASELINE.json north_star


def _kw(win, gamma=0.0, delta=0.0, lvl=1, alpha=20.0, beta=35.0):
    return dict(alpha=alpha, beta=beta, gamma=gamma, delta=delta, cur_pyr_lvl=lvl, n_pyr_lvls=5,
                sensor_size=win.sensor_size, scale_to_sensor_size_method='bilinear')


def _rel_inf(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


@pytest.fixture(scope='module')
def L():
    from eincm_b200 import losses
    yield losses
    losses.clear_cache()


@pytest.fixture(scope='module')
def tiny():
    return S.make_workload('tiny', seed=0)


@pytest.mark.parametrize('shape', [(1, 1), (2, 2), (4, 4), (16, 16), (48, 64)])
@pytest.mark.parametrize('point', ['zero', 'truth', 'perturbed'])
def test_value_and_grad_matches_oracle(L, tiny, shape, point):
    th = S.theta_test_points(tiny, shape)[point]
    kw = _kw(tiny)
    (loss, aux), grad = L.value_and_grad(L.loss_func, has_aux=True)(th, *tiny.args(), **kw)
    l_ref, g_ref = O.value_and_grad(th, *tiny.args(), **kw)
    assert abs(loss - l_ref) <= OBJ_RTOL * abs(l_ref)
    assert _rel_inf(grad, g_ref) <= GRAD_RTOL
    assert grad.shape == th.shape and grad.dtype == np.float64
    assert set(aux) == {'final_loss', 'scaled_theta', 'mean_rel_corr', 'mean_rel_contrast', 'mean_rel_iwe_divergence',
                        'theta_total_variation', 'multi_ref_weights'}       # losses.py:195-203


@pytest.mark.parametrize('shape', [(1, 1), (4, 4), (16, 16), (48, 64)])
@pytest.mark.parametrize('gamma,delta,lvl', [(0.0, 0.0, 1), (0.0025, 0.3, 0)])
def test_exact_f64_mode_is_tight(L, tiny, shape, gamma, delta, lvl):
    """EINCM_FLAG_EXACT_F64: float64 taps and scatter-adds - agrees with the float64 oracle to rounding."""
    th = S.theta_test_points(tiny, shape)['perturbed']
    kw = _kw(tiny, gamma=gamma, delta=delta, lvl=lvl)
    L.configure(exact_f64=True)
    try:
        loss, grad = L.value_and_grad(L.loss_func)(th, *tiny.args(), **kw)
    finally:
        L.configure(exact_f64=False)
    l_ref, g_ref = O.value_and_grad(th, *tiny.args(), **kw)
    assert abs(loss - l_ref) <= 1e-11 * abs(l_ref)
    assert _rel_inf(grad, g_ref) <= 1e-9


@pytest.mark.parametrize('gamma,delta,lvl', [(0.0025, 0.0, 0), (0.0, 0.3, 2), (0.0025, 0.3, 0), (0.0025, 0.0, 3)])
@pytest.mark.parametrize('shape', [(2, 2), (16, 16), (48, 64)])
def test_regulariser_and_divergence_terms(L, tiny, gamma, delta, lvl, shape):
    th = S.theta_test_points(tiny, shape)['perturbed']
    kw = _kw(tiny, gamma=gamma, delta=delta, lvl=lvl)
    loss, grad = L.value_and_grad(L.loss_func)(th, *tiny.args(), **kw)
    l_ref, g_ref = O.value_and_grad(th, *tiny.args(), **kw)
    assert abs(loss - l_ref) <= OBJ_RTOL * abs(l_ref)
    assert _rel_inf(grad, g_ref) <= GRAD_RTOL


@pytest.mark.parametrize('R', [1, 2, 3, 5])
def test_loss_at_zero_theta_known_answer(L, R):
    # loss(theta=0) = -(alpha+beta)/R
    w = S.make_window(48, 64, 3000, edge_ts=np.linspace(0, 1, R) if R > 1 else (0.0,), seed=3)
    loss, _ = L.loss_func(np.zeros((4, 4, 2)), *w.args(), **_kw(w))
    # fixed-point votes are order-independent: IWE_r == zero-IWE exactly at theta = 0
    assert loss == pytest.approx(-(20.0 + 35.0) / R, rel=1e-12)
    L.configure(exact_f64=True)
    try:
        loss, _ = L.loss_func(np.zeros((4, 4, 2)), *w.args(), **_kw(w))
    finally:
        L.configure(exact_f64=False)
    assert loss == pytest.approx(-(20.0 + 35.0) / R, rel=1e-9)


@pytest.mark.parametrize('exact', [False, True])
def test_intermediates_and_bit_exact_pixel_indices(L, tiny, exact):
    from eincm_b200 import plan as P
    th = S.theta_test_points(tiny, (4, 4))['perturbed']
    kw = _kw(tiny)
    l_ref, g_ref, inter = O.value_and_grad(th, *tiny.args(), **kw, return_intermediates=True)
    p = P.Plan(tiny.sensor_size, max_events=len(tiny.xs), max_refs=3, flags=P.FLAG_EXACT_F64 if exact else 0)
    p.set_window(*tiny.args())
    loss, grad = p.value_and_grad_host(th, P.make_hparams(20.0, 35.0, 0.0, 0.0, 1))
    # per-pixel image values: votes quantised to 2^-21 of the centre tap in the default mode, float64 in EXACT_F64
    rt, at, rd = (1e-12, 1e-14, 1e-9) if exact else (2e-5, 1e-7, 1e-4)
    np.testing.assert_allclose(p.zero_iwe().cpu().numpy(), inter['zero_iwe'], rtol=rt, atol=at)
    np.testing.assert_allclose(p.iwe().cpu().numpy(), inter['iwes'], rtol=rt, atol=at)
    assert _rel_inf(p.dldi().cpu().numpy(), inter['dLdI']) <= rd
    np.testing.assert_allclose(p.theta_full().cpu().numpy(), inter['aux']['scaled_theta'], rtol=1e-13, atol=1e-13)
    np.testing.assert_array_equal(p.event_mask().cpu().numpy().astype(bool), O.make_event_mask(tiny.xs, tiny.ys, tiny.sensor_size))
    obj = inter['objectives']
    for r in range(3):
        cols, rows = p.rounded_pixels(r)
        xr, yr = O.rounded_event_pixels(obj['warped_xs'][r], obj['warped_ys'][r])
        np.testing.assert_array_equal(cols, xr.astype(np.int32))      # bit-exact event->pixel indexing (both modes)
        np.testing.assert_array_equal(rows, yr.astype(np.int32))
    s = p.scalars()
    rs = 1e-11 if exact else 1e-6
    np.testing.assert_allclose(s['contrasts'], obj['contrasts'], rtol=rs)
    np.testing.assert_allclose(s['correlations'], obj['correlations'], rtol=rs)
    np.testing.assert_allclose(s['zero_correlations'], obj['zero_correlations'], rtol=rs)
    np.testing.assert_allclose(s['zero_contrast'], obj['zero_contrast'], rtol=rs)
    np.testing.assert_allclose(s['multi_ref_weights'], obj['multi_ref_weights'], rtol=1e-14)
    p.close()


def test_handover(L, tiny):
    pts = S.theta_test_points(tiny, (4, 4))
    prev, cur = pts['truth'], pts['perturbed']
    kw = _kw(tiny)
    a0 = 0.37
    loss, da = L.value_and_grad(L.handover_loss_func)(a0, prev, cur, *tiny.args(), **kw)
    l_ref, da_ref = O.handover_value_and_grad(a0, prev, cur, *tiny.args(), **kw)
    assert abs(loss - l_ref) <= OBJ_RTOL * abs(l_ref)
    assert abs(da - da_ref) <= GRAD_RTOL * abs(da_ref)
    assert L.handover_loss_func(a0, prev, cur, *tiny.args(), **kw) == pytest.approx(loss, rel=1e-6)


def test_partial_binding_like_hydra(L, tiny):
    # configs/theta_loss_func/default.yaml builds partial(loss_func, alpha=..., ...); solver.py:165 adds cur_pyr_lvl
    pf = functools.partial(L.loss_func, alpha=20.0, beta=35.0, gamma=0.0, delta=0.0, n_pyr_lvls=5,
                           sensor_size=tiny.sensor_size, scale_to_sensor_size_method='bilinear')
    th = S.theta_test_points(tiny, (2, 2))['truth']
    f = functools.partial(pf, cur_pyr_lvl=3)
    (loss, _), grad = L.value_and_grad(f, has_aux=True)(th, *tiny.args())
    l_ref, g_ref = O.value_and_grad(th, *tiny.args(), **_kw(tiny, lvl=3))
    assert abs(loss - l_ref) <= OBJ_RTOL * abs(l_ref)
    assert _rel_inf(grad, g_ref) <= GRAD_RTOL


def test_wrap_quirk_and_far_out_of_sensor_warps(tiny):
    """Events warped past the left/top border wrap (JAX negative-index normalisation); far ones are dropped."""
    from eincm_b200 import plan as P
    th = np.zeros((1, 1, 2)); th[..., 0] = 80.0; th[..., 1] = 60.0
    kw = _kw(tiny)
    for flags, wrap in ((0, True), (P.FLAG_NO_WRAP_NEGATIVE, False), (P.FLAG_EXACT_F64, True),
                        (P.FLAG_EXACT_F64 | P.FLAG_NO_WRAP_NEGATIVE, False)):
        p = P.Plan(tiny.sensor_size, max_events=len(tiny.xs), max_refs=3, flags=flags)
        p.set_window(*tiny.args())
        loss, grad = p.value_and_grad_host(th, P.make_hparams(20.0, 35.0, 0.0, 0.0, 1))
        l_ref, g_ref = O.value_and_grad(th, *tiny.args(), **kw, wrap_negative=wrap)
        assert abs(loss - l_ref) <= OBJ_RTOL * abs(l_ref)
        assert _rel_inf(grad, g_ref) <= GRAD_RTOL
        p.close()


def test_error_behaviour(tiny):
    from eincm_b200 import plan as P
    p = P.Plan(tiny.sensor_size, max_events=1000, max_refs=3)
    hp = P.make_hparams(20.0, 35.0, 0.0, 0.0, 1)
    with pytest.raises(P.EincmError) as e:
        p.value_and_grad_host(np.zeros((1, 1, 2)), hp)            # before set_window
    assert e.value.code == P.EINCM_ESTATE
    with pytest.raises(P.EincmError) as e:
        p.set_window(*tiny.args())                                  # more events than the plan holds
    assert e.value.code == P.EINCM_EINVAL
    xs = tiny.xs[:100].copy(); xs[5] = tiny.sensor_size[1]          # outside the sensor
    with pytest.raises(P.EincmError) as e:
        p.set_window(xs, tiny.ys[:100], tiny.ts[:100], tiny.edges, tiny.edge_ts)
    assert e.value.code == P.EINCM_ERANGE
    p.set_window(tiny.xs[:100], tiny.ys[:100], tiny.ts[:100], tiny.edges, tiny.edge_ts)
    with pytest.raises(P.EincmError) as e:
        p.value_and_grad_host(np.zeros((100, 100, 2)), hp)          # theta larger than the sensor
    assert e.value.code == P.EINCM_EINVAL
    with pytest.raises(P.EincmError):
        P.make_hparams(1, 1, 0, 0, 0, scale_to_sensor_size_method='lanczos3')
    p.close()


def test_empty_window_and_single_event(tiny):
    from eincm_b200 import plan as P
    p = P.Plan(tiny.sensor_size, max_events=16, max_refs=3)
    hp = P.make_hparams(20.0, 35.0, 0.0, 0.0, 1)
    e16 = np.zeros(0, dtype=np.int16)
    p.set_window(e16, e16, np.zeros(0), tiny.edges, tiny.edge_ts)
    loss, grad = p.value_and_grad_host(np.ones((2, 2, 2)), hp)
    l_ref, g_ref = O.value_and_grad(np.ones((2, 2, 2)), e16, e16, np.zeros(0), tiny.edges, tiny.edge_ts, **_kw(tiny))
    assert (np.isnan(loss) and np.isnan(l_ref)) or abs(loss - l_ref) <= OBJ_RTOL * abs(l_ref)
    assert (grad == 0).all() or np.isnan(grad).all()
    xs = np.array([10], dtype=np.int16); ys = np.array([20], dtype=np.int16); ts = np.array([0.3])
    p.set_window(xs, ys, ts, tiny.edges, tiny.edge_ts)
    th = np.full((1, 1, 2), 3.3)
    loss, grad = p.value_and_grad_host(th, hp)
    l_ref, g_ref = O.value_and_grad(th, xs, ys, ts, tiny.edges, tiny.edge_ts, **_kw(tiny))
    assert abs(loss - l_ref) <= OBJ_RTOL * abs(l_ref)
    assert _rel_inf(grad, g_ref) <= GRAD_RTOL
    p.close()


def test_stateless_host_form(tiny):
    from eincm_b200 import plan as P
    p = P.Plan(tiny.sensor_size, max_events=len(tiny.xs), max_refs=3)
    th = S.theta_test_points(tiny, (4, 4))['truth']
    loss, grad = p.value_and_grad_stateless_host(th, *tiny.args(), P.make_hparams(20.0, 35.0, 0.0, 0.0, 1))
    l_ref, g_ref = O.value_and_grad(th, *tiny.args(), **_kw(tiny))
    assert abs(loss - l_ref) <= OBJ_RTOL * abs(l_ref)
    assert _rel_inf(grad, g_ref) <= GRAD_RTOL
    p.close()


@pytest.mark.parametrize('name,shape', [('mvsec_dt4', (16, 16)), ('mvsec_dt1', (8, 8)), ('ecd', (4, 4)), ('mvsec_outdoor', (16, 16)),
                                        ('mvsec_raw_dt4', (16, 16))])
def test_mvsec_shaped_windows(L, name, shape):
    """BASELINE.json configs[2] (MVSEC-shaped, R = 2 / 5) at the reference's own N (oracle runs in seconds)."""
    w = S.make_workload(name, seed=1)
    th = S.theta_test_points(w, shape)['perturbed']
    kw = dict(w.hparams, cur_pyr_lvl=0, n_pyr_lvls=5, sensor_size=w.sensor_size, scale_to_sensor_size_method='bilinear')
    loss, grad = L.value_and_grad(L.loss_func)(th, *w.args(), **kw)
    l_ref, g_ref = O.value_and_grad(th, *w.args(), **kw)
    assert abs(loss - l_ref) <= OBJ_RTOL * abs(l_ref)
    assert _rel_inf(grad, g_ref) <= GRAD_RTOL


def test_dense_theta_mvsec(L):
    """BASELINE.json configs[2]: dense per-pixel flow (theta of shape (H, W, 2): identity resize)."""
    w = S.make_workload('mvsec_dt1', seed=2)
    th = S.theta_test_points(w, w.sensor_size)['perturbed']
    kw = dict(w.hparams, cur_pyr_lvl=0, n_pyr_lvls=5, sensor_size=w.sensor_size, scale_to_sensor_size_method='bilinear')
    kw['gamma'] = 0.0025
    loss, grad = L.value_and_grad(L.loss_func)(th, *w.args(), **kw)
    l_ref, g_ref = O.value_and_grad(th, *w.args(), **kw)
    assert abs(loss - l_ref) <= OBJ_RTOL * abs(l_ref)
    assert _rel_inf(grad, g_ref) <= GRAD_RTOL


def test_dsec_shaped_window_reduced_events(L):
    """BASELINE.json configs[0]/[1] shape (640x480, R=3) at an event count the NumPy oracle finishes in seconds."""
    w = S.make_workload('dsec_shipped', seed=0, n_events=200_000)
    kw = dict(w.hparams, cur_pyr_lvl=0, n_pyr_lvls=5, sensor_size=w.sensor_size, scale_to_sensor_size_method='bilinear')
    for shape in [(1, 1), (16, 16)]:
        th = S.theta_test_points(w, shape)['perturbed']
        loss, grad = L.value_and_grad(L.loss_func)(th, *w.args(), **kw)
        l_ref, g_ref = O.value_and_grad(th, *w.args(), **kw)
        assert abs(loss - l_ref) <= OBJ_RTOL * abs(l_ref)
        assert _rel_inf(grad, g_ref) <= GRAD_RTOL


def test_full_size_properties_dsec(L):
    """Size-independent properties at BASELINE's full DSEC size (N = 1.5 M, too slow for the NumPy oracle):
    loss(theta=0) = -(alpha+beta)/R exactly; splat mass conservation; event-split additivity of the IWE; the handover
    gradient equals <grad, prev - theta>; determinism of repeated evaluations within tolerance."""
    from eincm_b200 import plan as P
    w = S.make_workload('dsec_shipped', seed=0)
    N = len(w.xs)
    hp = P.make_hparams(w.hparams['alpha'], w.hparams['beta'], 0.0, 0.0, 0)
    p = P.Plan(w.sensor_size, max_events=N, max_refs=3)
    p.set_window(*w.args())
    loss0, g0 = p.value_and_grad_host(np.zeros((16, 16, 2)), hp)
    assert loss0 == pytest.approx(-(w.hparams['alpha'] + w.hparams['beta']) / 3, rel=1e-12)     # IWE_r == zero-IWE exactly
    # interior mass: each in-sensor, non-border event contributes 0.7794836797093877
    z = p.zero_iwe().cpu().numpy()
    interior = (w.xs >= 1) & (w.xs < w.sensor_size[1] - 1) & (w.ys >= 1) & (w.ys < w.sensor_size[0] - 1)
    lo = 0.7794836797093877 * interior.sum()
    assert lo <= z.sum() * (1 + 1e-9) and z.sum() <= 0.7794836797093877 * N * (1 + 1e-9) + 1e-6
    th = S.theta_test_points(w, (16, 16))['perturbed']
    loss, grad = p.value_and_grad_host(th, hp)
    iwe_full = p.iwe().cpu().numpy().copy()
    loss_b, grad_b = p.value_and_grad_host(th, hp)
    # fixed-point votes + fixed-order reductions: the objective is bit-reproducible; the gradient sums float64 atomics
    assert loss == loss_b and _rel_inf(grad_b, grad) <= 1e-11
    np.testing.assert_array_equal(p.iwe().cpu().numpy(), iwe_full)
    # handover: d/d alpha = <grad(theta_ho), prev - theta>
    prev = S.theta_test_points(w, (16, 16))['truth']
    a0 = 0.25
    th_ho = a0 * prev + (1 - a0) * th
    l_ho, da = p.handover_value_and_grad_host(a0, prev, th, hp)
    l_dir, g_dir = p.value_and_grad_host(th_ho, hp)
    assert l_ho == pytest.approx(l_dir, rel=1e-6)
    assert da == pytest.approx(float((g_dir * (prev - th)).sum()), rel=1e-4)
    # additivity: IWE(A u B) = IWE(A) + IWE(B)
    acc = np.zeros_like(iwe_full)
    for sl in (slice(0, None, 2), slice(1, None, 2)):
        p.set_window(w.xs[sl], w.ys[sl], w.ts[sl], w.edges, w.edge_ts)
        p.value_and_grad_host(th, hp, want_grad=False)
        acc += p.iwe().cpu().numpy()
    np.testing.assert_allclose(acc, iwe_full, rtol=1e-14, atol=1e-300)       # integer sums: additive to the last bit
    p.close()


@pytest.mark.parametrize('flow', [(300.0, -250.0), (-45.0, 70.0), (0.0, 1500.0), (60.0, -55.0), (110.0, 15.0), (-20.0, 190.0)])
def test_large_flow_takes_the_window_fallback(L, flow):
    """Flows larger than one shared-memory window can hold: the destination rectangle of a chunk is processed in row slices
    (60 .. 190 px), and beyond 16 slices / 1365 columns it is cropped and the remaining votes take the per-tap global
    reductions (forward) / global gathers (backward); many warps leave the sensor (wrap / drop rule)."""
    w = S.make_workload('dsec_shipped', seed=4, n_events=60_000)
    kw = dict(w.hparams, cur_pyr_lvl=1, n_pyr_lvls=5, sensor_size=w.sensor_size, scale_to_sensor_size_method='bilinear')
    th = np.zeros((2, 2, 2)); th[..., 0] = flow[0]; th[..., 1] = flow[1]
    th += np.random.default_rng(5).normal(0.0, 3.0, size=th.shape)
    loss, grad = L.value_and_grad(L.loss_func)(th, *w.args(), **kw)
    l_ref, g_ref = O.value_and_grad(th, *w.args(), **kw)
    assert abs(loss - l_ref) <= OBJ_RTOL * abs(l_ref)
    assert _rel_inf(grad, g_ref) <= GRAD_RTOL


@pytest.mark.parametrize('name,flow,shape', [('mvsec_dt4', (70.0, -90.0), (2, 2)), ('mvsec_dt1', (-120.0, 60.0), (1, 1)),
                                             ('mvsec_raw_dt4', (55.0, 140.0), (4, 4)), ('e00_single', (80.0, 80.0), (2, 2))])
def test_sliced_windows_with_other_reference_counts(L, name, flow, shape):
    """Sliced rectangles when the reference times do not fill whole passes: R = 5 (two passes of first slices, RB = 4), R = 2, R = 1,
    with different numbers of slices per reference time (the rectangle grows with |t - t_ref|)."""
    w = S.make_workload(name, seed=21, n_events=20_000)
    kw = dict(w.hparams, cur_pyr_lvl=1, n_pyr_lvls=5, sensor_size=w.sensor_size, scale_to_sensor_size_method='bilinear')
    th = np.zeros(shape + (2,)); th[..., 0] = flow[0]; th[..., 1] = flow[1]
    th += np.random.default_rng(22).normal(0.0, 2.0, size=th.shape)
    loss, grad = L.value_and_grad(L.loss_func)(th, *w.args(), **kw)
    l_ref, g_ref = O.value_and_grad(th, *w.args(), **kw)
    assert abs(loss - l_ref) <= OBJ_RTOL * abs(l_ref)
    assert _rel_inf(grad, g_ref) <= GRAD_RTOL


@pytest.mark.parametrize('c', [64.0, 96.0])
def test_sliced_windows_with_coordinates_on_rounding_boundaries(L, c):
    """Dyadic timestamps and a dyadic constant flow put thousands of warped coordinates exactly on k + 0.5 (round half to even,
    event_utils.py:33), some of them on the row that separates two slices of a sliced destination rectangle: those belong to
    neither slice and must take the fallback exactly once."""
    w = S.make_workload('dsec_shipped', seed=9, n_events=40_000)
    ts = np.sort(np.round(w.ts * 128.0) / 128.0)
    kw = dict(w.hparams, cur_pyr_lvl=4, n_pyr_lvls=5, sensor_size=w.sensor_size, scale_to_sensor_size_method='bilinear')
    th = np.zeros((1, 1, 2)); th[..., 0] = c; th[..., 1] = -c
    p = L.value_and_grad(L.loss_func)
    loss, grad = p(th, w.xs, w.ys, ts, w.edges, w.edge_ts, **kw)
    l_ref, g_ref = O.value_and_grad(th, w.xs, w.ys, ts, w.edges, w.edge_ts, **kw)
    assert abs(loss - l_ref) <= OBJ_RTOL * abs(l_ref)
    assert _rel_inf(grad, g_ref) <= GRAD_RTOL


def test_hot_pixel_and_ragged_tiles(L):
    """Adversarial event layouts: thousands of events on one pixel (several chunks of one tile, votes merged in
    registers), events on the sensor border, tiles with 1..3 events (padding sentinels), identical timestamps."""
    rng = np.random.default_rng(11)
    H, W = 96, 160
    base = S.make_window(H, W, 5000, seed=6)
    hot = 6000
    xs = np.concatenate([base.xs, np.full(hot, 77, np.int16), np.array([0, W - 1, 0, W - 1, 33], np.int16)])
    ys = np.concatenate([base.ys, np.full(hot, 41, np.int16), np.array([0, 0, H - 1, H - 1, 95], np.int16)])
    ts = np.concatenate([base.ts, rng.uniform(0, 1, hot), np.array([0.5, 0.5, 0.5, 0.5, 0.5])])
    order = np.argsort(ts, kind='stable')
    xs, ys, ts = np.ascontiguousarray(xs[order]), np.ascontiguousarray(ys[order]), np.ascontiguousarray(ts[order])
    kw = dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.0, cur_pyr_lvl=1, n_pyr_lvls=5, sensor_size=(H, W),
              scale_to_sensor_size_method='bilinear')
    th = rng.normal(0.0, 6.0, size=(4, 4, 2))
    loss, grad = L.value_and_grad(L.loss_func)(th, xs, ys, ts, base.edges, base.edge_ts, **kw)
    l_ref, g_ref = O.value_and_grad(th, xs, ys, ts, base.edges, base.edge_ts, **kw)
    assert abs(loss - l_ref) <= OBJ_RTOL * abs(l_ref)
    assert _rel_inf(grad, g_ref) <= GRAD_RTOL


@pytest.mark.parametrize('n_tile', [1, 3, 4, 1023, 1024, 1025, 1028, 2043, 2044, 2045, 2048, 4088, 4089, 6133])
def test_chunk_and_sub_chunk_boundaries(L, n_tile):
    """All events of one 16x16 source tile, at counts around the sub-chunk (1024) and chunk (2044) capacities: a chunk is voted in
    register-resident sub-chunks into one window; sentinels pad the last group; the IWE of every reference time must equal the
    oracle's and the objective must be bit-identical however the events are cut (integer votes)."""
    from eincm_b200 import plan as P
    rng = np.random.default_rng(n_tile)
    H, W = 64, 96
    base = S.make_window(H, W, 64, seed=2)                               # edges + a few events elsewhere (other tiles: 1 chunk each)
    xs = np.concatenate([base.xs, rng.integers(32, 48, n_tile).astype(np.int16)])
    ys = np.concatenate([base.ys, rng.integers(16, 32, n_tile).astype(np.int16)])
    ts = np.concatenate([base.ts, rng.uniform(0, 1, n_tile)])
    order = np.argsort(ts, kind='stable')
    xs, ys, ts = np.ascontiguousarray(xs[order]), np.ascontiguousarray(ys[order]), np.ascontiguousarray(ts[order])
    kw = dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.0, cur_pyr_lvl=1, n_pyr_lvls=5, sensor_size=(H, W),
              scale_to_sensor_size_method='bilinear')
    th = rng.normal(0.0, 5.0, size=(2, 3, 2))
    loss, grad = L.value_and_grad(L.loss_func)(th, xs, ys, ts, base.edges, base.edge_ts, **kw)
    l_ref, g_ref, inter = O.value_and_grad(th, xs, ys, ts, base.edges, base.edge_ts, **kw, return_intermediates=True)
    assert abs(loss - l_ref) <= OBJ_RTOL * abs(l_ref)
    assert _rel_inf(grad, g_ref) <= GRAD_RTOL
    # the same events in another order of arrival (ties in the per-pixel time order aside, another cut into groups): same bits
    perm = rng.permutation(len(xs))
    p = P.Plan((H, W), max_events=len(xs), max_refs=3)
    try:
        hp = P.make_hparams(20.0, 35.0, 0.0, 0.0, 1)
        p.set_window(xs, ys, ts, base.edges, base.edge_ts)
        la, _ = p.value_and_grad_host(th, hp)
        iwe_a = p.iwe().cpu().numpy().copy()
        np.testing.assert_allclose(iwe_a, inter['iwes'], rtol=2e-5, atol=1e-7)          # votes quantised to 2^-21 of the centre tap
        p.set_window(np.ascontiguousarray(xs[perm]), np.ascontiguousarray(ys[perm]), np.ascontiguousarray(ts[perm]), base.edges, base.edge_ts)
        lb, _ = p.value_and_grad_host(th, hp)
        assert la == lb == loss
        np.testing.assert_array_equal(iwe_a, p.iwe().cpu().numpy())
    finally:
        p.close()


@pytest.mark.parametrize('n_hot,n_other', [(1, 0), (31, 0), (32, 1), (33, 40), (64, 0), (100, 30), (200, 55), (255, 0), (256, 0), (252, 4), (257, 0), (300, 20)])
def test_small_chunks_one_event_per_thread(L, n_hot, n_other):
    """Chunks of at most 256 events take one event per thread (sparse windows): the run of a pixel with many events crosses warps, and the
    run's first warp collects the sums of the warps it continues into (k_events_tile.cuh, backward pass).  n_hot events on ONE pixel
    (+ n_other on its tile) at counts around the warp size and the mode's limit: objective and gradient against the oracle, and the
    gradient bit-identical from call to call (one reduction pair per pixel: no order of atomics to depend on)."""
    from eincm_b200 import plan as P
    rng = np.random.default_rng(1000 * n_hot + n_other)
    H, W = 64, 96
    base = S.make_window(H, W, 48, seed=3)                               # edges + a few events elsewhere (other tiles: small chunks too)
    keep = ~((base.xs >= 32) & (base.xs < 48) & (base.ys >= 16) & (base.ys < 32))       # the hot tile holds exactly n_hot + n_other events
    xs = np.concatenate([base.xs[keep], np.full(n_hot, 40, np.int16), rng.integers(32, 48, n_other).astype(np.int16)])
    ys = np.concatenate([base.ys[keep], np.full(n_hot, 21, np.int16), rng.integers(16, 32, n_other).astype(np.int16)])
    ts = np.concatenate([base.ts[keep], rng.uniform(0, 1, n_hot + n_other)])
    order = np.argsort(ts, kind='stable')
    xs, ys, ts = np.ascontiguousarray(xs[order]), np.ascontiguousarray(ys[order]), np.ascontiguousarray(ts[order])
    kw = dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.0, cur_pyr_lvl=1, n_pyr_lvls=5, sensor_size=(H, W),
              scale_to_sensor_size_method='bilinear')
    th = rng.normal(0.0, 4.0, size=(2, 3, 2))
    loss, grad = L.value_and_grad(L.loss_func)(th, xs, ys, ts, base.edges, base.edge_ts, **kw)
    l_ref, g_ref = O.value_and_grad(th, xs, ys, ts, base.edges, base.edge_ts, **kw)
    assert abs(loss - l_ref) <= OBJ_RTOL * abs(l_ref)
    assert _rel_inf(grad, g_ref) <= 1e-5                                 # float32 per-event gradients: far inside GRAD_RTOL (a lost carry is > 10 %)
    p = P.Plan((H, W), max_events=len(xs), max_refs=3)
    try:
        hp = P.make_hparams(20.0, 35.0, 0.0, 0.0, 1)
        p.set_window(xs, ys, ts, base.edges, base.edge_ts)
        la, ga = p.value_and_grad_host(th, hp)
        for _ in range(3):
            lb, gb = p.value_and_grad_host(th, hp)
            assert la == lb == loss
            if n_hot + n_other <= 256:
                np.testing.assert_array_equal(ga, gb)
            else:                                                        # four events per thread: a pixel's run may end in three reductions
                np.testing.assert_allclose(ga, gb, rtol=1e-11, atol=1e-300)
    finally:
        p.close()


def test_batched_host_call_matches_single_calls(tiny):
    """eincm_value_and_grad_host_batch: independent windows evaluated concurrently, each on its own stream, give the results
    of one-at-a-time calls (objective bit-identical: it does not depend on scheduling)."""
    from eincm_b200 import plan as P
    hp = P.make_hparams(20.0, 35.0, 0.0, 0.0, 1)
    wins = [tiny, S.make_workload('tiny', seed=7), S.make_workload('tiny', seed=8)]
    plans, thetas = [], []
    for k, w in enumerate(wins):
        p = P.Plan(w.sensor_size, max_events=len(w.xs), max_refs=3)
        p.set_window(*w.args())
        plans.append(p)
        thetas.append(S.theta_test_points(w, (4, 4), seed=k)['perturbed'])
    losses, grads = P.value_and_grad_host_batch(plans, thetas, hp)
    for k, w in enumerate(wins):
        l1, g1 = plans[k].value_and_grad_host(thetas[k], hp)
        assert losses[k] == l1
        assert _rel_inf(grads[k], g1) <= 1e-11
        l_ref, g_ref = O.value_and_grad(thetas[k], *w.args(), **_kw(w))
        assert abs(losses[k] - l_ref) <= OBJ_RTOL * abs(l_ref)
        assert _rel_inf(grads[k], g_ref) <= GRAD_RTOL
    losses_v, none = P.value_and_grad_host_batch(plans, thetas, hp, want_grad=False)
    assert none is None and (losses_v == losses).all()
    with pytest.raises(P.EincmError):
        P.value_and_grad_host_batch([plans[0], plans[0]], thetas[:2], hp)      # the same plan twice
    for p in plans:
        p.close()


def test_multi_level_solve_on_the_cuda_objective():
    """The solver mirror (reference src/eincm/solver.py) driven by the CUDA objective: level schedule, BFGS, handover.  BFGS
    trajectories are chaotic in the last digits of the objective, so instead of comparing two solves the oracle is evaluated AT
    the iterates the CUDA-driven solve produced ('theta test points: ... BFGS iterates'): the value the solver saw
    at every level is the oracle's value there, and no level ends above where it started."""
    from eincm_b200 import losses, solver as SV
    w0 = S.make_window(32, 48, 1500, seed=21, n_segments=12, flow_mag=4.0)
    w1 = S.make_window(32, 48, 1500, seed=22, n_segments=12, flow_mag=4.0)
    kwargs = dict(n_pyr_lvls=3, theta_opt_maxiters={'pyr_lvl_0': 8, 'pyr_lvl_1': 6, 'pyr_lvl_2': 5},
                  handover_opt_maxiters={'pyr_lvl_0': 4, 'pyr_lvl_1': 3, 'pyr_lvl_2': 2})
    gpu = losses.WindowObjective((32, 48), 20.0, 35.0, max_events=4096, max_refs=3)
    res = SV.solve_sequence(gpu, [w0, w1], kwargs)
    kw = dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.0, n_pyr_lvls=5, sensor_size=(32, 48), scale_to_sensor_size_method='bilinear')
    for w, r in zip((w0, w1), res):
        for k in range(3):
            key = f'pyr_lvl_{k}'
            l_end = O.loss_func(r['pre_handover_theta_pyr'][key], *w.args(), cur_pyr_lvl=k, **kw)[0]
            l_start = O.loss_func(r['pre_opt_theta_pyr'][key], *w.args(), cur_pyr_lvl=k, **kw)[0]
            assert r['theta_opt_state_pyr'][key].fun_val == pytest.approx(l_end, rel=OBJ_RTOL)
            assert l_end <= l_start + 1e-9 * abs(l_start)
    assert res[0]['ho_opt_state_pyr'] == {} and set(res[1]['ho_opt_state_pyr']) == {'pyr_lvl_1', 'pyr_lvl_0'}
    for k in (1, 0):                                   # the solved handover weight is what the blend used (solver.py:344-345)
        key = f'pyr_lvl_{k}'
        a = res[1]['final_handover_weight_pyr'][key]
        blend = a * res[1]['prior_theta_pyr'][key] + (1 - a) * res[1]['pre_handover_theta_pyr'][key]
        np.testing.assert_allclose(res[1]['final_theta_pyr'][key], blend, rtol=1e-12, atol=1e-12)
    gpu.close()


def test_native_solver_backend_matches_scipy_backend():
    """backend='native' (eincm_minimize_bfgs_host / eincm_minimize_handover_host: the optimizers run inside the library) against
    backend='scipy' on the same chained windows: the same level schedule and handover logic; every level ends at a point whose
    value the oracle confirms, and the native solve is at least as good as the scipy-driven one up to BFGS's path dependence."""
    from eincm_b200 import losses, solver as SV
    w0 = S.make_window(32, 48, 1500, seed=21, n_segments=12, flow_mag=4.0)
    w1 = S.make_window(32, 48, 1500, seed=22, n_segments=12, flow_mag=4.0)
    base = dict(n_pyr_lvls=3, theta_opt_maxiters={'pyr_lvl_0': 8, 'pyr_lvl_1': 6, 'pyr_lvl_2': 5},
                handover_opt_maxiters={'pyr_lvl_0': 4, 'pyr_lvl_1': 3, 'pyr_lvl_2': 2})
    kw = dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.0, n_pyr_lvls=5, sensor_size=(32, 48), scale_to_sensor_size_method='bilinear')
    out = {}
    for backend in ('scipy', 'native'):
        obj = losses.WindowObjective((32, 48), 20.0, 35.0, max_events=4096, max_refs=3)
        out[backend] = SV.solve_sequence(obj, [w0, w1], dict(base, backend=backend, own_stream=(backend == 'native')))
        assert obj.n_evals > 0
        obj.close()
    for w, r in zip((w0, w1), out['native']):
        for k in range(3):
            key = f'pyr_lvl_{k}'
            st = r['theta_opt_state_pyr'][key]
            assert st.status in (0, 1, 2) and st.iter_num <= base['theta_opt_maxiters'][key]
            l_end = O.loss_func(r['pre_handover_theta_pyr'][key], *w.args(), cur_pyr_lvl=k, **kw)[0]
            l_start = O.loss_func(r['pre_opt_theta_pyr'][key], *w.args(), cur_pyr_lvl=k, **kw)[0]
            assert st.fun_val == pytest.approx(l_end, rel=OBJ_RTOL)
            assert l_end <= l_start + 1e-9 * abs(l_start)
    assert set(out['native'][1]['ho_opt_state_pyr']) == {'pyr_lvl_1', 'pyr_lvl_0'}
    for k in (1, 0):
        a = out['native'][1]['final_handover_weight_pyr'][f'pyr_lvl_{k}']
        assert 0.0 <= a <= 1.0
    # first window, coarsest level: both start from theta = 0 with the same algorithm -> the same local minimum
    f_n = out['native'][0]['theta_opt_state_pyr']['pyr_lvl_2'].fun_val
    f_s = out['scipy'][0]['theta_opt_state_pyr']['pyr_lvl_2'].fun_val
    assert f_n == pytest.approx(f_s, rel=1e-3)


@pytest.mark.parametrize('burst', [100, 50])
def test_evaluation_group_solves_match_independent_solves(burst):
    """Sequences solved concurrently through an evaluation group (eincm_group: one thread launches the evaluations of all running
    minimisations in a burst) reach what the same sequences reach when solved one after the other: every optimizer sees exactly its
    own evaluations, so each level's end point is confirmed by the oracle and the coarsest level (one BFGS run from theta = 0)
    agrees with the ungrouped run."""
    import threading
    from eincm_b200 import losses, plan as P, solver as SV
    wins = [S.make_window(32, 48, 1500, seed=31 + k, n_segments=12, flow_mag=4.0) for k in range(3)]
    base = dict(n_pyr_lvls=3, theta_opt_maxiters={'pyr_lvl_0': 8, 'pyr_lvl_1': 6, 'pyr_lvl_2': 5},
                handover_opt_maxiters={'pyr_lvl_0': 4, 'pyr_lvl_1': 3, 'pyr_lvl_2': 2}, backend='native', own_stream=True)
    kw = dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.0, n_pyr_lvls=5, sensor_size=(32, 48), scale_to_sensor_size_method='bilinear')

    def run(grouped):
        objs = [losses.WindowObjective((32, 48), 20.0, 35.0, max_events=4096, max_refs=3) for _ in wins]
        grp = P.Group(burst) if grouped else None
        res = [None] * len(wins)

        def work(t):
            import torch
            torch.cuda.set_device(0)
            res[t] = SV.solve_sequence(objs[t], [wins[t], wins[(t + 1) % 3]], dict(base))

        if grouped:
            for o in objs:
                o.plan.set_group(grp)
            ths = [threading.Thread(target=work, args=(t,)) for t in range(len(wins))]
            for th in ths:
                th.start()
            for th in ths:
                th.join(timeout=120)
            assert not any(th.is_alive() for th in ths), 'evaluation group deadlocked'
        else:
            for t in range(len(wins)):
                work(t)
        for o in objs:
            o.plan.set_group(None)
            o.close()
        if grp is not None:
            grp.close()
        return res

    solo, grouped = run(False), run(True)
    for t in range(3):
        for wi, (w, r) in enumerate(zip((wins[t], wins[(t + 1) % 3]), grouped[t])):
            for k in range(3):
                key = f'pyr_lvl_{k}'
                st = r['theta_opt_state_pyr'][key]
                l_end = O.loss_func(r['pre_handover_theta_pyr'][key], *w.args(), cur_pyr_lvl=k, **kw)[0]
                assert st.fun_val == pytest.approx(l_end, rel=OBJ_RTOL)
        f_g = grouped[t][0]['theta_opt_state_pyr']['pyr_lvl_2'].fun_val
        f_s = solo[t][0]['theta_opt_state_pyr']['pyr_lvl_2'].fun_val
        assert f_g == pytest.approx(f_s, rel=1e-6)
