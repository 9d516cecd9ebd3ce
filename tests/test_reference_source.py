"""The reference's OWN source against the oracle.

tests/golden/refsrc/*.npz hold the outputs of the reference's unmodified ``eincm.losses`` (loss_func, compute_loss_objectives,
handover_loss_func - src/eincm/losses.py:49-276 and everything they import) executed over the float64 torch-backed JAX stand-in
``tests/_jaxshim`` on the inputs of the committed fixtures (tests/golden/make_golden_refsrc.py).  They pin the oracle's restatement of
the reference's composition, forward and reverse; the JAX primitives themselves stay restated (listed in tests/_jaxshim/jax/__init__.py).

Where /root/reference is present (this container, not the GPU box) the same is repeated live on further windows: events leaving the
frame on every side (negative-index wrap before mode='drop'), TV / divergence terms, 2 - 5 reference times, dense theta.
"""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

import eincm_b200.synth as S
from oracle import eincm_oracle as O
from tests import _golden as G

HERE = os.path.dirname(os.path.abspath(__file__))
REFSRC_DIR = os.path.join(G.GOLDEN_DIR, 'refsrc')
REFERENCE_SRC = '/root/reference/src'


def load_refsrc(name):
    z = np.load(os.path.join(REFSRC_DIR, name + '.npz'))
    return {k: z[k] for k in z.files}


def _rel_inf(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def test_every_fixture_has_reference_source_vectors():
    assert G.NAMES and sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(REFSRC_DIR, '*.npz'))) == G.NAMES


def _check_against_oracle(theta, prev, a_ho, args, hp, sensor_size, ref, tight=1e-11):
    kw = dict(n_pyr_lvls=5, sensor_size=sensor_size, scale_to_sensor_size_method='bilinear', **hp)
    loss, grad, inter = O.value_and_grad(theta, *args, return_intermediates=True, **kw)
    assert abs(loss - float(ref['loss'])) <= tight * abs(float(ref['loss']))
    assert _rel_inf(grad, ref['grad']) <= 1e-9
    np.testing.assert_allclose(O.scale_theta_to_sensor_size(theta, sensor_size), ref['scaled_theta'], rtol=1e-13, atol=1e-15)
    obj = inter['objectives']
    # the warped coordinates decide the event -> pixel index stream: their rint must agree exactly
    for k in ('warped_xs', 'warped_ys'):
        np.testing.assert_allclose(obj[k], ref['obj_' + k], rtol=0, atol=1e-11)
        np.testing.assert_array_equal(np.rint(obj[k]), np.rint(ref['obj_' + k]))
    for k in ('correlations', 'zero_correlations', 'contrasts', 'zero_contrast', 'theta_total_variation', 'theta_divergence',
              'iwe_divergences', 'zero_iwe_divergence', 'flow_warp_losses', 'multi_ref_weights'):
        if k == 'theta_total_variation' and hp['gamma'] == 0.0:
            continue    # unused by the loss; on (piecewise) constant flow its count of EXACTLY non-zero gradients depends on the
                        # summation order of the convolution (test_tv_count_depends_on_summation_order below)
        if k in obj:
            np.testing.assert_allclose(np.asarray(obj[k], dtype=np.float64), ref['obj_' + k], rtol=1e-10, atol=1e-14, err_msg=k)
    np.testing.assert_allclose(inter['zero_iwe'], ref['zero_iwe'], rtol=1e-12, atol=1e-15)           # the images of (warped) events
    np.testing.assert_allclose(inter['iwes'], ref['iwes'], rtol=1e-11, atol=1e-14)
    ho_loss, ho_dalpha = O.handover_value_and_grad(a_ho, prev, theta, *args, **kw)
    assert abs(ho_loss - float(ref['handover_loss'])) <= tight * abs(float(ref['handover_loss']))
    assert abs(ho_dalpha - float(ref['handover_dalpha'])) <= 1e-9 * max(abs(float(ref['handover_dalpha'])), np.abs(ref['grad']).max())


@pytest.mark.parametrize('name', G.NAMES)
def test_oracle_matches_reference_source_vectors(name):
    g, ref = G.load(name), load_refsrc(name)
    hp = dict(g['hp'])
    _check_against_oracle(g['theta'], g['prev_theta'], float(g['alpha_handover']), g['args'], hp, g['sensor_size'], ref)


@pytest.mark.parametrize('name', G.NAMES)
def test_committed_fixtures_match_reference_source_vectors(name):
    """the fixtures the C oracle and the CUDA path are tested against (tests/golden/*.npz) carry the reference source's numbers"""
    g, ref = G.load(name), load_refsrc(name)
    assert abs(float(g['loss']) - float(ref['loss'])) <= 1e-12 * abs(float(ref['loss']))
    assert _rel_inf(g['grad'], ref['grad']) <= 1e-10
    assert abs(float(g['handover_loss']) - float(ref['handover_loss'])) <= 1e-12 * abs(float(ref['handover_loss']))
    R = len(g['edge_ts'])
    np.testing.assert_allclose(g['zero_iwe'], ref['zero_iwe'], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(g['iwes'], ref['iwes'], rtol=1e-11, atol=1e-14)
    for r in range(R):
        np.testing.assert_array_equal(g['rounded'][r, 0], np.rint(ref['obj_warped_xs'][r]).astype(np.int32))
        np.testing.assert_array_equal(g['rounded'][r, 1], np.rint(ref['obj_warped_ys'][r]).astype(np.int32))


@pytest.mark.parametrize('name', G.NAMES)
def test_evaluation_metrics_match_reference_source_vectors(name):
    """src/evaluations/theta_eval.py:14-97 / flow_eval.py:14-75 executed as they are, against the oracle's evaluate_theta_array and the
    numbers stored in the fixtures (which the CUDA evaluation kernels are tested against, tests/test_gpu_golden_eval.py)"""
    g, ref = G.load(name), load_refsrc(name)
    hp = g['hp']
    theta_full = O.scale_theta_to_sensor_size(g['theta'], g['sensor_size'])
    ev = O.evaluate_theta_array(theta_full, *g['args'], g['gt_flow'], hp['alpha'], hp['beta'], hp['gamma'], hp['delta'], g['sensor_size'],
                                err_eval_event_mask=g['err_mask'])
    keys = [k[5:] for k in ref if k.startswith('eval_')]
    assert {'AEE', 'AREE', 'A1PE', 'A20PE', 'n_ee', 'n_pred', 'n_gt', 'n_pixels', 'loss', 'iwe_var', 'fwl', 'rel_contrasts'} <= set(keys)
    for k in keys:
        if k == 'theta_tot_var' and hp['gamma'] == 0.0:
            continue        # constant / linear stretches of a coarse theta: the count of exactly non-zero flow gradients depends on
                            # the summation order of the convolution (test_tv_count_depends_on_summation_order); not in the loss here
        np.testing.assert_allclose(np.asarray(ev[k], dtype=np.float64), ref['eval_' + k], rtol=1e-10, atol=1e-13, err_msg=k)
        if k in g['eval']:
            np.testing.assert_allclose(g['eval'][k], ref['eval_' + k], rtol=1e-10, atol=1e-13, err_msg=k)
    for k in ('n_ee', 'n_pred', 'n_gt', 'n_pixels'):
        assert int(ev[k]) == int(ref['eval_' + k])


@pytest.mark.parametrize('name', G.NAMES)
def test_c_oracle_matches_reference_source_vectors(name):
    from oracle import c_oracle as C
    if not C.available():
        subprocess.run(['make', '-C', os.path.join(os.path.dirname(HERE), 'oracle')], check=True, stdout=subprocess.DEVNULL)
    g, ref = G.load(name), load_refsrc(name)
    hp = g['hp']
    lc, gc = C.value_and_grad_raw(g['theta'], *g['args'], hp['alpha'], hp['beta'], hp['gamma'], hp['delta'], hp['cur_pyr_lvl'], g['sensor_size'])[:2]
    assert abs(lc - float(ref['loss'])) <= 1e-11 * abs(float(ref['loss']))
    assert _rel_inf(gc, ref['grad']) <= 1e-9


FULLSIZE_DIR = os.path.join(G.GOLDEN_DIR, 'refsrc_fullsize')
FULLSIZE = ['dsec_2m_theta16', 'mvsec_dt4_dense']


def load_fullsize(name):
    """(window regenerated from its seed, stored vector) - skips if the regenerated operands do not carry the stored checksum"""
    z = np.load(os.path.join(FULLSIZE_DIR, name + '.npz'))
    win = S.make_workload(str(z['workload']), seed=int(z['seed']))
    cs = np.array([float(np.asarray(win.xs, dtype=np.float64).sum()), float(np.asarray(win.ys, dtype=np.float64).sum()),
                   float(np.asarray(win.ts).sum()), float(np.asarray(win.edges).sum()), float(len(win.xs))])
    if not np.allclose(cs, z['checksum'], rtol=1e-12, atol=0):
        pytest.skip('eincm_b200.synth regenerates a different window on this machine (NumPy version?)')
    return win, z


@pytest.mark.parametrize('name', FULLSIZE)
def test_c_oracle_matches_reference_source_at_bench_configurations(name):
    """the checker of bench.py's `check` and of tests/test_gpu_fullsize.py, at the BENCH line's configuration (dsec, N = 2 M, theta 16 x 16,
    seed 0) and at MVSEC dt4 with dense theta, against the reference's own source (tests/golden/make_golden_refsrc_fullsize.py)"""
    from oracle import c_oracle as C
    if not C.available():
        subprocess.run(['make', '-C', os.path.join(os.path.dirname(HERE), 'oracle')], check=True, stdout=subprocess.DEVNULL)
    win, z = load_fullsize(name)
    hp = win.hparams
    lc, gc = C.value_and_grad_raw(z['theta'], *win.args(), hp['alpha'], hp['beta'], hp['gamma'], hp['delta'], int(z['cur_pyr_lvl']), win.sensor_size)[:2]
    assert abs(lc - float(z['loss'])) <= 1e-10 * abs(float(z['loss']))
    assert _rel_inf(gc, z['grad']) <= 1e-8


# ---------------------------------------------------------------------------------------------------------------------------------
# live: the reference source executed now (only where /root/reference exists)
# ---------------------------------------------------------------------------------------------------------------------------------

LIVE = [  # H, W, N, edge_ts, theta shape, point, flow scale, hyper-parameters
    (40, 56, 3000, (0.0, 0.5, 1.0), (2, 2), 'perturbed', 1.0, dict(alpha=2000.0, beta=4000.0, gamma=0.0, delta=0.0, cur_pyr_lvl=3)),
    (40, 56, 3000, (0.0, 1.0), (8, 8), 'truth', 1.0, dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.3, cur_pyr_lvl=1)),
    (33, 47, 2500, (0.0, 0.25, 0.5, 0.75, 1.0), (16, 16), 'perturbed', 1.0, dict(alpha=60.0, beta=60.0, gamma=0.0025, delta=0.3, cur_pyr_lvl=0)),
    (33, 47, 2000, (0.0, 0.5, 1.0), (33, 47), 'perturbed', 1.0, dict(alpha=20.0, beta=35.0, gamma=0.0025, delta=0.0, cur_pyr_lvl=0)),
    # flows far larger than the frame: events leave on every side; rows / columns -1 ... -H wrap before mode='drop' drops the rest
    (40, 56, 3000, (0.0, 0.5, 1.0), (4, 4), 'perturbed', 25.0, dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.0, cur_pyr_lvl=2)),
    (40, 56, 3000, (0.0, 0.5, 1.0), (4, 4), 'perturbed', -40.0, dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.3, cur_pyr_lvl=2)),
    # the shipped shapes at a tenth of their events: DSEC 480 x 640 with theta 16 x 16 (main.yaml weights), MVSEC 260 x 346 with dense theta, R = 5
    (480, 640, 200000, (0.0, 0.5, 1.0), (16, 16), 'perturbed', 1.0, dict(alpha=2000.0, beta=4000.0, gamma=0.0, delta=0.0, cur_pyr_lvl=0)),
    (260, 346, 30000, (0.0, 0.25, 0.5, 0.75, 1.0), (260, 346), 'perturbed', 1.0, dict(alpha=60.0, beta=60.0, gamma=0.0025, delta=0.0, cur_pyr_lvl=0)),
]


@pytest.fixture(scope='module')
def reference():
    if not os.path.isdir(REFERENCE_SRC):
        pytest.skip('/root/reference is not on this machine (the committed tests/golden/refsrc vectors stand in)')
    saved_path, saved_mods = list(sys.path), set(sys.modules)
    sys.path.insert(0, os.path.join(G.GOLDEN_DIR))
    import make_golden_refsrc as M
    jax, L = M.import_reference()
    yield M, jax, L
    sys.path[:] = saved_path
    for m in set(sys.modules) - saved_mods:          # the stand-in must not stay importable as "jax" for other tests
        del sys.modules[m]


@pytest.mark.parametrize('case', range(len(LIVE)))
def test_oracle_matches_reference_source_live(reference, case):
    M, jax, L = reference
    H, W, N, edge_ts, shape, point, fscale, hp = LIVE[case]
    win = S.make_window(seed=900 + case, H=H, W=W, N=N, edge_ts=edge_ts)
    pts = S.theta_test_points(win, shape, seed=case)
    theta = pts[point] * fscale
    prev = pts['zero'] if point == 'truth' else pts['truth']
    g = dict(xs=win.xs, ys=win.ys, ts=win.ts, edges=win.edges, edge_ts=np.asarray(win.edge_ts), theta=theta, prev_theta=prev, alpha_handover=0.61,
             hp_names=np.array(sorted(hp)), hp_values=np.array([float(hp[n]) for n in sorted(hp)]))
    ref = M.run_case(jax, L, g)
    if abs(fscale) > 1.0:   # the case must exercise what it is there for
        wx, wy = ref['obj_warped_xs'], ref['obj_warped_ys']
        assert (np.rint(wx) < -1).any() and (np.rint(wx) > W).any() and (np.rint(wy) < -1).any() and (np.rint(wy) > H).any()
    _check_against_oracle(theta, prev, 0.61, win.args(), hp, win.sensor_size, ref)


@pytest.mark.parametrize('seed,H,W', [(0, 48, 64), (1, 61, 83), (2, 120, 160)])
def test_edge_oracle_matches_reference_img_utils_live(reference, seed, H, W):
    """the reference's own utils.img_utils functions (importable over the stand-in; they run OpenCV / SciPy, both in this image), called
    the way exp_mgr.py:343-350 chains them, against oracle/edge_oracle.py - which tests/test_gpu_edges.py holds the device kernels to"""
    from oracle import edge_oracle as E
    import utils.img_utils as IU
    assert os.path.realpath(IU.__file__).startswith(os.path.realpath(REFERENCE_SRC))
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[:H, :W]
    f = (127 + 90 * np.sin(xx / 7.0 + seed) * np.cos(yy / 5.0) + rng.normal(0, 6, (H, W))).clip(0, 255).astype(np.uint8)
    edge = IU.image_to_edge(f)                                                             # img_utils.py:192-207
    assert np.array_equal(E.canny(f, 30, 80), edge)
    np.testing.assert_allclose(E.smoothen_edges(edge), IU.smoothen_edges(edge), rtol=0, atol=1e-10)            # :210-220
    np.testing.assert_allclose(E.eincm_inv_exp_dist_transform(edge, 6 / 5.541), IU.eincm_inv_exp_dist_transform(edge, 6 / 5.541), rtol=0, atol=1e-12)  # :229-235
    np.testing.assert_allclose(E.edge_map(f), IU.normalize_to_unit_range(IU.smoothen_edges(IU.image_to_edge(f))), rtol=0, atol=1e-10)
    # preprocess_image up to (not including) the bilateral filter, :147-181, in the reference's positional argument order
    import cv2 as cv
    d = E.fast_nl_means_denoising(f, 4, 3, 11)
    c = E.clahe_apply(d, 5.0, (10, 10))
    sharp = E.add_weighted_u8(c, 1.5, E.gaussian_blur_u8(c, 3), -0.5)
    assert np.array_equal(cv.bilateralFilter(sharp, 5, 15, 15), IU.preprocess_image(f))


def test_ingest_oracle_matches_reference_loaders_live(reference, capsys):
    """the reference's own loader methods (dsec_loader.py:145-186, 285-349; mvsec_loader.py:100-135), run on in-memory arrays - the
    loader objects are made without their file-opening constructors - against oracle/ingest_oracle.py, which tests/test_gpu_ingest.py
    holds the device kernels to"""
    from oracle import ingest_oracle as I
    import dataloaders.dsec_loader as DL
    import dataloaders.mvsec_loader as ML
    assert os.path.realpath(DL.__file__).startswith(os.path.realpath(REFERENCE_SRC))
    rng = np.random.default_rng(3)
    H, W, n = 48, 64, 20000
    ev = {'x': rng.integers(0, W, n).astype(np.uint16), 'y': rng.integers(0, H, n).astype(np.uint16),
          't': np.sort(rng.integers(0, 2_000_000, n)).astype(np.int64), 'p': rng.integers(0, 2, n).astype(np.uint8)}
    yy, xx = np.mgrid[:H, :W].astype(np.float32)
    rmap = np.stack([xx * 1.07 - 2.3 + 0.5 * np.sin(yy / 5), yy * 0.94 + 1.8 + rng.choice([0.0, 0.5], (H, W))], axis=-1).astype(np.float32)  # ties: .5
    ld = object.__new__(DL.DSECDataLoader)
    ld.rectify_map, ld.height, ld.width = rmap, H, W
    ld.l_events = {k: v.copy() for k, v in ev.items()}
    ld.rectify_events()
    want = I.rectify_events(ev['x'], ev['y'], ev['t'], ev['p'], rmap, H, W)
    for got, w_ in zip((ld.l_events['x'], ld.l_events['y'], ld.l_events['t'], ld.l_events['p']), want):
        assert got.dtype == w_.dtype and np.array_equal(got, w_)
    assert 0 < len(want[0]) < n                                                  # some events left the sensor

    # fixed-N window rule through get_sample: event "x" = its own index, so the slice tells (start, end)
    m = len(ld.l_events['x'])
    ld.l_events['x'] = np.arange(m)
    ld.t_offset, ld.data_split = 1_000, 'test'
    ld.eval_ts_us = np.array([[1_000, 101_000, 0], [500_000, 600_000, 1], [1_900_000, 2_001_000, 2], [700_000, 1_500_000, 3]], dtype=np.int64)
    ld.l_image_ts_us, ld.l_image_paths = np.array([0, 50_000, 2_100_000]), []
    ld.precompute_eval_event_indices()
    ld.precompute_eval_image_indices()
    for des, latest in [(None, False), (1500, False), (1500, True), (12000, False), (30000, True)]:
        ld.des_n_events, ld.prefer_latest_events, ld.n_event_deficiency = des, latest, 0
        for k in range(len(ld.eval_ts_us)):
            s_ = ld.get_sample(k)
            a0, a1 = int(ld.eval_event_start_idxs[k]), int(ld.eval_event_end_idxs[k])
            assert a0 == int(np.searchsorted(ld.l_events['t'], ld.eval_ts_us[k, 0] - ld.t_offset, side='left'))
            b0, b1, deficiency = I.window_event_range(a0, a1, m, des, latest)
            got = s_['events']['x']
            assert (int(got[0]), int(got[-1]) + 1) == (b0, b1) and len(got) == b1 - b0, (des, latest, k)
            if des is not None:
                assert int(s_['n_event_deficiency']) == deficiency
            np.testing.assert_array_equal(s_['events']['t'], ld.l_events['t'][b0:b1] + ld.t_offset)

    # MVSEC crop
    class Reader:
        def open_file(self): pass
        def close_file(self): pass
        def read_h5_dataset(self, name):
            return {'davis/left/events': raw, 'davis/left/image_raw': np.zeros((2, 260, 346), np.uint8)}.get(name, np.zeros(3))
    raw = np.stack([rng.integers(0, 346, 5000), rng.integers(0, 260, 5000), np.sort(rng.random(5000)) * 10, rng.choice([-1, 1], 5000)], axis=-1).astype(np.float64)
    mv = object.__new__(ML.MVSECDataLoader)
    mv.mvsec_h5_rdr = Reader()
    mv.load_left_data()
    want = I.crop_events(*raw.T)
    for k, w_ in zip('xytp', want):
        assert mv.l_events[k].dtype == w_.dtype and np.array_equal(mv.l_events[k], w_)
    assert mv.l_image_raw.shape == (2, 256, 336)
    capsys.readouterr()


def test_mirror_interfaces_match_the_reference_signatures(reference):
    """drop-in boundary: the host-side mirror takes the reference's argument names, order and defaults - read off the
    reference's own function objects, not restated"""
    import inspect
    import eincm.losses as RL
    import eincm.solver as RS
    import evaluations.theta_eval as RTE
    import evaluations.flow_eval as RFE
    import utils.img_utils as RIU
    from eincm_b200 import evaluations as ME, img_utils as MIU, losses as ML, solver as MS

    def params(f):
        return [(n, p.default) for n, p in inspect.signature(f).parameters.items()]

    for name in ('loss_func', 'handover_loss_func', 'compute_weights_for_multi_reference', 'compute_loss_objectives'):
        assert params(getattr(ML, name)) == params(getattr(RL, name)), name
    assert params(ME.sparse_flow_error) == params(RFE.sparse_flow_error)
    ref_eval, mir_eval = params(RTE.evaluate_theta_array), params(ME.evaluate_theta_array)
    assert mir_eval[:len(ref_eval)] == ref_eval                                  # the mirror may append device / stream keywords
    for name in ('image_to_edge', 'smoothen_edges', 'eincm_inv_exp_dist_transform', 'preprocess_image'):
        r, m = params(getattr(RIU, name)), params(getattr(MIU, name))
        assert m[:len(r)] == r, name
    # solver: constructor keywords (the mirror takes an objective in place of the two loss partials), state and result keys
    r = [n for n, _ in params(RS.MultipleLevelEINCMSolver.__init__)]
    m = [n for n, _ in params(MS.MultipleLevelEINCMSolver.__init__)]
    dropped = {'theta_loss_pfunc', 'handover_loss_pfunc'}
    assert [n for n in r if n not in dropped] == [n for n in m if n in r]
    for meth in ('set_datasample', 'solve', 'not_first_sample', '_pre_solve', '_stage_prior_theta_pyr', '_perform_handover_at_level', '_upscale_theta',
                 '_downscale_theta', '_initialize_theta_pyramids', '_initialize_handover_weights'):
        assert hasattr(MS.MultipleLevelEINCMSolver, meth), meth
    for meth in ('set_datasample', '_initialize_theta_pyramids', '_upscale_theta', '_downscale_theta', '_perform_handover_at_level'):
        assert params(getattr(MS.MultipleLevelEINCMSolver, meth)) == params(getattr(RS.MultipleLevelEINCMSolver, meth)), meth


def test_tv_count_depends_on_summation_order(reference, monkeypatch):
    """regularizers.py:26-29 counts pixels whose flow gradient is exactly non-zero.  On a constant flow field the reference source yields
    the oracle's value when antisymmetric taps are differenced first (the canonical order of the oracle and the kernels) and another one
    when the nine taps are accumulated in turn: XLA's order is not knowable from the source, which is why the order is fixed by decree."""
    M, jax, L = reference
    import jax.numpy as jnp
    win = S.make_window(seed=77, H=40, W=56, N=3000, edge_ts=(0.0, 1.0))
    theta_full = np.broadcast_to(np.array([3.7, -1.3]), (40, 56, 2)).copy()
    want = O.per_pix_total_variation(theta_full, win.xs, win.ys, win.ts)
    import eincm.regularizers as Rg
    monkeypatch.setenv('JAX_SHIM_CONV', 'pairs')
    got_pairs = float(Rg.per_pix_total_variation(jnp.array(theta_full), jnp.array(win.xs), jnp.array(win.ys), jnp.array(win.ts)))
    monkeypatch.delenv('JAX_SHIM_CONV')
    got_turn = float(Rg.per_pix_total_variation(jnp.array(theta_full), jnp.array(win.xs), jnp.array(win.ys), jnp.array(win.ts)))
    assert got_pairs == pytest.approx(want, rel=1e-12)
    assert got_turn != pytest.approx(want, rel=1e-3)


def test_reference_source_vectors_are_current(reference):
    """the committed vectors are what the reference source yields today"""
    M, jax, L = reference
    for name in G.NAMES:
        z = np.load(os.path.join(G.GOLDEN_DIR, name + '.npz'))
        out = M.run_case(jax, L, {k: z[k] for k in z.files})
        ref = load_refsrc(name)
        assert out['loss'] == pytest.approx(float(ref['loss']), rel=1e-13)          # not bitwise: torch's reductions split by thread count
        np.testing.assert_allclose(out['grad'], ref['grad'], rtol=1e-11, atol=1e-15)
