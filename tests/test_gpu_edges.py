"""Parity of the CUDA edge-image stage (eincm_edge_maps, through the C-ABI) against OpenCV / SciPy outputs (committed fixtures,
tests/golden/edges) and the NumPy restatement (oracle/edge_oracle.py).  cv.Canny's image bit-exact; float64 stages to 1e-12."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import eincm_b200.synth as S
from oracle import edge_oracle as E

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'edges', '*.npz')))
FTOL = 1e-12


@pytest.fixture(scope='module')
def I():
    from eincm_b200 import img_utils
    return img_utils


@pytest.mark.parametrize('path', GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_edge_maps_match_opencv_fixtures(I, path):
    z = np.load(path)
    th1, th2 = z['th']
    g, canny = I.edge_maps(z['frames'], th1, th2, smoothen='gaussian', k_size=1, return_canny=True)
    assert np.array_equal(canny.cpu().numpy(), z['canny'])
    np.testing.assert_allclose(g.cpu().numpy(), z['gauss'], rtol=0, atol=FTOL)
    d = I.edge_maps(z['frames'], th1, th2, smoothen='iedt', alpha=float(z['alpha']))
    np.testing.assert_allclose(d.cpu().numpy(), z['iedt'], rtol=0, atol=FTOL)


def test_dsec_sized_window_against_oracle(I):
    """Full DSEC frame size, R = 3 (BASELINE.json configs[1]); long contours cross many CTAs (hysteresis as union-find)."""
    frames = S.make_frames(480, 640, 3, seed=5)
    g, canny = I.edge_maps(frames, 30, 80, return_canny=True)
    ref = np.stack([E.canny(f, 30, 80) for f in frames])
    assert np.array_equal(canny.cpu().numpy(), ref)
    assert int((ref > 0).sum()) > 5000
    np.testing.assert_allclose(g.cpu().numpy(), np.stack([E.edge_map(f, 30, 80) for f in frames]), rtol=0, atol=FTOL)
    d = I.edge_maps(frames[:1], 30, 80, smoothen='iedt')
    np.testing.assert_allclose(d.cpu().numpy()[0], E.edge_map(frames[0], 30, 80, 'iedt'), rtol=0, atol=FTOL)


def test_hysteresis_follows_a_long_weak_contour(I):
    """A spiral whose contrast decays: only its start is above the high threshold, the rest is kept by connectivity alone."""
    H, W = 200, 240
    img = np.full((H, W), 100.0)
    t = np.linspace(0, 10 * np.pi, 20000)
    r = 4 + 2.6 * t
    x = np.rint(W / 2 + r * np.cos(t)).astype(int); y = np.rint(H / 2 + r * np.sin(t)).astype(int)
    ok = (x >= 0) & (x < W) & (y >= 0) & (y < H)
    amp = np.where(t < 2.0, 60.0, 12.0)                                       # gradient ~ 4 * amp: 240 at the start, 48 later
    img[y[ok], x[ok]] += amp[ok]
    f = np.clip(img, 0, 255).astype(np.uint8)
    ref = E.canny(f, 30, 80)
    m = E.canny_candidates(f, 30, 80)
    assert (m == 2).sum() < 100 and (ref > 0).sum() > 1000                 # ~28 strong candidates carry ~1250 weak ones
    got = I.image_to_edge(f, 3, 30, 80)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize('shape', [(1, 1), (1, 9), (7, 1), (3, 3), (17, 33), (16, 32), (31, 65)])
def test_small_and_ragged_sizes(I, shape):
    rng = np.random.default_rng(shape[0] * 100 + shape[1])
    f = rng.integers(0, 256, size=shape).astype(np.uint8)
    g, canny = I.edge_maps(f, 30, 80, return_canny=True)
    assert np.array_equal(canny.cpu().numpy(), E.canny(f, 30, 80))
    np.testing.assert_allclose(g.cpu().numpy(), E.edge_map(f, 30, 80), rtol=0, atol=FTOL)


def test_flat_frame_and_swapped_thresholds(I):
    f = np.full((40, 50), 10, np.uint8)
    g, canny = I.edge_maps(f, 30, 80, return_canny=True)
    assert int(canny.max()) == 0 and float(g.abs().max()) == 0.0
    fr = S.make_frames(64, 80, 1, seed=8)[0]
    assert np.array_equal(I.image_to_edge(fr, 3, 200, 100), E.canny(fr, 100, 200))


def test_stand_alone_stages(I):
    fr = S.make_frames(72, 96, 1, seed=9)[0]
    e = I.image_to_edge(fr, 3, 30, 80)
    np.testing.assert_allclose(I.smoothen_edges(e, 1, 1), E.smoothen_edges(e, 1, 1), rtol=0, atol=1e-10)
    np.testing.assert_allclose(I.eincm_inv_exp_dist_transform(e, 6), E.eincm_inv_exp_dist_transform(e, 6), rtol=0, atol=FTOL)
    np.testing.assert_allclose(I.normalize_to_unit_range(e.astype(np.float64)).cpu().numpy(), E.normalize_to_unit_range(e), rtol=0, atol=0)


def test_edges_feed_the_objective(I):
    """The device tensor is the `edges` operand of loss_func: same loss as with the oracle's edge images."""
    from eincm_b200 import losses
    w = S.make_workload('tiny', seed=3)
    H, W = w.sensor_size
    frames = S.make_frames(H, W, len(w.edge_ts), seed=4)
    edges = I.edge_maps(frames, 30, 80)
    ref_edges = np.stack([E.edge_map(f, 30, 80) for f in frames])
    kw = dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.0, cur_pyr_lvl=1, n_pyr_lvls=5, sensor_size=w.sensor_size,
              scale_to_sensor_size_method='bilinear')
    th = S.theta_test_points(w, (4, 4))['perturbed']
    l1, g1 = losses.value_and_grad(losses.loss_func)(th, w.xs, w.ys, w.ts, edges.cpu().numpy(), w.edge_ts, **kw)
    l2, g2 = losses.value_and_grad(losses.loss_func)(th, w.xs, w.ys, w.ts, ref_edges, w.edge_ts, **kw)
    assert abs(l1 - l2) <= 1e-9 * abs(l2)
    losses.clear_cache()


def test_error_behaviour(I):
    from eincm_b200 import plan
    with pytest.raises(plan.EincmError):
        I.edge_maps(np.zeros((4, 4), np.float64))                           # not uint8
    with pytest.raises(plan.EincmError):
        I.edge_maps(np.zeros((4, 4), np.uint8), smoothen='median')
    with pytest.raises(plan.EincmError):
        I.edge_maps(np.zeros((4, 4), np.uint8), k_size=0)                   # sigma must be positive
    with pytest.raises(plan.EincmError):
        I.image_to_edge(np.zeros((4, 4), np.uint8), apert_size=5)
    lib = plan.load_library()
    assert lib.eincm_edge_workspace_bytes(0, 4, 1) == 0
    p = plan.EdgeParams(30.0, 80.0, 0, 0, 1.0, 1.0)
    assert lib.eincm_edge_maps(0, None, 1, 4, 4, C.byref(p), None, None, None, 0, None) == plan.EINCM_EINVAL


@pytest.mark.parametrize('path', GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_nlm_denoise_matches_opencv_fixtures(I, path):
    z = np.load(path)
    got = I.fast_nl_means_denoising(z['frames'], 4, 3, 11)
    assert np.array_equal(got.cpu().numpy(), z['nlm'])                        # bit-exact with cv.fastNlMeansDenoising


@pytest.mark.parametrize('shape,h,t,sw', [((480, 640), 4, 3, 11), ((5, 7), 4, 3, 11), ((3, 40), 4, 3, 11), ((61, 83), 7.5, 5, 9),
                                          ((33, 65), 4, 4, 10), ((40, 40), 10, 7, 21), ((1, 1), 4, 3, 11)])
def test_nlm_denoise_matches_oracle(I, shape, h, t, sw):
    if shape == (480, 640):
        f = S.make_frames(480, 640, 1, seed=12, noise_sigma=4.0)[0]
    else:
        f = np.random.default_rng(shape[0] * 7 + shape[1]).integers(0, 256, size=shape).astype(np.uint8)
    got = I.fast_nl_means_denoising(f, h, t, sw).cpu().numpy()
    assert np.array_equal(got, E.fast_nl_means_denoising(f, h, t, sw))


def test_nlm_error_behaviour(I):
    from eincm_b200 import plan
    with pytest.raises(plan.EincmError):
        I.fast_nl_means_denoising(np.zeros((4, 4), np.float32))
    with pytest.raises(plan.EincmError):
        I.fast_nl_means_denoising(np.zeros((4, 4), np.uint8), 4, 3, 99)        # search window beyond the supported size


# ---- CLAHE and the Gaussian sharpen of preprocess_image (src/utils/img_utils.py:159-178): eincm_clahe / eincm_sharpen --------------------
@pytest.mark.parametrize('path', GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_clahe_sharpen_match_opencv_fixtures(I, path):
    z = np.load(path)
    clahe = I.clahe_apply(z['nlm'], 5, (10, 10))
    assert np.array_equal(clahe.cpu().numpy(), z['clahe'])                    # bit-exact with cv.createCLAHE(5, (10, 10)).apply
    sharp, blur = I.sharpen(clahe, 3, 1.5, -0.5, return_blur=True)            # chained on the device
    assert np.array_equal(blur.cpu().numpy(), z['blur'])                      # cv.GaussianBlur(uint8)
    assert np.array_equal(sharp.cpu().numpy(), z['sharp'])                    # cv.addWeighted


@pytest.mark.parametrize('shape,clip,grid', [((480, 640), 5, (10, 10)), ((260, 346), 5, (10, 10)), ((256, 336), 2.5, (8, 8)), ((61, 83), 40, (4, 6)),
                                             ((100, 100), 0, (10, 10)), ((33, 47), 5, (3, 2)), ((64, 64), 1000, (8, 8))])
def test_clahe_matches_oracle(I, shape, clip, grid):
    rng = np.random.default_rng(shape[0] * 13 + shape[1])
    frames = [S.make_frames(shape[0], shape[1], 1, seed=shape[1], noise_sigma=4.0)[0],
              rng.integers(0, 256, size=shape).astype(np.uint8),
              np.full(shape, 37, np.uint8),                                          # one bin holds everything: the whole tile is clipped
              (np.add.outer(np.arange(shape[0]), np.arange(shape[1])) % 256).astype(np.uint8)]
    got = I.clahe_apply(np.stack(frames), clip, grid).cpu().numpy()
    for g, f in zip(got, frames):
        assert np.array_equal(g, E.clahe_apply(f, clip, grid))


@pytest.mark.parametrize('shape,sigma,alpha,beta,gamma', [((480, 640), 3, 1.5, -0.5, 0.0), ((260, 346), 3, 1.5, -0.5, 0.0), ((61, 83), 1.0, 2.0, -1.0, 3.0),
                                                          ((20, 33), 2.2, 0.5, 0.5, 0.0), ((40, 40), 5, 1.7, -0.7, -2.5)])
def test_sharpen_matches_oracle(I, shape, sigma, alpha, beta, gamma):
    rng = np.random.default_rng(shape[0] * 17 + shape[1])
    frames = np.stack([S.make_frames(shape[0], shape[1], 1, seed=shape[0], noise_sigma=4.0)[0], rng.integers(0, 256, size=shape).astype(np.uint8)])
    sharp, blur = I.sharpen(frames, sigma, alpha, beta, gamma, return_blur=True)
    for k, f in enumerate(frames):
        b = E.gaussian_blur_u8(f, float(sigma))
        assert np.array_equal(blur[k].cpu().numpy(), b)
        assert np.array_equal(sharp[k].cpu().numpy(), E.add_weighted_u8(f, alpha, b, beta, gamma))
    one = I.sharpen(frames[0], sigma, alpha, beta, gamma)                            # a single (H, W) frame
    assert np.array_equal(one.cpu().numpy(), sharp[0].cpu().numpy())


def test_clahe_sharpen_error_behaviour(I):
    from eincm_b200 import plan
    with pytest.raises(plan.EincmError):
        I.clahe_apply(np.zeros((8, 8), np.float32))
    with pytest.raises(plan.EincmError):
        I.clahe_apply(np.zeros((8, 8), np.uint8), 5, (0, 4))
    with pytest.raises(plan.EincmError):
        I.clahe_apply(np.zeros((4, 4), np.uint8), 5, (10, 10))                        # tile grid larger than one reflection of the frame
    with pytest.raises(plan.EincmError):
        I.sharpen(np.zeros((8, 8), np.uint8), 50.0)                                   # beyond 63 taps
    with pytest.raises(plan.EincmError):
        I.sharpen(np.zeros((8, 8), np.uint8), 0.0)


def test_preprocess_image_matches_opencv_chain(I):
    """preprocess_image (src/utils/img_utils.py:131-191): denoise, CLAHE and sharpen on the device, the bilateral filter through OpenCV -
    against the all-OpenCV chain the reference runs."""
    cv = pytest.importorskip('cv2')
    f = S.make_frames(260, 346, 1, seed=9, noise_sigma=4.0)[0]
    d = cv.fastNlMeansDenoising(f, None, 4, 3, 11)
    c = cv.createCLAHE(clipLimit=5, tileGridSize=(10, 10)).apply(d)
    b = cv.GaussianBlur(c, None, 3, 2, 0)
    ref = cv.bilateralFilter(cv.addWeighted(c, 1.5, b, -0.5, 0), 5, 15, 15)
    assert np.array_equal(I.preprocess_image(f), ref)
