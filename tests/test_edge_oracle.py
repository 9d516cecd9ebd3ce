"""Pins the NumPy restatement of the edge-image stage (oracle/edge_oracle.py) against OpenCV / SciPy outputs: the committed golden
fixtures (tests/golden/edges/*.npz, generated with cv2 4.13.0 by tests/golden/make_golden_edges.py) and, where cv2 is importable,
cv2 itself on fresh inputs.  Canny is integer work: bit-exact.  Float64 stages: 1e-12 (summation order)."""
import glob
import os

import numpy as np
import pytest

import eincm_b200.synth as S
from oracle import edge_oracle as E

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'edges', '*.npz')))
FTOL = 1e-12


def test_fixtures_exist():
    assert len(GOLD) >= 3


@pytest.mark.parametrize('path', GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_opencv_fixtures(path):
    z = np.load(path)
    th1, th2 = z['th']
    for f, canny, gauss, iedt in zip(z['frames'], z['canny'], z['gauss'], z['iedt']):
        c = E.canny(f, th1, th2)
        assert np.array_equal(c, canny)                                       # bit-exact
        np.testing.assert_allclose(E.edge_map(f, th1, th2, 'gaussian', 1), gauss, rtol=0, atol=FTOL)
        np.testing.assert_allclose(E.edge_map(f, th1, th2, 'iedt', alpha=float(z['alpha'])), iedt, rtol=0, atol=FTOL)


@pytest.mark.parametrize('path', GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_clahe_sharpen_oracle_matches_opencv_fixtures(path):
    """CLAHE / uint8 Gaussian blur / addWeighted as preprocess_image calls them (img_utils.py:159-178), against the committed OpenCV outputs
    (frame sizes that do and do not divide by the 10 x 10 tile grid): bit-exact, no cv2 needed."""
    z = np.load(path)
    for f, clahe, blur, sharp in zip(z['nlm'], z['clahe'], z['blur'], z['sharp']):
        assert np.array_equal(E.clahe_apply(f, 5, (10, 10)), clahe)
        assert np.array_equal(E.gaussian_blur_u8(clahe, 3.0), blur)
        assert np.array_equal(E.add_weighted_u8(clahe, 1.5, blur, -0.5), sharp)


@pytest.mark.parametrize('path', GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_nlm_oracle_matches_opencv_fixtures(path):
    z = np.load(path)
    for f, ref in zip(z['frames'], z['nlm']):
        assert np.array_equal(E.fast_nl_means_denoising(f, 4, 3, 11), ref)    # bit-exact


def test_threshold_rule():
    assert E.canny_thresholds(30, 80) == (900, 6400)
    assert E.canny_thresholds(200, 100) == (10000, 40000)                     # swapped like cv.Canny does
    assert E.canny_thresholds(0, 1e9) == (0, 32767 * 32767)
    assert E.TG22 == 13573


def test_gaussian_taps_known_answer():
    w = E.gaussian_kernel_f64(1.0)
    assert len(w) == 9 and abs(w.sum() - 1.0) < 1e-15
    # centre tap of a 9-tap sigma-1 kernel: 1 / sum_k exp(-k^2 / 2), k = -4..4
    assert abs(w[4] - 1.0 / sum(np.exp(-0.5 * k * k) for k in range(-4, 5))) < 1e-16


def test_edt_known_answer():
    m = np.ones((5, 7), bool)
    m[2, 3] = False
    d = E.distance_transform_edt(m)
    yy, xx = np.mgrid[0:5, 0:7]
    np.testing.assert_array_equal(d, np.sqrt((yy - 2.0) ** 2 + (xx - 3.0) ** 2))


def test_flat_image_has_no_edges():
    f = np.full((20, 30), 77, np.uint8)
    assert E.canny(f, 30, 80).max() == 0
    assert np.all(E.edge_map(f, 30, 80) == 0.0)


try:
    import cv2 as cv
except ImportError:                                   # the fixtures above cover boxes without OpenCV
    cv = None
needs_cv = pytest.mark.skipif(cv is None, reason='OpenCV is the reference of this stage; not importable here')


@needs_cv
@pytest.mark.parametrize('seed,H,W,th', [(11, 72, 96, (30, 80)), (12, 50, 41, (100, 200)), (13, 33, 64, (10, 300))])
def test_oracle_matches_live_opencv(seed, H, W, th):
    from scipy import ndimage
    for f in S.make_frames(H, W, 2, seed=seed):
        ref = cv.Canny(f, th[0], th[1], None, 3, True)
        assert np.array_equal(E.canny(f, *th), ref)
        np.testing.assert_allclose(E.smoothen_edges(ref, 1, 1), cv.GaussianBlur(ref.astype(np.float64), None, 1, 1, 0), rtol=0, atol=1e-10)
        np.testing.assert_array_equal(E.distance_transform_edt(~ref.astype(bool)), ndimage.distance_transform_edt(~ref.astype(bool)))
    rng = np.random.default_rng(seed)
    f = (rng.integers(0, 5, size=(H, W)) * 60).astype(np.uint8)                # many exact ties in the magnitudes
    assert np.array_equal(E.canny(f, *th), cv.Canny(f, th[0], th[1], None, 3, True))


@needs_cv
@pytest.mark.parametrize('shape,h,t,sw', [((40, 56), 4, 3, 11), ((5, 7), 4, 3, 11), ((3, 40), 4, 3, 11), ((30, 31), 7.5, 5, 9),
                                          ((30, 31), 4, 4, 10), ((24, 24), 10, 7, 21)])
def test_nlm_oracle_matches_live_opencv(shape, h, t, sw):
    f = np.random.default_rng(shape[0] * 7 + shape[1]).integers(0, 256, size=shape).astype(np.uint8)
    assert np.array_equal(E.fast_nl_means_denoising(f, h, t, sw), cv.fastNlMeansDenoising(f, None, h, t, sw))
    g = S.make_frames(max(shape[0], 8), max(shape[1], 8), 1, seed=3, noise_sigma=5.0)[0]
    assert np.array_equal(E.fast_nl_means_denoising(g, h, t, sw), cv.fastNlMeansDenoising(g, None, h, t, sw))


@needs_cv
@pytest.mark.parametrize('H,W,seed', [(480, 640, 0), (96, 128, 1), (61, 83, 2), (100, 100, 3)])
def test_clahe_sharpen_restatements_match_live_opencv(H, W, seed):
    """The steps of preprocess_image between the denoise and the bilateral filter (img_utils.py:159-181), called as the reference calls
    them (CUDA counterparts: eincm_clahe / eincm_sharpen, tests/test_gpu_edges.py): bit-exact."""
    f = S.make_frames(H, W, 1, seed=seed, noise_sigma=4.0)[0]
    clahe = cv.createCLAHE(clipLimit=5, tileGridSize=(10, 10)).apply(f)
    assert np.array_equal(E.clahe_apply(f, 5, (10, 10)), clahe)
    blur = cv.GaussianBlur(clahe, None, 3, 2, 0)                              # positional like img_utils.py:165-169: sigmaX = 3
    assert np.array_equal(E.gaussian_blur_u8(clahe, 3.0), blur)
    assert np.array_equal(E.add_weighted_u8(clahe, 1.5, blur, -0.5), cv.addWeighted(clahe, 1.5, blur, -0.5, 0))
    assert int(E.gaussian_kernel_u8_fixed_point(3.0).sum()) == 256 and len(E.gaussian_kernel_u8_fixed_point(3.0)) == 19
