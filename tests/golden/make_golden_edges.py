"""Generates the golden fixtures of the edge-image stage under tests/golden/edges/ (committed together with this script).

    python tests/golden/make_golden_edges.py        # needs cv2 (4.13.0 in this image) and scipy

Unlike the objective (whose reference needs JAX), this stage's reference arithmetic IS runnable here: the fixtures hold the outputs
of OpenCV and SciPy themselves, called exactly as the reference calls them (src/utils/img_utils.py:194-235,
src/experiments/e00/exp_mgr.py:343-350):

    canny  = cv.Canny(img, th1, th2, None, 3, True)
    gauss  = normalize_to_unit_range(cv.GaussianBlur(canny.astype(float64), None, 1, 1, 0))
    iedt   = normalize_to_unit_range(eincm_inv_exp_dist_transform(canny, alpha))
    nlm    = cv.fastNlMeansDenoising(img, None, 4, 3, 11)              (first step of preprocess_image, img_utils.py:147-157)
    clahe  = cv.createCLAHE(clipLimit=5, tileGridSize=(10, 10)).apply(nlm)                              (img_utils.py:159-161)
    blur   = cv.GaussianBlur(clahe, None, 3, 2, 0)                      (img_utils.py:165-169, positional like the reference)
    sharp  = cv.addWeighted(clahe, 1.5, blur, -0.5, 0)                                                  (img_utils.py:171-178)

They pin the NumPy restatement (oracle/edge_oracle.py, tests/test_edge_oracle.py) and the CUDA path (tests/test_gpu_edges.py).
"""
import os
import sys

import cv2 as cv
import numpy as np
from scipy import ndimage

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from eincm_b200 import synth  # noqa: E402

EPS = np.finfo(np.float64).eps
ALPHA = 6.0 / 5.541           # configs/edge_extraction/smoothen/iedt.yaml


def normalize_to_unit_range(a):                       # src/utils/img_utils.py:24-25
    return (a - a.min()) / (a.max() - a.min() + EPS)


def ref_iedt(edge_img, alpha):                        # src/utils/img_utils.py:231-235, verbatim semantics
    d = ndimage.distance_transform_edt(~(edge_img.astype('bool')))
    e = 1 - np.exp(-d / alpha)
    return 1 - normalize_to_unit_range(e)


CASES = [
    # name, H, W, R, seed, th1, th2, noise
    ('frames_96x128_dsec_thresholds', 96, 128, 3, 1, 30, 80, 3.0),
    ('frames_61x83_mvsec_thresholds', 61, 83, 2, 2, 100, 200, 2.0),
    ('noise_40x56', 40, 56, 1, 3, 30, 80, None),
]


def main():
    out_dir = os.path.join(HERE, 'edges')
    os.makedirs(out_dir, exist_ok=True)
    for name, H, W, R, seed, th1, th2, noise in CASES:
        if noise is None:
            frames = np.random.default_rng(seed).integers(0, 256, size=(R, H, W)).astype(np.uint8)
        else:
            frames = synth.make_frames(H, W, R, seed=seed, noise_sigma=noise)
        canny = np.stack([cv.Canny(f, th1, th2, None, 3, True) for f in frames])
        gauss = np.stack([normalize_to_unit_range(cv.GaussianBlur(c.astype(np.float64), None, 1, 1, 0)) for c in canny])
        iedt = np.stack([normalize_to_unit_range(ref_iedt(c, ALPHA)) for c in canny])
        nlm = np.stack([cv.fastNlMeansDenoising(f, None, 4, 3, 11) for f in frames])     # denoise/default.yaml: h 4, template 3, search 11
        clahe = np.stack([cv.createCLAHE(clipLimit=5, tileGridSize=(10, 10)).apply(f) for f in nlm])
        blur = np.stack([cv.GaussianBlur(f, None, 3, 2, 0) for f in clahe])
        sharp = np.stack([cv.addWeighted(a, 1.5, b, -0.5, 0) for a, b in zip(clahe, blur)])
        np.savez_compressed(os.path.join(out_dir, name + '.npz'), frames=frames, canny=canny, gauss=gauss, iedt=iedt, nlm=nlm,
                            clahe=clahe, blur=blur, sharp=sharp,
                            th=np.array([th1, th2], np.float64), alpha=np.float64(ALPHA), cv_version=np.array(cv.__version__))
        print(name, frames.shape, 'edge pixels', int((canny > 0).sum()))


if __name__ == '__main__':
    main()
