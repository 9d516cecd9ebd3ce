"""CPU restatement of the reference's edge-map production - TEST INFRASTRUCTURE, not product.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.

The reference builds the edge images of a window with OpenCV and SciPy (reference src/experiments/e00/exp_mgr.py:343-350):

    edges_r = normalize_to_unit_range(smoothen_edges(image_to_edge(u8 image)))

* ``image_to_edge``  = ``cv.Canny(img, th1, th2, None, 3, L2gradient=True)``            (src/utils/img_utils.py:194-211)
* ``smoothen_edges`` = ``cv.GaussianBlur(edge.astype(f64), None, k_size, sigma, 0)``      (src/utils/img_utils.py:213-222);
  the positional arguments land on (src, ksize=None, sigmaX=k_size, dst=sigma, sigmaY=0): the kernel size is derived from
  sigmaX = k_size (9 taps for k_size = 1 at float64 depth), ``sigma`` is ignored.  Restated as OpenCV behaves, not as named.
* alternative smoothing ``eincm_inv_exp_dist_transform`` (src/utils/img_utils.py:231-235):
  ``1 - normalize(1 - exp(-EDT(~edge) / alpha))`` with ``scipy.ndimage.distance_transform_edt``
* ``normalize_to_unit_range`` (src/utils/img_utils.py:24-25)

The arithmetic lives in third-party libraries that are not under /root/reference: OpenCV (unpinned, ``pip install
opencv-python``; 4.13.0 is importable in this image) and SciPy.  OpenCV's published Canny algorithm (modules/imgproc/src/canny.cpp)
is restated here in NumPy; the restatement is PINNED against ``cv2`` itself in tests/test_edge_oracle.py (run wherever cv2 is
importable) and against golden fixtures generated with cv2 4.13.0 (tests/golden/edges_*.npz, tests/golden/make_golden_edges.py).
"""
import numpy as np

EPS = np.finfo(np.float64).eps          # jnp.finfo(jnp.float64).eps, img_utils.py:25
CANNY_SHIFT = 15
TG22 = int(0.4142135623730950488016887242097 * (1 << CANNY_SHIFT) + 0.5)     # 13573


def normalize_to_unit_range(a):
    """src/utils/img_utils.py:24-25"""
    a = np.asarray(a, dtype=np.float64)
    return (a - a.min()) / (a.max() - a.min() + EPS)


def sobel3_replicate(img_u8):
    """cv.Sobel(src, CV_16S, 1|0, 0|1, ksize=3, borderType=BORDER_REPLICATE) as cv.Canny calls it: (dx, dy) int16."""
    p = np.pad(np.asarray(img_u8).astype(np.int32), 1, mode='edge')
    dx = (p[:-2, 2:] - p[:-2, :-2]) + 2 * (p[1:-1, 2:] - p[1:-1, :-2]) + (p[2:, 2:] - p[2:, :-2])
    dy = (p[2:, :-2] - p[:-2, :-2]) + 2 * (p[2:, 1:-1] - p[:-2, 1:-1]) + (p[2:, 2:] - p[:-2, 2:])
    return dx.astype(np.int16), dy.astype(np.int16)


def canny_thresholds(th1, th2):
    """cv.Canny with L2gradient: thresholds are clamped to 32767, squared and floored; th1 > th2 are swapped."""
    lo, hi = (th2, th1) if th1 > th2 else (th1, th2)
    lo, hi = min(32767.0, float(lo)), min(32767.0, float(hi))
    if lo > 0:
        lo *= lo
    if hi > 0:
        hi *= hi
    return int(np.floor(lo)), int(np.floor(hi))


def canny_candidates(img_u8, th1, th2):
    """Non-maximum suppression of cv.Canny (L2 gradient, aperture 3).  Returns the map of OpenCV's first stage without its push
    shortcuts: 0 = local maximum above the low threshold (edge if connected to a strong one), 1 = not an edge, 2 = local maximum
    above the high threshold."""
    dx, dy = sobel3_replicate(img_u8)
    H, W = dx.shape
    xs, ys = dx.astype(np.int64), dy.astype(np.int64)
    mag = xs * xs + ys * ys
    lo, hi = canny_thresholds(th1, th2)
    mp = np.pad(mag, 1)                                        # zero border of the magnitude buffer
    c = mp[1:-1, 1:-1]
    left, right = mp[1:-1, :-2], mp[1:-1, 2:]
    up, down = mp[:-2, 1:-1], mp[2:, 1:-1]
    x, y = np.abs(xs), np.abs(ys) << CANNY_SHIFT
    tg22x = x * TG22
    tg67x = tg22x + (x << (CANNY_SHIFT + 1))
    horiz = y < tg22x
    vert = (~horiz) & (y > tg67x)
    diag = ~(horiz | vert)
    s_neg = (xs ^ ys) < 0                                       # s = -1: compare (row-1, col+1) and (row+1, col-1)
    up_l, up_r = mp[:-2, :-2], mp[:-2, 2:]
    dn_l, dn_r = mp[2:, :-2], mp[2:, 2:]
    d_prev = np.where(s_neg, up_r, up_l)                        # _mag_p[j - s]
    d_next = np.where(s_neg, dn_l, dn_r)                        # _mag_n[j + s]
    is_max = (horiz & (c > left) & (c >= right)) | (vert & (c > up) & (c >= down)) | (diag & (c > d_prev) & (c > d_next))
    cand = (c > lo) & is_max
    out = np.ones((H, W), np.uint8)
    out[cand] = 0
    out[cand & (c > hi)] = 2
    return out


def canny(img_u8, th1, th2):
    """cv.Canny(img, th1, th2, None, 3, True): 255 on candidates 8-connected to a strong candidate (src/utils/img_utils.py:194-211)."""
    m = canny_candidates(img_u8, th1, th2)
    H, W = m.shape
    edge = (m == 2)
    stack = list(zip(*np.nonzero(edge)))
    while stack:                                                # hysteresis: flood fill from the strong pixels
        r, c = stack.pop()
        for rr in range(max(r - 1, 0), min(r + 2, H)):
            for cc in range(max(c - 1, 0), min(c + 2, W)):
                if m[rr, cc] == 0 and not edge[rr, cc]:
                    edge[rr, cc] = True
                    stack.append((rr, cc))
    return np.where(edge, 255, 0).astype(np.uint8)


def gaussian_kernel_f64(sigma):
    """cv.getGaussianKernel(ksize, sigma, CV_64F) with ksize derived from sigma for a float64 image: round(sigma * 8 + 1) | 1."""
    k = int(round(sigma * 8 + 1)) | 1
    x = np.arange(k, dtype=np.float64) - (k - 1) * 0.5
    w = np.exp(-0.5 / (sigma * sigma) * x * x)
    return w / w.sum()


def smoothen_edges(edge_img, k_size=1, sigma=1):
    """src/utils/img_utils.py:213-222 as OpenCV executes it: separable Gaussian with sigmaX = sigmaY = k_size, BORDER_REFLECT_101."""
    del sigma                                                   # lands on the dst parameter of cv.GaussianBlur
    a = np.asarray(edge_img, dtype=np.float64)
    w = gaussian_kernel_f64(float(k_size))
    h = len(w) // 2
    p = np.pad(a, ((0, 0), (h, h)), mode='reflect')
    rowf = sum(w[i] * p[:, i:i + a.shape[1]] for i in range(len(w)))
    p = np.pad(rowf, ((h, h), (0, 0)), mode='reflect')
    return sum(w[i] * p[i:i + a.shape[0], :] for i in range(len(w)))


def distance_transform_edt(mask):
    """scipy.ndimage.distance_transform_edt(mask): exact Euclidean distance of every non-zero pixel to the nearest zero pixel."""
    mask = np.asarray(mask, dtype=bool)
    H, W = mask.shape
    INF = np.int64(1) << 40
    g = np.full((H, W), INF, np.int64)                          # vertical distance to the nearest zero pixel of the column
    run = np.full(W, INF, np.int64)
    for r in range(H):
        run = np.where(mask[r], np.minimum(run + 1, INF), 0)
        g[r] = run
    run = np.full(W, INF, np.int64)
    for r in range(H - 1, -1, -1):
        run = np.where(mask[r], np.minimum(run + 1, INF), 0)
        g[r] = np.minimum(g[r], run)
    g2 = np.where(g >= INF, INF, g * g)
    cols = np.arange(W, dtype=np.int64)
    dx2 = (cols[:, None] - cols[None, :]) ** 2                   # [x, x']
    d2 = np.empty((H, W), np.int64)
    for r in range(H):
        d2[r] = np.min(dx2 + g2[r][None, :], axis=1)
    return np.sqrt(d2.astype(np.float64))


def eincm_inv_exp_dist_transform(edge_img, alpha=6):
    """src/utils/img_utils.py:231-235"""
    d = distance_transform_edt(~(np.asarray(edge_img).astype(bool)))
    e = 1.0 - np.exp(-d / alpha)
    return 1.0 - normalize_to_unit_range(e)


def edge_map(img_u8, th1=30, th2=80, smoothen='gaussian', k_size=1, alpha=6 / 5.541):
    """One edge image of a window as exp_mgr.py:343-350 stages it (the u8 image is the pre-processed grayscale frame)."""
    e = canny(img_u8, th1, th2)
    s = smoothen_edges(e, k_size) if smoothen == 'gaussian' else eincm_inv_exp_dist_transform(e, alpha)
    return normalize_to_unit_range(s)


def fast_nl_means_denoising(src, h=4, template_win_size=3, search_win_size=11):
    """cv.fastNlMeansDenoising(src, None, h, template_win_size, search_win_size) for a uint8 single-channel image, as
    preprocess_image calls it (src/utils/img_utils.py:147-157).  Restates OpenCV's FastNlMeansDenoisingInvoker<uchar, int, unsigned,
    DistSquared, int> (modules/photo/src/fast_nlmeans_denoising_invoker.hpp): reflect-101 border of search / 2 + template / 2 pixels,
    integer patch distances, the fixed-point weight table indexed by dist >> log2ceil(template^2), rounded integer average.
    Pinned bit-exactly against cv2 (tests/test_edge_oracle.py, tests/golden/edges)."""
    import math
    src = np.asarray(src, np.uint8)
    th, sh = int(template_win_size) // 2, int(search_win_size) // 2
    tw, sw = 2 * th + 1, 2 * sh + 1
    b = th + sh
    H, W = src.shape
    idx_r = np.arange(-b, H + b); idx_c = np.arange(-b, W + b)

    def refl(i, n):                                             # cv::borderInterpolate, BORDER_REFLECT_101
        if n == 1:
            return np.zeros_like(i)
        i = i.copy()
        while True:
            bad = (i < 0) | (i >= n)
            if not bad.any():
                return i
            i = np.where(i < 0, -i, i)
            i = np.where(i >= n, 2 * (n - 1) - i, i)

    ext = src.astype(np.int64)[refl(idx_r, H)][:, refl(idx_c, W)]
    fixed_point_mult = min((2 ** 31 - 1) // (sw * sw * 255), 2 ** 31 - 1)
    tsq = tw * tw
    shift = 0
    while (1 << shift) < tsq:
        shift += 1
    mult = float(1 << shift) / tsq
    almost_max = int(255 * 255 / mult + 1)
    hh = float(np.float32(h) * np.float32(h))
    wt = np.empty(almost_max, np.int64)
    for a in range(almost_max):
        w = math.exp(-(a * mult) / hh)
        v = int(round(fixed_point_mult * w))                    # cvRound: half to even, like Python's round
        wt[a] = 0 if v < 0.001 * fixed_point_mult else v
    est = np.zeros((H, W), np.int64)
    wsum = np.zeros((H, W), np.int64)

    def at(dy, dx):
        return ext[b + dy: b + dy + H, b + dx: b + dx + W]

    own = [[at(ty, tx) for tx in range(-th, th + 1)] for ty in range(-th, th + 1)]
    for y in range(-sh, sh + 1):
        for x in range(-sh, sh + 1):
            d = np.zeros((H, W), np.int64)
            for ty in range(-th, th + 1):
                for tx in range(-th, th + 1):
                    diff = at(y + ty, x + tx) - own[ty + th][tx + th]
                    d += diff * diff
            wgt = wt[d >> shift]
            est += wgt * at(y, x)
            wsum += wgt
    return np.clip((est + wsum // 2) // wsum, 0, 255).astype(np.uint8)


# ---- the steps of preprocess_image between the denoise and the bilateral filter (src/utils/img_utils.py:159-181) ---------------
# Restated and pinned against cv2 as preparation for moving them to the device; no CUDA counterpart yet.
# The bilateral filter that follows them (img_utils.py:183-189) is NOT restated: with OpenCV 4.13.0 from the opencv-python wheel its
# interior comes from Intel IPP (truncating) and its left / right border columns from OpenCV's own code (rounding), so its output is a
# property of the build, not of an algorithm one can cite.

def clahe_apply(src, clip_limit=5.0, tile_grid_size=(10, 10)):
    """cv.createCLAHE(clipLimit, tileGridSize).apply(src) for a uint8 image (modules/imgproc/src/clahe.cpp): per-tile histogram,
    clip + redistribution (batch + residual with stride), LUT = round(cdf * 255 / tile area) in float32, bilinear blend of the four
    neighbouring tile LUTs in float32, round half to even."""
    src = np.asarray(src, np.uint8)
    H, W = src.shape
    tx, ty = int(tile_grid_size[0]), int(tile_grid_size[1])
    if W % tx == 0 and H % ty == 0:
        ext = src
    else:                                                       # BORDER_REFLECT_101 padding on the right / bottom
        ext = np.pad(src, ((0, ty - (H % ty)), (0, tx - (W % tx))), mode='reflect')
    tw, th = ext.shape[1] // tx, ext.shape[0] // ty
    area = tw * th
    lut_scale = np.float32(255) / np.float32(area)
    clip = max(int(clip_limit * area / 256), 1) if clip_limit > 0.0 else 0
    luts = np.zeros((ty, tx, 256), np.uint8)
    for j in range(ty):
        for i in range(tx):
            hist = np.bincount(ext[j * th:(j + 1) * th, i * tw:(i + 1) * tw].ravel(), minlength=256).astype(np.int64)
            if clip > 0:
                clipped = int(np.maximum(hist - clip, 0).sum())
                hist = np.minimum(hist, clip)
                batch = clipped // 256
                residual = clipped - batch * 256
                hist += batch
                if residual != 0:
                    step = max(256 // residual, 1)
                    k = 0
                    while k < 256 and residual > 0:
                        hist[k] += 1
                        k += step
                        residual -= 1
            luts[j, i] = np.clip(np.rint(np.cumsum(hist).astype(np.float32) * lut_scale), 0, 255).astype(np.uint8)
    inv_tw, inv_th = np.float32(1.0) / np.float32(tw), np.float32(1.0) / np.float32(th)
    txf = np.arange(W).astype(np.float32) * inv_tw - np.float32(0.5)
    tx1 = np.floor(txf).astype(np.int64)
    xa = (txf - tx1.astype(np.float32)).astype(np.float32)
    xa1 = (np.float32(1.0) - xa).astype(np.float32)
    tx2 = np.minimum(tx1 + 1, tx - 1)
    tx1 = np.maximum(tx1, 0)
    out = np.zeros((H, W), np.uint8)
    for y in range(H):
        tyf = np.float32(y) * inv_th - np.float32(0.5)
        ty1 = int(np.floor(tyf))
        ya = np.float32(tyf - np.float32(ty1))
        ya1 = np.float32(1.0) - ya
        ty2 = min(ty1 + 1, ty - 1)
        ty1 = max(ty1, 0)
        v = src[y]
        res = ((luts[ty1, tx1, v].astype(np.float32) * xa1 + luts[ty1, tx2, v].astype(np.float32) * xa) * ya1
               + (luts[ty2, tx1, v].astype(np.float32) * xa1 + luts[ty2, tx2, v].astype(np.float32) * xa) * ya)
        out[y] = np.clip(np.rint(res.astype(np.float32)), 0, 255).astype(np.uint8)
    return out


def gaussian_kernel_u8_fixed_point(sigma):
    """The 8-bit fixed-point kernel cv.GaussianBlur uses for uint8 images (getGaussianKernelFixedPoint_ED): size round(sigma * 6 + 1) | 1,
    taps rounded to 1/256 with error diffusion from the ends inwards, centre tap = 256 - the rest."""
    k = int(round(sigma * 3 * 2 + 1)) | 1
    x = np.arange(k, dtype=np.float64) - (k - 1) * 0.5
    w = np.exp(-0.5 / (sigma * sigma) * x * x)
    w /= w.sum()
    taps = np.zeros(k, np.int64)
    err, total = 0.0, 0
    for i in range(k // 2):
        adj = w[i] * 256.0 + err
        v = int(round(adj))
        err = adj - v
        taps[i] = taps[k - 1 - i] = v
        total += v
    taps[k // 2] = 256 - 2 * total
    return taps


def gaussian_blur_u8(src, sigma):
    """cv.GaussianBlur(uint8 image, None, sigma, ...) - preprocess_image's sharpening blur (img_utils.py:165-169; as OpenCV reads that
    call, sigmaX = sharpen_kernel_size = 3, 19 taps): separable integer filter with the 8-bit kernel, BORDER_REFLECT_101,
    (sum + 2^15) >> 16."""
    taps = gaussian_kernel_u8_fixed_point(float(sigma))
    h = len(taps) // 2
    a = np.asarray(src).astype(np.int64)
    p = np.pad(a, ((0, 0), (h, h)), mode='reflect')
    row = sum(taps[i] * p[:, i:i + a.shape[1]] for i in range(len(taps)))
    p = np.pad(row, ((h, h), (0, 0)), mode='reflect')
    col = sum(taps[i] * p[i:i + a.shape[0], :] for i in range(len(taps)))
    return np.clip((col + (1 << 15)) >> 16, 0, 255).astype(np.uint8)


def add_weighted_u8(a, alpha, b, beta, gamma=0.0):
    """cv.addWeighted for uint8 images (img_utils.py:171-178): float32 arithmetic, round half to even, saturate."""
    r = np.asarray(a).astype(np.float32) * np.float32(alpha) + np.asarray(b).astype(np.float32) * np.float32(beta) + np.float32(gamma)
    return np.clip(np.rint(r), 0, 255).astype(np.uint8)
