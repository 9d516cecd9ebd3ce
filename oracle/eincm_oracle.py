"""CPU oracle for the EINCM contrast-correlation objective (float64, NumPy).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker (or the timed
CPU baseline), never as a fallback for the CUDA path.

PARITY: the reference (robotic-vision-lab/Edge-Informed-Contrast-Maximization)
is pure Python on JAX/jaxopt and ships no tests, golden vectors or fixtures for
this path; JAX/jaxlib/jaxopt are not installable in this image (no network).
Two layers:

* pinned against the reference's OWN SOURCE: its unmodified ``eincm.losses``,
  ``evaluations.theta_eval`` and ``eincm.solver`` are imported from
  /root/reference and executed over a float64 torch stand-in for the JAX entry
  points they call (tests/_jaxshim); the outputs are committed under
  tests/golden/refsrc/ (tests/golden/make_golden_refsrc.py) and this file is
  held to them - loss 1e-11, gradient 1e-9, every intermediate
  (tests/test_reference_source.py, tests/test_reference_solver_source.py);
* PARITY UNPINNED for the JAX primitives themselves (list below): the stand-in
  and this file both restate them from published JAX behaviour; only vectors
  from a real JAX can settle those.

Older pins, kept (tests/test_oracle_*.py): known answers derived by hand from
the reference source, scipy/torch cross checks of every JAX primitive restated
here, an independent torch-autograd re-derivation of the gradient, and central
finite differences.

Every function cites the reference file:line (relative to the reference repo
root) that it restates.  JAX semantics that are not visible in the reference
source are restated from the published JAX behaviour:

* ``jnp.round``                    -> round-half-to-even (``np.rint``)
* ``x.at[r, c].add(v, mode='drop')`` -> negative indices in [-N, -1] wrap
  (NumPy-style normalisation happens first), anything still outside [0, N)
  is dropped (whole update dropped if either axis is out of range)
* ``jax.image.scale_and_translate(method='bilinear')`` -> separable triangle
  kernel, half-pixel centres, column-normalised weights (``compute_weight_mat``)
* ``jax.scipy.signal.convolve(mode='same')`` -> ``scipy.signal.convolve2d``
  (kernel flipped, zero padding)
* ``jax.scipy.stats.multivariate_normal.pdf`` with identity covariance
  -> ``exp(-0.5*|q|^2 - log(2*pi))``
* reverse-mode of ``min``/``max`` -> cotangent split evenly among ties
* reverse-mode of ``abs`` -> ``sign`` (0 at 0); ``round`` -> zero gradient
"""
from __future__ import annotations

import math
import sys
from typing import Dict, Optional, Tuple

import numpy as np
from scipy.signal import convolve2d, correlate2d

EPSN = sys.float_info.epsilon  # src/eincm/losses.py:24, src/utils/img_utils.py:18

SCHARR_GX = np.array([[3.0, 0.0, -3.0], [10.0, 0.0, -10.0], [3.0, 0.0, -3.0]])    # img_utils.py:417
SCHARR_GY = np.array([[3.0, 10.0, 3.0], [0.0, 0.0, 0.0], [-3.0, -10.0, -3.0]])    # img_utils.py:418
DIV_KERN = np.array([[1 / 12, 1 / 6, 1 / 12], [1 / 6, 0.0, 1 / 6], [1 / 12, 1 / 6, 1 / 12]])  # event_collapse_objectives.py:14
LOG_2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------- #
# theta resize: src/utils/theta_utils.py:10-37  (jax.image.scale_and_translate)
# --------------------------------------------------------------------------- #
def compute_weight_mat(input_size: int, output_size: int, scale: float,
                       translation: float = 0.0, antialias: bool = True) -> np.ndarray:
    """Per-axis resize weights ``W[i_in, j_out]`` of jax.image.scale_and_translate with
    the 'linear' (triangle) kernel.  Restates jax._src.image.scale.compute_weight_mat
    (SURVEY.md A.1); called by the reference at src/utils/theta_utils.py:25-35."""
    inv_scale = 1.0 / scale
    kernel_scale = max(inv_scale, 1.0) if antialias else 1.0
    sample_f = (np.arange(output_size, dtype=np.float64) + 0.5) * inv_scale - translation * inv_scale - 0.5
    x = np.abs(sample_f[np.newaxis, :] - np.arange(input_size, dtype=np.float64)[:, np.newaxis]) / kernel_scale
    weights = np.maximum(0.0, 1.0 - np.abs(x))
    total = weights.sum(axis=0, keepdims=True)
    weights = np.where(np.abs(total) > 1000.0 * float(np.finfo(np.float32).eps),
                       weights / np.where(total != 0, total, 1.0), 0.0)
    inside = np.logical_and(sample_f >= -0.5, sample_f <= input_size - 0.5)
    return np.where(inside[np.newaxis, :], weights, 0.0)


def resize_weights(theta_shape: Tuple[int, int], sensor_size: Tuple[int, int]) -> Tuple[np.ndarray, np.ndarray]:
    """(Wy[h,H], Wx[w,W]) used by scale_theta_to_sensor_size (theta_utils.py:28-31: scale = H/h, W/w)."""
    h, w = theta_shape
    H, W = sensor_size
    return compute_weight_mat(h, H, H / h), compute_weight_mat(w, W, W / w)


def scale_theta_to_sensor_size(theta: np.ndarray, sensor_size: Tuple[int, int], method: str = 'bilinear') -> np.ndarray:
    """src/utils/theta_utils.py:10-37.  theta (h,w,2) -> (H,W,2); channel axis has scale 1 (identity)."""
    if method not in ('bilinear', 'linear'):
        raise NotImplementedError(f'oracle restates only the shipped method (bilinear), got {method!r}')
    theta = np.asarray(theta, dtype=np.float64)
    Wy, Wx = resize_weights(theta.shape[:2], sensor_size)
    return np.einsum('ijc,iy,jx->yxc', theta, Wy, Wx, optimize=True)


# --------------------------------------------------------------------------- #
# warp: src/eincm/event_warpers.py:28-35
# --------------------------------------------------------------------------- #
def per_pix_warp(per_pix_theta, xs, ys, ts, t_ref, delta_time=1.0):
    xi = np.rint(xs).astype(np.int16)                       # event_warpers.py:28
    yi = np.rint(ys).astype(np.int16)                       # event_warpers.py:29
    dts = np.asarray(ts, dtype=np.float64) - t_ref          # event_warpers.py:30
    yy = yi.astype(np.int64)
    xx = xi.astype(np.int64)
    warped_xs = xi - per_pix_theta[yy, xx, 0] * dts * delta_time   # event_warpers.py:34
    warped_ys = yi - per_pix_theta[yy, xx, 1] * dts * delta_time   # event_warpers.py:35
    return warped_xs, warped_ys


# --------------------------------------------------------------------------- #
# splat: src/utils/event_utils.py:13-61
# --------------------------------------------------------------------------- #
def _drop_index(rs: np.ndarray, cs: np.ndarray, H: int, W: int, wrap_negative: bool):
    """Index rule of ``frame.at[rs, cs].add(v, mode='drop')``.
    Returns (flat_index, valid_mask)."""
    if wrap_negative:
        rs = np.where(rs < 0, rs + H, rs)
        cs = np.where(cs < 0, cs + W, cs)
    valid = (rs >= 0) & (rs < H) & (cs >= 0) & (cs < W)
    return rs * W + cs, valid


def rounded_event_pixels(xs, ys):
    """Xs_rounded of event_utils.py:32-33 — the bit-exact event->pixel index stream."""
    Xw = np.asarray(xs, dtype=np.float64)
    Yw = np.asarray(ys, dtype=np.float64)
    with np.errstate(invalid='ignore'):
        xr = np.rint(Xw)
        yr = np.rint(Yw)
    # astype(int32) of out-of-range doubles is implementation defined; such votes are dropped
    # by every implementation, so saturate well outside any sensor instead.
    big = 2.0 ** 30
    xr = np.where(np.isfinite(xr), np.clip(xr, -big, big), big)
    yr = np.where(np.isfinite(yr), np.clip(yr, -big, big), big)
    return xr.astype(np.int64), yr.astype(np.int64)


def events_to_pdf_frame(xs, ys, sensor_size=(260, 346), window_size=3, wrap_negative=True) -> np.ndarray:
    H, W = sensor_size
    Xw = np.asarray(xs, dtype=np.float64)                   # event_utils.py:32
    Yw = np.asarray(ys, dtype=np.float64)
    xr, yr = rounded_event_pixels(Xw, Yw)                   # event_utils.py:33
    frame = np.zeros(H * W, dtype=np.float64)               # event_utils.py:36
    w = window_size // 2
    for dx in range(-w, w + 1):                             # event_utils.py:42
        for dy in range(-w, w + 1):                         # event_utils.py:43
            cs = xr + dx                                    # event_utils.py:46
            rs = yr + dy
            qx = cs - Xw                                    # event_utils.py:55
            qy = rs - Yw
            pdf_val = np.exp(-0.5 * (qx * qx + qy * qy) - LOG_2PI)   # event_utils.py:56
            flat, valid = _drop_index(rs, cs, H, W, wrap_negative)   # event_utils.py:59
            frame += np.bincount(flat[valid], weights=pdf_val[valid], minlength=H * W)
    return frame.reshape(H, W)


def make_event_mask(xs, ys, sensor_size) -> np.ndarray:
    """src/utils/event_utils.py:64-77 (and the mask implied by theta_utils.py:70-71)."""
    H, W = sensor_size
    m = np.zeros((H, W), dtype=bool)
    m[np.asarray(ys).astype(np.int64), np.asarray(xs).astype(np.int64)] = True
    return m


# --------------------------------------------------------------------------- #
# image ops: src/utils/img_utils.py:24-25, 414-425
# --------------------------------------------------------------------------- #
def normalize_to_unit_range(arr: np.ndarray) -> np.ndarray:
    return (arr - arr.min()) / (arr.max() - arr.min() + EPSN)        # img_utils.py:25


def _shift(img: np.ndarray, di: int, dj: int) -> np.ndarray:
    """out[i, j] = img[i + di, j + dj], zero outside the image (the 'same' zero padding)."""
    H, W = img.shape
    p = np.pad(img, 1)
    return p[1 + di:1 + di + H, 1 + dj:1 + dj + W]


def sobel_scharr_literal(image: np.ndarray) -> np.ndarray:
    """Literal img_utils.py:414-425 through scipy (``jax.scipy.signal.convolve(mode='same')`` ==
    ``scipy.signal.convolve2d(mode='same')``).  Kept as the pin for the canonical form below."""
    I_x = convolve2d(image, SCHARR_GX, mode='same')                  # img_utils.py:420
    I_y = convolve2d(image, SCHARR_GY, mode='same')                  # img_utils.py:421
    return np.stack([I_x, I_y], axis=-1)                             # img_utils.py:423


def sobel_scharr_optimized_image_grads(image: np.ndarray) -> np.ndarray:
    """img_utils.py:414-425 in a *canonical summation order* (difference first, no FMA):
        Gx = (3*(I[i+1,j+1]-I[i+1,j-1]) + 10*(I[i,j+1]-I[i,j-1])) + 3*(I[i-1,j+1]-I[i-1,j-1])
        Gy = (3*(I[i+1,j+1]-I[i-1,j+1]) + 10*(I[i+1,j]-I[i-1,j])) + 3*(I[i+1,j-1]-I[i-1,j-1])
    Mathematically identical to the literal convolution (tests pin it to 1e-13); the order is fixed
    because the reference's TV regulariser counts pixels whose gradient is *exactly* non-zero
    (regularizers.py:26-29) and the sign of exact zeros feeds d|.|, so XLA's (unknowable) summation
    order leaks into the result.  Difference-first yields exact zeros on locally constant flow, the
    mathematically correct answer; the CUDA kernels use the same order with non-contracted mul/add."""
    I = np.asarray(image, dtype=np.float64)
    gx = (3.0 * (_shift(I, 1, 1) - _shift(I, 1, -1)) + 10.0 * (_shift(I, 0, 1) - _shift(I, 0, -1))) \
        + 3.0 * (_shift(I, -1, 1) - _shift(I, -1, -1))
    gy = (3.0 * (_shift(I, 1, 1) - _shift(I, -1, 1)) + 10.0 * (_shift(I, 1, 0) - _shift(I, -1, 0))) \
        + 3.0 * (_shift(I, 1, -1) - _shift(I, -1, -1))
    return np.stack([gx, gy], axis=-1)


def div_kern_conv(a: np.ndarray) -> np.ndarray:
    """convolve(a, DIV_KERN, 'same') (event_collapse_objectives.py:15-16) in canonical order:
    ((corner sum)/12 + (edge sum)/6) with corners/edges summed row-major."""
    corners = ((_shift(a, 1, 1) + _shift(a, 1, -1)) + _shift(a, -1, 1)) + _shift(a, -1, -1)
    edges_ = ((_shift(a, 1, 0) + _shift(a, 0, 1)) + _shift(a, 0, -1)) + _shift(a, -1, 0)
    return corners * (1.0 / 12.0) + edges_ * (1.0 / 6.0)


def compute_mean_gradient_magnitude(arr: np.ndarray) -> float:
    g = sobel_scharr_optimized_image_grads(arr.astype(np.float64))   # contrast_objectives.py:22
    return float((g[..., 0] ** 2 + g[..., 1] ** 2).mean())           # contrast_objectives.py:23-25


def compute_mean_squared_error(arr_1: np.ndarray, arr_2: np.ndarray) -> float:
    return float(((arr_1 - arr_2) ** 2).mean())                      # correlation_objectives.py:25-26


def _iwe_div_field(iwe: np.ndarray) -> np.ndarray:
    g = sobel_scharr_optimized_image_grads(iwe)                      # event_collapse_objectives.py:10
    dx = div_kern_conv(g[..., 0])                                    # :15
    dy = div_kern_conv(g[..., 1])                                    # :16
    return dx + dy


def iwe_divergence(iwe: np.ndarray) -> float:
    return float(np.abs(_iwe_div_field(iwe)).mean())                 # event_collapse_objectives.py:18-20


def compute_fwl(iwe: np.ndarray, zero_iwe: np.ndarray) -> float:
    return float(np.var(iwe) / np.var(zero_iwe))                     # contrast_metrics.py:16


# --------------------------------------------------------------------------- #
# regulariser: src/eincm/regularizers.py:14-38, src/utils/theta_utils.py:40-73
# --------------------------------------------------------------------------- #
def per_pix_theta_to_flow(theta_full, xs, ys, ts) -> np.ndarray:
    m = make_event_mask(xs, ys, theta_full.shape[:2])                # theta_utils.py:66-71 (dt == 1)
    return theta_full * m[..., None]


def _tv_fields(flow):
    gx = sobel_scharr_optimized_image_grads(flow[..., 0])            # regularizers.py:20
    gy = sobel_scharr_optimized_image_grads(flow[..., 1])            # regularizers.py:21
    return gx[..., 0], gx[..., 1], gy[..., 0], gy[..., 1]


def per_pix_total_variation(theta_full, xs, ys, ts) -> float:
    flow = per_pix_theta_to_flow(theta_full, xs, ys, ts)             # regularizers.py:16
    a, b, c, d = _tv_fields(flow)
    nz = (np.abs(a) > 0) | (np.abs(b) > 0) | (np.abs(c) > 0) | (np.abs(d) > 0)   # regularizers.py:26-29
    tot = np.sum((np.abs(a) * 0.25 + np.abs(b) * 0.25) + (np.abs(c) * 0.25 + np.abs(d) * 0.25))
    return float(tot / (nz.sum() + EPSN))                            # regularizers.py:31-36


def per_pix_theta_divergence(theta_full) -> float:
    """src/eincm/regularizers.py:41-58 (reported in aux / eval only, never in the loss)."""
    a, b, c, d = _tv_fields(theta_full)
    s = ((div_kern_conv(a) + div_kern_conv(b)) + div_kern_conv(c)) + div_kern_conv(d)
    return float(np.abs(s).mean())


# --------------------------------------------------------------------------- #
# loss assembly: src/eincm/losses.py:39-205
# --------------------------------------------------------------------------- #
def compute_weights_for_multi_reference(n_refs: int, n_sigma: float = 1.5) -> np.ndarray:
    x = np.linspace(-n_sigma, n_sigma, n_refs)                       # losses.py:44
    w = np.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)              # stats.norm.pdf(x, 0, 1)
    return w / w.sum()                                               # losses.py:46


def compute_loss_objectives(theta_full, xs, ys, ts, edges, edge_ts, sensor_size,
                            wrap_negative=True) -> Dict[str, np.ndarray]:
    """src/eincm/losses.py:49-105 (theta_full is the (H,W,2) field)."""
    sensor_size = tuple(sensor_size)
    edges = np.asarray(edges, dtype=np.float64)
    edge_ts = np.asarray(edge_ts, dtype=np.float64)
    R = len(edge_ts)
    zero_iwe = events_to_pdf_frame(xs, ys, sensor_size, wrap_negative=wrap_negative)      # :54
    normalized_zero_iwe = normalize_to_unit_range(zero_iwe)                               # :55
    warped = [per_pix_warp(theta_full, xs, ys, ts, edge_ts[r], 1.0) for r in range(R)]    # :58
    warped_xs = np.stack([w_[0] for w_ in warped])
    warped_ys = np.stack([w_[1] for w_ in warped])
    iwes = np.stack([events_to_pdf_frame(warped_xs[r], warped_ys[r], sensor_size,
                                         wrap_negative=wrap_negative) for r in range(R)])  # :61
    normalized_iwes = np.stack([normalize_to_unit_range(i) for i in iwes])                # :62
    corrs = np.array([compute_mean_squared_error(edges[r], normalized_iwes[r]) for r in range(R)]) * (-1)       # :65
    zero_corrs = np.array([compute_mean_squared_error(edges[r], normalized_zero_iwe) for r in range(R)]) * (-1)  # :66
    rel_corrs = corrs / (zero_corrs + EPSN)                                               # :67
    contrasts = np.array([compute_mean_gradient_magnitude(i) for i in iwes])              # :70
    zero_contrast = compute_mean_gradient_magnitude(zero_iwe)                             # :71
    rel_contrasts = contrasts / (zero_contrast + EPSN)                                    # :72
    theta_total_variation = per_pix_total_variation(theta_full, xs, ys, ts)               # :75
    theta_divergence = per_pix_theta_divergence(theta_full)                               # :76
    iwe_divergences = np.array([iwe_divergence(n) for n in normalized_iwes])              # :79
    zero_iwe_divergence = iwe_divergence(normalized_zero_iwe)                             # :80
    rel_iwe_divergences = iwe_divergences / (zero_iwe_divergence + EPSN)                  # :81
    flow_warp_losses = np.array([compute_fwl(i, zero_iwe) for i in iwes])                 # :84
    multi_ref_weights = compute_weights_for_multi_reference(R)                            # :87
    return {
        'warped_xs': warped_xs, 'warped_ys': warped_ys,
        'correlations': corrs, 'zero_correlations': zero_corrs, 'rel_correlations': rel_corrs,
        'contrasts': contrasts, 'zero_contrast': zero_contrast, 'rel_contrasts': rel_contrasts,
        'theta_total_variation': theta_total_variation, 'theta_divergence': theta_divergence,
        'iwe_divergences': iwe_divergences, 'zero_iwe_divergence': zero_iwe_divergence,
        'rel_iwe_divergences': rel_iwe_divergences, 'flow_warp_losses': flow_warp_losses,
        'multi_ref_weights': multi_ref_weights,
        # extras for the parity tests (not in the reference dict)
        '_iwes': iwes, '_zero_iwe': zero_iwe, '_normalized_iwes': normalized_iwes,
        '_normalized_zero_iwe': normalized_zero_iwe,
    }


def loss_func(theta, xs, ys, ts, edges, edge_ts, alpha, beta, gamma, delta,
              cur_pyr_lvl, n_pyr_lvls, sensor_size, scale_to_sensor_size_method='bilinear',
              wrap_negative=True, _keep=False):
    """src/eincm/losses.py:108-205.  Returns (final_loss, aux_info)."""
    sensor_size = tuple(sensor_size)
    scaled_theta = scale_theta_to_sensor_size(theta, sensor_size, scale_to_sensor_size_method)   # :158-160
    lo = compute_loss_objectives(scaled_theta, xs, ys, ts, edges, edge_ts, sensor_size,
                                 wrap_negative=wrap_negative)                                   # :162-165
    corrs, zero_corrs = lo['correlations'], lo['zero_correlations']
    contrasts, zero_contrast = lo['contrasts'], lo['zero_contrast']
    theta_total_variation = lo['theta_total_variation'] if cur_pyr_lvl <= 0 else 0.0             # :171
    iwe_divergences, zero_iwe_divergence = lo['iwe_divergences'], lo['zero_iwe_divergence']
    w = lo['multi_ref_weights']
    rel_corrs = (w * corrs) / (zero_corrs + EPSN)                                                # :176
    rel_contrasts = (w * contrasts) / (zero_contrast + EPSN)                                     # :177
    rel_iwe_divergences = (w * iwe_divergences) / (zero_iwe_divergence + EPSN)                   # :178
    mean_rel_corr = rel_corrs.mean()
    mean_rel_contrast = rel_contrasts.mean()
    mean_rel_iwe_divergence = rel_iwe_divergences.mean()
    contrast_loss = mean_rel_contrast * (-1)                                                     # :187
    correlation_loss = mean_rel_corr * (-1)                                                      # :188
    contrast_correlation_loss = (alpha * contrast_loss + beta * correlation_loss) ** 1           # :190
    regularization_loss = gamma * theta_total_variation + delta * mean_rel_iwe_divergence        # :191
    final_loss = contrast_correlation_loss + regularization_loss                                 # :193
    aux = {
        'final_loss': final_loss, 'scaled_theta': scaled_theta, 'mean_rel_corr': mean_rel_corr,
        'mean_rel_contrast': mean_rel_contrast, 'mean_rel_iwe_divergence': mean_rel_iwe_divergence,
        'theta_total_variation': theta_total_variation, 'multi_ref_weights': w,
    }
    if _keep:
        aux['_objectives'] = lo
    return float(final_loss), aux


# --------------------------------------------------------------------------- #
# evaluation metrics: src/evaluations/flow_eval.py:14-75, src/evaluations/theta_eval.py:14-95
# --------------------------------------------------------------------------- #
N_PE_THRESHOLDS = (1, 2, 3, 5, 10, 20)                                    # flow_eval.py:72


def sparse_flow_error(pred_flow, gt_flow, event_mask=None) -> Dict:
    """src/evaluations/flow_eval.py:14-75: end-point errors over the pixels where both flows are valid."""
    pred_flow = np.asarray(pred_flow, dtype=np.float64)
    gt_flow = np.asarray(gt_flow, dtype=np.float64)
    mask_pred = (~np.isinf(pred_flow[..., 0]) & ~np.isinf(pred_flow[..., 1])) & (np.linalg.norm(pred_flow, axis=-1) > 0)   # :32-36
    if event_mask is not None:
        mask_pred = mask_pred & np.asarray(event_mask, dtype=bool)                                                    # :38-39
    mask_gt = (~np.isinf(gt_flow[..., 0]) & ~np.isinf(gt_flow[..., 1])) & (np.linalg.norm(gt_flow, axis=-1) > 0)       # :42-46
    inter = mask_pred & mask_gt                                                                                        # :49
    pm, gm = pred_flow[inter], gt_flow[inter]                                                                          # :52-53
    ee = np.linalg.norm(pm - gm, axis=-1)                                                                              # :56
    ree = ee / (np.linalg.norm(gm, axis=-1) + EPSN)                                                                    # :59
    cnts = {'n_ee': int(ee.shape[0]), 'n_pred': int(mask_pred.sum()), 'n_gt': int(mask_gt.sum())}                      # :65-67
    with np.errstate(invalid='ignore'), __import__('warnings').catch_warnings():
        __import__('warnings').simplefilter('ignore')
        errs = {'AEE': float(ee.mean()) if ee.size else float('nan'),                                                  # :70 (mean of empty = nan)
                'AREE': float(ree.mean()) if ree.size else float('nan')}                                               # :71
    for n in N_PE_THRESHOLDS:
        errs[f'A{n}PE'] = float((ee > n).sum() * 100 / (cnts['n_ee'] + EPSN))                                          # :73-74
    return {'errors': errs, 'counts': cnts}


def evaluate_theta_array(theta_array, xs, ys, ts, edges, edge_ts, gt_flow, alpha, beta, gamma, delta, sensor_size,
                         err_eval_event_mask=None, wrap_negative=True) -> Dict:
    """src/evaluations/theta_eval.py:14-95 (the numbers; the reference also formats a log line).  theta_array is the
    sensor-size field (H, W, 2).  Note the un-weighted means (:27-29) and that TV / divergence enter regardless of the
    pyramid level (:37-42), unlike loss_func."""
    sensor_size = tuple(sensor_size)
    theta_array = np.asarray(theta_array, dtype=np.float64)
    lo = compute_loss_objectives(theta_array, xs, ys, ts, edges, edge_ts, sensor_size, wrap_negative=wrap_negative)   # :21-24
    mean_rel_contrast = float(lo['rel_contrasts'].mean())                                                             # :27
    mean_rel_corr = float(lo['rel_correlations'].mean())                                                              # :28
    mean_rel_iwe_div = float(lo['rel_iwe_divergences'].mean())                                                        # :29
    tot_var = float(lo['theta_total_variation'])                                                                      # :30
    theta_div = float(lo['theta_divergence'])                                                                         # :31
    fwl = float(lo['flow_warp_losses'][0])                                                                            # :32
    l_iwe = events_to_pdf_frame(lo['warped_xs'][0], lo['warped_ys'][0], sensor_size, wrap_negative=wrap_negative)     # :36
    loss = alpha * (-mean_rel_contrast) + beta * (-mean_rel_corr) + gamma * tot_var + delta * mean_rel_iwe_div        # :37-42
    evals = {}
    if gt_flow is not None:
        pred_flow = per_pix_theta_to_flow(theta_array, xs, ys, ts)                                                    # :47
        fe = sparse_flow_error(pred_flow, gt_flow, err_eval_event_mask)                                               # :48
        evals.update(fe['errors']); evals.update(fe['counts'])                                                        # :61-62
        evals['n_pixels'] = sensor_size[0] * sensor_size[1]                                                           # :59,63
    evals.update({'loss': float(loss), 'iwe_var': float(np.var(l_iwe)), 'mean_rel_contrast': mean_rel_contrast,      # :80-94
                  'mean_rel_corr': mean_rel_corr, 'theta_tot_var': tot_var, 'theta_div': theta_div, 'fwl': fwl,
                  'mean_rel_iwe_div': mean_rel_iwe_div, 'rel_iwe_divergences': lo['rel_iwe_divergences'],
                  'rel_contrasts': lo['rel_contrasts'], 'rel_correlations': lo['rel_correlations'],
                  'flow_warp_losses': lo['flow_warp_losses'], 'multi_ref_weights': lo['multi_ref_weights']})
    return evals


def handover_loss_func(alpha_handover, prev_theta, theta, xs, ys, ts, edges, edge_ts,
                       alpha, beta, gamma, delta, cur_pyr_lvl, n_pyr_lvls, sensor_size,
                       scale_to_sensor_size_method='bilinear', wrap_negative=True) -> float:
    """src/eincm/losses.py:208-276."""
    theta_ho = alpha_handover * np.asarray(prev_theta) + (1 - alpha_handover) * np.asarray(theta)   # :269
    loss, _ = loss_func(theta_ho, xs, ys, ts, edges, edge_ts, alpha, beta, gamma, delta,
                        cur_pyr_lvl, n_pyr_lvls, sensor_size, scale_to_sensor_size_method,
                        wrap_negative=wrap_negative)
    return loss


# --------------------------------------------------------------------------- #
# analytic reverse mode (what jax.value_and_grad(loss_func) yields inside jaxopt;
# SURVEY.md §3.3).  Independently re-derived with torch autograd in
# tests/test_oracle_autograd.py.
# --------------------------------------------------------------------------- #
def _scharr_adjoint(gx_bar: np.ndarray, gy_bar: np.ndarray) -> np.ndarray:
    """Adjoint of I -> (Gx, Gy) above: 'same' *correlation* with the same kernels
    (== correlate2d(gx_bar, Kx, 'same') + correlate2d(gy_bar, Ky, 'same'); pinned in tests)."""
    ax = (3.0 * (_shift(gx_bar, -1, -1) - _shift(gx_bar, -1, 1)) + 10.0 * (_shift(gx_bar, 0, -1) - _shift(gx_bar, 0, 1))) \
        + 3.0 * (_shift(gx_bar, 1, -1) - _shift(gx_bar, 1, 1))
    ay = (3.0 * (_shift(gy_bar, -1, -1) - _shift(gy_bar, 1, -1)) + 10.0 * (_shift(gy_bar, -1, 0) - _shift(gy_bar, 1, 0))) \
        + 3.0 * (_shift(gy_bar, -1, 1) - _shift(gy_bar, 1, 1))
    return ax + ay


def _div_kern_adjoint(a: np.ndarray) -> np.ndarray:
    """DIV_KERN is symmetric under 180-degree rotation, so the adjoint is the same stencil."""
    return div_kern_conv(a)


def _minmax_normalize_backward(I: np.ndarray, gN: np.ndarray) -> np.ndarray:
    """Cotangent of I through N = (I - min)/(max - min + eps)   (img_utils.py:25)."""
    m, M = I.min(), I.max()
    D = M - m + EPSN
    s1 = gN.sum()
    s2 = (gN * (I - m)).sum()
    g_M = -s2 / (D * D)
    g_m = -s1 / D + s2 / (D * D)
    is_min = (I == m)
    is_max = (I == M)
    return gN / D + g_m * is_min / is_min.sum() + g_M * is_max / is_max.sum()


def splat_backward(xw, yw, dLdI: np.ndarray, wrap_negative=True):
    """dL/dx', dL/dy' of events_to_pdf_frame (rint has zero gradient)."""
    H, W = dLdI.shape
    xr, yr = rounded_event_pixels(xw, yw)
    flatI = dLdI.reshape(-1)
    gx = np.zeros_like(xw, dtype=np.float64)
    gy = np.zeros_like(yw, dtype=np.float64)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            cs = xr + dx
            rs = yr + dy
            qx = cs - xw
            qy = rs - yw
            v = np.exp(-0.5 * (qx * qx + qy * qy) - LOG_2PI)
            flat, valid = _drop_index(rs, cs, H, W, wrap_negative)
            g = np.where(valid, flatI[np.where(valid, flat, 0)], 0.0) * v
            # d v / d x' = v * (c - x')
            gx += g * qx
            gy += g * qy
    return gx, gy


def value_and_grad(theta, xs, ys, ts, edges, edge_ts, alpha, beta, gamma, delta,
                   cur_pyr_lvl, n_pyr_lvls, sensor_size, scale_to_sensor_size_method='bilinear',
                   wrap_negative=True, return_intermediates=False):
    """(loss, dloss/dtheta) — the pair jaxopt's ``jit(value_and_grad(fun))`` hands scipy
    (src/eincm/solver.py:165-173).  Gradient has the shape of ``theta``."""
    theta = np.asarray(theta, dtype=np.float64)
    sensor_size = tuple(sensor_size)
    H, W = sensor_size
    HW = float(H * W)
    edges = np.asarray(edges, dtype=np.float64)
    edge_ts = np.asarray(edge_ts, dtype=np.float64)
    ts = np.asarray(ts, dtype=np.float64)
    R = len(edge_ts)
    loss, aux = loss_func(theta, xs, ys, ts, edges, edge_ts, alpha, beta, gamma, delta,
                          cur_pyr_lvl, n_pyr_lvls, sensor_size, scale_to_sensor_size_method,
                          wrap_negative=wrap_negative, _keep=True)
    lo = aux['_objectives']
    theta_full = aux['scaled_theta']
    w = lo['multi_ref_weights']
    C0 = lo['zero_contrast']
    M0 = -lo['zero_correlations']              # MSE(edge_r, N0)
    D0 = lo['zero_iwe_divergence']
    iwes, nrm = lo['_iwes'], lo['_normalized_iwes']

    xi = np.rint(xs).astype(np.int64)
    yi = np.rint(ys).astype(np.int64)
    G_full = np.zeros((H, W, 2), dtype=np.float64)
    dLdI_all = np.zeros((R, H, W), dtype=np.float64)
    for r in range(R):
        # loss = sum_r [ a_r * C_r + b_r * M_r + d_r * Div_r ]
        a_r = -alpha * w[r] / ((C0 + EPSN) * R)
        b_r = beta * w[r] / ((-M0[r] + EPSN) * R)
        d_r = delta * w[r] / ((D0 + EPSN) * R)
        I = iwes[r]
        g = sobel_scharr_optimized_image_grads(I)
        dI = a_r * (2.0 / HW) * _scharr_adjoint(g[..., 0], g[..., 1])
        gN = b_r * (-2.0 / HW) * (edges[r] - nrm[r])
        if delta != 0.0:
            S = _iwe_div_field(nrm[r])
            sbar = d_r * np.sign(S) / HW
            kbar = _div_kern_adjoint(sbar)
            gN = gN + _scharr_adjoint(kbar, kbar)
        dI = dI + _minmax_normalize_backward(I, gN)
        dLdI_all[r] = dI
        gx, gy = splat_backward(lo['warped_xs'][r], lo['warped_ys'][r], dI, wrap_negative)
        dts = ts - edge_ts[r]
        np.add.at(G_full[..., 0], (yi, xi), -dts * gx)
        np.add.at(G_full[..., 1], (yi, xi), -dts * gy)
    if gamma != 0.0 and cur_pyr_lvl <= 0:
        mask = make_event_mask(xs, ys, sensor_size)
        flow = theta_full * mask[..., None]
        a, b, c, d = _tv_fields(flow)
        nz = (np.abs(a) > 0) | (np.abs(b) > 0) | (np.abs(c) > 0) | (np.abs(d) > 0)
        k = gamma * 0.25 / (nz.sum() + EPSN)
        G_full[..., 0] += k * _scharr_adjoint(np.sign(a), np.sign(b)) * mask
        G_full[..., 1] += k * _scharr_adjoint(np.sign(c), np.sign(d)) * mask
    Wy, Wx = resize_weights(theta.shape[:2], sensor_size)
    grad = np.einsum('yxc,iy,jx->ijc', G_full, Wy, Wx, optimize=True)
    if return_intermediates:
        return loss, grad, {'iwes': iwes, 'dLdI': dLdI_all, 'G_full': G_full, 'aux': aux,
                            'zero_iwe': lo['_zero_iwe'], 'objectives': lo}
    return loss, grad


def handover_value_and_grad(alpha_handover, prev_theta, theta, xs, ys, ts, edges, edge_ts,
                            alpha, beta, gamma, delta, cur_pyr_lvl, n_pyr_lvls, sensor_size,
                            scale_to_sensor_size_method='bilinear', wrap_negative=True):
    """(loss, dloss/dalpha_handover): what ScipyBoundedMinimize differentiates
    (src/eincm/solver.py:175-183, losses.py:269)."""
    prev_theta = np.asarray(prev_theta, dtype=np.float64)
    theta = np.asarray(theta, dtype=np.float64)
    theta_ho = alpha_handover * prev_theta + (1 - alpha_handover) * theta
    loss, g = value_and_grad(theta_ho, xs, ys, ts, edges, edge_ts, alpha, beta, gamma, delta,
                             cur_pyr_lvl, n_pyr_lvls, sensor_size, scale_to_sensor_size_method,
                             wrap_negative=wrap_negative)
    return loss, float((g * (prev_theta - theta)).sum())


# --------------------------------------------------------------------------- #
# event-split restatement: the same arithmetic as value_and_grad, phrased over
# COMPLETE images (sums over all ranks) and LOCAL events.  Used by the gloo tests of
# eincm_b200.parallel.EventSplitObjective; tests pin it to value_and_grad.
# --------------------------------------------------------------------------- #
def partial_images(theta, xs, ys, ts, edge_ts, sensor_size, wrap_negative=True):
    """Local contribution to the R images of warped events (additive over event subsets)."""
    sensor_size = tuple(sensor_size)
    theta_full = scale_theta_to_sensor_size(np.asarray(theta, dtype=np.float64), sensor_size)
    out = []
    for t_ref in np.asarray(edge_ts, dtype=np.float64):
        xw, yw = per_pix_warp(theta_full, xs, ys, ts, t_ref, 1.0)
        out.append(events_to_pdf_frame(xw, yw, sensor_size, wrap_negative=wrap_negative))
    return np.stack(out)


def split_value_and_grad(theta, iwes, zero_iwe, mask, xs, ys, ts, edges, edge_ts, alpha, beta, gamma, delta,
                         cur_pyr_lvl, n_pyr_lvls, sensor_size, include_replicated_grad=True, wrap_negative=True):
    """(loss, partial grad): loss from the complete images ``iwes`` / ``zero_iwe`` / ``mask``; gradient contribution of the
    LOCAL events ``xs, ys, ts`` (plus, if ``include_replicated_grad``, the TV term that is identical on every rank)."""
    theta = np.asarray(theta, dtype=np.float64)
    sensor_size = tuple(sensor_size)
    H, W = sensor_size
    HW = float(H * W)
    edges = np.asarray(edges, dtype=np.float64)
    edge_ts = np.asarray(edge_ts, dtype=np.float64)
    ts = np.asarray(ts, dtype=np.float64)
    R = len(edge_ts)
    theta_full = scale_theta_to_sensor_size(theta, sensor_size)
    w = compute_weights_for_multi_reference(R)
    nz0 = normalize_to_unit_range(zero_iwe)
    C0 = compute_mean_gradient_magnitude(zero_iwe)
    M0 = np.array([compute_mean_squared_error(edges[r], nz0) for r in range(R)])
    D0 = iwe_divergence(nz0)
    nrm = np.stack([normalize_to_unit_range(i) for i in iwes])
    contrasts = np.array([compute_mean_gradient_magnitude(i) for i in iwes])
    corrs = -np.array([compute_mean_squared_error(edges[r], nrm[r]) for r in range(R)])
    divs = np.array([iwe_divergence(n) for n in nrm])
    use_tv = cur_pyr_lvl <= 0
    tv = 0.0
    if use_tv:
        flow = theta_full * mask[..., None]
        a, b, c, d = _tv_fields(flow)
        nzp = (np.abs(a) > 0) | (np.abs(b) > 0) | (np.abs(c) > 0) | (np.abs(d) > 0)
        tv = float(np.sum((np.abs(a) * 0.25 + np.abs(b) * 0.25) + (np.abs(c) * 0.25 + np.abs(d) * 0.25)) / (nzp.sum() + EPSN))
    mean_rel_corr = ((w * corrs) / (-M0 + EPSN)).mean()
    mean_rel_contrast = ((w * contrasts) / (C0 + EPSN)).mean()
    mean_rel_div = ((w * divs) / (D0 + EPSN)).mean()
    loss = (alpha * (-mean_rel_contrast) + beta * (-mean_rel_corr)) + (gamma * tv + delta * mean_rel_div)

    xi = np.rint(xs).astype(np.int64)
    yi = np.rint(ys).astype(np.int64)
    G_full = np.zeros((H, W, 2), dtype=np.float64)
    for r in range(R):
        a_r = -alpha * w[r] / ((C0 + EPSN) * R)
        b_r = beta * w[r] / ((-M0[r] + EPSN) * R)
        d_r = delta * w[r] / ((D0 + EPSN) * R)
        I = iwes[r]
        g = sobel_scharr_optimized_image_grads(I)
        dI = a_r * (2.0 / HW) * _scharr_adjoint(g[..., 0], g[..., 1])
        gN = b_r * (-2.0 / HW) * (edges[r] - nrm[r])
        if delta != 0.0:
            sbar = d_r * np.sign(_iwe_div_field(nrm[r])) / HW
            kbar = _div_kern_adjoint(sbar)
            gN = gN + _scharr_adjoint(kbar, kbar)
        dI = dI + _minmax_normalize_backward(I, gN)
        xw, yw = per_pix_warp(theta_full, xs, ys, ts, edge_ts[r], 1.0)
        gx, gy = splat_backward(xw, yw, dI, wrap_negative)
        dts = ts - edge_ts[r]
        np.add.at(G_full[..., 0], (yi, xi), -dts * gx)
        np.add.at(G_full[..., 1], (yi, xi), -dts * gy)
    if gamma != 0.0 and use_tv and include_replicated_grad:
        k = gamma * 0.25 / (nzp.sum() + EPSN)
        G_full[..., 0] += k * _scharr_adjoint(np.sign(a), np.sign(b)) * mask
        G_full[..., 1] += k * _scharr_adjoint(np.sign(c), np.sign(d)) * mask
    Wy, Wx = resize_weights(theta.shape[:2], sensor_size)
    return float(loss), np.einsum('yxc,iy,jx->ijc', G_full, Wy, Wx, optimize=True)
