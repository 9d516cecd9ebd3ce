"""Independent re-derivation of the reference objective with torch autograd (CPU, float64).

Mirrors the *forward* of the reference op by op (same citations as oracle/eincm_oracle.py) and lets
torch's reverse mode produce d loss / d theta, the way ``jax.value_and_grad`` does inside jaxopt.  Used
only by tests to pin the oracle's hand-written backward.  torch's ``amin``/``amax`` split the cotangent
evenly among ties and ``abs`` differentiates to ``sign`` - the same conventions as JAX.
"""
import math
import sys

import numpy as np
import torch
import torch.nn.functional as F

from oracle import eincm_oracle as O

EPSN = sys.float_info.epsilon


def _shift(img, di, dj):
    H, W = img.shape
    p = F.pad(img, (1, 1, 1, 1))
    return p[1 + di:1 + di + H, 1 + dj:1 + dj + W]


def _scharr(I):
    # same canonical order as oracle.sobel_scharr_optimized_image_grads (img_utils.py:414-425)
    gx = (3.0 * (_shift(I, 1, 1) - _shift(I, 1, -1)) + 10.0 * (_shift(I, 0, 1) - _shift(I, 0, -1))) \
        + 3.0 * (_shift(I, -1, 1) - _shift(I, -1, -1))
    gy = (3.0 * (_shift(I, 1, 1) - _shift(I, -1, 1)) + 10.0 * (_shift(I, 1, 0) - _shift(I, -1, 0))) \
        + 3.0 * (_shift(I, 1, -1) - _shift(I, -1, -1))
    return gx, gy


def _divk(a):
    corners = ((_shift(a, 1, 1) + _shift(a, 1, -1)) + _shift(a, -1, 1)) + _shift(a, -1, -1)
    edges_ = ((_shift(a, 1, 0) + _shift(a, 0, 1)) + _shift(a, 0, -1)) + _shift(a, -1, 0)
    return corners * (1.0 / 12.0) + edges_ * (1.0 / 6.0)


def _splat(xw, yw, H, W, wrap_negative=True):
    xr = torch.round(xw.detach()).to(torch.int64)   # torch.round is half-to-even
    yr = torch.round(yw.detach()).to(torch.int64)
    frame = torch.zeros(H * W, dtype=torch.float64)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            cs = xr + dx
            rs = yr + dy
            qx = cs - xw
            qy = rs - yw
            v = torch.exp(-0.5 * (qx * qx + qy * qy) - math.log(2 * math.pi))
            if wrap_negative:
                rs = torch.where(rs < 0, rs + H, rs)
                cs = torch.where(cs < 0, cs + W, cs)
            ok = (rs >= 0) & (rs < H) & (cs >= 0) & (cs < W)
            frame = frame.index_add(0, (rs * W + cs)[ok], v[ok])
    return frame.reshape(H, W)


def _normalize(a):
    return (a - a.amin()) / (a.amax() - a.amin() + EPSN)


def _contrast(a):
    gx, gy = _scharr(a)
    return (gx ** 2 + gy ** 2).mean()


def _iwe_div(a):
    gx, gy = _scharr(a)
    return (_divk(gx) + _divk(gy)).abs().mean()


def loss_torch(theta, xs, ys, ts, edges, edge_ts, alpha, beta, gamma, delta, cur_pyr_lvl, sensor_size,
               wrap_negative=True):
    H, W = sensor_size
    Wy, Wx = O.resize_weights(tuple(theta.shape[:2]), (H, W))
    Wy = torch.as_tensor(Wy); Wx = torch.as_tensor(Wx)
    theta_full = torch.einsum('ijc,iy,jx->yxc', theta, Wy, Wx)
    xi = torch.as_tensor(np.asarray(xs).astype(np.int64))
    yi = torch.as_tensor(np.asarray(ys).astype(np.int64))
    tt = torch.as_tensor(np.asarray(ts, dtype=np.float64))
    E = torch.as_tensor(np.asarray(edges, dtype=np.float64))
    R = len(edge_ts)
    w = torch.as_tensor(O.compute_weights_for_multi_reference(R))
    zero_iwe = _splat(xi.double(), yi.double(), H, W, wrap_negative)
    nz = _normalize(zero_iwe)
    zc = _contrast(zero_iwe)
    zd = _iwe_div(nz)
    rel_c, rel_m, rel_d = [], [], []
    for r in range(R):
        dts = tt - float(edge_ts[r])
        xw = xi - theta_full[yi, xi, 0] * dts * 1.0
        yw = yi - theta_full[yi, xi, 1] * dts * 1.0
        iwe = _splat(xw, yw, H, W, wrap_negative)
        n = _normalize(iwe)
        corr = -((E[r] - n) ** 2).mean()
        zcorr = -((E[r] - nz) ** 2).mean()
        rel_m.append(w[r] * corr / (zcorr + EPSN))
        rel_c.append(w[r] * _contrast(iwe) / (zc + EPSN))
        rel_d.append(w[r] * _iwe_div(n) / (zd + EPSN))
    mean_rel_corr = torch.stack(rel_m).mean()
    mean_rel_contrast = torch.stack(rel_c).mean()
    mean_rel_div = torch.stack(rel_d).mean()
    tv = torch.zeros((), dtype=torch.float64)
    if cur_pyr_lvl <= 0:
        mask = torch.zeros(H, W, dtype=torch.float64)
        mask[yi, xi] = 1.0
        flow = theta_full * mask[..., None]
        terms = []
        for c in range(2):
            terms.extend(_scharr(flow[..., c]))
        nzm = torch.zeros(H, W, dtype=torch.bool)
        for t_ in terms:
            nzm |= t_.detach().abs() > 0
        tv = sum(t_.abs() * 0.25 for t_ in terms).sum() / (nzm.sum() + EPSN)
    return alpha * (-mean_rel_contrast) + beta * (-mean_rel_corr) + gamma * tv + delta * mean_rel_div


def value_and_grad_torch(theta, xs, ys, ts, edges, edge_ts, alpha, beta, gamma, delta, cur_pyr_lvl,
                         sensor_size, wrap_negative=True):
    th = torch.tensor(np.asarray(theta, dtype=np.float64), requires_grad=True)
    loss = loss_torch(th, xs, ys, ts, edges, edge_ts, alpha, beta, gamma, delta, cur_pyr_lvl, sensor_size,
                      wrap_negative)
    loss.backward()
    return float(loss.detach()), th.grad.numpy().copy()
