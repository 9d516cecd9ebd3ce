"""jax.image.scale_and_translate of the stand-in: the separable weight matrices of jax._src.image.scale (published algorithm)."""
import numpy as _np
import torch as _torch

from ._array import _unwrap, _wrap


def _triangle(x):
    return _torch.clamp(1.0 - x.abs(), min=0.0)


def _keys_cubic(x):
    x = x.abs()
    out = ((1.5 * x - 2.5) * x) * x + 1.0
    out = _torch.where(x >= 1.0, ((-0.5 * x + 2.5) * x - 4.0) * x + 2.0, out)
    return _torch.where(x >= 2.0, _torch.zeros_like(x), out)


def _lanczos(radius):
    def k(x):
        x = x.abs()
        y = radius * _torch.sin(_np.pi * x) * _torch.sin(_np.pi * x / radius)
        out = _torch.where(x > 1e-3, y / _torch.where(x != 0, _np.pi ** 2 * x ** 2, _torch.ones_like(x)), _torch.ones_like(x))
        return _torch.where(x > radius, _torch.zeros_like(x), out)
    return k


_KERNELS = {'lanczos3': _lanczos(3.0), 'lanczos5': _lanczos(5.0),
            'linear': _triangle, 'bilinear': _triangle, 'trilinear': _triangle, 'triangle': _triangle,
            'cubic': _keys_cubic, 'bicubic': _keys_cubic, 'tricubic': _keys_cubic}


def _weight_mat(n_in, n_out, scale, translation, kernel, antialias):
    inv = 1.0 / scale
    kscale = max(inv, 1.0) if antialias else 1.0
    sample = (_torch.arange(n_out, dtype=_torch.float64) + 0.5) * inv - translation * inv - 0.5
    x = (sample[None, :] - _torch.arange(n_in, dtype=_torch.float64)[:, None]).abs() / kscale
    w = kernel(x)
    tot = w.sum(dim=0, keepdim=True)
    eps = float(_np.finfo(_np.float64).eps)
    w = _torch.where(tot.abs() > 1000.0 * eps, w / _torch.where(tot != 0, tot, _torch.ones_like(tot)), _torch.zeros_like(w))
    inside = (sample >= -0.5) & (sample <= n_in - 0.5)
    return _torch.where(inside[None, :], w, _torch.zeros_like(w))          # (n_in, n_out)


def scale_and_translate(image, shape, spatial_dims, scale, translation, method, antialias=True, precision=None):
    t = _unwrap(image).to(_torch.float64)
    scale = [float(s) for s in _unwrap(scale)]
    translation = [float(s) for s in _unwrap(translation)]
    kernel = _KERNELS[method]
    for k, d in enumerate(spatial_dims):
        w = _weight_mat(t.shape[d], int(shape[d]), scale[k], translation[k], kernel, antialias)
        t = _torch.movedim(_torch.tensordot(t, w, dims=([d], [0])), -1, d)
    return _wrap(t)
