"""jax.numpy entry points used by the reference's objective path (see tests/_jaxshim/jax/__init__.py)."""
import numpy as _np
import torch as _torch

from ._array import Array, _dtype, _unwrap, _wrap, float64, float32, int64, int32, int16, bool_  # noqa: F401

ndarray = Array
pi = _np.pi


def array(x, dtype=None):
    t = _unwrap(x)
    return _wrap(t.to(_dtype(dtype)) if dtype is not None else t)


asarray = array


def zeros(shape, dtype=None):
    return _wrap(_torch.zeros(tuple(shape) if not isinstance(shape, int) else (shape,), dtype=_dtype(dtype) if dtype is not None else float64))


def ones(shape, dtype=None):
    return _wrap(_torch.ones(tuple(shape) if not isinstance(shape, int) else (shape,), dtype=_dtype(dtype) if dtype is not None else float64))


def zeros_like(x):
    return _wrap(_torch.zeros_like(_unwrap(x)))


def ones_like(x):
    return _wrap(_torch.ones_like(_unwrap(x)))


def round(x):  # noqa: A001  (half to even; integers pass through)
    t = _unwrap(x)
    return _wrap(_torch.round(t) if t.is_floating_point() else t)


def abs(x):  # noqa: A001
    return _wrap(_unwrap(x).abs())


def sum(x, axis=None):  # noqa: A001
    return _wrap(x).sum(axis)


def mean(x, axis=None):
    return _wrap(x).mean(axis)


def var(x, axis=None):
    return _wrap(x).var(axis)


def stack(xs, axis=0):
    return _wrap(_torch.stack([_unwrap(x) for x in xs], dim=axis))


def where(c, a, b):
    return _wrap(_torch.where(_unwrap(c), _unwrap(a), _unwrap(b)))


def sqrt(x):
    return _wrap(_unwrap(x).sqrt())


def exp(x):
    return _wrap(_unwrap(x).exp())


def logical_and(a, b):
    return _wrap(_torch.logical_and(_unwrap(a), _unwrap(b)))


def logical_or(a, b):
    return _wrap(_torch.logical_or(_unwrap(a), _unwrap(b)))


def isinf(x):
    return _wrap(_torch.isinf(_unwrap(x)))


class linalg:
    @staticmethod
    def norm(x, axis=None):
        t = _unwrap(x).to(float64)
        return _wrap(_torch.sqrt((t * t).sum() if axis is None else (t * t).sum(dim=axis)))


def repeat(x, repeats, axis=None):
    return _wrap(_torch.repeat_interleave(_unwrap(x), int(repeats), dim=axis))


def clip(x, lo=None, hi=None):
    return _wrap(_torch.clamp(_unwrap(x), min=lo, max=hi))
