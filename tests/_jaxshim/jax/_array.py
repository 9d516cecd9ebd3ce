"""Array wrapper of the JAX stand-in (tests/_jaxshim/jax/__init__.py): NumPy-style methods over a torch tensor."""
import numpy as np
import torch

float64, float32, int64, int32, int16, bool_ = torch.float64, torch.float32, torch.int64, torch.int32, torch.int16, torch.bool


def _dtype(d):
    if d is bool:
        return torch.bool
    if d is float:
        return torch.float64
    if d is int:
        return torch.int64
    if isinstance(d, torch.dtype):
        return d
    return {np.dtype('float64'): float64, np.dtype('float32'): float32, np.dtype('int64'): int64, np.dtype('int32'): int32,
            np.dtype('int16'): int16, np.dtype('bool'): bool_}[np.dtype(d)]


def _unwrap(x):
    """anything -> torch tensor (float64 for Python / NumPy floats: jax_enable_x64 is on in the reference's configs)"""
    if isinstance(x, Array):
        return x.t
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (list, tuple)) and any(isinstance(e, (Array, torch.Tensor)) for e in x):
        return torch.stack([_unwrap(e) for e in x])
    a = np.asarray(x)
    if a.dtype == np.float32 and not isinstance(x, np.ndarray):
        a = a.astype(np.float64)
    return torch.from_numpy(np.ascontiguousarray(a)) if a.ndim else torch.tensor(a.item(), dtype=_dtype(a.dtype))


def _wrap(t):
    return t if isinstance(t, Array) else Array(t)


def _gather_index(i, n):
    """x[idx]: negative indices wrap, out-of-range indices are clamped"""
    i = _unwrap(i).long()
    i = torch.where(i < 0, i + n, i)
    return i.clamp(0, n - 1)


class _AtIndex:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx if isinstance(idx, tuple) else (idx,)

    def _scatter(self, vals, mode, accumulate):
        assert mode == 'drop', 'only mode="drop" is used by the reference path'
        x = self.arr.t
        idx = [_unwrap(i).long() for i in self.idx]
        assert len(idx) == x.dim()
        idx = list(torch.broadcast_tensors(*idx))
        vals = _unwrap(vals).to(x.dtype) if not isinstance(vals, (int, float)) else torch.tensor(vals, dtype=x.dtype)
        vals = vals.expand(idx[0].shape)
        ok = torch.ones(idx[0].shape, dtype=torch.bool)
        for d, i in enumerate(idx):
            n = x.shape[d]
            i = torch.where(i < 0, i + n, i)        # NumPy rule for negative indices, BEFORE the bounds test
            ok &= (i >= 0) & (i < n)
            idx[d] = i
        if not bool(ok.all()):
            idx = [i[ok] for i in idx]
            vals = vals[ok]
        if not accumulate and idx[0].numel():
            # .set with duplicate targets: ONE update wins and it alone receives the cotangent (JAX's scatter transpose masks the
            # losers; torch.index_put would hand the cotangent to every duplicate).  The last update in index order is taken.
            lin = torch.zeros_like(idx[0])
            for d, i in enumerate(idx):
                lin = lin * x.shape[d] + i
            uniq, inv = torch.unique(lin, return_inverse=True)
            win = torch.zeros(uniq.numel(), dtype=torch.long).scatter_reduce(0, inv, torch.arange(lin.numel()), 'amax', include_self=False)
            idx, vals = [i[win] for i in idx], vals[win]
        return _wrap(torch.index_put(x, tuple(idx), vals, accumulate=accumulate))

    def add(self, vals, mode=None):
        return self._scatter(vals, mode, True)

    def set(self, vals, mode=None):
        return self._scatter(vals, mode, False)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIndex(self.arr, idx)


class Array:
    __array_ufunc__ = None          # numpy_array * Array -> Array.__rmul__
    __array_priority__ = 1000

    def __init__(self, t):
        self.t = _unwrap(t)

    # -- NumPy-style attributes ---------------------------------------------------------------------------------------------
    shape = property(lambda s: tuple(s.t.shape))
    ndim = property(lambda s: s.t.dim())
    dtype = property(lambda s: s.t.dtype)
    size = property(lambda s: s.t.numel())
    T = property(lambda s: _wrap(s.t.permute(*reversed(range(s.t.dim())))))
    at = property(lambda s: _At(s))

    def __len__(self):
        return self.t.shape[0]

    def __iter__(self):
        return (_wrap(self.t[i]) for i in range(self.t.shape[0]))

    def __array__(self, dtype=None, copy=None):
        a = self.t.detach().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a

    def __float__(self):
        return float(self.t.detach())

    def __int__(self):
        return int(self.t)

    def __bool__(self):
        return bool(self.t)

    def __format__(self, spec):
        return format(self.t.detach().item(), spec)

    def __repr__(self):
        return f'ShimArray({self.t!r})'

    def astype(self, d):
        return _wrap(self.t.to(_dtype(d)))

    def transpose(self, *axes):
        axes = axes[0] if len(axes) == 1 and isinstance(axes[0], (tuple, list)) else axes
        return _wrap(self.t.permute(*axes)) if axes else self.T

    def reshape(self, *shape):
        shape = shape[0] if len(shape) == 1 and isinstance(shape[0], (tuple, list)) else shape
        return _wrap(self.t.reshape(*shape))

    def _reduce(self, fn, axis, **kw):
        return _wrap(fn(self.t, **kw) if axis is None else fn(self.t, dim=axis, **kw))

    def sum(self, axis=None):
        return self._reduce(torch.sum, axis)

    def mean(self, axis=None):
        return self._reduce(torch.mean, axis)

    def min(self, axis=None):
        return self._reduce(torch.amin, axis)       # even split of the cotangent over ties, as in JAX

    def max(self, axis=None):
        return self._reduce(torch.amax, axis)

    def var(self, axis=None):
        return self._reduce(torch.var, axis, correction=0)

    def __getitem__(self, idx):
        tup = idx if isinstance(idx, tuple) else (idx,)
        if len(tup) == 1 and isinstance(tup[0], (Array, torch.Tensor)) and _unwrap(tup[0]).dtype == torch.bool:
            return _wrap(self.t[_unwrap(tup[0])])          # boolean mask over the leading axes
        if any(isinstance(i, (Array, torch.Tensor, np.ndarray, list)) for i in tup):
            dims = [d for d, i in enumerate(tup)]
            assert Ellipsis not in tup and len(tup) == self.t.dim(), 'advanced index must name every axis in this stand-in'
            conv = []
            for d, i in zip(dims, tup):
                n = self.t.shape[d]
                conv.append(_gather_index(i if not isinstance(i, int) else torch.tensor(i), n))
            return _wrap(self.t[tuple(conv)])
        return _wrap(self.t[idx])

    # -- arithmetic ----------------------------------------------------------------------------------------------------------
    def _bin(self, other, fn, rev=False):
        o = _unwrap(other)
        if o.is_floating_point() and o.dtype != torch.float64 and not isinstance(other, (Array, torch.Tensor, np.ndarray)):
            o = o.double()
        a, b = (o, self.t) if rev else (self.t, o)
        return _wrap(fn(a, b))

    __add__ = lambda s, o: s._bin(o, torch.add)
    __radd__ = lambda s, o: s._bin(o, torch.add, True)
    __sub__ = lambda s, o: s._bin(o, torch.sub)
    __rsub__ = lambda s, o: s._bin(o, torch.sub, True)
    __mul__ = lambda s, o: s._bin(o, torch.mul)
    __rmul__ = lambda s, o: s._bin(o, torch.mul, True)
    __truediv__ = lambda s, o: s._bin(o, torch.true_divide)
    __rtruediv__ = lambda s, o: s._bin(o, torch.true_divide, True)
    __pow__ = lambda s, o: s._bin(o, torch.pow)
    __gt__ = lambda s, o: s._bin(o, torch.gt)
    __ge__ = lambda s, o: s._bin(o, torch.ge)
    __lt__ = lambda s, o: s._bin(o, torch.lt)
    __le__ = lambda s, o: s._bin(o, torch.le)
    __or__ = lambda s, o: s._bin(o, torch.logical_or)
    __and__ = lambda s, o: s._bin(o, torch.logical_and)
    __invert__ = lambda s: _wrap(~s.t)
    __neg__ = lambda s: _wrap(-s.t)
    __abs__ = lambda s: _wrap(s.t.abs())
