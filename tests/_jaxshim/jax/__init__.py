"""TEST INFRASTRUCTURE ONLY - a float64 torch-backed stand-in for the handful of JAX entry points that the reference's objective
path calls, so that the reference's OWN source files (``/root/reference/src/eincm/losses.py`` and what it imports) can be executed
unmodified in an image without jax / jaxlib, forward and - through torch autograd - reverse mode.

It is NOT JAX: each primitive below is written from JAX's published behaviour (docstrings of jax.numpy / jax.image, "sharp bits"),
one short function per primitive, independently of ``oracle/``:

  * ``x.at[idx].add(v, mode='drop')`` / ``.set``: negative indices count from the end (NumPy rule) FIRST, what is still outside is
    dropped; duplicates accumulate.
  * ``x[idx]`` (gather): negative indices wrap, the rest is clamped to the array.
  * ``jnp.round``: half to even; integer arrays pass through.  ``astype(int)`` of a float truncates.
  * ``min`` / ``max`` cotangents are split evenly over ties (torch.amin / amax do the same).
  * ``jax.scipy.signal.convolve(a, k, mode='same')``: true convolution (flipped kernel), zero padding, centred crop.
  * ``jax.image.scale_and_translate``: separable weight matrices of ``jax._src.image.scale._compute_weight_mat`` (triangle / Keys cubic
    kernel, antialias=True, columns renormalised, samples outside [-0.5, n - 0.5] zeroed).
  * ``jax.vmap``: a Python loop over the mapped axis.  ``jax.jit``: identity.  ``jax.value_and_grad``: torch autograd.

What running the reference over this stand-in pins is the reference's own COMPOSITION (which primitive is applied to what, in which
order, with which constants, signs, weights and normalisations); the primitives stay restated.  Only tests/ and
tests/golden/make_golden_refsrc.py import it.
"""
import functools

import numpy as _np
import torch as _torch

from . import _array
from ._array import Array, _unwrap, _wrap

__all__ = ['Array', 'jit', 'vmap', 'value_and_grad', 'grad']


def jit(fn=None, **_kw):
    if fn is None:
        return lambda f: f
    return fn


def vmap(fn, in_axes=0, out_axes=0):
    assert out_axes == 0

    @functools.wraps(fn)
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        assert len(axes) == len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                m = _np.shape(a)[ax] if not isinstance(a, Array) else a.shape[ax]
                assert n is None or n == m
                n = m
        outs = []
        for i in range(n):
            call = []
            for a, ax in zip(args, axes):
                if ax is None:
                    call.append(a)
                else:
                    t = _unwrap(a)
                    call.append(_wrap(t.select(ax, i)))
            outs.append(fn(*call))
        if isinstance(outs[0], tuple):
            return tuple(_wrap(_torch.stack([_unwrap(o[k]) for o in outs])) for k in range(len(outs[0])))
        return _wrap(_torch.stack([_unwrap(o) for o in outs]))
    return mapped


def value_and_grad(fn, argnums=0, has_aux=False):
    def vg(*args, **kw):
        args = list(args)
        x = _unwrap(args[argnums]).detach().clone().to(_torch.float64).requires_grad_(True)
        args[argnums] = _wrap(x)
        out = fn(*args, **kw)
        val, aux = out if has_aux else (out, None)
        v = _unwrap(val)
        g, = _torch.autograd.grad(v, x)
        val = _wrap(v.detach())
        return ((val, aux), _wrap(g)) if has_aux else (val, _wrap(g))
    return vg


def grad(fn, argnums=0, has_aux=False):
    vg = value_and_grad(fn, argnums, has_aux)

    def g(*a, **k):
        v, gr = vg(*a, **k)
        return (gr, v[1]) if has_aux else gr
    return g


from . import numpy, image, scipy, typing  # noqa: E402,F401
