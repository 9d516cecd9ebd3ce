"""JAX side of the XLA FFI binding (integration/xla_ffi/eincm_xla_ffi.cc): a drop-in for ``eincm.losses.loss_func`` of the reference
(src/eincm/losses.py:108-205) whose value AND gradient come from one launch sequence of the CUDA library.

NOT RUN IN THIS REPOSITORY'S IMAGE (jax / jaxlib / jaxopt are not installed; no network).  Written against the public ``jax.ffi`` API
(jax >= 0.4.38: ``jax.ffi.register_ffi_target``, ``jax.ffi.ffi_call``, ``jax.ffi.pycapsule``).

Usage inside the reference (configs/theta_loss_func/default.yaml points ``_target_`` at ``loss_func``)::

    from eincm_jax import loss_func, new_window            # instead of: from eincm.losses import loss_func
    ...
    new_window()                                           # in MultipleLevelEINCMSolver.set_datasample (solver.py:185-194)

jaxopt's ``ScipyMinimize(fun=partial(loss_func, cur_pyr_lvl=l, ...), has_aux=True, jit=True)`` then builds
``jit(value_and_grad(fun, has_aux=True))`` as before; the ``custom_vjp`` below makes the forward pass return the loss and stash the
gradient, so the backward pass is a multiplication by the incoming cotangent.
"""
import ctypes
import functools
import itertools
import os

import jax
import jax.numpy as jnp
import numpy as np

_LIB = ctypes.CDLL(os.environ.get('EINCM_XLA_FFI_LIB', os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libeincm_xla_ffi.so')))
jax.ffi.register_ffi_target('eincm_value_and_grad', jax.ffi.pycapsule(_LIB.EincmValueAndGrad), platform='CUDA')

_window_counter = itertools.count(1)
_window_id = 0


def new_window():
    """Call whenever a new datasample is staged: the next evaluation packs and sorts the events again (once per window)."""
    global _window_id
    _window_id = next(_window_counter)


def _call(theta, xs, ys, ts, edges, edge_ts_host, alpha, beta, gamma, delta, cur_pyr_lvl, n_pyr_lvls):
    out_types = (jax.ShapeDtypeStruct((), jnp.float64), jax.ShapeDtypeStruct(theta.shape, jnp.float64))
    return jax.ffi.ffi_call('eincm_value_and_grad', out_types)(
        theta, xs, ys, ts, edges,
        edge_ts=np.asarray(edge_ts_host, dtype=np.float64), alpha=np.float64(alpha), beta=np.float64(beta), gamma=np.float64(gamma),
        delta=np.float64(delta), cur_pyr_lvl=np.int32(cur_pyr_lvl), n_pyr_lvls=np.int32(n_pyr_lvls), window_id=np.int64(_window_id))


def loss_func(theta, xs, ys, ts, edges, edge_ts, alpha, beta, gamma, delta, cur_pyr_lvl, n_pyr_lvls, sensor_size,
              scale_to_sensor_size_method='bilinear'):
    """Same signature as the reference's ``loss_func`` (losses.py:108-123); returns ``(final_loss, aux)``.  ``edge_ts`` must be a
    concrete (NumPy) array - it is, the loaders deliver it as one - because the reference times become kernel constants."""
    if scale_to_sensor_size_method != 'bilinear':
        raise NotImplementedError('the reference ships only bilinear (configs/main.yaml:27)')
    edge_ts_host = tuple(float(t) for t in np.asarray(edge_ts))
    static = (edge_ts_host, float(alpha), float(beta), float(gamma), float(delta), int(cur_pyr_lvl), int(n_pyr_lvls))

    @functools.partial(jax.custom_vjp)
    def objective(theta_):
        return _call(theta_, xs, ys, ts, edges, *static)[0]

    def fwd(theta_):
        loss, grad = _call(theta_, xs, ys, ts, edges, *static)
        return loss, grad

    def bwd(grad, g):
        return (g * grad,)

    objective.defvjp(fwd, bwd)
    final_loss = objective(theta)
    # the solver drops aux (has_aux=True only unpacks it, solver.py:165-183); the keys the reference fills (losses.py:195-203) that
    # cost nothing are provided, the rest are available through eincm_get_scalars
    aux = {'final_loss': final_loss}
    return final_loss, aux
