"""JAX side of the XLA FFI binding (integration/xla_ffi/eincm_xla_ffi.cc): a drop-in for ``eincm.losses.loss_func`` of the reference
(src/eincm/losses.py:108-205) whose value AND gradient come from one launch sequence of the CUDA library.

NOT RUN IN THIS REPOSITORY'S IMAGE (jax / jaxlib / jaxopt are not installed; no network).  Written against the public ``jax.ffi`` API
(jax >= 0.4.38: ``jax.ffi.register_ffi_target``, ``jax.ffi.ffi_call``, ``jax.ffi.pycapsule``).  The C++ handler bodies behind it ARE
exercised on a GPU through a stand-in for the FFI header (tests/test_gpu_ffi_handler.py).

Two changes inside the reference (configs/theta_loss_func/default.yaml points ``_target_`` at ``loss_func``)::

    from eincm_jax import loss_func, set_datasample        # instead of: from eincm.losses import loss_func

    # MultipleLevelEINCMSolver.set_datasample (src/eincm/solver.py:185-194): stage the window ONCE, eagerly, and carry the token
    def set_datasample(self, xs, ys, ts, edges, edge_ts):
        self.datasample = {'events': {'x': xs, 'y': ys, 't': ts}, 'edges': edges,
                           'edge_ts': set_datasample(xs, ys, ts, edges, edge_ts)}          # (edge_ts, token): a pytree operand

jaxopt's ``ScipyMinimize(fun=partial(loss_func, cur_pyr_lvl=l, ...), has_aux=True, jit=True).run(theta, xs, ys, ts, edges, edge_ts)``
(solver.py:165-173, :209-216) then works unchanged: ``edge_ts`` arrives in ``loss_func`` as the traced pair, the token is a RUN-TIME
operand of the cached executable (nothing about the window is read at trace time), and the ``custom_vjp`` below makes the forward
pass return the loss and stash the gradient, so the backward pass is a multiplication by the incoming cotangent.
"""
import ctypes
import os

import jax
import jax.numpy as jnp
import numpy as np

_LIB = ctypes.CDLL(os.environ.get('EINCM_XLA_FFI_LIB', os.path.join(os.path.dirname(os.path.abspath(__file__)), 'libeincm_xla_ffi.so')))
jax.ffi.register_ffi_target('eincm_set_window', jax.ffi.pycapsule(_LIB.EincmSetWindow), platform='CUDA')
jax.ffi.register_ffi_target('eincm_value_and_grad', jax.ffi.pycapsule(_LIB.EincmValueAndGrad), platform='CUDA')


def set_datasample(xs, ys, ts, edges, edge_ts, slot=0):
    """Stages one window in the library (events packed and sorted, zero-warp statistics: everything that does not depend on theta) and
    returns ``(edge_ts, token)`` to store where the solver keeps ``edge_ts``.  Call it eagerly, once per window; ``slot`` separates
    solvers that share a device.  ``edge_ts`` stays a device array: the handler copies the R reference times to the host itself."""
    token = jax.ffi.ffi_call('eincm_set_window', jax.ShapeDtypeStruct((2,), jnp.int64), has_side_effect=True)(
        jnp.asarray(xs, jnp.int16), jnp.asarray(ys, jnp.int16), jnp.asarray(ts, jnp.float64), jnp.asarray(edges, jnp.float64),
        jnp.asarray(edge_ts, jnp.float64), slot=np.int32(slot))
    return edge_ts, token


def _call(theta, token, alpha, beta, gamma, delta, cur_pyr_lvl, n_pyr_lvls, slot):
    out_types = (jax.ShapeDtypeStruct((), jnp.float64), jax.ShapeDtypeStruct(theta.shape, jnp.float64))
    return jax.ffi.ffi_call('eincm_value_and_grad', out_types, has_side_effect=True)(
        theta, token, alpha=np.float64(alpha), beta=np.float64(beta), gamma=np.float64(gamma), delta=np.float64(delta),
        cur_pyr_lvl=np.int32(cur_pyr_lvl), n_pyr_lvls=np.int32(n_pyr_lvls), slot=np.int32(slot))


def loss_func(theta, xs, ys, ts, edges, edge_ts, alpha, beta, gamma, delta, cur_pyr_lvl, n_pyr_lvls, sensor_size,
              scale_to_sensor_size_method='bilinear', slot=0):
    """Same signature as the reference's ``loss_func`` (losses.py:108-123); returns ``(final_loss, aux)``.  ``edge_ts`` is the pair
    made by ``set_datasample``; xs / ys / ts / edges are not touched here (they were staged by ``set_datasample``), so XLA does not
    even pass them to the custom call.  The hyper-parameters are Python numbers bound by ``functools.partial`` in the reference
    (solver.py:165-173): static attributes of the call."""
    if scale_to_sensor_size_method != 'bilinear':
        raise NotImplementedError('the reference ships only bilinear (configs/main.yaml:27)')
    _, token = edge_ts
    static = (float(alpha), float(beta), float(gamma), float(delta), int(cur_pyr_lvl), int(n_pyr_lvls), int(slot))

    @jax.custom_vjp
    def objective(theta_, token_):
        return _call(theta_, token_, *static)[0]

    def fwd(theta_, token_):
        loss, grad = _call(theta_, token_, *static)
        return loss, (grad, token_)

    def bwd(res, g):
        grad, token_ = res
        return g * grad, np.zeros(token_.shape, dtype=jax.dtypes.float0)      # integer operand: no cotangent

    objective.defvjp(fwd, bwd)
    final_loss = objective(theta, token)
    # the solver drops aux (has_aux=True only unpacks it, solver.py:165-183); the keys the reference fills (losses.py:195-203) that
    # cost nothing are provided, the rest are available through eincm_get_scalars
    aux = {'final_loss': final_loss}
    return final_loss, aux
