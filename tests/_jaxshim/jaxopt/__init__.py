"""TEST INFRASTRUCTURE ONLY - stand-in for jaxopt's SciPy wrappers (jaxopt is absent from this image and not vendored by the reference),
so that the reference's own ``src/eincm/solver.py`` can be executed.  Written from jaxopt's published behaviour
(``jaxopt._src.scipy_wrappers.ScipyMinimize`` / ``ScipyBoundedMinimize``): parameters are raveled to one float64 vector, the objective
handed to ``scipy.optimize.minimize`` is ``value_and_grad(fun)`` (aux dropped) with ``jac=True``, ``options['maxiter'] = maxiter``, and
the result comes back as ``(params, ScipyMinimizeInfo)``.  The callback hook is the one the reference's README (lines 92-126) tells its
users to patch into jaxopt: SciPy's ``OptimizeResult`` is passed through as ``intermediate_result`` with ``.x`` converted."""
from collections import namedtuple

import numpy as _np
import scipy.optimize as _opt

import jax as _jax
import jax.numpy as _jnp

ScipyMinimizeInfo = namedtuple('ScipyMinimizeInfo', 'fun_val success status iter_num hess_inv num_fun_eval num_jac_eval num_hess_eval')
OptStep = namedtuple('OptStep', 'params state')


class ScipyMinimize:
    def __init__(self, fun, callback=None, tol=None, options=None, method=None, dtype=_np.float64, jit=True, implicit_diff_solve=None,
                 has_aux=False, maxiter=500, value_and_grad=False):
        self.fun, self.callback, self.tol, self.method, self.dtype, self.has_aux, self.maxiter = fun, callback, tol, method, dtype, has_aux, maxiter
        self.options = dict(options or {})
        self.options['maxiter'] = maxiter
        f = (lambda x, *a, **k: fun(x, *a, **k)[0]) if has_aux else fun
        self._value_and_grad_fun = fun if value_and_grad else _jax.value_and_grad(f)

    def _run(self, init_params, bounds, *args, **kwargs):
        shape = _np.shape(init_params)

        def to_jnp(x):
            return _jnp.array(_np.asarray(x, self.dtype).reshape(shape))

        def scipy_fun(x_onp):
            value, grads = self._value_and_grad_fun(to_jnp(x_onp), *args, **kwargs)
            return _np.asarray(value, self.dtype), _np.asarray(grads, self.dtype).ravel()

        def scipy_callback(intermediate_result):          # reference README.md:113-121
            intermediate_result.x = to_jnp(intermediate_result.x)
            return self.callback(intermediate_result)

        res = _opt.minimize(scipy_fun, _np.asarray(init_params, self.dtype).ravel(), jac=True, tol=self.tol, bounds=bounds, method=self.method,
                            callback=scipy_callback if self.callback is not None else None, options=self.options)
        info = ScipyMinimizeInfo(fun_val=_jnp.array(res.fun), success=res.success, status=res.status, iter_num=res.nit,
                                 hess_inv=getattr(res, 'hess_inv', None), num_fun_eval=_jnp.array(res.nfev), num_jac_eval=_jnp.array(res.njev),
                                 num_hess_eval=_jnp.array(getattr(res, 'nhev', 0)))
        return OptStep(to_jnp(res.x), info)

    def run(self, init_params, *args, **kwargs):
        return self._run(init_params, None, *args, **kwargs)


class ScipyBoundedMinimize(ScipyMinimize):
    def run(self, init_params, bounds, *args, **kwargs):
        lo, hi = bounds
        n = int(_np.size(init_params))
        b = _opt.Bounds(lb=_np.broadcast_to(_np.asarray(lo, self.dtype), _np.shape(init_params)).reshape(n),
                        ub=_np.broadcast_to(_np.asarray(hi, self.dtype), _np.shape(init_params)).reshape(n))
        return self._run(init_params, b, *args, **kwargs)
