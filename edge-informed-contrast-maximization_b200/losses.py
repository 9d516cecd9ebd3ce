"""Host-side mirror of the reference's ``eincm.losses`` interface, backed by the CUDA plan.

Same names, argument order and meaning as reference src/eincm/losses.py:
``loss_func`` (:108-205), ``handover_loss_func`` (:208-276), ``compute_weights_for_multi_reference`` (:39-46),
plus ``value_and_grad`` - the transformation jaxopt applies to them (``jit(value_and_grad(fun))``, built from
reference src/eincm/solver.py:165-183) - so that ``functools.partial(loss_func, alpha=..., ...)`` objects written for
the reference keep working.  The big operands (``xs, ys, ts, edges, edge_ts``) are staged on the device once per
window: consecutive calls that pass the *same array objects* (as the solver's BFGS loop does) reuse the staged window.
Operands are therefore treated as immutable, like the ``jnp`` arrays of the reference.

Nothing here computes on the CPU; every call needs the CUDA library and a B200 and raises otherwise.
"""
from __future__ import annotations

import functools
import math
from typing import Callable, Dict, Optional, Tuple

import numpy as np

from . import plan as _plan

EPSN = 2.220446049250313e-16     # sys.float_info.epsilon, reference src/eincm/losses.py:24


def compute_weights_for_multi_reference(n_refs: int, n_sigma: float = 1.5) -> np.ndarray:
    """reference src/eincm/losses.py:39-46 (host constant; the plan computes the same numbers on its side)."""
    x = np.linspace(-n_sigma, n_sigma, n_refs)
    w = np.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)
    return w / w.sum()


# ---------------------------------------------------------------------------------------------------------------
# staged-window cache
# ---------------------------------------------------------------------------------------------------------------
class _WindowCache:
    """One plan per (device, sensor size); the window staged in it is identified by the operand objects."""

    def __init__(self):
        self.plans: Dict[Tuple[int, int, int, int], _plan.Plan] = {}
        self.staged: Dict[Tuple[int, int, int, int], Tuple] = {}

    def plan_for(self, xs, ys, ts, edges, edge_ts, sensor_size, flags=None) -> _plan.Plan:
        import torch
        if not torch.cuda.is_available():
            raise _plan.EincmError(_plan.EINCM_ECUDA, 'no CUDA device: the EINCM objective runs only on a B200 (no CPU fallback)')
        dev = torch.cuda.current_device()
        flags = _default_flags if flags is None else flags
        H, W = int(sensor_size[0]), int(sensor_size[1])
        key = (dev, H, W, int(flags))
        n = int(np.shape(xs)[0]) if not hasattr(xs, 'numel') else int(xs.numel())
        R = int(np.shape(edge_ts)[0]) if not hasattr(edge_ts, 'numel') else int(edge_ts.numel())
        p = self.plans.get(key)
        if p is None or p.max_events < n or p.max_refs < R:
            if p is not None:
                p.close()
            p = _plan.Plan((H, W), max_events=max(n, 1 << 16), max_refs=max(R, 5), device=dev, flags=flags)
            self.plans[key] = p
            self.staged.pop(key, None)
        ops = (xs, ys, ts, edges, edge_ts)
        cur = self.staged.get(key)
        if cur is None or any(a is not b for a, b in zip(cur, ops)):
            p.set_window(xs, ys, ts, edges, edge_ts)
            self.staged[key] = ops          # strong references: the ids stay unique while the window is staged
        return p

    def clear(self):
        for p in self.plans.values():
            p.close()
        self.plans.clear()
        self.staged.clear()


_cache = _WindowCache()
_default_flags = 0


def configure(exact_f64: Optional[bool] = None, wrap_negative: Optional[bool] = None) -> int:
    """Plan flags used by the module-level callables: ``exact_f64`` selects the float64 nine-tap scatter
    (EINCM_FLAG_EXACT_F64) instead of the default fixed-point tile splat; ``wrap_negative=False`` drops votes with negative
    row / column instead of wrapping them as JAX does (EINCM_FLAG_NO_WRAP_NEGATIVE).  Returns the flag word."""
    global _default_flags
    if exact_f64 is not None:
        _default_flags = (_default_flags | _plan.FLAG_EXACT_F64) if exact_f64 else (_default_flags & ~_plan.FLAG_EXACT_F64)
    if wrap_negative is not None:
        _default_flags = (_default_flags & ~_plan.FLAG_NO_WRAP_NEGATIVE) if wrap_negative else (_default_flags | _plan.FLAG_NO_WRAP_NEGATIVE)
    return _default_flags


def clear_cache():
    """Releases every cached plan (device memory) - the analogue of reference src/utils/jax_helpers.py:15-18."""
    _cache.clear()


def _as_theta(theta) -> np.ndarray:
    if hasattr(theta, 'detach'):
        theta = theta.detach().cpu().numpy()
    theta = np.asarray(theta, dtype=np.float64)
    if theta.ndim != 3 or theta.shape[2] != 2:
        raise _plan.EincmError(_plan.EINCM_EINVAL, f'theta must have shape (h, w, 2), got {theta.shape}')
    return theta


def _aux_from_plan(p: _plan.Plan, loss: float) -> Dict[str, object]:
    """aux_info of reference src/eincm/losses.py:195-203 ('scaled_theta' stays on the device).  One deviation: a regulariser whose
    weight is zero is not evaluated on the device, so 'mean_rel_iwe_divergence' (delta == 0) and 'theta_total_variation' (gamma == 0)
    read 0.0 here where the reference reports the value it goes on to multiply by zero; the loss and its gradient are unaffected
    (tests/test_gpu_reference_source.py).  eincm_b200.evaluations.evaluate_theta_array reports both regardless of the weights."""
    s = p.scalars()
    return {'final_loss': loss, 'scaled_theta': p.theta_full(), 'mean_rel_corr': s['mean_rel_corr'],
            'mean_rel_contrast': s['mean_rel_contrast'], 'mean_rel_iwe_divergence': s['mean_rel_iwe_divergence'],
            'theta_total_variation': s['theta_total_variation'], 'multi_ref_weights': s['multi_ref_weights']}


# ---------------------------------------------------------------------------------------------------------------
# the reference's two loss callables
# ---------------------------------------------------------------------------------------------------------------
def loss_func(theta, xs, ys, ts, edges, edge_ts, alpha, beta, gamma, delta, cur_pyr_lvl, n_pyr_lvls, sensor_size,
              scale_to_sensor_size_method) -> Tuple[float, Dict]:
    """reference src/eincm/losses.py:108-205: ``(final_loss, aux_info)``."""
    hp = _plan.make_hparams(alpha, beta, gamma, delta, cur_pyr_lvl, n_pyr_lvls, scale_to_sensor_size_method)
    p = _cache.plan_for(xs, ys, ts, edges, edge_ts, sensor_size)
    loss, _ = p.value_and_grad_host(_as_theta(theta), hp, want_grad=False)
    return loss, _aux_from_plan(p, loss)


def handover_loss_func(alpha_handover, prev_theta, theta, xs, ys, ts, edges, edge_ts, alpha, beta, gamma, delta, cur_pyr_lvl,
                       n_pyr_lvls, sensor_size, scale_to_sensor_size_method) -> float:
    """reference src/eincm/losses.py:208-276: loss of ``alpha_handover*prev_theta + (1-alpha_handover)*theta``."""
    hp = _plan.make_hparams(alpha, beta, gamma, delta, cur_pyr_lvl, n_pyr_lvls, scale_to_sensor_size_method)
    p = _cache.plan_for(xs, ys, ts, edges, edge_ts, sensor_size)
    loss, _ = p.handover_value_and_grad_host(float(np.asarray(alpha_handover)), _as_theta(prev_theta), _as_theta(theta), hp,
                                             want_grad=False)
    return loss


def _loss_func_vg(theta, xs, ys, ts, edges, edge_ts, alpha, beta, gamma, delta, cur_pyr_lvl, n_pyr_lvls, sensor_size,
                  scale_to_sensor_size_method, has_aux=False):
    hp = _plan.make_hparams(alpha, beta, gamma, delta, cur_pyr_lvl, n_pyr_lvls, scale_to_sensor_size_method)
    p = _cache.plan_for(xs, ys, ts, edges, edge_ts, sensor_size)
    loss, grad = p.value_and_grad_host(_as_theta(theta), hp, want_grad=True)
    if has_aux:
        return (loss, _aux_from_plan(p, loss)), grad
    return loss, grad


def _handover_vg(alpha_handover, prev_theta, theta, xs, ys, ts, edges, edge_ts, alpha, beta, gamma, delta, cur_pyr_lvl,
                 n_pyr_lvls, sensor_size, scale_to_sensor_size_method, has_aux=False):
    hp = _plan.make_hparams(alpha, beta, gamma, delta, cur_pyr_lvl, n_pyr_lvls, scale_to_sensor_size_method)
    p = _cache.plan_for(xs, ys, ts, edges, edge_ts, sensor_size)
    loss, da = p.handover_value_and_grad_host(float(np.asarray(alpha_handover)), _as_theta(prev_theta), _as_theta(theta), hp)
    return loss, da


def value_and_grad(fun: Callable, has_aux: bool = False) -> Callable:
    """``jax.value_and_grad(fun, has_aux=has_aux)`` for the two callables above (differentiated w.r.t. argument 0,
    as jaxopt does inside ``ScipyMinimize`` / ``ScipyBoundedMinimize``): value and analytic gradient come from ONE
    evaluation on the device.  ``fun`` may be ``loss_func`` / ``handover_loss_func`` or a ``functools.partial`` of
    them (the Hydra ``_partial_`` objects of reference configs/theta_loss_func/default.yaml)."""
    base, args0, kw0 = fun, (), {}
    while isinstance(base, functools.partial):
        args0 = base.args + args0
        kw0 = {**base.keywords, **kw0}
        base = base.func
    if base is loss_func:
        impl = _loss_func_vg
    elif base is handover_loss_func:
        impl = _handover_vg
    else:
        raise TypeError('value_and_grad supports eincm_b200.losses.loss_func / handover_loss_func (or partials of them)')

    def vg(*args, **kw):
        return impl(*args0, *args, **kw0, **kw, has_aux=has_aux)

    return vg


def compute_loss_objectives(theta, xs, ys, ts, edges, edge_ts, sensor_size) -> Dict[str, object]:
    """reference src/eincm/losses.py:49-105 for a dense ``theta`` of shape (H, W, 2) (what
    ``evaluate_theta_array`` feeds it, reference src/evaluations/theta_eval.py:14-24).  Entries the reference
    derives but the loss never uses (warped coordinates, theta divergence, FWL) are not materialised here."""
    hp = _plan.make_hparams(1.0, 1.0, 1.0, 1.0, 0, 1, 'bilinear')
    p = _cache.plan_for(xs, ys, ts, edges, edge_ts, sensor_size)
    p.value_and_grad_host(_as_theta(theta), hp, want_grad=False)
    s = p.scalars()
    out = {
        'correlations': s['correlations'], 'zero_correlations': s['zero_correlations'],
        'rel_correlations': s['correlations'] / (s['zero_correlations'] + EPSN),
        'contrasts': s['contrasts'], 'zero_contrast': s['zero_contrast'],
        'rel_contrasts': s['contrasts'] / (s['zero_contrast'] + EPSN),
        'theta_total_variation': s['theta_total_variation'],
        'iwe_divergences': s['iwe_divergences'], 'zero_iwe_divergence': s['zero_iwe_divergence'],
        'rel_iwe_divergences': s['iwe_divergences'] / (s['zero_iwe_divergence'] + EPSN),
        'multi_ref_weights': s['multi_ref_weights'],
    }
    return out


# ---------------------------------------------------------------------------------------------------------------
# object form used by the solver mirror and the benchmark
# ---------------------------------------------------------------------------------------------------------------
class WindowObjective:
    """A staged event window plus the bound hyper-parameters: exposes exactly what jaxopt hands scipy,
    ``scipy_fun(x_flat) -> (float value, float64 grad_flat)`` (``jac=True``)."""

    def __init__(self, sensor_size, alpha, beta, gamma=0.0, delta=0.0, n_pyr_lvls=5, scale_to_sensor_size_method='bilinear',
                 max_events: int = 1 << 16, max_refs: int = 5, device: Optional[int] = None, flags: int = 0):
        self.sensor_size = (int(sensor_size[0]), int(sensor_size[1]))
        self.kw = dict(alpha=alpha, beta=beta, gamma=gamma, delta=delta, n_pyr_lvls=n_pyr_lvls,
                       scale_to_sensor_size_method=scale_to_sensor_size_method)
        self.plan = _plan.Plan(self.sensor_size, max_events=max_events, max_refs=max_refs, device=device, flags=flags)
        self.n_evals = 0

    def hparams(self, cur_pyr_lvl: int) -> _plan.HParams:
        return _plan.make_hparams(cur_pyr_lvl=cur_pyr_lvl, **self.kw)

    def set_datasample(self, xs, ys, ts, edges, edge_ts):
        """reference src/eincm/solver.py:185-194."""
        self.plan.set_window(xs, ys, ts, edges, edge_ts)

    def value_and_grad(self, theta, cur_pyr_lvl: int):
        self.n_evals += 1
        return self.plan.value_and_grad_host(_as_theta(theta), self.hparams(cur_pyr_lvl))

    def value(self, theta, cur_pyr_lvl: int) -> float:
        self.n_evals += 1
        return self.plan.value_and_grad_host(_as_theta(theta), self.hparams(cur_pyr_lvl), want_grad=False)[0]

    def handover_value_and_grad(self, alpha_handover, prev_theta, theta, cur_pyr_lvl: int):
        self.n_evals += 1
        return self.plan.handover_value_and_grad_host(alpha_handover, _as_theta(prev_theta), _as_theta(theta),
                                                      self.hparams(cur_pyr_lvl))

    def scipy_fun(self, shape, cur_pyr_lvl: int) -> Callable:
        hp = self.hparams(cur_pyr_lvl)

        def fun(x_flat):
            self.n_evals += 1
            v, g = self.plan.value_and_grad_host(np.asarray(x_flat, dtype=np.float64).reshape(shape), hp)
            return v, g.ravel()

        return fun

    # native optimizers (eincm_minimize_*_host): the BFGS / bounded loops run inside the library
    def minimize_bfgs(self, theta0, cur_pyr_lvl: int, maxiter: int, gtol: float, own_stream: bool = False):
        theta, res = self.plan.minimize_bfgs_host(_as_theta(theta0), self.hparams(cur_pyr_lvl), maxiter, gtol, own_stream=own_stream)
        self.n_evals += int(res.nfev)
        return theta, res

    def minimize_bfgs_graph(self, theta0, cur_pyr_lvl: int, maxiter: int, gtol: float):
        """The level solve as ONE CUDA graph with the loop on the device (eincm_minimize_bfgs_graph_host)."""
        theta, res = self.plan.minimize_bfgs_graph_host(_as_theta(theta0), self.hparams(cur_pyr_lvl), maxiter, gtol)
        self.n_evals += int(res.nfev)
        return theta, res

    def minimize_handover(self, alpha0, bounds, prev_theta, theta, cur_pyr_lvl: int, maxiter: int, pgtol: float, own_stream: bool = False):
        a, res = self.plan.minimize_handover_host(alpha0, bounds, _as_theta(prev_theta), _as_theta(theta), self.hparams(cur_pyr_lvl),
                                                  maxiter, pgtol, own_stream=own_stream)
        self.n_evals += int(res.nfev)
        return a, res

    def close(self):
        self.plan.close()
