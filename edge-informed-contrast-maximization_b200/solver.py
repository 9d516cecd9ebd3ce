"""Host-side mirror of the reference's multi-level solver (reference src/eincm/solver.py), driven by scipy.

The reference builds, per pyramid level, ``jaxopt.ScipyMinimize(fun=partial(loss, cur_pyr_lvl=l), method='BFGS', maxiter, jit=True,
has_aux=True, options={'gtol', 'return_all'})`` and ``jaxopt.ScipyBoundedMinimize(fun=partial(handover_loss, ...), 'L-BFGS-B')``
(solver.py:165-183).  Underneath, jaxopt calls ``scipy.optimize.minimize(scipy_fun, x0.ravel(), jac=True, method, callback,
options)`` with ``scipy_fun = jit(value_and_grad(fun))``.  jaxopt is not installable in this image, so this module makes the
same scipy calls directly with ``scipy_fun`` backed by the CUDA plan (``losses.WindowObjective``): the level schedule,
retries, handover, ``repeat`` up-scaling and the ``jax.image.scale_and_translate`` down-scaling of the priors follow
solver.py line by line, so complete solves - and the BASELINE metric "windows/s" - can be run and measured here.

The solver is the CALLER of the hot path: it stays host Python in the reference and here.
"""
from __future__ import annotations

import collections
import functools
import math
from typing import Dict, Optional, Sequence

import numpy as np

OptState = collections.namedtuple('OptState', 'fun_val success status iter_num n_evals')


def growing_maxiters(miniter, maxiter, n_pyr_lvls, order):
    """reference src/experiments/e00/exp_mgr.py:169-187."""
    out = {}
    for lvl in range(n_pyr_lvls):
        p = lvl / (n_pyr_lvls - 1)
        out[lvl] = int(math.ceil(miniter * p ** order + maxiter * (1 - p) ** order))
    return out


# ---------------------------------------------------------------------------------------------------------------
# jax.image.scale_and_translate on the host (used by the solver for the priors only)
# ---------------------------------------------------------------------------------------------------------------
def _kernel(method: str):
    if method in ('linear', 'bilinear', 'trilinear'):
        return lambda x: np.maximum(0.0, 1.0 - np.abs(x))
    if method in ('lanczos3', 'lanczos5'):
        radius = 3.0 if method == 'lanczos3' else 5.0

        def lanczos(x):
            y = radius * np.sin(np.pi * x) * np.sin(np.pi * x / radius)
            with np.errstate(divide='ignore', invalid='ignore'):
                out = np.where(x > 1e-3, y / np.where(x != 0, np.pi ** 2 * x ** 2, 1.0), 1.0)
            return np.where(x > radius, 0.0, out)

        return lanczos
    if method in ('cubic', 'bicubic', 'tricubic'):
        def keys(x):                                    # Keys cubic, a = -0.5
            out = ((1.5 * x - 2.5) * x) * x + 1.0
            out = np.where(x >= 1.0, ((-0.5 * x + 2.5) * x - 4.0) * x + 2.0, out)
            return np.where(x >= 2.0, 0.0, out)

        return keys
    raise NotImplementedError(f'resize method "{method}" is not implemented')


@functools.lru_cache(maxsize=256)
def compute_weight_mat(input_size: int, output_size: int, scale: float, method: str, antialias: bool = True) -> np.ndarray:
    """``jax._src.image.scale.compute_weight_mat`` with translation 0: ``(input_size, output_size)`` float64 (cached per shape: the
    solver resizes the same five pyramid shapes for every window; treat the result as read-only)."""
    inv_scale = 1.0 / scale
    kernel_scale = max(inv_scale, 1.0) if antialias else 1.0
    sample_f = (np.arange(output_size) + 0.5) * inv_scale - 0.5
    x = np.abs(sample_f[None, :] - np.arange(input_size)[:, None]) / kernel_scale
    weights = _kernel(method)(x)
    total = weights.sum(axis=0, keepdims=True)
    weights = np.where(np.abs(total) > 1000.0 * np.finfo(np.float32).eps, weights / np.where(total != 0, total, 1.0), 0.0)
    ok = (sample_f >= -0.5) & (sample_f <= input_size - 0.5)
    out = np.where(ok[None, :], weights, 0.0)
    out.setflags(write=False)
    return out


def scale_and_translate(theta: np.ndarray, shape: Sequence[int], method: str) -> np.ndarray:
    """``jim.scale_and_translate(theta, shape, spatial_dims=(0, 1, 2), scale=[s, s, 1], translation=0, method)``
    (reference solver.py:356-361, :370-375); the channel axis has scale 1 and is the identity."""
    h, w = theta.shape[:2]
    H, W = int(shape[0]), int(shape[1])
    wy = compute_weight_mat(h, H, H / h, method)
    wx = compute_weight_mat(w, W, W / w, method)
    return np.einsum('ijc,iy,jx->yxc', theta, wy, wx)


class EmptySolverCallback:
    """The callback protocol the reference's solver drives (src/eincm/callbacks.py); does nothing."""

    def __init__(self):
        self.iters: Dict[str, int] = {}
        self.cur = 0

    def reset(self):
        self.iters = {}

    def set_cur_pyr_lvl(self, pyr_lvl):
        self.cur = pyr_lvl

    def reset_opt_iter(self):
        pass

    def set_prior_and_current_thetas(self, prior, current):
        pass

    def get_iters(self):
        return self.iters

    def __call__(self, intermediate_result=None):
        key = f'pyr_lvl_{self.cur}'
        self.iters[key] = self.iters.get(key, 0) + 1


DEFAULT_THETA_OPT = {'method': 'BFGS', 'maxiter': 40, 'miniter': 8, 'n_extra_attempts': {'pyr_lvl_0': 1, 'pyr_lvl_1': 1},
                     'options': {'gtol': 1e-7}}                                                   # main.yaml:31-40
DEFAULT_HANDOVER_OPT = {'method': 'L-BFGS-B', 'maxiter': 20, 'miniter': 4, 'options': {'gtol': 1e-6}}   # main.yaml:41-46
DEFAULT_HANDOVER_SETTINGS = {'use_handover': True, 'solve_handover_for_levels': [1, 0], 'use_downscaled_finest_priors': True,
                             'handover_limits': [0.0, 1.0], 'clip_solved_handover': False,
                             'clip_solved_handover_limits': [0.1, 0.9], 'alpha_handover': 0.67}    # main.yaml:51-59


class MultipleLevelEINCMSolver:
    """reference src/eincm/solver.py:10-383.  ``objective`` is a ``losses.WindowObjective`` (staged window + bound
    hyper-parameters); everything else keeps the reference's names and defaults (configs/main.yaml)."""

    def __init__(self, objective, n_pyr_lvls: int = 5, theta_opt_maxiters: Optional[Dict[str, int]] = None,
                 theta_opt_solver_params: Optional[dict] = None, handover_opt_maxiters: Optional[Dict[str, int]] = None,
                 handover_opt_solver_params: Optional[dict] = None, handover_settings: Optional[dict] = None,
                 pyramid_downscale_method: str = 'lanczos3', pyramid_upscale_method: str = 'repeat',
                 pyramid_bases: Optional[Sequence[int]] = None, theta_solver_callback=None, handover_solver_callback=None,
                 maxiters_grow_order: float = 1.413, backend: str = 'scipy', own_stream: bool = False):
        """``backend='scipy'``: the scipy calls jaxopt makes, every evaluation a host call; ``backend='native'``: the same two
        optimizers inside the library (``eincm_minimize_bfgs_host`` / ``eincm_minimize_handover_host``), one host call per
        level - per-iteration callbacks are not invoked.  ``own_stream``: run on the plan's own CUDA stream (several solvers
        driven from several host threads overlap on the GPU; the native calls release the GIL)."""
        assert backend in ('scipy', 'native', 'graph')
        self.backend, self.own_stream = backend, own_stream
        self.objective = objective
        self.n_pyr_lvls = n_pyr_lvls
        self.theta_opt_solver_params = theta_opt_solver_params or DEFAULT_THETA_OPT
        self.handover_opt_solver_params = handover_opt_solver_params or DEFAULT_HANDOVER_OPT
        self.handover_settings = dict(DEFAULT_HANDOVER_SETTINGS, **(handover_settings or {}))
        if theta_opt_maxiters is None:      # exp_mgr.py:169-187 with use_growing_maxiters
            m = growing_maxiters(self.theta_opt_solver_params.get('miniter', 8), self.theta_opt_solver_params['maxiter'],
                                 n_pyr_lvls, maxiters_grow_order)
            theta_opt_maxiters = {f'pyr_lvl_{k}': v for k, v in m.items()}
        if handover_opt_maxiters is None:
            m = growing_maxiters(self.handover_opt_solver_params.get('miniter', 4), self.handover_opt_solver_params['maxiter'],
                                 n_pyr_lvls, maxiters_grow_order)
            handover_opt_maxiters = {f'pyr_lvl_{k}': v for k, v in m.items()}
        assert len(theta_opt_maxiters) == n_pyr_lvls, 'theta_opt_maxiters should be provided for each pyramid level'
        assert len(handover_opt_maxiters) == n_pyr_lvls, 'handover_opt_maxiters should be provided for each pyramid level'
        self.theta_opt_maxiters = theta_opt_maxiters
        self.handover_opt_maxiters = handover_opt_maxiters
        hs = self.handover_settings
        self.use_handover = hs['use_handover']
        self.solve_handover_switch_per_level = {f'pyr_lvl_{l}': (l in hs['solve_handover_for_levels']) for l in range(n_pyr_lvls)}
        self.use_downscaled_finest_priors = hs['use_downscaled_finest_priors']
        self.clip_solved_handover = hs['clip_solved_handover']
        self.clip_solved_handover_limits = hs['clip_solved_handover_limits'] if self.clip_solved_handover else None
        self.alpha_handover = hs['alpha_handover']
        self.pyramid_downscale_method = pyramid_downscale_method
        self.pyramid_upscale_method = pyramid_upscale_method
        self.pyramid_bases = list(pyramid_bases) if pyramid_bases is not None else [2] * (n_pyr_lvls - 1)
        self.theta_solver_callback = theta_solver_callback or EmptySolverCallback()
        self.handover_solver_callback = handover_solver_callback or EmptySolverCallback()

        self.pre_opt_theta_pyr, self.opt_theta_pyr, self.handover_opt_theta_pyr, self.prior_theta_pyr = {}, {}, {}, {}
        self._initialize_theta_pyramids()
        self.init_handover_weight_pyr, self.final_handover_weight_pyr = {}, {}
        self._initialize_handover_weights()
        self.theta_opt_state_pyr, self.ho_opt_state_pyr = {}, {}
        self._IS_FIRST_SAMPLE = True

    # -- state ------------------------------------------------------------------------------------------------
    def not_first_sample(self):
        self._IS_FIRST_SAMPLE = False

    def _initialize_theta_pyramids(self, theta_pyr_init=None):                                    # solver.py:129-148
        """``theta_pyr_init``: a prior pyramid to start from (a resumed run hands in the last solved one) instead of zeros."""
        top = f'pyr_lvl_{self.n_pyr_lvls - 1}'
        pyrs = [self.pre_opt_theta_pyr, self.opt_theta_pyr, self.handover_opt_theta_pyr]
        if theta_pyr_init is not None:
            self.prior_theta_pyr = theta_pyr_init
        else:
            pyrs.append(self.prior_theta_pyr)
        for pyr in pyrs:
            pyr[top] = np.zeros((1, 1, 2))
        for pyr_lvl in reversed(range(self.n_pyr_lvls - 1)):
            key, key_coarser = f'pyr_lvl_{pyr_lvl}', f'pyr_lvl_{pyr_lvl + 1}'
            base = self.pyramid_bases[-pyr_lvl - 1]
            for pyr in pyrs:
                pyr[key] = self._upscale_theta(pyr[key_coarser], base=base)

    def _initialize_handover_weights(self):                                                       # solver.py:151-155
        for pyr_lvl in range(self.n_pyr_lvls):
            self.init_handover_weight_pyr[f'pyr_lvl_{pyr_lvl}'] = 0.5
            self.final_handover_weight_pyr[f'pyr_lvl_{pyr_lvl}'] = 0.5

    def set_datasample(self, xs, ys, ts, edges, edge_ts):                                         # solver.py:185-194
        self.objective.set_datasample(xs, ys, ts, edges, edge_ts)

    # -- the two scipy calls jaxopt makes -------------------------------------------------------------------------
    def _run_theta_solver(self, pyr_lvl: int, theta0: np.ndarray):
        """ScipyMinimize.run (solver.py:165-173, :209-216): BFGS on the raveled theta, jac=True."""
        key = f'pyr_lvl_{pyr_lvl}'
        if self.backend == 'graph' and int(np.prod(np.shape(theta0))) <= 1024:
            # the whole level as one CUDA graph: the BFGS loop runs on the device (levels beyond 1024 parameters: native host loop)
            theta, r = self.objective.minimize_bfgs_graph(theta0, pyr_lvl, self.theta_opt_maxiters[key],
                                                          self.theta_opt_solver_params['options']['gtol'])
            return theta, OptState(float(r.fun), r.status == 0, int(r.status), int(r.nit), int(r.nfev))
        if self.backend in ('native', 'graph'):
            theta, r = self.objective.minimize_bfgs(theta0, pyr_lvl, self.theta_opt_maxiters[key],
                                                    self.theta_opt_solver_params['options']['gtol'], own_stream=self.own_stream)
            return theta, OptState(float(r.fun), r.status == 0, int(r.status), int(r.nit), int(r.nfev))
        import scipy.optimize
        shape = theta0.shape
        n0 = self.objective.n_evals
        res = scipy.optimize.minimize(self.objective.scipy_fun(shape, pyr_lvl), np.asarray(theta0, dtype=np.float64).ravel(),
                                      jac=True, method=self.theta_opt_solver_params['method'], callback=self.theta_solver_callback,
                                      options={'maxiter': self.theta_opt_maxiters[key],
                                               'gtol': self.theta_opt_solver_params['options']['gtol'], 'return_all': True})
        state = OptState(float(res.fun), bool(res.success), int(res.status), int(res.nit), self.objective.n_evals - n0)
        return res.x.reshape(shape), state

    def _run_handover_solver(self, pyr_lvl: int, alpha0: float, bounds, prior_theta, theta):
        """ScipyBoundedMinimize.run (solver.py:175-183, :325-335): L-BFGS-B on the scalar handover weight."""
        key = f'pyr_lvl_{pyr_lvl}'
        if self.backend in ('native', 'graph'):
            a, r = self.objective.minimize_handover(alpha0, bounds, prior_theta, theta, pyr_lvl, self.handover_opt_maxiters[key],
                                                    self.handover_opt_solver_params['options']['gtol'], own_stream=self.own_stream)
            return a, OptState(float(r.fun), r.status == 0, int(r.status), int(r.nit), int(r.nfev))
        import scipy.optimize
        n0 = self.objective.n_evals

        def fun(a):
            v, da = self.objective.handover_value_and_grad(float(a[0]), prior_theta, theta, pyr_lvl)
            return v, np.array([da])

        res = scipy.optimize.minimize(fun, np.array([alpha0], dtype=np.float64), jac=True,
                                      method=self.handover_opt_solver_params['method'], bounds=[tuple(bounds)],
                                      callback=self.handover_solver_callback,
                                      options={'maxiter': self.handover_opt_maxiters[key],
                                               'gtol': self.handover_opt_solver_params['options']['gtol']})
        state = OptState(float(res.fun), bool(res.success), int(res.status), int(res.nit), self.objective.n_evals - n0)
        return float(res.x[0]), state

    # -- solve (solver.py:197-267) --------------------------------------------------------------------------------
    def solve(self) -> dict:
        self._pre_solve()
        extra = self.theta_opt_solver_params.get('n_extra_attempts', {})
        for pyr_lvl in reversed(range(self.n_pyr_lvls)):
            key, next_key = f'pyr_lvl_{pyr_lvl}', f'pyr_lvl_{pyr_lvl - 1}'
            self._update_callback_pyr_lvl(pyr_lvl)
            n_extra_attempts = 0
            self.opt_theta_pyr[key], self.theta_opt_state_pyr[key] = self._run_theta_solver(pyr_lvl, self.pre_opt_theta_pyr[key])
            while ((not self.theta_opt_state_pyr[key].success) and self.theta_opt_state_pyr[key].iter_num > 0
                   and key in extra and n_extra_attempts < extra[key]):
                n_extra_attempts += 1
                self.opt_theta_pyr[key], self.theta_opt_state_pyr[key] = self._run_theta_solver(pyr_lvl, self.opt_theta_pyr[key])
            self.handover_opt_theta_pyr[key] = self._perform_handover_at_level(pyr_lvl)
            if pyr_lvl != 0:
                self.pre_opt_theta_pyr[next_key] = self._upscale_theta(self.handover_opt_theta_pyr[key],
                                                                       base=self.pyramid_bases[-pyr_lvl])
        old_prior_theta_pyr = dict(self.prior_theta_pyr)
        self.prior_theta_pyr = dict(self.handover_opt_theta_pyr)
        self._IS_FIRST_SAMPLE = False
        return {
            'prior_theta_pyr': old_prior_theta_pyr,
            'pre_opt_theta_pyr': dict(self.pre_opt_theta_pyr),
            'theta_opt_state_pyr': dict(self.theta_opt_state_pyr),
            'pre_handover_theta_pyr': dict(self.opt_theta_pyr),
            'ho_opt_state_pyr': dict(self.ho_opt_state_pyr),
            'final_handover_weight_pyr': dict(self.final_handover_weight_pyr),
            'final_theta_pyr': dict(self.handover_opt_theta_pyr),
        }

    def _pre_solve(self):                                                                         # solver.py:270-281
        self._stage_prior_theta_pyr()
        key_coarsest = f'pyr_lvl_{self.n_pyr_lvls - 1}'
        self.pre_opt_theta_pyr[key_coarsest] = self.prior_theta_pyr[key_coarsest]
        self.theta_solver_callback.reset()
        self.handover_solver_callback.reset()
        self.theta_opt_state_pyr, self.ho_opt_state_pyr = {}, {}

    def _stage_prior_theta_pyr(self):                                                             # solver.py:283-289
        if self.use_downscaled_finest_priors:
            for pyr_lvl in range(1, self.n_pyr_lvls):
                key, key_finer = f'pyr_lvl_{pyr_lvl}', f'pyr_lvl_{pyr_lvl - 1}'
                self.prior_theta_pyr[key] = self._downscale_theta(self.prior_theta_pyr[key_finer],
                                                                  base=self.pyramid_bases[-(pyr_lvl - 1) - 1])

    def _update_callback_pyr_lvl(self, pyr_lvl):                                                  # solver.py:292-299
        for cb in (self.theta_solver_callback, self.handover_solver_callback):
            cb.set_cur_pyr_lvl(pyr_lvl)
            cb.reset_opt_iter()

    def _perform_handover_at_level(self, cur_pyr_lvl):                                            # solver.py:302-347
        key, key_finer = f'pyr_lvl_{cur_pyr_lvl}', f'pyr_lvl_{cur_pyr_lvl - 1}'
        self.handover_solver_callback.set_prior_and_current_thetas(self.prior_theta_pyr[key], self.opt_theta_pyr[key])
        if self._IS_FIRST_SAMPLE or not self.use_handover:
            return self.opt_theta_pyr[key]
        if self.solve_handover_switch_per_level[key]:
            # the up-scaling follows the handover, so the weight is solved at the finer scale unless already finest
            if cur_pyr_lvl > 0:
                prior_theta = self.prior_theta_pyr[key_finer]
                theta = self._upscale_theta(self.opt_theta_pyr[key], self.pyramid_bases[-cur_pyr_lvl])
                solver_lvl = cur_pyr_lvl - 1
            else:
                prior_theta, theta, solver_lvl = self.prior_theta_pyr[key], self.opt_theta_pyr[key], cur_pyr_lvl
            w, self.ho_opt_state_pyr[key] = self._run_handover_solver(solver_lvl, self.init_handover_weight_pyr[key],
                                                                      self.handover_settings['handover_limits'], prior_theta, theta)
            if self.clip_solved_handover:
                w = float(np.clip(w, *self.clip_solved_handover_limits))
            self.final_handover_weight_pyr[key] = w
        else:
            self.final_handover_weight_pyr[key] = self.alpha_handover
        a = self.final_handover_weight_pyr[key]
        return a * self.prior_theta_pyr[key] + (1 - a) * self.opt_theta_pyr[key]

    def _upscale_theta(self, theta, base=2):                                                      # solver.py:350-364
        if self.pyramid_upscale_method == 'repeat':
            return np.repeat(np.repeat(theta, base, axis=0), base, axis=1)
        return scale_and_translate(theta, (int(theta.shape[0] * base), int(theta.shape[1] * base)), self.pyramid_upscale_method)

    def _downscale_theta(self, theta, base=2):                                                    # solver.py:366-377
        return scale_and_translate(theta, (int(theta.shape[0] / base), int(theta.shape[1] / base)), self.pyramid_downscale_method)


def solve_sequence(objective, windows, solver_kwargs: Optional[dict] = None, n_repeat_solve: int = 1):
    """The run loop of reference src/experiments/e00/exp_mgr.py:615-659 for a sequence of staged windows: consecutive windows
    are chained through ``prior_theta_pyr`` (handover).  Returns the list of ``solve()`` results."""
    solver = MultipleLevelEINCMSolver(objective, **(solver_kwargs or {}))
    out = []
    for w in windows:
        solver.set_datasample(*w.args())
        for _ in range(n_repeat_solve):
            res = solver.solve()
        out.append(res)
    return out


class BatchedMultipleLevelEINCMSolver:
    """B independent windows (different sequences, or windows solved without handover between them) taken through the level schedule
    of ``MultipleLevelEINCMSolver.solve`` (reference src/eincm/solver.py:197-267) in LOCKSTEP: per pyramid level ONE call solves the level
    for every window on the device (``eincm_batch_minimize_bfgs_graph_host``: batched evaluation kernels + one optimizer CTA per window inside
    a CUDA graph).  Each window keeps its own ``MultipleLevelEINCMSolver`` state (pyramids, priors, handover weights), so
    consecutive batches chain through the handover prior exactly like consecutive windows of a sequence; retries (solver.py:218-226) re-solve
    the subset of windows that asked for one; the scalar handover solves stay per window."""

    def __init__(self, objectives, handover_threads: int = 4, **solver_kwargs):
        """``handover_threads``: host threads that run the per-window scalar handover solves of a level side by side (each on its plan's own
        stream; the native call releases the GIL) - they are independent, and one at a time they are bound by the latency of a small evaluation."""
        from . import plan as _plan
        solver_kwargs = dict(solver_kwargs, backend='graph', own_stream=True)
        self.handover_threads = max(1, int(handover_threads))
        self._pool = None
        self.solvers = [MultipleLevelEINCMSolver(o, **solver_kwargs) for o in objectives]
        self.objectives = list(objectives)
        self.batch = _plan.Batch([o.plan for o in objectives])
        self.n_pyr_lvls = self.solvers[0].n_pyr_lvls
        self.graph_launches = 0

    def close(self):
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
        self.batch.close()

    def _handover_all(self, pyr_lvl: int):
        """``_perform_handover_at_level`` of every window; the scalar solves (first window / unsolved levels: a blend, no solve) in parallel."""
        sv = self.solvers
        key = f'pyr_lvl_{pyr_lvl}'
        solved = (not sv[0]._IS_FIRST_SAMPLE) and sv[0].use_handover and sv[0].solve_handover_switch_per_level[key]
        if not solved or self.handover_threads == 1 or len(sv) == 1:
            return [s._perform_handover_at_level(pyr_lvl) for s in sv]
        if self._pool is None:
            import concurrent.futures
            self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=self.handover_threads)
        dev = self.objectives[0].plan.device

        def one(s):
            import torch
            torch.cuda.set_device(dev)
            return s._perform_handover_at_level(pyr_lvl)

        return list(self._pool.map(one, sv))

    def set_datasamples(self, windows):
        """``windows[k]``: the ``(xs, ys, ts, edges, edge_ts)`` operands of window k."""
        for s, w in zip(self.solvers, windows):
            s.set_datasample(*w)

    def _run_theta_solvers(self, pyr_lvl: int, thetas0, active=None):
        s0 = self.solvers[0]
        key = f'pyr_lvl_{pyr_lvl}'
        hp = self.objectives[0].hparams(pyr_lvl)
        if hp.delta != 0.0 or (hp.gamma != 0.0 and pyr_lvl <= 0) or int(np.prod(np.shape(thetas0[0]))) > 1024:
            # what the batched kernels do not evaluate (the TV regulariser at the finest level - gamma != 0: MVSEC outdoor, run.sh:85 -, delta != 0)
            # or hold (more than 1024 parameters): one window at a time through the single-window solver
            thetas, states = [], []
            for k, s in enumerate(self.solvers):
                if active is None or active[k]:
                    th, st = s._run_theta_solver(pyr_lvl, thetas0[k])
                else:
                    th, st = np.asarray(thetas0[k]), None
                thetas.append(th)
                states.append(st)
            return thetas, states
        thetas, res = self.batch.minimize_bfgs_graph_host(np.stack(thetas0), self.objectives[0].hparams(pyr_lvl), s0.theta_opt_maxiters[key],
                                                          s0.theta_opt_solver_params['options']['gtol'], active=active)
        self.graph_launches += self.batch.solve_launches()
        states = [OptState(float(r.fun), r.status == 0, int(r.status), int(r.nit), int(r.nfev)) for r in res]
        for k, (o, r) in enumerate(zip(self.objectives, res)):
            if active is None or active[k]:
                o.n_evals += int(r.nfev)
        return thetas, states

    def solve(self):
        sv = self.solvers
        for s in sv:
            s._pre_solve()
        extra = sv[0].theta_opt_solver_params.get('n_extra_attempts', {})
        for pyr_lvl in reversed(range(self.n_pyr_lvls)):
            key, next_key = f'pyr_lvl_{pyr_lvl}', f'pyr_lvl_{pyr_lvl - 1}'
            for s in sv:
                s._update_callback_pyr_lvl(pyr_lvl)
            thetas, states = self._run_theta_solvers(pyr_lvl, [s.pre_opt_theta_pyr[key] for s in sv])
            for k, s in enumerate(sv):
                s.opt_theta_pyr[key], s.theta_opt_state_pyr[key] = thetas[k], states[k]
            n_extra_attempts = 0
            while key in extra and n_extra_attempts < extra[key]:
                again = np.array([(not s.theta_opt_state_pyr[key].success) and s.theta_opt_state_pyr[key].iter_num > 0 for s in sv], dtype=np.int32)
                if not again.any():
                    break
                n_extra_attempts += 1
                thetas, states = self._run_theta_solvers(pyr_lvl, [s.opt_theta_pyr[key] for s in sv], active=again)
                for k, s in enumerate(sv):
                    if again[k]:
                        s.opt_theta_pyr[key], s.theta_opt_state_pyr[key] = thetas[k], states[k]
            for s, th_ho in zip(sv, self._handover_all(pyr_lvl)):
                s.handover_opt_theta_pyr[key] = th_ho
                if pyr_lvl != 0:
                    s.pre_opt_theta_pyr[next_key] = s._upscale_theta(s.handover_opt_theta_pyr[key], base=s.pyramid_bases[-pyr_lvl])
        out = []
        for s in sv:
            old_prior = dict(s.prior_theta_pyr)
            s.prior_theta_pyr = dict(s.handover_opt_theta_pyr)
            s._IS_FIRST_SAMPLE = False
            out.append({'prior_theta_pyr': old_prior, 'pre_opt_theta_pyr': dict(s.pre_opt_theta_pyr),
                        'theta_opt_state_pyr': dict(s.theta_opt_state_pyr), 'pre_handover_theta_pyr': dict(s.opt_theta_pyr),
                        'ho_opt_state_pyr': dict(s.ho_opt_state_pyr), 'final_handover_weight_pyr': dict(s.final_handover_weight_pyr),
                        'final_theta_pyr': dict(s.handover_opt_theta_pyr)})
        return out
