"""ctypes binding of the C-ABI in ``include/eincm.h`` (``lib/libeincm_b200.so``).

PyTorch is used for plumbing only: device buffers (``torch.empty(..., device='cuda')``), the current CUDA
stream, and ``torch.distributed`` in ``parallel.py``.  All arithmetic happens in the CUDA library; there is no
CPU fallback - importing this module without the built library, or creating a plan without a B200, raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('EINCM_B200_LIB') or os.path.join(_PKG_DIR, 'lib', 'libeincm_b200.so')     # the override is for A/B builds

EINCM_OK = 0
EINCM_EINVAL, EINCM_ECUDA, EINCM_ENOMEM, EINCM_ESTATE, EINCM_ERANGE, EINCM_EUNSUPPORTED = -1, -2, -3, -4, -5, -6
FLAG_NO_WRAP_NEGATIVE = 0x1
FLAG_EVENT_SPLIT = 0x2
FLAG_EXACT_F64 = 0x4
FLAG_BLOCKING_SYNC = 0x8
METHOD_BILINEAR = 0
S_HEADER = 8

# every symbol include/eincm.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = (
    'eincm_plan_create', 'eincm_plan_destroy', 'eincm_last_error', 'eincm_abi_version', 'eincm_plan_set_window',
    'eincm_plan_set_window_device_ts', 'eincm_plan_set_split_fixed_point',
    'eincm_value_and_grad', 'eincm_handover_value_and_grad', 'eincm_value_and_grad_host',
    'eincm_handover_value_and_grad_host', 'eincm_value_and_grad_stateless_host', 'eincm_value_and_grad_host_batch',
    'eincm_window_finalize',
    'eincm_forward_events', 'eincm_backward', 'eincm_zero_iwe_ptr', 'eincm_iwe_ptr', 'eincm_dldi_ptr',
    'eincm_theta_full_ptr', 'eincm_mask_ptr', 'eincm_get_scalars', 'eincm_debug_rounded_pixels', 'eincm_plan_info',
    'eincm_plan_set_event_split', 'eincm_plan_launch_count', 'eincm_plan_set_timing', 'eincm_plan_get_timing',
    'eincm_plan_ipc_handle', 'eincm_plan_set_peers', 'eincm_plan_set_peer_pointers', 'eincm_iwe_fix_ptr', 'eincm_split_prepare',
    'eincm_split_window_images', 'eincm_minimize_bfgs_host', 'eincm_minimize_bfgs_graph_host', 'eincm_minimize_handover_host', 'eincm_sparse_flow_error',
    'eincm_evaluate_theta', 'eincm_group_create', 'eincm_group_destroy', 'eincm_plan_set_group', 'eincm_group_set_burst_percent', 'eincm_plan_host_times',
    'eincm_edge_workspace_bytes', 'eincm_edge_maps', 'eincm_edge_maps_host',
    'eincm_nlm_workspace_bytes', 'eincm_nlm_denoise', 'eincm_clahe_workspace_bytes', 'eincm_clahe', 'eincm_sharpen',
    'eincm_rectify_workspace_bytes', 'eincm_rectify_events', 'eincm_normalize_times', 'eincm_window_event_range',
    'eincm_batch_create', 'eincm_batch_destroy', 'eincm_batch_last_error', 'eincm_batch_value_and_grad', 'eincm_batch_value_and_grad_host',
    'eincm_batch_launch_count', 'eincm_batch_set_timing', 'eincm_batch_get_timing', 'eincm_batch_minimize_bfgs_graph_host',
    'eincm_batch_solve_launches',
)


class EincmError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f'eincm error {code}: {message}')
        self.code = code


class OptResult(C.Structure):
    """``eincm_opt_result``: what ``jaxopt`` reports as ``state.fun_val / iter_num / status`` (reference src/eincm/solver.py:379-384)."""
    _fields_ = [('fun', C.c_double), ('nit', C.c_int32), ('nfev', C.c_int32), ('status', C.c_int32), ('reserved', C.c_int32)]


OWN_STREAM = C.c_void_p(-1)     # stream argument selecting the plan's own stream
EVAL_MAX_REFS = 8


class FlowErrors(C.Structure):
    """``eincm_flow_errors`` (reference src/evaluations/flow_eval.py:14-75)."""
    _fields_ = [('AEE', C.c_double), ('AREE', C.c_double), ('ANPE', C.c_double * 6), ('n_ee', C.c_int64), ('n_pred', C.c_int64),
                ('n_gt', C.c_int64)]

    def as_dict(self):
        errs = {'AEE': self.AEE, 'AREE': self.AREE}
        for k, n in enumerate((1, 2, 3, 5, 10, 20)):
            errs[f'A{n}PE'] = self.ANPE[k]
        return {'errors': errs, 'counts': {'n_ee': int(self.n_ee), 'n_pred': int(self.n_pred), 'n_gt': int(self.n_gt)}}


class EvalMetrics(C.Structure):
    """``eincm_eval_metrics`` (reference src/evaluations/theta_eval.py:80-94)."""
    _fields_ = [('loss', C.c_double), ('iwe_var', C.c_double), ('mean_rel_contrast', C.c_double), ('mean_rel_corr', C.c_double),
                ('mean_rel_iwe_div', C.c_double), ('theta_tot_var', C.c_double), ('theta_div', C.c_double), ('fwl', C.c_double),
                ('rel_contrasts', C.c_double * EVAL_MAX_REFS), ('rel_correlations', C.c_double * EVAL_MAX_REFS),
                ('rel_iwe_divergences', C.c_double * EVAL_MAX_REFS), ('flow_warp_losses', C.c_double * EVAL_MAX_REFS),
                ('multi_ref_weights', C.c_double * EVAL_MAX_REFS), ('n_refs', C.c_int32), ('has_flow', C.c_int32),
                ('n_pixels', C.c_int64), ('flow', FlowErrors)]


class HParams(C.Structure):
    """``eincm_hparams``: keyword-bound arguments of the loss partial (reference src/eincm/losses.py:115-122)."""
    _fields_ = [('alpha', C.c_double), ('beta', C.c_double), ('gamma', C.c_double), ('delta', C.c_double),
                ('cur_pyr_lvl', C.c_int32), ('n_pyr_lvls', C.c_int32), ('method', C.c_int32), ('reserved', C.c_int32)]


def make_hparams(alpha, beta, gamma, delta, cur_pyr_lvl, n_pyr_lvls=5, scale_to_sensor_size_method='bilinear') -> HParams:
    if scale_to_sensor_size_method not in ('bilinear', 'linear'):
        raise EincmError(EINCM_EUNSUPPORTED, f'scale_to_sensor_size_method {scale_to_sensor_size_method!r}: only the '
                                              f'shipped default (bilinear) is implemented')
    return HParams(float(alpha), float(beta), float(gamma), float(delta), int(cur_pyr_lvl), int(n_pyr_lvls), METHOD_BILINEAR, 0)


_lib = None


class EdgeParams(C.Structure):
    """``eincm_edge_params`` of include/eincm.h."""
    _fields_ = [('canny_th1', C.c_double), ('canny_th2', C.c_double), ('smoothen', C.c_int32), ('stages', C.c_int32),
                ('gauss_sigma', C.c_double), ('iedt_alpha', C.c_double)]


EINCM_SMOOTHEN_GAUSSIAN, EINCM_SMOOTHEN_IEDT = 0, 1
EINCM_EDGE_STAGE_CANNY, EINCM_EDGE_STAGE_SMOOTHEN, EINCM_EDGE_STAGE_NORMALIZE = 1, 2, 4


def load_library(path: Optional[str] = None) -> C.CDLL:
    """Loads the CUDA library.  Fails loudly when it has not been built (``python -c 'import __graft_entry__ as g;
    g.build()'``): the product has no other code path."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise ImportError(f'{path} is missing - build it with __graft_entry__.build(); there is no CPU fallback')
    lib = C.CDLL(path)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    hp = C.POINTER(HParams)
    sig = {
        'eincm_plan_create': (i32, [C.POINTER(vp), i32, i32, i32, i64, i32, C.c_uint]),
        'eincm_plan_destroy': (None, [vp]),
        'eincm_last_error': (C.c_char_p, [vp]),
        'eincm_abi_version': (i32, []),
        'eincm_plan_set_window': (i32, [vp, vp, vp, vp, i64, vp, C.POINTER(dbl), i32, vp]),
        'eincm_plan_set_window_device_ts': (i32, [vp, vp, vp, vp, i64, vp, vp, i32, vp]),
        'eincm_value_and_grad': (i32, [vp, vp, i32, i32, hp, vp, vp, vp]),
        'eincm_handover_value_and_grad': (i32, [vp, dbl, vp, vp, i32, i32, hp, vp, vp, vp]),
        'eincm_value_and_grad_host': (i32, [vp, vp, i32, i32, hp, C.POINTER(dbl), vp, vp]),
        'eincm_handover_value_and_grad_host': (i32, [vp, dbl, vp, vp, i32, i32, hp, C.POINTER(dbl), C.POINTER(dbl), vp]),
        'eincm_value_and_grad_stateless_host': (i32, [vp, vp, i32, i32, vp, vp, vp, i64, vp, vp, i32, hp, C.POINTER(dbl), vp, vp]),
        'eincm_value_and_grad_host_batch': (i32, [C.POINTER(vp), i32, C.POINTER(vp), i32, i32, hp, C.POINTER(dbl), C.POINTER(vp)]),
        'eincm_window_finalize': (i32, [vp, vp]),
        'eincm_forward_events': (i32, [vp, vp, i32, i32, hp, vp]),
        'eincm_backward': (i32, [vp, hp, vp, vp, vp]),
        'eincm_zero_iwe_ptr': (vp, [vp]),
        'eincm_iwe_ptr': (vp, [vp]),
        'eincm_dldi_ptr': (vp, [vp]),
        'eincm_theta_full_ptr': (vp, [vp]),
        'eincm_mask_ptr': (vp, [vp]),
        'eincm_get_scalars': (i32, [vp, C.POINTER(dbl), i32, vp]),
        'eincm_debug_rounded_pixels': (i32, [vp, i32, vp, vp, vp]),
        'eincm_plan_set_event_split': (i32, [vp, i32, i32]),
        'eincm_plan_set_split_fixed_point': (i32, [vp, i32]),
        'eincm_plan_ipc_handle': (i32, [vp, vp, i32]),
        'eincm_plan_set_peers': (i32, [vp, vp, i32]),
        'eincm_plan_set_peer_pointers': (i32, [vp, C.POINTER(vp), i32]),
        'eincm_iwe_fix_ptr': (vp, [vp]),
        'eincm_split_prepare': (i32, [vp, vp]),
        'eincm_split_window_images': (i32, [vp, vp]),
        'eincm_minimize_bfgs_host': (i32, [vp, vp, i32, i32, hp, i32, dbl, C.POINTER(OptResult), vp]),
        'eincm_minimize_bfgs_graph_host': (i32, [vp, vp, i32, i32, hp, i32, dbl, C.POINTER(OptResult), vp]),
        'eincm_minimize_handover_host': (i32, [vp, C.POINTER(dbl), dbl, dbl, vp, vp, i32, i32, hp, i32, dbl, C.POINTER(OptResult), vp]),
        'eincm_sparse_flow_error': (i32, [i32, i32, i32, vp, vp, vp, C.POINTER(FlowErrors), vp]),
        'eincm_evaluate_theta': (i32, [vp, vp, i32, i32, hp, vp, vp, C.POINTER(EvalMetrics), vp]),
        'eincm_group_create': (i32, [C.POINTER(vp)]),
        'eincm_group_destroy': (None, [vp]),
        'eincm_plan_set_group': (i32, [vp, vp]),
        'eincm_group_set_burst_percent': (i32, [vp, i32]),
        'eincm_plan_host_times': (i32, [vp, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(i64), i32]),
        'eincm_plan_launch_count': (i64, [vp]),
        'eincm_plan_set_timing': (i32, [vp, i32]),
        'eincm_plan_get_timing': (i32, [vp, C.c_char_p, i32, C.POINTER(dbl), C.POINTER(i64), i32, C.POINTER(i32)]),
        'eincm_plan_info': (i32, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i64), C.POINTER(i32), C.POINTER(i32)]),
        'eincm_edge_workspace_bytes': (C.c_size_t, [i32, i32, i32]),
        'eincm_edge_maps': (i32, [i32, vp, i32, i32, i32, C.POINTER(EdgeParams), vp, vp, vp, C.c_size_t, vp]),
        'eincm_edge_maps_host': (i32, [i32, vp, i32, i32, i32, C.POINTER(EdgeParams), vp, vp]),
        'eincm_nlm_workspace_bytes': (C.c_size_t, [i32, i32]),
        'eincm_nlm_denoise': (i32, [i32, vp, i32, i32, i32, C.c_float, i32, i32, vp, vp, C.c_size_t, vp]),
        'eincm_clahe_workspace_bytes': (C.c_size_t, [i32, i32, i32]),
        'eincm_clahe': (i32, [i32, vp, i32, i32, i32, dbl, i32, i32, vp, vp, C.c_size_t, vp]),
        'eincm_sharpen': (i32, [i32, vp, i32, i32, i32, dbl, dbl, dbl, dbl, vp, vp, vp]),
        'eincm_rectify_workspace_bytes': (C.c_size_t, [i64]),
        'eincm_rectify_events': (i32, [i32, vp, vp, vp, vp, i64, vp, i32, i32, vp, vp, vp, vp, C.POINTER(i64), vp, C.c_size_t, vp]),
        'eincm_normalize_times': (i32, [i32, vp, i64, i64, i64, vp, vp]),
        'eincm_window_event_range': (i32, [i64, i64, i64, i64, i32, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]),
        'eincm_batch_create': (i32, [C.POINTER(vp), C.POINTER(vp), i32]),
        'eincm_batch_destroy': (None, [vp]),
        'eincm_batch_last_error': (C.c_char_p, [vp]),
        'eincm_batch_value_and_grad': (i32, [vp, C.POINTER(vp), i32, i32, hp, C.POINTER(vp), C.POINTER(vp), vp]),
        'eincm_batch_value_and_grad_host': (i32, [vp, C.POINTER(vp), i32, i32, hp, C.POINTER(dbl), C.POINTER(vp), vp]),
        'eincm_batch_launch_count': (i64, [vp]),
        'eincm_batch_set_timing': (i32, [vp, i32]),
        'eincm_batch_get_timing': (i32, [vp, C.POINTER(dbl), C.POINTER(i64)]),
        'eincm_batch_minimize_bfgs_graph_host': (i32, [vp, vp, i32, i32, hp, i32, dbl, vp, vp, vp]),
        'eincm_batch_solve_launches': (i64, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path == LIB_PATH:
        _lib = lib
    return lib


def _torch():
    import torch
    return torch


def _stream_ptr(stream=None) -> int:
    torch = _torch()
    s = stream if stream is not None else torch.cuda.current_stream()
    return int(s.cuda_stream)


def value_and_grad_host_batch(plans: Sequence['Plan'], thetas: Sequence[np.ndarray], hp: HParams, want_grad: bool = True):
    """``eincm_value_and_grad_host_batch``: one synchronous call evaluating independent windows (one plan each) concurrently.
    Returns ``(losses float64[n], grads list of (h, w, 2) arrays or None)``."""
    n = len(plans)
    if n == 0 or len(thetas) != n:
        raise EincmError(EINCM_EINVAL, 'plans and thetas must be non-empty and of equal length')
    ths = [np.ascontiguousarray(t, dtype=np.float64) for t in thetas]
    shape = ths[0].shape
    if len(shape) != 3 or shape[2] != 2 or any(t.shape != shape for t in ths):
        raise EincmError(EINCM_EINVAL, 'every theta of a batch must have the same shape (h, w, 2)')
    lib = plans[0].lib
    hs = (C.c_void_p * n)(*[p._h.value for p in plans])
    tp = (C.c_void_p * n)(*[t.ctypes.data for t in ths])
    losses = np.empty(n, dtype=np.float64)
    grads = [np.empty_like(t) for t in ths] if want_grad else None
    gp = (C.c_void_p * n)(*[g.ctypes.data for g in grads]) if want_grad else None
    rc = lib.eincm_value_and_grad_host_batch(hs, n, tp, shape[0], shape[1], C.byref(hp), losses.ctypes.data_as(C.POINTER(C.c_double)), gp)
    if rc != EINCM_OK:
        raise EincmError(rc, (lib.eincm_last_error(plans[0]._h) or b'').decode())
    return losses, grads


class Group:
    """``eincm_group``: plans that joined the same group rendezvous their objective evaluations inside ``minimize_bfgs_host`` - one
    thread launches the evaluations of all concurrently running minimisations in a burst, the others sleep (include/eincm.h)."""

    def __init__(self, burst_percent: int = 100):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.eincm_group_create(C.byref(h))
        if rc != 0:
            raise EincmError(rc, 'eincm_group_create failed')
        self._h = h
        rc = self.lib.eincm_group_set_burst_percent(self._h, int(burst_percent))
        if rc != 0:
            raise EincmError(rc, 'burst_percent must be in 1..100')

    def close(self):
        if getattr(self, '_h', None):
            self.lib.eincm_group_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """``eincm_batch``: B staged windows (one ``Plan`` each) evaluated with one launch per kernel of the evaluation
    (include/eincm.h, "batched evaluation").  Results are identical to the per-plan calls."""

    def __init__(self, plans: Sequence['Plan']):
        self.lib = plans[0].lib
        self.plans = list(plans)
        hs = (C.c_void_p * len(plans))(*[p._h.value for p in plans])
        h = C.c_void_p()
        rc = self.lib.eincm_batch_create(C.byref(h), hs, len(plans))
        if rc != EINCM_OK:
            raise EincmError(rc, (self.lib.eincm_last_error(plans[0]._h) or b'').decode())
        self._h = h
        self._keep = []

    def _check(self, rc: int):
        if rc != EINCM_OK:
            raise EincmError(rc, (self.lib.eincm_batch_last_error(self._h) or b'').decode())

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            self.lib.eincm_batch_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def value_and_grad_device(self, thetas_d, hp: HParams, losses_d, grads_d, stream=None):
        """Device operands (lists of CUDA tensors, one per window), asynchronous on the current (or given) stream."""
        n = len(self.plans)
        if not (len(thetas_d) == len(losses_d) == len(grads_d) == n):
            raise EincmError(EINCM_EINVAL, 'one theta / loss / gradient tensor per window of the batch')
        h, w = int(thetas_d[0].shape[0]), int(thetas_d[0].shape[1])
        tp = (C.c_void_p * n)(*[t.data_ptr() for t in thetas_d])
        lp = (C.c_void_p * n)(*[t.data_ptr() for t in losses_d])
        gp = (C.c_void_p * n)(*[t.data_ptr() for t in grads_d])
        self._keep = [thetas_d, losses_d, grads_d]
        self._check(self.lib.eincm_batch_value_and_grad(self._h, tp, h, w, C.byref(hp), lp, gp, _stream_ptr(stream)))

    def value_and_grad_host(self, thetas: Sequence[np.ndarray], hp: HParams, want_grad: bool = True, stream=None):
        """Host operands, synchronous: returns ``(losses float64[n], grads list of (h, w, 2) arrays or None)``."""
        n = len(self.plans)
        ths = [np.ascontiguousarray(t, dtype=np.float64) for t in thetas]
        shape = ths[0].shape
        if len(ths) != n or len(shape) != 3 or shape[2] != 2 or any(t.shape != shape for t in ths):
            raise EincmError(EINCM_EINVAL, 'one theta of shape (h, w, 2) per window of the batch')
        tp = (C.c_void_p * n)(*[t.ctypes.data for t in ths])
        losses = np.empty(n, dtype=np.float64)
        grads = [np.empty_like(t) for t in ths] if want_grad else None
        gp = (C.c_void_p * n)(*[g.ctypes.data for g in grads]) if want_grad else None
        self._check(self.lib.eincm_batch_value_and_grad_host(self._h, tp, shape[0], shape[1], C.byref(hp),
                                                             losses.ctypes.data_as(C.POINTER(C.c_double)), gp, _stream_ptr(stream)))
        return losses, grads

    def launch_count(self) -> int:
        return int(self.lib.eincm_batch_launch_count(self._h))

    def minimize_bfgs_graph_host(self, thetas0, hp: HParams, maxiter: int, gtol: float, active=None, stream=None):
        """One BFGS level solve for every window of the batch, in lockstep, with the loop on the device (``eincm_batch_minimize_bfgs_graph_host``):
        ``thetas0`` is ``(B, h, w, 2)``; returns ``(thetas (B, h, w, 2), [OptResult] * B)``.  ``active``: optional ``B`` flags, windows with a
        zero flag take no part (theta returned unchanged, result zero).  ``stream`` None: the first plan's own stream."""
        n = len(self.plans)
        thetas = np.array(thetas0, dtype=np.float64, order='C', copy=True)
        if thetas.ndim != 4 or thetas.shape[0] != n or thetas.shape[3] != 2:
            raise EincmError(EINCM_EINVAL, f'thetas must have shape ({n}, h, w, 2), got {thetas.shape}')
        res = (OptResult * n)()
        act = None if active is None else np.ascontiguousarray(active, dtype=np.int32)
        if act is not None and act.shape != (n,):
            raise EincmError(EINCM_EINVAL, f'active must have {n} entries')
        st = OWN_STREAM if stream is None else _stream_ptr(stream)
        self._check(self.lib.eincm_batch_minimize_bfgs_graph_host(self._h, thetas.ctypes.data, thetas.shape[1], thetas.shape[2], C.byref(hp),
                                                                  int(maxiter), float(gtol), None if act is None else act.ctypes.data,
                                                                  C.cast(res, C.c_void_p), st))
        return thetas, list(res)

    def solve_launches(self) -> int:
        return int(self.lib.eincm_batch_solve_launches(self._h))

    KERNELS = ('k_splat', 'k_image_stats', 'k_image_grad', 'k_backward_events', 'k_theta_grad')

    def set_timing(self, enabled: bool):
        self._check(self.lib.eincm_batch_set_timing(self._h, 1 if enabled else 0))

    def get_timing(self) -> Dict[str, Tuple[float, int]]:
        ms = (C.c_double * 5)()
        cnt = (C.c_int64 * 5)()
        self._check(self.lib.eincm_batch_get_timing(self._h, ms, cnt))
        return {k: (ms[i], int(cnt[i])) for i, k in enumerate(self.KERNELS) if cnt[i] > 0}


class _DevView:
    """Exposes a plan-owned device buffer through ``__cuda_array_interface__`` so that torch can alias it."""

    def __init__(self, ptr: int, shape: Tuple[int, ...], typestr: str, owner):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr, 'data': (int(ptr), False), 'version': 2}
        self._owner = owner


class Plan:
    """One ``eincm_plan``: all device state for windows of up to ``max_events`` events and ``max_refs`` reference
    times on an ``H x W`` sensor.  Methods mirror the C entry points one to one."""

    def __init__(self, sensor_size: Tuple[int, int], max_events: int, max_refs: int = 8, device: Optional[int] = None,
                 flags: int = 0):
        torch = _torch()
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise EincmError(EINCM_ECUDA, 'no CUDA device: the EINCM objective runs only on a B200 (no CPU fallback)')
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.H, self.W = int(sensor_size[0]), int(sensor_size[1])
        self.max_events, self.max_refs, self.flags = int(max_events), int(max_refs), int(flags)
        h = C.c_void_p()
        rc = self.lib.eincm_plan_create(C.byref(h), self.device, self.H, self.W, self.max_events, self.max_refs, self.flags)
        if rc != EINCM_OK:
            raise EincmError(rc, (self.lib.eincm_last_error(None) or b'').decode())
        self._h = h
        self.n_events = 0
        self.n_refs = 0
        self._keep = []      # operand tensors of the last call (kept alive until the next one)
        with torch.cuda.device(self.device):
            self._out = torch.zeros(2, dtype=torch.float64, device='cuda')

    # -- helpers ----------------------------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != EINCM_OK:
            raise EincmError(rc, (self.lib.eincm_last_error(self._h) or b'').decode())

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            self.lib.eincm_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _dev(self, a, dtype):
        """numpy / torch operand -> contiguous CUDA tensor of ``dtype`` on this plan's device."""
        torch = _torch()
        if isinstance(a, torch.Tensor):
            t = a
        else:
            arr = np.ascontiguousarray(np.asarray(a))
            if not arr.flags.writeable:          # a read-only mapping (eincm_b200.shards): only read here, on its way to the device
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter('ignore')
                    t = torch.from_numpy(arr)
            else:
                t = torch.from_numpy(arr)
        return t.to(device=f'cuda:{self.device}', dtype=dtype).contiguous()

    # -- window -----------------------------------------------------------------------------------------------
    def set_window(self, xs, ys, ts, edges, edge_ts, stream=None):
        """``eincm_plan_set_window`` (replaces ``MultipleLevelEINCMSolver.set_datasample``, reference
        src/eincm/solver.py:185-194)."""
        torch = _torch()
        xs_d, ys_d = self._dev(xs, torch.int16), self._dev(ys, torch.int16)
        ts_d, edges_d = self._dev(ts, torch.float64), self._dev(edges, torch.float64)
        ets = np.ascontiguousarray(np.asarray(edge_ts.cpu() if isinstance(edge_ts, torch.Tensor) else edge_ts, dtype=np.float64))
        n, R = int(xs_d.numel()), int(ets.shape[0])
        if ys_d.numel() != n or ts_d.numel() != n:
            raise EincmError(EINCM_EINVAL, 'xs, ys, ts must have the same length')
        if tuple(edges_d.shape) != (R, self.H, self.W):
            raise EincmError(EINCM_EINVAL, f'edges must have shape {(R, self.H, self.W)}, got {tuple(edges_d.shape)}')
        self._check(self.lib.eincm_plan_set_window(self._h, xs_d.data_ptr(), ys_d.data_ptr(), ts_d.data_ptr(), n,
                                                   edges_d.data_ptr(), ets.ctypes.data_as(C.POINTER(C.c_double)), R,
                                                   _stream_ptr(stream)))
        self.n_events, self.n_refs = n, R

    def set_event_split(self, rank: int, world: int):
        self._check(self.lib.eincm_plan_set_event_split(self._h, int(rank), int(world)))

    def set_split_fixed_point(self, on: bool = True):
        """The caller all-reduces ``iwe_fix()`` (int64) instead of ``iwe()`` (float64) between forward_events and backward."""
        self._check(self.lib.eincm_plan_set_split_fixed_point(self._h, 1 if on else 0))

    def iwe_fix(self):
        """(R, H, W) int64 CUDA tensor aliasing the fixed-point images of warped events (2^21 * 2 pi * value)."""
        return self._view(self.lib.eincm_iwe_fix_ptr(self._h), (self.n_refs, self.H, self.W), '<i8')

    # -- event split with peer access (fused splat + all-reduce over NVLink) ---------------------------------------
    IPC_HANDLE_BYTES = 64

    def ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(self.IPC_HANDLE_BYTES)
        self._check(self.lib.eincm_plan_ipc_handle(self._h, buf, self.IPC_HANDLE_BYTES))
        return buf.raw

    def set_peers(self, handles: Sequence[bytes]):
        """``handles``: the ``ipc_handle()`` of every rank of the split, in rank order (this rank's own entry is ignored)."""
        blob = b''.join(handles)
        self._check(self.lib.eincm_plan_set_peers(self._h, blob, len(handles)))

    def set_peer_pointers(self, plans: Sequence['Plan']):
        """In-process form: the plans of all ranks of the split live in this process (same device or peer-accessible)."""
        ptrs = (C.c_void_p * len(plans))(*[p.lib.eincm_iwe_fix_ptr(p._h) for p in plans])
        self._check(self.lib.eincm_plan_set_peer_pointers(self._h, ptrs, len(plans)))

    def split_prepare(self, stream=None):
        self._check(self.lib.eincm_split_prepare(self._h, _stream_ptr(stream)))

    def split_window_images(self, stream=None):
        self._check(self.lib.eincm_split_window_images(self._h, _stream_ptr(stream)))

    def window_finalize(self, stream=None):
        self._check(self.lib.eincm_window_finalize(self._h, _stream_ptr(stream)))

    # -- evaluation, device operands (asynchronous) ---------------------------------------------------------------
    def value_and_grad_device(self, theta_d, hp: HParams, loss_out_d, grad_out_d=None, stream=None):
        h, w = int(theta_d.shape[0]), int(theta_d.shape[1])
        self._keep = [theta_d, loss_out_d, grad_out_d]
        self._check(self.lib.eincm_value_and_grad(self._h, theta_d.data_ptr(), h, w, C.byref(hp), loss_out_d.data_ptr(),
                                                  grad_out_d.data_ptr() if grad_out_d is not None else None,
                                                  _stream_ptr(stream)))

    def handover_value_and_grad_device(self, alpha_handover: float, prev_d, theta_d, hp: HParams, loss_out_d, dalpha_out_d=None,
                                       stream=None):
        h, w = int(theta_d.shape[0]), int(theta_d.shape[1])
        self._keep = [prev_d, theta_d, loss_out_d, dalpha_out_d]
        self._check(self.lib.eincm_handover_value_and_grad(self._h, float(alpha_handover), prev_d.data_ptr(), theta_d.data_ptr(),
                                                           h, w, C.byref(hp), loss_out_d.data_ptr(),
                                                           dalpha_out_d.data_ptr() if dalpha_out_d is not None else None,
                                                           _stream_ptr(stream)))

    def forward_events(self, theta_d, hp: HParams, stream=None):
        h, w = int(theta_d.shape[0]), int(theta_d.shape[1])
        self._keep = [theta_d]
        self._check(self.lib.eincm_forward_events(self._h, theta_d.data_ptr(), h, w, C.byref(hp), _stream_ptr(stream)))

    def backward(self, hp: HParams, loss_out_d, grad_out_d=None, stream=None):
        self._keep += [loss_out_d, grad_out_d]
        self._check(self.lib.eincm_backward(self._h, C.byref(hp), loss_out_d.data_ptr(),
                                            grad_out_d.data_ptr() if grad_out_d is not None else None, _stream_ptr(stream)))

    # -- evaluation, host operands (synchronous: what jaxopt's scipy_fun does per line-search step) ------------------
    def value_and_grad_host(self, theta: np.ndarray, hp: HParams, want_grad: bool = True, stream=None):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        if theta.ndim != 3 or theta.shape[2] != 2:
            raise EincmError(EINCM_EINVAL, f'theta must have shape (h, w, 2), got {theta.shape}')
        h, w = theta.shape[:2]
        loss = C.c_double()
        grad = np.empty_like(theta) if want_grad else None
        self._check(self.lib.eincm_value_and_grad_host(self._h, theta.ctypes.data, h, w, C.byref(hp), C.byref(loss),
                                                       grad.ctypes.data if want_grad else None, _stream_ptr(stream)))
        return loss.value, grad

    def handover_value_and_grad_host(self, alpha_handover: float, prev_theta: np.ndarray, theta: np.ndarray, hp: HParams,
                                     want_grad: bool = True, stream=None):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        prev_theta = np.ascontiguousarray(prev_theta, dtype=np.float64)
        if theta.shape != prev_theta.shape or theta.ndim != 3 or theta.shape[2] != 2:
            raise EincmError(EINCM_EINVAL, 'prev_theta and theta must both have shape (h, w, 2)')
        h, w = theta.shape[:2]
        loss, da = C.c_double(), C.c_double()
        self._check(self.lib.eincm_handover_value_and_grad_host(self._h, float(alpha_handover), prev_theta.ctypes.data,
                                                                theta.ctypes.data, h, w, C.byref(hp), C.byref(loss),
                                                                C.byref(da) if want_grad else None, _stream_ptr(stream)))
        return loss.value, (da.value if want_grad else None)

    # -- native optimizers (no Python between evaluations; ctypes releases the GIL for the whole solve) ---------------
    def minimize_bfgs_host(self, theta0: np.ndarray, hp: HParams, maxiter: int, gtol: float, own_stream: bool = False, stream=None):
        """``ScipyMinimize(method='BFGS').run`` on this plan's objective: returns ``(theta, OptResult)``."""
        theta = np.array(theta0, dtype=np.float64, order='C', copy=True)
        if theta.ndim != 3 or theta.shape[2] != 2:
            raise EincmError(EINCM_EINVAL, f'theta must have shape (h, w, 2), got {theta.shape}')
        res = OptResult()
        st = OWN_STREAM if own_stream else _stream_ptr(stream)
        self._check(self.lib.eincm_minimize_bfgs_host(self._h, theta.ctypes.data, theta.shape[0], theta.shape[1], C.byref(hp),
                                                      int(maxiter), float(gtol), C.byref(res), st))
        return theta, res

    def minimize_bfgs_graph_host(self, theta0: np.ndarray, hp: HParams, maxiter: int, gtol: float, stream=None):
        """The same level solve with the loop on the device (one CUDA graph, no host round trip per evaluation): returns
        ``(theta, OptResult)``.  ``stream`` None: the plan's own stream."""
        theta = np.array(theta0, dtype=np.float64, order='C', copy=True)
        if theta.ndim != 3 or theta.shape[2] != 2:
            raise EincmError(EINCM_EINVAL, f'theta must have shape (h, w, 2), got {theta.shape}')
        res = OptResult()
        st = OWN_STREAM if stream is None else _stream_ptr(stream)
        self._check(self.lib.eincm_minimize_bfgs_graph_host(self._h, theta.ctypes.data, theta.shape[0], theta.shape[1], C.byref(hp),
                                                            int(maxiter), float(gtol), C.byref(res), st))
        return theta, res

    def minimize_handover_host(self, alpha0: float, bounds, prev_theta: np.ndarray, theta: np.ndarray, hp: HParams, maxiter: int,
                               pgtol: float, own_stream: bool = False, stream=None):
        """``ScipyBoundedMinimize(method='L-BFGS-B').run`` on the scalar handover weight: returns ``(alpha, OptResult)``."""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        prev_theta = np.ascontiguousarray(prev_theta, dtype=np.float64)
        if theta.shape != prev_theta.shape or theta.ndim != 3 or theta.shape[2] != 2:
            raise EincmError(EINCM_EINVAL, 'prev_theta and theta must both have shape (h, w, 2)')
        a = C.c_double(float(alpha0))
        res = OptResult()
        st = OWN_STREAM if own_stream else _stream_ptr(stream)
        self._check(self.lib.eincm_minimize_handover_host(self._h, C.byref(a), float(bounds[0]), float(bounds[1]), prev_theta.ctypes.data,
                                                          theta.ctypes.data, theta.shape[0], theta.shape[1], C.byref(hp), int(maxiter),
                                                          float(pgtol), C.byref(res), st))
        return a.value, res

    def value_and_grad_stateless_host(self, theta, xs, ys, ts, edges, edge_ts, hp: HParams, stream=None):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        xs = np.ascontiguousarray(xs, dtype=np.int16); ys = np.ascontiguousarray(ys, dtype=np.int16)
        ts = np.ascontiguousarray(ts, dtype=np.float64); edges = np.ascontiguousarray(edges, dtype=np.float64)
        edge_ts = np.ascontiguousarray(edge_ts, dtype=np.float64)
        h, w = theta.shape[:2]
        loss = C.c_double()
        grad = np.empty_like(theta)
        self._check(self.lib.eincm_value_and_grad_stateless_host(
            self._h, theta.ctypes.data, h, w, xs.ctypes.data, ys.ctypes.data, ts.ctypes.data, xs.shape[0], edges.ctypes.data,
            edge_ts.ctypes.data, edge_ts.shape[0], C.byref(hp), C.byref(loss), grad.ctypes.data, _stream_ptr(stream)))
        self.n_events, self.n_refs = int(xs.shape[0]), int(edge_ts.shape[0])
        return loss.value, grad

    # -- read-outs --------------------------------------------------------------------------------------------
    def set_group(self, group: Optional['Group']):
        """Joins (or, with ``None``, leaves) an evaluation group: see ``Group``."""
        self._check(self.lib.eincm_plan_set_group(self._h, group._h if group is not None else None))
        self._group = group               # keeps the group alive as long as the plan refers to it

    # -- evaluation metrics of a solved window (reference src/evaluations/theta_eval.py) ----------------------------
    def evaluate_theta(self, theta, hp: HParams, gt_flow=None, err_eval_event_mask=None, stream=None) -> EvalMetrics:
        torch = _torch()
        th = self._dev(theta, torch.float64)
        gt = self._dev(gt_flow, torch.float64) if gt_flow is not None else None
        em = self._dev(np.asarray(err_eval_event_mask.cpu() if isinstance(err_eval_event_mask, torch.Tensor) else err_eval_event_mask).astype(np.uint8),
                       torch.uint8) if err_eval_event_mask is not None else None
        if gt is not None and tuple(gt.shape) != (self.H, self.W, 2):
            raise EincmError(EINCM_EINVAL, f'gt_flow must have shape {(self.H, self.W, 2)}')
        out = EvalMetrics()
        self._check(self.lib.eincm_evaluate_theta(self._h, th.data_ptr(), int(th.shape[0]), int(th.shape[1]), C.byref(hp),
                                                  gt.data_ptr() if gt is not None else None, em.data_ptr() if em is not None else None,
                                                  C.byref(out), _stream_ptr(stream)))
        return out

    def _view(self, ptr, shape, dtype_str):
        torch = _torch()
        with torch.cuda.device(self.device):
            return torch.as_tensor(_DevView(ptr, shape, dtype_str, self), device=f'cuda:{self.device}')

    def iwe(self):
        """(R, H, W) float64 CUDA tensor aliasing the images of warped events of the last evaluation."""
        return self._view(self.lib.eincm_iwe_ptr(self._h), (self.n_refs, self.H, self.W), '<f8')

    def zero_iwe(self):
        return self._view(self.lib.eincm_zero_iwe_ptr(self._h), (self.H, self.W), '<f8')

    def dldi(self):
        return self._view(self.lib.eincm_dldi_ptr(self._h), (self.n_refs, self.H, self.W), '<f8')

    def theta_full(self):
        """aux 'scaled_theta' (reference src/eincm/losses.py:197)."""
        return self._view(self.lib.eincm_theta_full_ptr(self._h), (self.H, self.W, 2), '<f8')

    def event_mask(self):
        return self._view(self.lib.eincm_mask_ptr(self._h), (self.H, self.W), '|u1')

    def scalars(self, stream=None) -> Dict[str, object]:
        n = S_HEADER + 5 * self.max_refs
        buf = (C.c_double * n)()
        self._check(self.lib.eincm_get_scalars(self._h, buf, n, _stream_ptr(stream)))
        a = np.frombuffer(buf, dtype=np.float64).copy()
        M, R = self.max_refs, self.n_refs
        per = a[S_HEADER:].reshape(5, M)[:, :R]
        return {'final_loss': a[0], 'mean_rel_corr': a[1], 'mean_rel_contrast': a[2], 'mean_rel_iwe_divergence': a[3],
                'theta_total_variation': a[4], 'zero_contrast': a[5], 'zero_iwe_divergence': a[6], 'dalpha_handover': a[7],
                'contrasts': per[0].copy(), 'correlations': per[1].copy(), 'zero_correlations': per[2].copy(),
                'iwe_divergences': per[3].copy(), 'multi_ref_weights': per[4].copy()}

    def host_times(self, reset: bool = False):
        """(seconds launching, seconds waiting, evaluations) of the synchronous host entry points since the last reset."""
        a, b, n = C.c_double(), C.c_double(), C.c_int64()
        self._check(self.lib.eincm_plan_host_times(self._h, C.byref(a), C.byref(b), C.byref(n), int(reset)))
        return a.value, b.value, n.value

    def launch_count(self) -> int:
        return int(self.lib.eincm_plan_launch_count(self._h))

    def set_timing(self, enabled: bool):
        self._check(self.lib.eincm_plan_set_timing(self._h, 1 if enabled else 0))

    def get_timing(self) -> Dict[str, Tuple[float, int]]:
        """{kernel name: (total ms, launches)} of the spans recorded since the last call (synchronises)."""
        cap = 64
        names = C.create_string_buffer(4096)
        ms = (C.c_double * cap)()
        cnt = (C.c_int64 * cap)()
        n = C.c_int()
        self._check(self.lib.eincm_plan_get_timing(self._h, names, 4096, ms, cnt, cap, C.byref(n)))
        ks = names.value.decode().split('\n')[:n.value]
        return {k: (ms[i], int(cnt[i])) for i, k in enumerate(ks)}

    def rounded_pixels(self, ref: int, stream=None) -> Tuple[np.ndarray, np.ndarray]:
        """Bit-exact event->pixel index stream (``Xs_rounded``, reference src/utils/event_utils.py:33) of reference
        ``ref`` for the last evaluated theta, in the original event order."""
        torch = _torch()
        with torch.cuda.device(self.device):
            cols = torch.empty(self.n_events, dtype=torch.int32, device='cuda')
            rows = torch.empty(self.n_events, dtype=torch.int32, device='cuda')
        self._check(self.lib.eincm_debug_rounded_pixels(self._h, int(ref), cols.data_ptr(), rows.data_ptr(), _stream_ptr(stream)))
        return cols.cpu().numpy(), rows.cpu().numpy()
