"""Host mirror of the reference's evaluation entry points (``src/evaluations/flow_eval.py``, ``src/evaluations/theta_eval.py``),
backed by the CUDA library: same names, argument meaning and result keys, so the reference's evaluation loop
(``src/experiments/e00/exp_mgr.py``) can call these instead.  SURVEY.md 8f rank 2.

There is no CPU fallback: without the built library the import of ``eincm_b200.plan`` fails.
"""
import ctypes as C
import time
from typing import Dict, Optional

import numpy as np

from . import plan as _plan
from . import losses as _losses

__all__ = ['sparse_flow_error', 'evaluate_theta_array']


def _cuda(a, dtype, device=None):
    import torch
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(a)))
    dev = f'cuda:{torch.cuda.current_device() if device is None else device}'
    return t.to(device=dev, dtype=dtype).contiguous()


def sparse_flow_error(pred_flow, gt_flow, event_mask=None) -> Dict:
    """reference src/evaluations/flow_eval.py:14-75: ``{'errors': {AEE, AREE, A1PE..A20PE}, 'counts': {n_ee, n_pred, n_gt}}``."""
    import torch
    pf, gf = _cuda(pred_flow, torch.float64), _cuda(gt_flow, torch.float64)
    if pf.dim() != 3 or pf.shape[2] != 2 or tuple(pf.shape) != tuple(gf.shape):
        raise _plan.EincmError(_plan.EINCM_EINVAL, 'pred_flow and gt_flow must both have shape (H, W, 2)')
    em = None
    if event_mask is not None:
        em = _cuda(np.asarray(event_mask.cpu() if isinstance(event_mask, torch.Tensor) else event_mask).astype(np.uint8), torch.uint8)
        if tuple(em.shape) != tuple(pf.shape[:2]):
            raise _plan.EincmError(_plan.EINCM_EINVAL, 'event_mask must have shape (H, W)')
    lib = _plan.load_library()
    out = _plan.FlowErrors()
    rc = lib.eincm_sparse_flow_error(pf.device.index, int(pf.shape[0]), int(pf.shape[1]), pf.data_ptr(), gf.data_ptr(),
                                     em.data_ptr() if em is not None else None, C.byref(out), _plan._stream_ptr(None))
    if rc != 0:
        raise _plan.EincmError(rc, 'eincm_sparse_flow_error failed')
    return out.as_dict()


def evaluate_theta_array(theta_array, eval_xs, eval_ys, eval_ts, edges, edge_ts, gt_flow, alpha, beta, gamma, delta, sensor_size,
                         err_eval_event_mask=None):
    """reference src/evaluations/theta_eval.py:14-95.  Returns ``(time_str, eval_str, evals, loss_obj)`` like the reference;
    ``loss_obj`` carries the per-reference objectives the evaluation derives its numbers from (the (R, N) warped coordinates of
    the reference's dict stay on the device: use ``Plan.rounded_pixels`` / ``Plan.iwe`` for them)."""
    sensor_size = (int(sensor_size[0]), int(sensor_size[1]))
    pl = _losses._cache.plan_for(eval_xs, eval_ys, eval_ts, edges, edge_ts, sensor_size)
    hp = _plan.make_hparams(alpha, beta, gamma, delta, 0)
    m = pl.evaluate_theta(theta_array, hp, gt_flow=gt_flow, err_eval_event_mask=err_eval_event_mask)
    R = int(m.n_refs)
    arr = lambda a: np.array(list(a)[:R], dtype=np.float64)
    evals = {}
    acc_eval_str = ''
    if gt_flow is not None:
        fe = m.flow.as_dict()
        evals.update(fe['errors']); evals.update(fe['counts'])                      # theta_eval.py:61-62
        evals['n_pixels'] = int(m.n_pixels)
        e, c = fe['errors'], fe['counts']
        acc_eval_str = (f', AEE(↓): {e["AEE"]:8.6f}, AREE(↓): {e["AREE"]:8.6f}, A1PE(↓): {e["A1PE"]:8.6f}, A2PE(↓): {e["A2PE"]:8.6f}, '
                        f'A3PE(↓): {e["A3PE"]:8.6f}, A5PE(↓): {e["A5PE"]:8.6f}, A10PE(↓): {e["A10PE"]:8.6f}, A20PE(↓): {e["A20PE"]:8.6f}, '
                        f'| n_pixels:{int(m.n_pixels):,}, n_gt_mask:{c["n_gt"]:,}, n_event_mask:{c["n_pred"]:,}, n_ee: {c["n_ee"]:,}\n')
    time_str = f'[{time.strftime("%Y-%m-%d %H:%M:%S")}]'
    eval_str = (f'total_loss(↓): {m.loss:8.6f}, iwe_var(↑): {m.iwe_var:8.6f}, mean_rel_contrast(↑): {m.mean_rel_contrast:8.6f}, '
                f'mean_rel_corr(↑): {m.mean_rel_corr:8.6f}, theta_tot_var(↓): {m.theta_tot_var:8.6f}, theta_div(↓): {m.theta_div:8.6f}, '
                f'mean_rel_iwe_div(↓): {m.mean_rel_iwe_div:8.6f}, FWL(↑): {m.fwl:8.6f}{acc_eval_str}')
    loss_obj = {'rel_iwe_divergences': arr(m.rel_iwe_divergences), 'rel_contrasts': arr(m.rel_contrasts),
                'rel_correlations': arr(m.rel_correlations), 'flow_warp_losses': arr(m.flow_warp_losses),
                'multi_ref_weights': arr(m.multi_ref_weights), 'theta_total_variation': m.theta_tot_var, 'theta_divergence': m.theta_div}
    evals.update({'loss': m.loss, 'iwe_var': m.iwe_var, 'mean_rel_contrast': m.mean_rel_contrast, 'mean_rel_corr': m.mean_rel_corr,
                  'theta_tot_var': m.theta_tot_var, 'theta_div': m.theta_div, 'fwl': m.fwl, 'mean_rel_iwe_div': m.mean_rel_iwe_div,
                  'rel_iwe_divergences': loss_obj['rel_iwe_divergences'], 'rel_contrasts': loss_obj['rel_contrasts'],
                  'rel_correlations': loss_obj['rel_correlations'], 'flow_warp_losses': loss_obj['flow_warp_losses'],
                  'multi_ref_weights': loss_obj['multi_ref_weights']})
    return time_str, eval_str, evals, loss_obj
