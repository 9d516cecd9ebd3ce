"""Host mirror of the event-ingest steps of the reference's DSEC loader and experiment manager, backed by the CUDA library
(SURVEY.md 8f rank 4): ``rectify_events`` (src/dataloaders/dsec_loader.py:145-170), the fixed-N window rule of ``get_sample``
(dsec_loader.py:293-311) and the time normalisation of ``stage_datasample`` (src/experiments/e00/exp_mgr.py:313-321).  Reading the
h5 files stays with the reference (h5py); these functions take the arrays it has read.

There is no CPU fallback: without the built library the import of ``eincm_b200.plan`` fails.
"""
import ctypes as C

import numpy as np

from . import plan as _plan

__all__ = ['rectify_events', 'crop_events', 'window_event_range', 'normalize_times', 'stage_window_events']


def _dev(a, dtype, device=None):
    import torch
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a)))
    dev = f'cuda:{torch.cuda.current_device() if device is None else device}'
    return t.to(device=dev, dtype=dtype).contiguous()


def rectify_events(x, y, t, p, rectify_map, height, width):
    """dsec_loader.py:145-170 on the device: returns ``(x, y, t, p)`` CUDA tensors (int16, int16, int64, bool) of the events whose
    rectified pixel lies inside the sensor, in the original order."""
    import torch
    rm = _dev(rectify_map, torch.float32)
    if tuple(rm.shape) != (int(height), int(width), 2):
        raise _plan.EincmError(_plan.EINCM_EINVAL, f'rectify_map must have shape ({height}, {width}, 2), not {tuple(rm.shape)}')
    dx, dy = _dev(x, torch.int16), _dev(y, torch.int16)
    dt = _dev(np.asarray(t).astype(np.int64) if not isinstance(t, torch.Tensor) else t, torch.int64)
    dp = _dev(np.asarray(p).astype(np.uint8) if not isinstance(p, torch.Tensor) else p, torch.uint8)
    n = int(dx.numel())
    if not (dy.numel() == n and dt.numel() == n and dp.numel() == n):
        raise _plan.EincmError(_plan.EINCM_EINVAL, 'x, y, t, p must have the same length')
    lib = _plan.load_library()
    ox, oy = torch.empty(n, dtype=torch.int16, device=dx.device), torch.empty(n, dtype=torch.int16, device=dx.device)
    ot, op = torch.empty(n, dtype=torch.int64, device=dx.device), torch.empty(n, dtype=torch.uint8, device=dx.device)
    wsb = int(lib.eincm_rectify_workspace_bytes(n))
    ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device=dx.device)
    n_out = C.c_int64(0)
    rc = lib.eincm_rectify_events(dx.device.index, dx.data_ptr(), dy.data_ptr(), dt.data_ptr(), dp.data_ptr(), n, rm.data_ptr(),
                                  int(height), int(width), ox.data_ptr(), oy.data_ptr(), ot.data_ptr(), op.data_ptr(), C.byref(n_out),
                                  ws.data_ptr(), wsb, _plan._stream_ptr(None))
    if rc != 0:
        raise _plan.EincmError(rc, 'eincm_rectify_events failed')
    k = int(n_out.value)
    return ox[:k], oy[:k], ot[:k], op[:k].bool()


def crop_events(xs, ys, ts, ps, x_offset=5, y_offset=2, height=256, width=336, in_height=260, in_width=346):
    """The event crop of the reference's MVSEC loader (src/dataloaders/mvsec_loader.py:113-129): ``xs - 5, ys - 2``, events outside the
    ``width x height`` window dropped, order kept.  Runs through ``eincm_rectify_events`` with a shift map whose out-of-window pixels
    point outside the sensor.  ``ts`` may be float64 (MVSEC) or int64; returned with the dtype it came with."""
    import torch
    dev = f'cuda:{torch.cuda.current_device()}'
    xx = torch.arange(in_width, device=dev, dtype=torch.float32) - float(x_offset)
    yy = torch.arange(in_height, device=dev, dtype=torch.float32) - float(y_offset)
    xx = torch.where((xx >= 0) & (xx < width), xx, torch.full_like(xx, -1.0))
    yy = torch.where((yy >= 0) & (yy < height), yy, torch.full_like(yy, -1.0))
    rmap = torch.stack([xx[None, :].expand(in_height, in_width), yy[:, None].expand(in_height, in_width)], dim=-1).contiguous()
    t = ts if isinstance(ts, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(ts)))
    is_float = t.dtype == torch.float64
    t_bits = t.contiguous().view(torch.int64) if is_float else t.to(torch.int64)
    x, y, tt, p = rectify_events(xs, ys, t_bits, ps, rmap, in_height, in_width)
    return x, y, (tt.view(torch.float64) if is_float else tt), p


def window_event_range(idx_evt_start, idx_evt_end, n_total, des_n_events=None, prefer_latest_events=False):
    """dsec_loader.py:293-311: ``(start, end, n_event_deficiency)`` of the fixed-N window."""
    lib = _plan.load_library()
    a, b, d = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    rc = lib.eincm_window_event_range(int(idx_evt_start), int(idx_evt_end), int(n_total), int(des_n_events or 0),
                                      1 if prefer_latest_events else 0, C.byref(a), C.byref(b), C.byref(d))
    if rc != 0:
        raise _plan.EincmError(rc, 'eincm_window_event_range: bad index range')
    return int(a.value), int(b.value), int(d.value)


def normalize_times(ts_us, start_time, end_time):
    """exp_mgr.py:313-321 (time_scaler = 1): float64 CUDA tensor ``(ts - start) / (end - start + eps)``."""
    import torch
    dt = _dev(np.asarray(ts_us).astype(np.int64) if not isinstance(ts_us, torch.Tensor) else ts_us, torch.int64)
    out = torch.empty(dt.numel(), dtype=torch.float64, device=dt.device)
    lib = _plan.load_library()
    rc = lib.eincm_normalize_times(dt.device.index, dt.data_ptr(), int(dt.numel()), int(start_time), int(end_time), out.data_ptr(),
                                   _plan._stream_ptr(None))
    if rc != 0:
        raise _plan.EincmError(rc, 'eincm_normalize_times failed')
    return out


def stage_window_events(x, y, t_us, idx_evt_start, idx_evt_end, eval_ts_us, des_n_events=None, prefer_latest_events=False, t_offset=0):
    """One window of an (already rectified) device-resident stream: the slice ``get_sample`` takes and the normalised timestamps
    ``stage_datasample`` derives - the ``(xs, ys, ts)`` operands of ``loss_func`` / ``Plan.set_window`` as CUDA tensors."""
    a, b, deficiency = window_event_range(idx_evt_start, idx_evt_end, int(x.shape[0]), des_n_events, prefer_latest_events)
    ts = normalize_times(t_us[a:b] + int(t_offset), int(eval_ts_us[0]), int(eval_ts_us[1]))
    return x[a:b], y[a:b], ts, deficiency
