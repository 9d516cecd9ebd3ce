"""CPU restatement of the reference's event ingest between the DSEC h5 stream and loss_func's operands
- TEST INFRASTRUCTURE, not product.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.

The reference code for this step is plain NumPy / jax.numpy (no third-party algorithm involved), restated line by line:
  rectify_events        src/dataloaders/dsec_loader.py:145-170
  get_sample (range)    src/dataloaders/dsec_loader.py:293-311
  time normalisation    src/experiments/e00/exp_mgr.py:313-321
Pinned by hand-derived known answers (tests/test_ingest_oracle.py) and, since round 2, by the reference's own loader methods
(DSECDataLoader.rectify_events / precompute_eval_event_indices / get_sample, MVSECDataLoader.load_left_data) executed on in-memory arrays:
the loader modules import over empty h5py / imageio stand-ins (tests/_jaxshim) and the objects are made without their file-opening
constructors (tests/test_reference_source.py::test_ingest_oracle_matches_reference_loaders_live, build container only).  The time
normalisation (exp_mgr.py, needs hydra / omegaconf / matplotlib to import) stays pinned by known answers only.
"""
import sys

import numpy as np


def rectify_events(x, y, t, p, rectify_map, height, width):
    """dsec_loader.py:145-170"""
    assert rectify_map.shape == (height, width, 2), rectify_map.shape
    rect_event_coords = rectify_map[y, x]
    rec_x, rec_y = rect_event_coords.T
    rec_x = np.round(rec_x).astype('int16')
    rec_y = np.round(rec_y).astype('int16')
    rec_x_mask = np.logical_and(rec_x >= 0, rec_x < width)
    rec_y_mask = np.logical_and(rec_y >= 0, rec_y < height)
    m = np.logical_and(rec_x_mask, rec_y_mask)
    return rec_x[m], rec_y[m], t[m], p[m]


def window_event_range(idx_evt_start, idx_evt_end, n_total, des_n_events, prefer_latest_events=False):
    """dsec_loader.py:293-311: returns (start, end, n_event_deficiency)"""
    n_event_deficiency = 0
    if des_n_events is not None:
        n_event_deficiency = des_n_events - (idx_evt_end - idx_evt_start)
        if n_event_deficiency > 0:
            idx_evt_start -= np.ceil(n_event_deficiency / 2).astype(int)
            idx_evt_end += np.floor(n_event_deficiency / 2).astype(int)
            idx_evt_start = max(0, idx_evt_start)
            idx_evt_end = min(idx_evt_end, n_total)
        elif n_event_deficiency < 0:
            if prefer_latest_events:
                idx_evt_start = idx_evt_end - des_n_events
            else:
                idx_evt_end = idx_evt_start + des_n_events
    return int(idx_evt_start), int(idx_evt_end), int(n_event_deficiency)


def normalize_times(ts_us, start_time, end_time):
    """exp_mgr.py:313-321 with time_scaler = 1: ts is uint64, start / end int64 -> float64 arithmetic (NumPy and jax.numpy promote
    uint64 with int64 to float64)."""
    ts = np.asarray(ts_us).astype(np.uint64)
    return (ts - np.int64(start_time)) / (np.int64(end_time) - np.int64(start_time) + sys.float_info.epsilon)


def crop_events(xs, ys, ts, ps, x_offset=5, y_offset=2, height=256, width=336):
    """src/dataloaders/mvsec_loader.py:113-129"""
    xs = xs - x_offset
    ys = ys - y_offset
    m = (xs >= 0) & (xs < width) & (ys >= 0) & (ys < height)
    return xs[m].astype('int16'), ys[m].astype('int16'), ts[m].astype('float64'), ps[m].astype('bool')
