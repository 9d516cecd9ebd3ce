"""ctypes binding of oracle/eincm_oracle_c.c (TEST INFRASTRUCTURE: checker / timed CPU baseline only, never a product path)."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libeincm_oracle_c.so')
_lib = None


def available() -> bool:
    return os.path.exists(_SO)


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(_SO)
        d, i16 = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int16)
        lib.eincm_oracle_c_value_and_grad.restype = ctypes.c_int
        lib.eincm_oracle_c_value_and_grad.argtypes = [d, ctypes.c_int, ctypes.c_int, i16, i16, d, ctypes.c_int64, d, d, ctypes.c_int,
                                                      ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                                      ctypes.c_double, ctypes.c_int, ctypes.c_int, d, d, d]
        lib.eincm_oracle_c_num_threads.restype = ctypes.c_int
        lib.eincm_oracle_c_set_num_threads.argtypes = [ctypes.c_int]
        _lib = lib
    return _lib


def num_threads() -> int:
    return int(_load().eincm_oracle_c_num_threads())


def set_num_threads(n: int) -> None:
    _load().eincm_oracle_c_set_num_threads(int(n))


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def value_and_grad_raw(theta, xs, ys, ts, edges, edge_ts, alpha, beta, gamma, delta, cur_pyr_lvl, sensor_size, wrap_negative=True,
                       want_grad=True, want_iwes=False):
    lib = _load()
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    xs = np.ascontiguousarray(xs, dtype=np.int16)
    ys = np.ascontiguousarray(ys, dtype=np.int16)
    ts = np.ascontiguousarray(ts, dtype=np.float64)
    edges = np.ascontiguousarray(edges, dtype=np.float64)
    edge_ts = np.ascontiguousarray(edge_ts, dtype=np.float64)
    H, W = sensor_size
    R = len(edge_ts)
    h, w = theta.shape[:2]
    loss = np.zeros(1)
    grad = np.zeros_like(theta) if want_grad else None
    iwes = np.zeros((R, H, W)) if want_iwes else None
    rc = lib.eincm_oracle_c_value_and_grad(_p(theta, ctypes.c_double), h, w, _p(xs, ctypes.c_int16), _p(ys, ctypes.c_int16),
                                           _p(ts, ctypes.c_double), len(xs), _p(edges, ctypes.c_double), _p(edge_ts, ctypes.c_double), R,
                                           H, W, alpha, beta, gamma, delta, int(cur_pyr_lvl), int(bool(wrap_negative)),
                                           _p(loss, ctypes.c_double), _p(grad, ctypes.c_double) if want_grad else None,
                                           _p(iwes, ctypes.c_double) if want_iwes else None)
    if rc != 0:
        raise RuntimeError(f'eincm_oracle_c_value_and_grad failed ({rc})')
    return float(loss[0]), grad, iwes


def value_and_grad(theta, win, hp, lvl):
    """bench.py's CPU-baseline signature: (theta, Window, hparams dict, cur_pyr_lvl) -> (loss, grad)."""
    loss, grad, _ = value_and_grad_raw(theta, win.xs, win.ys, win.ts, win.edges, win.edge_ts, hp['alpha'], hp['beta'], hp['gamma'],
                                       hp['delta'], lvl, win.sensor_size)
    return loss, grad
