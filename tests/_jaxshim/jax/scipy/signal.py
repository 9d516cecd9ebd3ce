"""jax.scipy.signal.convolve(mode='same') of the stand-in: true 2-D convolution, zero padding, centred crop.  Taps are accumulated in
the row-major order of the flipped kernel, zero taps skipped (JAX_SHIM_CONV=pairs: antisymmetric tap pairs are differenced first, the
canonical order of oracle/eincm_oracle.py; the order is only observable through exact zeros - regularizers.py:26-29)."""
import os

import torch as _torch

from .._array import _unwrap, _wrap


def convolve(a, k, mode='full', method='auto', precision=None):
    assert mode == 'same'
    a, k = _unwrap(a).to(_torch.float64), _unwrap(k).to(_torch.float64)
    assert a.dim() == 2 and k.dim() == 2 and k.shape[0] % 2 == 1 and k.shape[1] % 2 == 1
    kh, kw = k.shape
    H, W = a.shape
    p = _torch.nn.functional.pad(a, (kw // 2, kw // 2, kh // 2, kh // 2))
    taps = []
    for i in range(kh):
        for j in range(kw):
            c = float(k[kh - 1 - i, kw - 1 - j])            # out[r, c] = sum_ij k[kh-1-i, kw-1-j] * a[r + i - kh//2, c + j - kw//2]
            if c != 0.0:
                taps.append((c, p[i:i + H, j:j + W]))
    if os.environ.get('JAX_SHIM_CONV') == 'pairs':
        out, used = None, [False] * len(taps)
        for m, (c, v) in enumerate(taps):
            if used[m]:
                continue
            used[m] = True
            term = None
            for n in range(m + 1, len(taps)):
                if not used[n] and taps[n][0] == -c:
                    used[n] = True
                    term = c * (v - taps[n][1])
                    break
            if term is None:
                term = c * v
            out = term if out is None else out + term
        return _wrap(out)
    out = None
    for c, v in taps:
        out = c * v if out is None else out + c * v
    return _wrap(out)
