"""jax.scipy.stats.multivariate_normal.pdf of the stand-in (closed form, any positive-definite covariance)."""
import math

import torch as _torch

from .._array import _unwrap, _wrap


class multivariate_normal:
    @staticmethod
    def pdf(x, mean, cov):
        x, mean, cov = (_unwrap(v).to(_torch.float64) for v in (x, mean, cov))
        d = x - mean
        k = d.shape[-1]
        q = ((d @ _torch.linalg.inv(cov)) * d).sum(-1)
        return _wrap(_torch.exp(-0.5 * q) / math.sqrt((2.0 * math.pi) ** k * float(_torch.linalg.det(cov))))
