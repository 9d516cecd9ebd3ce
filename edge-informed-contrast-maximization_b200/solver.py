"""placeholder - filled in below"""
import math


def growing_maxiters(miniter, maxiter, n_pyr_lvls, order):
    """reference src/experiments/e00/exp_mgr.py:169-187."""
    out = {}
    for lvl in range(n_pyr_lvls):
        p = lvl / (n_pyr_lvls - 1)
        out[lvl] = int(math.ceil(miniter * p ** order + maxiter * (1 - p) ** order))
    return out
