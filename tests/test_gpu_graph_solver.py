"""The device-side solve loop (eincm_minimize_bfgs_graph_host: one CUDA graph per pyramid level, WHILE conditional node, k_bfgs_step)
against the host-driven native BFGS on the same windows: the same line-search machines and the same objective bits, so the same
schedule - exactly for two parameters (sums of two terms do not depend on their order), to rounding beyond."""
import numpy as np
import pytest

import eincm_b200.synth as S

pytestmark = pytest.mark.gpu


def _plan(P, win):
    p = P.Plan(win.sensor_size, max_events=len(win.xs), max_refs=max(3, len(win.edge_ts)))
    p.set_window(*win.args())
    return p


def test_one_by_one_level_is_identical_to_the_host_loop():
    from eincm_b200 import plan as P
    win = S.make_workload('tiny', seed=4)
    hp = P.make_hparams(win.hparams['alpha'], win.hparams['beta'], 0.0, 0.0, 4)
    p = _plan(P, win)
    try:
        th0 = np.zeros((1, 1, 2))
        ta, ra = p.minimize_bfgs_host(th0, hp, 8, 1e-7, own_stream=True)
        tb, rb = p.minimize_bfgs_graph_host(th0, hp, 8, 1e-7)
        assert (rb.status, rb.nit, rb.nfev) == (ra.status, ra.nit, ra.nfev)
        assert rb.fun == ra.fun
        np.testing.assert_allclose(tb, ta, rtol=1e-15, atol=0)                      # the device code may contract a * b + c into an FMA: an ulp or two
        # again on the same window: the instantiated graph is re-used; from another start
        th1 = np.full((1, 1, 2), 0.7)
        ta, ra = p.minimize_bfgs_host(th1, hp, 8, 1e-7, own_stream=True)
        tb, rb = p.minimize_bfgs_graph_host(th1, hp, 8, 1e-7)
        assert (rb.status, rb.nit, rb.nfev) == (ra.status, ra.nit, ra.nfev) and rb.fun == ra.fun
        # maxiter 0 and an already converged start
        _, r0 = p.minimize_bfgs_graph_host(th1, hp, 0, 1e-7)
        assert (r0.status, r0.nit, r0.nfev) == (1, 0, 1)
        _, r1 = p.minimize_bfgs_graph_host(th1, hp, 8, 1e30)
        assert (r1.status, r1.nit, r1.nfev) == (0, 0, 1)
    finally:
        p.close()


@pytest.mark.parametrize('name,shape,lvl,maxiter', [('tiny', (2, 2), 3, 11), ('tiny', (4, 4), 2, 19), ('mvsec_dt4', (8, 8), 1, 12), ('mvsec_dt4', (16, 16), 0, 10)])
def test_levels_follow_the_host_loop(name, shape, lvl, maxiter):
    from eincm_b200 import plan as P
    win = S.make_workload(name, seed=6)
    hp = P.make_hparams(win.hparams['alpha'], win.hparams['beta'], 0.0, 0.0, lvl)
    p = _plan(P, win)
    try:
        th0 = 0.25 * S.theta_test_points(win, shape)['truth']
        ta, ra = p.minimize_bfgs_host(th0, hp, maxiter, 1e-7, own_stream=True)
        tb, rb = p.minimize_bfgs_graph_host(th0, hp, maxiter, 1e-7)
        l0, _ = p.value_and_grad_host(th0, hp)
        assert rb.fun < l0 and ra.fun < l0                                        # both descend
        # the inverse-Hessian products are summed in another order: iterates agree to rounding until a line search amplifies it
        # (the float64 reductions of the gradient are atomics: their order, and with it the last bits, vary from run to run - a line search
        # near a tie can take another branch, so the schedules are compared loosely and the descents tightly)
        assert abs(rb.fun - ra.fun) <= 5e-2 * abs(ra.fun)
        assert 0 < rb.nit <= maxiter and rb.nit < rb.nfev <= 25 * (maxiter + 1)
        lb, _ = p.value_and_grad_host(tb, hp)
        assert lb == rb.fun                                                        # the reported value is the objective at the reported point
    finally:
        p.close()


def test_new_window_rebuilds_the_graph_and_errors_are_reported():
    from eincm_b200 import plan as P
    wins = [S.make_workload('tiny', seed=s) for s in (8, 9)]
    hp = P.make_hparams(wins[0].hparams['alpha'], wins[0].hparams['beta'], 0.0, 0.0, 3)
    p = _plan(P, wins[0])
    try:
        th0 = np.zeros((2, 2, 2))
        _, r_first = p.minimize_bfgs_graph_host(th0, hp, 6, 1e-7)
        p.set_window(*wins[1].args())
        tb, rb = p.minimize_bfgs_graph_host(th0, hp, 6, 1e-7)
        ref = _plan(P, wins[1])
        ta, ra = ref.minimize_bfgs_host(th0, hp, 6, 1e-7, own_stream=True)
        ref.close()
        assert abs(rb.fun - ra.fun) <= 1e-6 * abs(ra.fun) and rb.fun != r_first.fun      # the second window's solve, not the first's
        with pytest.raises(P.EincmError):
            p.minimize_bfgs_graph_host(np.zeros((32, 32, 2)), hp, 6, 1e-7)               # beyond 1024 parameters: the host loop's job
    finally:
        p.close()
