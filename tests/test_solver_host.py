"""Host logic of the multi-level solver mirror (no GPU): resize weights, schedule, level loop, retries and handover, driven
by the CPU oracle as the objective (test infrastructure only)."""
import numpy as np
import pytest

import eincm_b200.synth as S
from eincm_b200 import solver as SV
from oracle import eincm_oracle as O


class OracleObjective:
    """Same surface as losses.WindowObjective, evaluated by the float64 oracle."""

    def __init__(self, sensor_size, alpha, beta, gamma=0.0, delta=0.0):
        self.sensor_size = sensor_size
        self.kw = dict(alpha=alpha, beta=beta, gamma=gamma, delta=delta, n_pyr_lvls=5, sensor_size=sensor_size,
                       scale_to_sensor_size_method='bilinear')
        self.n_evals = 0
        self.ops = None

    def set_datasample(self, xs, ys, ts, edges, edge_ts):
        self.ops = (xs, ys, ts, edges, edge_ts)

    def scipy_fun(self, shape, cur_pyr_lvl):
        def fun(x):
            self.n_evals += 1
            v, g = O.value_and_grad(np.asarray(x).reshape(shape), *self.ops, cur_pyr_lvl=cur_pyr_lvl, **self.kw)
            return v, g.ravel()
        return fun

    def handover_value_and_grad(self, a, prev, theta, cur_pyr_lvl):
        self.n_evals += 1
        return O.handover_value_and_grad(a, prev, theta, *self.ops, cur_pyr_lvl=cur_pyr_lvl, **self.kw)


def test_weight_mat_matches_the_oracle_resize_and_is_a_partition_of_unity():
    for n_in, n_out in [(1, 48), (2, 64), (4, 48), (16, 480)]:
        w = SV.compute_weight_mat(n_in, n_out, n_out / n_in, 'bilinear')
        np.testing.assert_allclose(w, O.compute_weight_mat(n_in, n_out, n_out / n_in), rtol=0, atol=1e-15)
    for method in ('lanczos3', 'bilinear', 'bicubic'):
        for n_in in (16, 8, 4, 2):
            w = SV.compute_weight_mat(n_in, n_in // 2, 0.5, method)
            np.testing.assert_allclose(w.sum(axis=0), 1.0, rtol=1e-12)          # constants are preserved
    theta = np.random.default_rng(0).normal(size=(16, 16, 2))
    down = SV.scale_and_translate(theta, (8, 8), 'lanczos3')
    assert down.shape == (8, 8, 2)
    const = SV.scale_and_translate(np.full((8, 8, 2), 3.25), (4, 4), 'lanczos3')
    np.testing.assert_allclose(const, 3.25, rtol=1e-12)
    # lanczos3 with antialias (kernel_scale = 2 when halving): 12 input taps per output sample
    w = SV.compute_weight_mat(16, 8, 0.5, 'lanczos3')
    assert (np.abs(w[:, 4]) > 0).sum() == 11 or (np.abs(w[:, 4]) > 0).sum() == 12


def test_default_schedules_match_main_yaml():
    sol = SV.MultipleLevelEINCMSolver(OracleObjective((48, 64), 20.0, 35.0))
    assert [sol.theta_opt_maxiters[f'pyr_lvl_{k}'] for k in range(5)] == [40, 28, 19, 11, 8]          # exp_mgr.py:177-184
    assert [sol.handover_opt_maxiters[f'pyr_lvl_{k}'] for k in range(5)] == [20, 14, 10, 6, 4]
    assert sol.solve_handover_switch_per_level == {'pyr_lvl_0': True, 'pyr_lvl_1': True, 'pyr_lvl_2': False,
                                                   'pyr_lvl_3': False, 'pyr_lvl_4': False}            # main.yaml:54
    assert [sol.pre_opt_theta_pyr[f'pyr_lvl_{k}'].shape for k in range(5)] == [(16, 16, 2), (8, 8, 2), (4, 4, 2), (2, 2, 2), (1, 1, 2)]


def test_two_window_solve_with_handover():
    """Coarse-to-fine solve of two chained windows on a small sensor: the loss decreases level by level from theta = 0, the
    first window skips the handover (solver.py:305-306), the second solves alpha at levels 1 and 0 and blends with the prior."""
    w0 = S.make_window(32, 48, 1500, seed=21, n_segments=12, flow_mag=4.0)
    w1 = S.make_window(32, 48, 1500, seed=21, n_segments=12, flow_mag=4.0)          # same scene: the prior is a good guess
    obj = OracleObjective((32, 48), 20.0, 35.0)
    sol = SV.MultipleLevelEINCMSolver(obj, n_pyr_lvls=3, theta_opt_maxiters={'pyr_lvl_0': 6, 'pyr_lvl_1': 5, 'pyr_lvl_2': 4},
                                      handover_opt_maxiters={'pyr_lvl_0': 4, 'pyr_lvl_1': 3, 'pyr_lvl_2': 2})
    sol.set_datasample(*w0.args())
    r0 = sol.solve()
    assert set(r0) == {'prior_theta_pyr', 'pre_opt_theta_pyr', 'theta_opt_state_pyr', 'pre_handover_theta_pyr', 'ho_opt_state_pyr',
                       'final_handover_weight_pyr', 'final_theta_pyr'}                              # solver.py:259-267
    assert r0['ho_opt_state_pyr'] == {}                                                            # first sample: no handover
    assert [r0['final_theta_pyr'][f'pyr_lvl_{k}'].shape for k in range(3)] == [(4, 4, 2), (2, 2, 2), (1, 1, 2)]
    kw = dict(obj.kw)
    l_zero = O.loss_func(np.zeros((1, 1, 2)), *w0.args(), cur_pyr_lvl=2, **kw)[0]
    losses = [r0['theta_opt_state_pyr'][f'pyr_lvl_{k}'].fun_val for k in (2, 1, 0)]
    assert losses[0] < l_zero and losses[2] <= losses[0] + 1e-9
    # pre-opt theta of a finer level is the repeat-upscaled result of the coarser one (solver.py:245-248)
    np.testing.assert_array_equal(r0['pre_opt_theta_pyr']['pyr_lvl_1'], np.repeat(np.repeat(r0['final_theta_pyr']['pyr_lvl_2'], 2, 0), 2, 1))
    sol.set_datasample(*w1.args())
    r1 = sol.solve()
    assert set(r1['ho_opt_state_pyr']) == {'pyr_lvl_1', 'pyr_lvl_0'}
    for k in (1, 0):
        a = r1['final_handover_weight_pyr'][f'pyr_lvl_{k}']
        assert 0.0 <= a <= 1.0
        blend = a * r1['prior_theta_pyr'][f'pyr_lvl_{k}'] + (1 - a) * r1['pre_handover_theta_pyr'][f'pyr_lvl_{k}']
        np.testing.assert_allclose(r1['final_theta_pyr'][f'pyr_lvl_{k}'], blend, rtol=1e-12, atol=1e-12)
    assert r1['final_handover_weight_pyr']['pyr_lvl_2'] == 0.67                                    # fixed weight (main.yaml:59)
    # the priors of the coarser levels are the lanczos3-downscaled finest prior (solver.py:283-289)
    np.testing.assert_allclose(r1['prior_theta_pyr']['pyr_lvl_1'],
                               SV.scale_and_translate(r0['final_theta_pyr']['pyr_lvl_0'], (2, 2), 'lanczos3'), rtol=1e-12)


def test_theta_pyramids_can_start_from_a_given_prior():
    """solver.py:129-148: `theta_pyr_init` replaces the zero prior pyramid (how a resumed run hands in its last solved pyramid,
    exp_mgr.py:236-243 does the same by assignment); the other three pyramids start from zeros"""
    sol = SV.MultipleLevelEINCMSolver(OracleObjective((48, 64), 20.0, 35.0), n_pyr_lvls=3)
    prior = {f'pyr_lvl_{k}': np.full((4 >> k, 4 >> k, 2), 1.5 + k) for k in range(3)}
    sol._initialize_theta_pyramids(theta_pyr_init=prior)
    assert sol.prior_theta_pyr is prior
    for pyr in (sol.pre_opt_theta_pyr, sol.opt_theta_pyr, sol.handover_opt_theta_pyr):
        assert [pyr[f'pyr_lvl_{k}'].shape for k in range(3)] == [(4, 4, 2), (2, 2, 2), (1, 1, 2)]
        assert all(not pyr[f'pyr_lvl_{k}'].any() for k in range(3))
    assert sol.init_handover_weight_pyr == {f'pyr_lvl_{k}': 0.5 for k in range(3)} == sol.final_handover_weight_pyr
