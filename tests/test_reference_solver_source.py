"""The reference's OWN multi-level solver (src/eincm/solver.py:10-383, callbacks.py, losses.py - imported unmodified from
/root/reference and executed over the stand-ins of tests/_jaxshim: jax on float64 torch, jaxopt's SciPy wrappers, easydict) against
the solver mirror of this repo driven by the CPU oracle (eincm_b200.solver.MultipleLevelEINCMSolver, backend 'scipy').  Same SciPy
underneath both, so what is compared is everything around it: pyramid set-up, maxiter schedule, retries, lanczos3 down-scaling of the
prior, 'repeat' up-scaling, the handover solves at levels 1 / 0 and the blend - window after window.

Needs /root/reference: runs in the build container, skipped on the GPU box.  """
import os
import sys
from functools import partial

import numpy as np
import pytest

import eincm_b200.synth as S
from eincm_b200 import solver as SV
from tests import _golden as G
from tests.test_solver_host import OracleObjective

REFERENCE_SRC = '/root/reference/src'
HP = dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.0)


@pytest.fixture(scope='module')
def reference():
    if not os.path.isdir(REFERENCE_SRC):
        pytest.skip('/root/reference is not on this machine')
    saved_path, saved_mods = list(sys.path), set(sys.modules)
    sys.path.insert(0, G.GOLDEN_DIR)
    import make_golden_refsrc as M
    jax, L = M.import_reference()
    import torch
    n_threads = torch.get_num_threads()
    torch.set_num_threads(1)            # thousands of small tensor operations: thread pools only cost
    import eincm.solver as RS
    import eincm.callbacks as RC
    from easydict import EasyDict
    yield jax, L, RS, RC, EasyDict
    torch.set_num_threads(n_threads)
    sys.path[:] = saved_path
    for m in set(sys.modules) - saved_mods:
        del sys.modules[m]


def _reference_solver(ref, sensor_size, n_lvls, theta_maxiters, ho_maxiters, theta_params, ho_params, settings):
    jax, L, RS, RC, EasyDict = ref
    import jax.numpy as jnp                                                     # the stand-in
    from utils.theta_utils import scale_theta_to_sensor_size
    kw = dict(n_pyr_lvls=n_lvls, sensor_size=sensor_size, scale_to_sensor_size_method='bilinear', **HP)
    scale = partial(scale_theta_to_sensor_size, sensor_size=sensor_size, method='bilinear')
    tcb = RC.EINCMThetaSolverCallback(n_lvls, scale, None, EasyDict(print_intermediate_loss=False, collect_thetas_and_losses=True, eval_thetas=False,
                                                                    collect_eval_results=False, print_eval_results=False))
    hcb = RC.EINCMHandoverSolverCallback(n_lvls, scale, None, EasyDict(print_intermediate_loss=False, collect_ho_weights_and_losses=True,
                                                                       collect_thetas=True, eval_ho_weights=False, collect_eval_results=False,
                                                                       print_eval_results=False))
    sol = RS.MultipleLevelEINCMSolver(n_pyr_lvls=n_lvls, theta_opt_maxiters=theta_maxiters, theta_loss_pfunc=partial(L.loss_func, **kw),
                                      theta_opt_solver_params=theta_params, handover_opt_maxiters=ho_maxiters,
                                      handover_loss_pfunc=partial(L.handover_loss_func, **kw), handover_opt_solver_params=ho_params,
                                      handover_settings=EasyDict(settings), pyramid_downscale_method='lanczos3', pyramid_upscale_method='repeat',
                                      pyramid_bases=[2] * (n_lvls - 1), theta_solver_callback=tcb, handover_solver_callback=hcb)
    return sol, jnp


SLOW = pytest.mark.skipif(os.environ.get('EINCM_SLOW_TESTS') != '1', reason='~2 min each on the float64 torch stand-in: EINCM_SLOW_TESTS=1')


@pytest.mark.parametrize('settings_over,n_windows', [
    ({}, 2),                                                                    # main.yaml:51-59 as shipped
    pytest.param({}, 3, marks=SLOW),
    pytest.param({'solve_handover_for_levels': [0], 'clip_solved_handover': True, 'use_downscaled_finest_priors': False}, 2, marks=SLOW)])
def test_solver_mirror_follows_the_reference_solver(reference, settings_over, n_windows, capsys):
    n_lvls, sensor = 3, (32, 48)
    theta_maxiters = {'pyr_lvl_0': 6, 'pyr_lvl_1': 5, 'pyr_lvl_2': 4}
    ho_maxiters = {'pyr_lvl_0': 4, 'pyr_lvl_1': 3, 'pyr_lvl_2': 2}
    theta_params = {'method': 'BFGS', 'maxiter': 6, 'n_extra_attempts': {'pyr_lvl_0': 1, 'pyr_lvl_1': 1}, 'options': {'gtol': 1e-7}}
    ho_params = {'method': 'L-BFGS-B', 'maxiter': 4, 'options': {'gtol': 1e-6}}
    settings = dict(SV.DEFAULT_HANDOVER_SETTINGS, **settings_over)

    ref_sol, jnp = _reference_solver(reference, sensor, n_lvls, theta_maxiters, ho_maxiters, theta_params, ho_params, settings)
    obj = OracleObjective(sensor, HP['alpha'], HP['beta'])
    obj.kw['n_pyr_lvls'] = n_lvls
    mir = SV.MultipleLevelEINCMSolver(obj, n_pyr_lvls=n_lvls, theta_opt_maxiters=dict(theta_maxiters), theta_opt_solver_params=theta_params,
                                      handover_opt_maxiters=dict(ho_maxiters), handover_opt_solver_params=ho_params, handover_settings=settings,
                                      backend='scipy')
    for k in range(n_lvls):
        key = f'pyr_lvl_{k}'
        assert mir.pre_opt_theta_pyr[key].shape == tuple(ref_sol.pre_opt_theta_pyr[key].shape)
        assert mir.solve_handover_switch_per_level[key] == ref_sol.solve_handover_switch_per_level[key]

    n_handover_solves = 0
    rs0 = np.random.default_rng(5)
    truth0, drift = rs0.uniform(-4, 4, size=(2, 2, 2)), rs0.normal(0, 0.3, size=(2, 2, 2))
    for w in range(n_windows):      # chained windows of one edge-dense scene whose flow drifts (a sparse scene sends BFGS to flows of
        #                     hundreds of pixels, where round-off decides the line search and no two runs agree: synth.make_sequence)
        win = S.make_window(32, 48, 4000, seed=300 + w, n_segments=40, scene_seed=77, truth_theta=truth0 + w * drift)
        ref_sol.set_datasample(*(jnp.array(a) for a in win.args()))
        mir.set_datasample(*win.args())
        r, m = ref_sol.solve(), mir.solve()
        capsys.readouterr()
        assert set(r) == set(m)
        for k in reversed(range(n_lvls)):
            key = f'pyr_lvl_{k}'
            rs, ms = r['theta_opt_state_pyr'][key], m['theta_opt_state_pyr'][key]
            assert (int(rs.iter_num), int(rs.status), bool(rs.success)) == (ms.iter_num, ms.status, ms.success), (w, key)
            assert float(rs.fun_val) == pytest.approx(ms.fun_val, rel=1e-7), (w, key)
            for name in ('prior_theta_pyr', 'pre_opt_theta_pyr', 'pre_handover_theta_pyr', 'final_theta_pyr'):
                np.testing.assert_allclose(np.asarray(m[name][key]), np.asarray(r[name][key]), rtol=1e-6, atol=1e-6, err_msg=f'window {w} {name} {key}')
            assert float(m['final_handover_weight_pyr'][key]) == pytest.approx(float(r['final_handover_weight_pyr'][key]), abs=1e-6), (w, key)
            assert (key in r['ho_opt_state_pyr']) == (key in m['ho_opt_state_pyr'])
            if key in r['ho_opt_state_pyr']:
                n_handover_solves += 1
                assert int(r['ho_opt_state_pyr'][key].iter_num) == m['ho_opt_state_pyr'][key].iter_num
    assert n_handover_solves == (n_windows - 1) * len(settings['solve_handover_for_levels'])   # none for the first window (solver.py:305-306)
