"""The batched device-side solve loop (eincm_batch_minimize_bfgs_graph_host: B windows in lockstep inside one unrolled CUDA graph, one
optimizer CTA per window, finished windows skipped) against the single-window device loop on the same windows: the same step kernel body
and the same evaluation kernel bodies, so the same schedule - exactly for two parameters, to rounding beyond - and the lockstep solver
mirror against one MultipleLevelEINCMSolver per window (reference src/eincm/solver.py:197-267)."""
import numpy as np
import pytest

import eincm_b200.synth as S

pytestmark = pytest.mark.gpu


def _plans(P, wins):
    ps = []
    for win in wins:
        p = P.Plan(win.sensor_size, max_events=max(len(w.xs) for w in wins), max_refs=max(3, len(win.edge_ts)))
        p.set_window(*win.args())
        ps.append(p)
    return ps


def test_one_by_one_level_is_identical_to_the_single_window_loop():
    from eincm_b200 import plan as P
    wins = [S.make_workload('tiny', seed=s) for s in (4, 5, 6, 7, 8)]
    hp = P.make_hparams(wins[0].hparams['alpha'], wins[0].hparams['beta'], 0.0, 0.0, 4)
    ps = _plans(P, wins)
    b = P.Batch(ps)
    try:
        th0 = np.stack([np.full((1, 1, 2), 0.1 * k) for k in range(len(ps))])
        single = [p.minimize_bfgs_graph_host(th0[k], hp, 8, 1e-7) for k, p in enumerate(ps)]
        tb, rb = b.minimize_bfgs_graph_host(th0, hp, 8, 1e-7)
        assert b.solve_launches() >= 1
        for k, (ta, ra) in enumerate(single):
            assert (rb[k].status, rb[k].nit, rb[k].nfev) == (ra.status, ra.nit, ra.nfev), k
            assert rb[k].fun == ra.fun
            np.testing.assert_array_equal(tb[k], ta)                              # the same device code: identical
        assert len({r.nfev for r in rb}) > 1                 # the windows end at different steps: the skip flags were exercised
        # a subset: the others are left untouched
        act = np.array([1, 0, 1, 0, 0], dtype=np.int32)
        tc, rc = b.minimize_bfgs_graph_host(th0, hp, 8, 1e-7, active=act)
        for k in range(len(ps)):
            if act[k]:
                assert (rc[k].status, rc[k].nit, rc[k].nfev, rc[k].fun) == (rb[k].status, rb[k].nit, rb[k].nfev, rb[k].fun)
                np.testing.assert_array_equal(tc[k], tb[k])
            else:
                assert rc[k].nfev == 0
                np.testing.assert_array_equal(tc[k], th0[k])
        # maxiter 0 and an already converged start
        _, r0 = b.minimize_bfgs_graph_host(th0, hp, 0, 1e-7)
        assert all((r.status, r.nit, r.nfev) == (1, 0, 1) for r in r0)
        _, r1 = b.minimize_bfgs_graph_host(th0, hp, 8, 1e30)
        assert all((r.status, r.nit, r.nfev) == (0, 0, 1) for r in r1)
    finally:
        b.close()
        for p in ps:
            p.close()


@pytest.mark.parametrize('name,shape,lvl,maxiter', [('tiny', (4, 4), 2, 19), ('mvsec_dt4', (16, 16), 0, 10)])
def test_levels_follow_the_single_window_loop(name, shape, lvl, maxiter):
    from eincm_b200 import plan as P
    wins = [S.make_workload(name, seed=s) for s in (6, 7, 8)]
    hp = P.make_hparams(wins[0].hparams['alpha'], wins[0].hparams['beta'], 0.0, 0.0, lvl)
    ps = _plans(P, wins)
    b = P.Batch(ps)
    try:
        th0 = np.stack([0.25 * S.theta_test_points(w, shape)['truth'] for w in wins])
        single = [p.minimize_bfgs_graph_host(th0[k], hp, maxiter, 1e-7) for k, p in enumerate(ps)]
        tb, rb = b.minimize_bfgs_graph_host(th0, hp, maxiter, 1e-7)
        for k, (ta, ra) in enumerate(single):
            l0, _ = ps[k].value_and_grad_host(th0[k], hp)
            assert rb[k].fun < l0
            # the gradient's float64 reductions are summed in another order: iterates agree to rounding until a line search amplifies it
            # and their order varies from run to run: schedules compared loosely, descents tightly (see test_gpu_graph_solver.py)
            assert abs(rb[k].fun - ra.fun) <= 5e-2 * abs(ra.fun)
            assert 0 < rb[k].nit <= maxiter and rb[k].nit < rb[k].nfev <= 25 * (maxiter + 1)
            lb, _ = ps[k].value_and_grad_host(tb[k], hp)
            assert lb == rb[k].fun                                                 # the reported value is the objective at the reported point
    finally:
        b.close()
        for p in ps:
            p.close()


def test_new_windows_reuse_the_graph_and_errors_are_reported():
    from eincm_b200 import plan as P
    first = [S.make_workload('tiny', seed=s) for s in (8, 9)]
    second = [S.make_workload('tiny', seed=s) for s in (10, 11)]
    hp = P.make_hparams(first[0].hparams['alpha'], first[0].hparams['beta'], 0.0, 0.0, 3)
    ps = _plans(P, first + second)[:2]
    b = P.Batch(ps)
    try:
        th0 = np.zeros((2, 2, 2, 2))
        _, r_first = b.minimize_bfgs_graph_host(th0, hp, 6, 1e-7)
        for p, w in zip(ps, second):
            p.set_window(*w.args())
        tb, rb = b.minimize_bfgs_graph_host(th0, hp, 6, 1e-7)
        refs = _plans(P, second)
        for k, ref in enumerate(refs):
            ta, ra = ref.minimize_bfgs_graph_host(th0[k], hp, 6, 1e-7)
            ref.close()
            assert abs(rb[k].fun - ra.fun) <= 1e-6 * abs(ra.fun) and rb[k].fun != r_first[k].fun    # the second windows' solves
        with pytest.raises(P.EincmError):
            b.minimize_bfgs_graph_host(np.zeros((2, 32, 32, 2)), hp, 6, 1e-7)                       # beyond 1024 parameters
        with pytest.raises(P.EincmError):
            b.minimize_bfgs_graph_host(np.zeros((3, 2, 2, 2)), hp, 6, 1e-7)                         # one theta per window
    finally:
        b.close()
        for p in ps:
            p.close()


def test_lockstep_solver_matches_one_solver_per_window():
    from eincm_b200 import losses, solver as SV
    seqs = [S.make_sequence('mvsec_dt4', 2, seed=t) for t in range(3)]
    w0 = seqs[0][0]
    H, W = w0.sensor_size
    hpd = w0.hparams

    def objs():
        return [losses.WindowObjective((H, W), hpd['alpha'], hpd['beta'], 0.0, hpd['delta'], max_events=len(w0.xs), max_refs=max(3, len(w0.edge_ts)))
                for _ in seqs]

    oa, ob = objs(), objs()
    single = [SV.MultipleLevelEINCMSolver(o, backend='graph') for o in oa]
    lock = SV.BatchedMultipleLevelEINCMSolver(ob)
    try:
        for step in range(2):                                # the second window of every sequence goes through the handover
            ra = []
            for s, seq in zip(single, seqs):
                s.set_datasample(*seq[step].args())
                ra.append(s.solve())
            lock.set_datasamples([seq[step].args() for seq in seqs])
            rb = lock.solve()
            for k in range(len(seqs)):
                fa = ra[k]['theta_opt_state_pyr']['pyr_lvl_0'].fun_val
                fb = rb[k]['theta_opt_state_pyr']['pyr_lvl_0'].fun_val
                # five levels of BFGS amplify the rounding differences of the gradient reductions: a descent of the same size, not the same digits
                f_start = ob[k].value(rb[k]['pre_opt_theta_pyr']['pyr_lvl_0'], 0)      # BFGS never returns a point above its start
                assert fb <= f_start and abs(fa - fb) <= 0.5 * abs(fa), (step, k, fa, fb, f_start)
                ta, tb = ra[k]['final_theta_pyr']['pyr_lvl_0'], rb[k]['final_theta_pyr']['pyr_lvl_0']
                assert tb.shape == ta.shape == (16, 16, 2) and np.isfinite(tb).all()
                # every level's state is reported per window, with scipy's status codes
                for lvl in range(5):
                    sa, sb = ra[k]['theta_opt_state_pyr'][f'pyr_lvl_{lvl}'], rb[k]['theta_opt_state_pyr'][f'pyr_lvl_{lvl}']
                    assert sb.status in (0, 1, 2) and sb.n_evals >= 1 and sb.iter_num >= 0
                    if lvl == 4:                             # the coarsest level of the first window starts from the same point: 2 parameters, identical
                        assert step > 0 or (sb.status, sb.iter_num, sb.n_evals, sb.fun_val) == (sa.status, sa.iter_num, sa.n_evals, sa.fun_val)
        assert lock.graph_launches > 0
    finally:
        lock.close()
        for o in oa + ob:
            o.close()


def test_lockstep_solver_hands_the_tv_level_to_the_single_window_solver():
    """gamma != 0 (MVSEC outdoor, run.sh:85): the TV regulariser is active at the finest level, which the batched kernels do not evaluate -
    that level runs per window, the coarser levels in lockstep."""
    from eincm_b200 import losses, solver as SV
    wins = [S.make_workload('tiny', seed=s) for s in (21, 22)]
    H, W = wins[0].sensor_size
    hpd = wins[0].hparams
    objs = [losses.WindowObjective((H, W), hpd['alpha'], hpd['beta'], 0.0025, 0.0, max_events=max(len(w.xs) for w in wins), max_refs=max(3, len(wins[0].edge_ts)))
            for _ in wins]
    lock = SV.BatchedMultipleLevelEINCMSolver(objs)
    try:
        lock.set_datasamples([w.args() for w in wins])
        res = lock.solve()
        assert lock.graph_launches > 0                                          # levels 4 .. 1 in lockstep
        for k, r in enumerate(res):
            st = r['theta_opt_state_pyr']['pyr_lvl_0']
            assert st.status in (0, 1, 2) and st.n_evals >= 1
            th = r['final_theta_pyr']['pyr_lvl_0']
            assert th.shape == (16, 16, 2) and np.isfinite(th).all()
            assert objs[k].value(th, 0) == st.fun_val                            # the finest level's value includes the TV term
    finally:
        lock.close()
        for o in objs:
            o.close()
