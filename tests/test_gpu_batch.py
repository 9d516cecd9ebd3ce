"""Batched evaluation (eincm_batch: one launch per kernel for B windows, blockIdx.y = window) against the per-plan calls of the same
library (bit-identical by construction: same kernel bodies) and against the oracle per window."""
import numpy as np
import pytest

import eincm_b200.synth as S
from oracle import eincm_oracle as O

pytestmark = pytest.mark.gpu

OBJ_RTOL = 1e-5      # BASELINE.json north_star
GRAD_RTOL = 1e-4     # BASELINE.json north_star


def _rel_inf(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def _make(P, name, n_win, n_events=None):
    wins = [S.make_workload(name, seed=10 + k, n_events=n_events) for k in range(n_win)]
    R = len(wins[0].edge_ts)
    plans = []
    for k, w in enumerate(wins):
        # ragged batch: windows of different event counts (the last one has a third of the events)
        n = len(w.xs) if k < n_win - 1 else len(w.xs) // 3
        p = P.Plan(w.sensor_size, max_events=len(w.xs), max_refs=max(3, R))
        p.set_window(w.xs[:n], w.ys[:n], w.ts[:n], w.edges, w.edge_ts)
        wins[k] = S.Window(xs=w.xs[:n], ys=w.ys[:n], ts=w.ts[:n], edges=w.edges, edge_ts=w.edge_ts, sensor_size=w.sensor_size,
                           truth_theta=w.truth_theta, hparams=w.hparams)
        plans.append(p)
    return wins, plans


@pytest.mark.parametrize('name,shape,n_win', [('tiny', (4, 4), 5), ('tiny', (1, 1), 3), ('mvsec_dt4', (16, 16), 6), ('mvsec_dt1', (16, 16), 4),
                                              ('tiny', 'dense', 4), ('mvsec_dt1', 'dense', 3)])
def test_batch_matches_per_plan_calls_and_oracle(name, shape, n_win):
    import torch
    from eincm_b200 import plan as P
    wins, plans = _make(P, name, n_win)
    if shape == 'dense':
        shape = wins[0].sensor_size
    hp = wins[0].hparams
    hpc = P.make_hparams(hp['alpha'], hp['beta'], 0.0, 0.0, 1)
    thetas = [S.theta_test_points(w, shape, seed=k)['perturbed'] for k, w in enumerate(wins)]
    batch = P.Batch(plans)
    try:
        single = [p.value_and_grad_host(th, hpc) for p, th in zip(plans, thetas)]
        losses, grads = batch.value_and_grad_host(thetas, hpc)
        for k in range(n_win):
            assert losses[k] == single[k][0]                                  # same kernels: bit-identical objective
            assert _rel_inf(grads[k], single[k][1]) <= 1e-10                   # float64 atomics in a different order
            l_ref, g_ref = O.value_and_grad(thetas[k], *wins[k].args(), hp['alpha'], hp['beta'], 0.0, 0.0, 1, 5, wins[k].sensor_size)
            assert abs(losses[k] - l_ref) <= OBJ_RTOL * abs(l_ref)
            assert _rel_inf(grads[k], g_ref) <= GRAD_RTOL
        # device operands, queued back to back (programmatic dependent launches between the batches), repeated
        th_d = [torch.from_numpy(t).cuda() for t in thetas]
        lo_d = [torch.zeros(1, dtype=torch.float64, device='cuda') for _ in thetas]
        gr_d = [torch.full(t.shape, 3.0, dtype=torch.float64, device='cuda') for t in thetas]
        for rep in range(3):
            batch.value_and_grad_device(th_d, hpc, lo_d, gr_d)
        torch.cuda.synchronize()
        for k in range(n_win):
            assert float(lo_d[k][0]) == single[k][0]
            assert _rel_inf(gr_d[k].cpu().numpy(), single[k][1]) <= 1e-10
        # the per-plan state is as after a single evaluation: read-outs keep working
        iwe = plans[0].iwe().cpu().numpy()
        assert iwe.shape == (len(wins[0].edge_ts),) + tuple(wins[0].sensor_size) and np.isfinite(iwe).all() and iwe.sum() > 0
        # and a per-plan evaluation after the batch still gives the same answer
        l2, g2 = plans[1].value_and_grad_host(thetas[1], hpc)
        assert l2 == single[1][0] and _rel_inf(g2, single[1][1]) <= 1e-10
    finally:
        batch.close()
        for p in plans:
            p.close()


def test_batch_new_windows_and_shapes_between_calls():
    """Windows are re-staged and the theta shape changes between calls (what a sweep over sequences / pyramid levels does)."""
    from eincm_b200 import plan as P
    wins, plans = _make(P, 'tiny', 3)
    hp = wins[0].hparams
    batch = P.Batch(plans)
    try:
        for rnd, shape in enumerate([(2, 2), (4, 4), (2, 2), (1, 1)]):
            hpc = P.make_hparams(hp['alpha'], hp['beta'], 0.0, 0.0, 2)
            if rnd == 2:                                        # new windows in the same plans
                for k, p in enumerate(plans):
                    w = S.make_workload('tiny', seed=50 + k)
                    wins[k] = w
                    p.set_window(*w.args())
            thetas = [S.theta_test_points(w, shape, seed=rnd)['perturbed'] for w in wins]
            losses, grads = batch.value_and_grad_host(thetas, hpc)
            for k in range(3):
                l_ref, g_ref = O.value_and_grad(thetas[k], *wins[k].args(), hp['alpha'], hp['beta'], 0.0, 0.0, 2, 5, wins[k].sensor_size)
                assert abs(losses[k] - l_ref) <= OBJ_RTOL * abs(l_ref)
                assert _rel_inf(grads[k], g_ref) <= GRAD_RTOL
    finally:
        batch.close()
        for p in plans:
            p.close()


def test_batch_rejects_what_it_does_not_implement():
    from eincm_b200 import plan as P
    wins, plans = _make(P, 'tiny', 2)
    batch = P.Batch(plans)
    try:
        th = [np.zeros((4, 4, 2))] * 2
        with pytest.raises(P.EincmError) as e:
            batch.value_and_grad_host(th, P.make_hparams(20.0, 35.0, 0.0, 0.3, 1))        # delta != 0
        assert e.value.code == P.EINCM_EUNSUPPORTED
        with pytest.raises(P.EincmError) as e:
            batch.value_and_grad_host(th, P.make_hparams(20.0, 35.0, 0.0025, 0.0, 0))     # TV term at level 0
        assert e.value.code == P.EINCM_EUNSUPPORTED
        with pytest.raises(P.EincmError):
            batch.value_and_grad_host([np.zeros((4, 4, 2))], P.make_hparams(20.0, 35.0, 0.0, 0.0, 1))   # one theta for two windows
    finally:
        batch.close()
        for p in plans:
            p.close()
