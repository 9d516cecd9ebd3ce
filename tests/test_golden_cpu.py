"""The committed golden vectors (tests/golden/, outputs of the NumPy oracle at generation time - see make_golden.py for why they
are not outputs of the JAX reference) against the oracle as it is now, and against the C restatement used as the timed CPU
baseline.  Guards the checker itself against regressions."""
import os
import subprocess

import numpy as np
import pytest

from oracle import c_oracle as C
from oracle import eincm_oracle as O
from tests import _golden as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fixtures_are_committed():
    assert len(G.NAMES) >= 4


@pytest.mark.parametrize('name', G.NAMES)
def test_numpy_oracle_reproduces_golden(name):
    g = G.load(name)
    kw = dict(n_pyr_lvls=5, sensor_size=g['sensor_size'], scale_to_sensor_size_method='bilinear', **g['hp'])
    loss, grad, inter = O.value_and_grad(g['theta'], *g['args'], return_intermediates=True, **kw)
    assert loss == pytest.approx(float(g['loss']), rel=1e-13)
    assert np.abs(grad - g['grad']).max() <= 1e-11 * np.abs(g['grad']).max()
    np.testing.assert_allclose(inter['iwes'], g['iwes'], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(inter['zero_iwe'], g['zero_iwe'], rtol=1e-13, atol=1e-15)
    obj = inter['objectives']
    for r in range(len(g['edge_ts'])):
        xr, yr = O.rounded_event_pixels(obj['warped_xs'][r], obj['warped_ys'][r])
        np.testing.assert_array_equal(xr.astype(np.int32), g['rounded'][r, 0])      # index stream: bit-exact
        np.testing.assert_array_equal(yr.astype(np.int32), g['rounded'][r, 1])
    hl, hd = O.handover_value_and_grad(float(g['alpha_handover']), g['prev_theta'], g['theta'], *g['args'], **kw)
    assert hl == pytest.approx(float(g['handover_loss']), rel=1e-13)
    assert hd == pytest.approx(float(g['handover_dalpha']), rel=1e-10, abs=1e-12)


@pytest.mark.parametrize('name', G.NAMES)
def test_evaluation_metrics_reproduce_golden(name):
    g = G.load(name)
    hp = g['hp']
    theta_full = O.scale_theta_to_sensor_size(g['theta'], g['sensor_size'])
    ev = O.evaluate_theta_array(theta_full, *g['args'], g['gt_flow'], hp['alpha'], hp['beta'], hp['gamma'], hp['delta'], g['sensor_size'],
                                err_eval_event_mask=g['err_mask'])
    for k, v in g['eval'].items():
        assert float(ev[k]) == pytest.approx(v, rel=1e-12, abs=1e-14), k
    np.testing.assert_allclose(ev['flow_warp_losses'], g['eval_flow_warp_losses'], rtol=1e-12)


@pytest.mark.parametrize('name', G.NAMES)
def test_c_oracle_matches_golden(name):
    if not C.available():
        subprocess.run(['make', '-C', os.path.join(ROOT, 'oracle')], check=True, stdout=subprocess.DEVNULL)
    g = G.load(name)
    hp = g['hp']
    # the TV count of exactly-non-zero gradients depends on the summation order of the resize on piecewise-linear fields
    # (DESIGN.md section 2); the golden TV cases use a generic 'perturbed' theta, where it does not
    lc, gc, iw = C.value_and_grad_raw(g['theta'], *g['args'], hp['alpha'], hp['beta'], hp['gamma'], hp['delta'], hp['cur_pyr_lvl'],
                                      g['sensor_size'], want_iwes=True)
    assert lc == pytest.approx(float(g['loss']), rel=1e-12)
    assert np.abs(gc - g['grad']).max() <= 1e-10 * np.abs(g['grad']).max()
    assert np.abs(iw - g['iwes']).max() <= 1e-12 * np.abs(g['iwes']).max()


# ---- sparse_flow_error known answers (reference src/evaluations/flow_eval.py:14-75, derived by hand) ---------------------------
def test_sparse_flow_error_known_answers():
    pred = np.zeros((2, 3, 2))
    gt = np.zeros((2, 3, 2))
    pred[0, 0] = (3.0, 4.0); gt[0, 0] = (0.0, 4.0)        # ee = 3, |gt| = 4
    pred[0, 1] = (1.0, 0.0); gt[0, 1] = (1.0, 0.5)        # ee = 0.5, |gt| = sqrt(1.25)
    pred[0, 2] = (0.0, 0.0); gt[0, 2] = (1.0, 1.0)        # zero prediction: invalid
    pred[1, 0] = (2.0, 2.0); gt[1, 0] = (0.0, 0.0)        # zero ground truth: invalid
    pred[1, 1] = (np.inf, 1.0); gt[1, 1] = (1.0, 1.0)     # infinite prediction: invalid
    pred[1, 2] = (30.0, 0.0); gt[1, 2] = (5.0, 0.0)       # ee = 25, |gt| = 5
    r = O.sparse_flow_error(pred, gt)
    assert r['counts'] == {'n_ee': 3, 'n_pred': 4, 'n_gt': 5}
    assert r['errors']['AEE'] == pytest.approx((3.0 + 0.5 + 25.0) / 3.0, rel=1e-15)
    assert r['errors']['AREE'] == pytest.approx((3.0 / 4.0 + 0.5 / np.sqrt(1.25) + 5.0) / 3.0, rel=1e-14)
    assert [round(r['errors'][f'A{n}PE'], 9) for n in (1, 2, 3, 5, 10, 20)] == [round(200 / 3, 9), round(200 / 3, 9), round(100 / 3, 9),
                                                                                  round(100 / 3, 9), round(100 / 3, 9), round(100 / 3, 9)]
    m = np.ones((2, 3), dtype=bool); m[1, 2] = False       # event mask removes the large error
    r = O.sparse_flow_error(pred, gt, m)
    assert r['counts'] == {'n_ee': 2, 'n_pred': 3, 'n_gt': 5}
    assert r['errors']['AEE'] == pytest.approx(1.75, rel=1e-15)
    r = O.sparse_flow_error(np.zeros((2, 2, 2)), np.ones((2, 2, 2)))       # empty intersection: mean of nothing, 0 %
    assert r['counts']['n_ee'] == 0 and np.isnan(r['errors']['AEE']) and r['errors']['A1PE'] == 0.0


def test_evaluation_loss_is_ungated_and_unweighted():
    """theta_eval.py:27-42: plain means over the reference times and TV / divergence always included, unlike loss_func."""
    import eincm_b200.synth as S
    win = S.make_workload('tiny', seed=2)
    th = O.scale_theta_to_sensor_size(S.theta_test_points(win, (4, 4))['perturbed'], win.sensor_size)
    ev = O.evaluate_theta_array(th, *win.args(), None, 20.0, 35.0, 0.5, 0.25, win.sensor_size)
    lo = O.compute_loss_objectives(th, *win.args(), win.sensor_size)
    want = (20.0 * -lo['rel_contrasts'].mean() + 35.0 * -lo['rel_correlations'].mean() + 0.5 * lo['theta_total_variation']
            + 0.25 * lo['rel_iwe_divergences'].mean())
    assert ev['loss'] == pytest.approx(want, rel=1e-14)
    assert ev['fwl'] == pytest.approx(np.var(lo['_iwes'][0]) / np.var(lo['_zero_iwe']), rel=1e-14)
    assert 'AEE' not in ev
