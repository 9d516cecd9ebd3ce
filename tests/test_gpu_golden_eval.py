"""CUDA path (through the C-ABI) against the committed golden vectors, and the evaluation metrics (SURVEY.md 8f rank 2:
reference src/evaluations/theta_eval.py, flow_eval.py) against the oracle."""
import numpy as np
import pytest

import eincm_b200.synth as S
from oracle import eincm_oracle as O
from tests import _golden as G

pytestmark = pytest.mark.gpu

OBJ_RTOL = 1e-5      # BASELINE.json north_star
GRAD_RTOL = 1e-4


def _rel_inf(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


@pytest.mark.parametrize('exact', [False, True])
@pytest.mark.parametrize('name', G.NAMES)
def test_cuda_matches_golden(name, exact):
    from eincm_b200 import plan as P
    g = G.load(name)
    hp = P.make_hparams(**g['hp'])
    R = len(g['edge_ts'])
    p = P.Plan(g['sensor_size'], max_events=len(g['xs']), max_refs=max(R, 3), flags=P.FLAG_EXACT_F64 if exact else 0)
    p.set_window(*g['args'])
    loss, grad = p.value_and_grad_host(g['theta'], hp)
    rl, rg = (1e-11, 1e-9) if exact else (OBJ_RTOL, GRAD_RTOL)
    assert abs(loss - float(g['loss'])) <= rl * abs(float(g['loss']))
    assert _rel_inf(grad, g['grad']) <= rg
    for r in range(R):
        cols, rows = p.rounded_pixels(r)
        np.testing.assert_array_equal(cols, g['rounded'][r, 0])          # bit-exact event -> pixel indexing
        np.testing.assert_array_equal(rows, g['rounded'][r, 1])
    rt, at = (1e-12, 1e-14) if exact else (2e-5, 1e-7)
    np.testing.assert_allclose(p.zero_iwe().cpu().numpy(), g['zero_iwe'], rtol=rt, atol=at)
    if exact or g['hp']['delta'] != 0.0:
        np.testing.assert_allclose(p.iwe().cpu().numpy(), g['iwes'], rtol=rt, atol=at)
    hl, hd = p.handover_value_and_grad_host(float(g['alpha_handover']), g['prev_theta'], g['theta'], hp)
    assert abs(hl - float(g['handover_loss'])) <= rl * abs(float(g['handover_loss']))
    assert abs(hd - float(g['handover_dalpha'])) <= rg * max(abs(float(g['handover_dalpha'])), np.abs(g['grad']).max())
    p.close()


@pytest.mark.parametrize('exact', [False, True])
@pytest.mark.parametrize('name', G.NAMES)
def test_evaluate_theta_matches_golden(name, exact):
    from eincm_b200 import plan as P
    g = G.load(name)
    hpd = g['hp']
    R = len(g['edge_ts'])
    p = P.Plan(g['sensor_size'], max_events=len(g['xs']), max_refs=max(R, 3), flags=P.FLAG_EXACT_F64 if exact else 0)
    p.set_window(*g['args'])
    m = p.evaluate_theta(g['theta'], P.make_hparams(**hpd), gt_flow=g['gt_flow'], err_eval_event_mask=g['err_mask'])
    e = g['eval']
    rt = 1e-10 if exact else 2e-5
    for k in ('loss', 'iwe_var', 'mean_rel_contrast', 'mean_rel_corr', 'mean_rel_iwe_div', 'theta_tot_var', 'theta_div', 'fwl'):
        assert getattr(m, k) == pytest.approx(e[k], rel=rt), k
    np.testing.assert_allclose(list(m.flow_warp_losses)[:R], g['eval_flow_warp_losses'], rtol=rt)
    np.testing.assert_allclose(list(m.rel_contrasts)[:R], g['eval_rel_contrasts'], rtol=rt)
    np.testing.assert_allclose(list(m.rel_correlations)[:R], g['eval_rel_correlations'], rtol=rt)
    np.testing.assert_allclose(list(m.rel_iwe_divergences)[:R], g['eval_rel_iwe_divergences'], rtol=rt)
    # flow errors depend on theta and masks only: float64 everywhere, exact counts
    fe = m.flow.as_dict()
    assert fe['counts'] == {k: int(e[k]) for k in ('n_ee', 'n_pred', 'n_gt')}
    assert int(m.n_pixels) == int(e['n_pixels'])
    for k in ('AEE', 'AREE', 'A1PE', 'A2PE', 'A3PE', 'A5PE', 'A10PE', 'A20PE'):
        assert fe['errors'][k] == pytest.approx(e[k], rel=1e-12, abs=1e-13), k
    p.close()


def test_mirror_entry_points_match_oracle():
    """eincm_b200.evaluations keeps the reference's names, arguments and result keys."""
    from eincm_b200 import evaluations as E, losses
    win = S.make_workload('mvsec_dt4', seed=3, n_events=12000)
    pts = S.theta_test_points(win, (8, 8))
    th = O.scale_theta_to_sensor_size(pts['perturbed'], win.sensor_size)
    gt = O.scale_theta_to_sensor_size(pts['truth'], win.sensor_size)
    gt[::7, ::5] = 0.0
    ref = O.evaluate_theta_array(th, *win.args(), gt, 20.0, 35.0, 0.0025, 0.0, win.sensor_size)
    time_str, eval_str, evals, loss_obj = E.evaluate_theta_array(th, *win.args(), gt, 20.0, 35.0, 0.0025, 0.0, win.sensor_size)
    assert time_str.startswith('[') and 'FWL' in eval_str and 'AEE' in eval_str
    for k, v in ref.items():
        if np.isscalar(v):
            assert float(evals[k]) == pytest.approx(float(v), rel=2e-5, abs=1e-12), k
        else:
            np.testing.assert_allclose(evals[k], v, rtol=2e-5)
    assert set(ref) == set(evals)
    # stand-alone flow error, no mask / with mask / empty intersection
    rng = np.random.default_rng(0)
    pred = rng.normal(size=(40, 50, 2)) * 5; pred[rng.random((40, 50)) < 0.2] = 0.0; pred[3, 3, 0] = np.inf; pred[4, 4, 1] = np.nan
    gtf = rng.normal(size=(40, 50, 2)) * 5; gtf[rng.random((40, 50)) < 0.2] = 0.0; gtf[5, 5] = -np.inf
    mask = rng.random((40, 50)) < 0.7
    for mk in (None, mask):
        a, b = E.sparse_flow_error(pred, gtf, mk), O.sparse_flow_error(pred, gtf, mk)
        assert a['counts'] == b['counts']
        for k in b['errors']:
            assert a['errors'][k] == pytest.approx(b['errors'][k], rel=1e-12), k
    a = E.sparse_flow_error(np.zeros((4, 4, 2)), np.ones((4, 4, 2)))
    assert a['counts']['n_ee'] == 0 and np.isnan(a['errors']['AEE']) and a['errors']['A1PE'] == 0.0
    losses.clear_cache()
