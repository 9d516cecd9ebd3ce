"""CUDA path (through the C-ABI) against the vectors of the reference's own source (tests/golden/refsrc/*.npz: the reference's
unmodified eincm.losses executed over the float64 JAX stand-in of tests/_jaxshim, see tests/golden/make_golden_refsrc.py).
Tolerances are BASELINE.json's: objective 1e-5, gradient 1e-4 relative (inf-norm); the event -> pixel index stream bit-exact."""
import numpy as np
import pytest

from tests import _golden as G
from tests.test_reference_source import FULLSIZE, load_fullsize, load_refsrc, _rel_inf

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('exact', [False, True])
@pytest.mark.parametrize('name', G.NAMES)
def test_cuda_matches_reference_source_vectors(name, exact):
    from eincm_b200 import plan as P
    g, ref = G.load(name), load_refsrc(name)
    hp = P.make_hparams(**g['hp'])
    R = len(g['edge_ts'])
    p = P.Plan(g['sensor_size'], max_events=len(g['xs']), max_refs=max(R, 3), flags=P.FLAG_EXACT_F64 if exact else 0)
    p.set_window(*g['args'])
    loss, grad = p.value_and_grad_host(g['theta'], hp)
    rl, rg = (1e-11, 1e-9) if exact else (1e-5, 1e-4)
    assert abs(loss - float(ref['loss'])) <= rl * abs(float(ref['loss']))
    assert _rel_inf(grad, ref['grad']) <= rg
    for r in range(R):
        cols, rows = p.rounded_pixels(r)
        np.testing.assert_array_equal(cols, np.rint(ref['obj_warped_xs'][r]).astype(np.int32))
        np.testing.assert_array_equal(rows, np.rint(ref['obj_warped_ys'][r]).astype(np.int32))
    rt, at = (1e-10, 1e-13) if exact else (2e-5, 1e-7)                            # the images of (warped) events themselves
    np.testing.assert_allclose(p.zero_iwe().cpu().numpy(), ref['zero_iwe'], rtol=rt, atol=at)
    if exact or g['hp']['delta'] != 0.0:                                          # default mode keeps fixed-point images unless delta needs them
        np.testing.assert_allclose(p.iwe().cpu().numpy(), ref['iwes'], rtol=rt, atol=at)
    hl, hd = p.handover_value_and_grad_host(float(g['alpha_handover']), g['prev_theta'], g['theta'], hp)
    assert abs(hl - float(ref['handover_loss'])) <= rl * abs(float(ref['handover_loss']))
    assert abs(hd - float(ref['handover_dalpha'])) <= rg * max(abs(float(ref['handover_dalpha'])), np.abs(ref['grad']).max())
    p.close()


@pytest.mark.parametrize('name', FULLSIZE)
def test_cuda_matches_reference_source_at_bench_configurations(name):
    """the BENCH line's configuration (dsec, N = 2 M, R = 3, theta 16 x 16, window of seed 0, the 'perturbed' point bench.py evaluates) and
    MVSEC dt4 with dense theta, against loss and gradient of the reference's own source (tests/golden/refsrc_fullsize/)"""
    from eincm_b200 import plan as P
    win, z = load_fullsize(name)
    hp = win.hparams
    R = len(win.edge_ts)
    p = P.Plan(win.sensor_size, max_events=len(win.xs), max_refs=max(R, 3))
    try:
        p.set_window(*win.args())
        loss, grad = p.value_and_grad_host(z['theta'], P.make_hparams(hp['alpha'], hp['beta'], hp['gamma'], hp['delta'], int(z['cur_pyr_lvl'])))
    finally:
        p.close()
    assert abs(loss - float(z['loss'])) <= 1e-5 * abs(float(z['loss'])), (loss, float(z['loss']))
    assert _rel_inf(grad, z['grad']) <= 1e-4


@pytest.mark.parametrize('name', G.NAMES)
def test_python_mirror_matches_reference_source_vectors(name):
    """the call a user of the reference makes - eincm_b200.losses.loss_func / handover_loss_func under value_and_grad, with the
    reference's argument list - against what the reference's own functions returned for the same arguments, aux_info included"""
    from eincm_b200 import losses as L
    g, ref = G.load(name), load_refsrc(name)
    kw = dict(n_pyr_lvls=5, sensor_size=g['sensor_size'], scale_to_sensor_size_method='bilinear', **g['hp'])
    try:
        (loss, aux), grad = L.value_and_grad(L.loss_func, has_aux=True)(g['theta'], *g['args'], **kw)
        assert abs(loss - float(ref['loss'])) <= 1e-5 * abs(float(ref['loss']))
        assert _rel_inf(grad, ref['grad']) <= 1e-4
        for k in ('mean_rel_corr', 'mean_rel_contrast'):
            assert abs(float(aux[k]) - float(ref[k])) <= 1e-5 * max(abs(float(ref[k])), 1e-12), k
        # KNOWN DEVIATION in aux_info (not in the loss): the device skips a regulariser whose weight is zero, so its aux entry reads 0.0
        # where the reference still reports the value it then multiplies by zero (losses.py:165-167, 188).  With the weight set they agree.
        if g['hp']['delta'] != 0.0:
            assert abs(float(aux['mean_rel_iwe_divergence']) - float(ref['mean_rel_iwe_divergence'])) <= 1e-5 * abs(float(ref['mean_rel_iwe_divergence']))
        else:
            assert float(aux['mean_rel_iwe_divergence']) == 0.0
        if g['hp']['gamma'] != 0.0:
            assert abs(float(aux['theta_total_variation']) - float(ref['theta_total_variation_used'])) <= 1e-5 * abs(float(ref['theta_total_variation_used']))
        np.testing.assert_allclose(np.asarray(aux['multi_ref_weights']), ref['obj_multi_ref_weights'], rtol=1e-12)
        scaled = aux['scaled_theta']
        scaled = scaled.cpu().numpy() if hasattr(scaled, 'cpu') else np.asarray(scaled)
        np.testing.assert_allclose(scaled, ref['scaled_theta'], rtol=1e-12, atol=1e-13)
        hl, hd = L.value_and_grad(L.handover_loss_func)(float(g['alpha_handover']), g['prev_theta'], g['theta'], *g['args'], **kw)
        assert abs(hl - float(ref['handover_loss'])) <= 1e-5 * abs(float(ref['handover_loss']))
        assert abs(float(hd) - float(ref['handover_dalpha'])) <= 1e-4 * max(abs(float(ref['handover_dalpha'])), np.abs(ref['grad']).max())
        obj = L.compute_loss_objectives(ref['scaled_theta'], *g['args'], g['sensor_size'])
        for k in ('correlations', 'zero_correlations', 'rel_correlations', 'contrasts', 'zero_contrast', 'rel_contrasts'):
            np.testing.assert_allclose(np.asarray(obj[k], dtype=np.float64), ref['obj_' + k], rtol=2e-5, err_msg=k)
        assert set(obj) <= {k[4:] for k in ref if k.startswith('obj_')}          # nothing the reference does not return
    finally:
        L.clear_cache()
