"""Evaluation-path contracts that are not plain parity: several pyramid-level shapes against the oracle through the host calls (with
the handover gradient), device-resident evaluations queued back to back on one stream (programmatic dependent launches chain the five
kernels of consecutive evaluations), flows that push patches over every border and beyond one shared-memory window, and the
stream-ordering contract of set_window / window_finalize."""
import numpy as np
import pytest

import eincm_b200.synth as S
from oracle import eincm_oracle as O

pytestmark = pytest.mark.gpu

OBJ_RTOL = 1e-5      # BASELINE.json north_star
GRAD_RTOL = 1e-4     # BASELINE.json north_star


def _rel_inf(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


@pytest.mark.parametrize('name,shape', [('tiny', (1, 1)), ('tiny', (2, 2)), ('tiny', (3, 4)), ('mvsec_dt4', (16, 16)), ('mvsec_raw_dt4', (16, 16)),
                                        ('mvsec_dt1', (8, 8)), ('ecd', (8, 8)), ('ecd', (16, 16))])
def test_levels_and_handover_against_oracle(name, shape):
    from eincm_b200 import plan as P
    win = S.make_workload(name, seed=3)
    hp = win.hparams
    hpc = P.make_hparams(hp['alpha'], hp['beta'], 0.0, 0.0, 1)
    pts = S.theta_test_points(win, shape)
    p = P.Plan(win.sensor_size, max_events=len(win.xs), max_refs=max(3, len(win.edge_ts)))
    try:
        p.set_window(*win.args())
        for th in (pts['perturbed'], pts['zero'], pts['truth']):
            la, ga = p.value_and_grad_host(th, hpc)
            l_ref, g_ref = O.value_and_grad(th, *win.args(), hp['alpha'], hp['beta'], 0.0, 0.0, 1, 5, win.sensor_size)
            assert abs(la - l_ref) <= OBJ_RTOL * abs(l_ref)
            assert _rel_inf(ga, g_ref) <= GRAD_RTOL
        # handover: d / d alpha = <grad(theta_ho), prev - theta>, delivered by the same kernel
        a0 = 0.3
        l_ho, da = p.handover_value_and_grad_host(a0, pts['truth'], pts['perturbed'], hpc)
        l_ref, da_ref = O.handover_value_and_grad(a0, pts['truth'], pts['perturbed'], *win.args(), hp['alpha'], hp['beta'], 0.0, 0.0, 1, 5, win.sensor_size)
        assert abs(l_ho - l_ref) <= OBJ_RTOL * abs(l_ref)
        assert abs(da - da_ref) <= GRAD_RTOL * abs(da_ref)
    finally:
        p.close()


def test_device_operands_and_repeated_evaluations():
    """Device-resident form (eincm_value_and_grad): the caller's gradient buffer is the accumulator; evaluations queued back to
    back on one stream (programmatic dependent launches chain them) give the same answer as one at a time."""
    import torch
    from eincm_b200 import plan as P
    win = S.make_workload('tiny', seed=5)
    hp = win.hparams
    hpc = P.make_hparams(hp['alpha'], hp['beta'], 0.0, 0.0, 0)
    ths = [S.theta_test_points(win, (4, 4), seed=s)['perturbed'] for s in range(6)]
    p = P.Plan(win.sensor_size, max_events=len(win.xs), max_refs=3)
    try:
        p.set_window(*win.args())
        ref = [p.value_and_grad_host(th, hpc) for th in ths]
        th_d = [torch.from_numpy(th).cuda() for th in ths]
        losses = [torch.zeros(1, dtype=torch.float64, device='cuda') for _ in ths]
        grads = [torch.full((4, 4, 2), 7.0, dtype=torch.float64, device='cuda') for _ in ths]      # garbage: the pass must clear it
        for rep in range(3):
            for k in range(len(ths)):
                p.value_and_grad_device(th_d[k], hpc, losses[k], grads[k])
        torch.cuda.synchronize()
        for k in range(len(ths)):
            assert float(losses[k][0]) == ref[k][0]
            assert _rel_inf(grads[k].cpu().numpy(), ref[k][1]) <= 1e-9
    finally:
        p.close()


def test_large_flow_and_border_wrap():
    """Flows that push patches over every border (wrap / drop index rule) and rectangles beyond one shared-memory window."""
    from eincm_b200 import plan as P
    win = S.make_workload('ecd', seed=1, n_events=20_000)          # 176 x 240
    hp = win.hparams
    hpc = P.make_hparams(hp['alpha'], hp['beta'], 0.0, 0.0, 2)
    th = np.full((2, 2, 2), 37.0)
    th[0, 0] = (-81.0, 95.0)
    th[1, 1] = (140.0, -20.0)
    p = P.Plan(win.sensor_size, max_events=len(win.xs), max_refs=max(3, len(win.edge_ts)))
    try:
        p.set_window(*win.args())
        la, ga = p.value_and_grad_host(th, hpc)
        l_ref, g_ref = O.value_and_grad(th, *win.args(), hp['alpha'], hp['beta'], 0.0, 0.0, 2, 5, win.sensor_size)
        assert abs(la - l_ref) <= OBJ_RTOL * abs(l_ref) and _rel_inf(ga, g_ref) <= GRAD_RTOL
    finally:
        p.close()


def test_evaluation_on_another_stream_waits_for_window_finalize():
    """eincm_window_finalize is asynchronous on the caller's stream; an evaluation enqueued on ANOTHER stream must
    not overtake its kernels.  A long kernel occupies the staging stream while the evaluation is enqueued on a second stream."""
    import torch
    from eincm_b200 import plan as P
    win = S.make_workload('tiny', seed=7)
    hp = win.hparams
    hpc = P.make_hparams(hp['alpha'], hp['beta'], 0.0, 0.0, 1)
    th = S.theta_test_points(win, (4, 4))['perturbed']
    ref = P.Plan(win.sensor_size, max_events=len(win.xs), max_refs=3)
    ref.set_window(*win.args())
    l_ref, g_ref = ref.value_and_grad_host(th, hpc)
    ref.close()
    p = P.Plan(win.sensor_size, max_events=len(win.xs), max_refs=3, flags=P.FLAG_EVENT_SPLIT)
    try:
        p.set_event_split(0, 1)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        p.set_window(*win.args(), stream=s1)
        th_d = torch.from_numpy(th).cuda()
        loss = torch.zeros(1, dtype=torch.float64, device='cuda'); grad = torch.zeros_like(th_d)
        torch.cuda.synchronize()
        with torch.cuda.stream(s1):
            torch.cuda._sleep(400_000_000)                          # ~0.2 s: the finalize kernels queue behind it
        p.window_finalize(stream=s1)
        p.forward_events(th_d, hpc, stream=s2)
        p.backward(hpc, loss, grad, stream=s2)
        torch.cuda.synchronize()
        assert float(loss[0]) == pytest.approx(l_ref, rel=1e-12)
        assert _rel_inf(grad.cpu().numpy(), g_ref) <= 1e-6
    finally:
        p.close()
