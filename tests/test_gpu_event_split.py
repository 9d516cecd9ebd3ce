"""Event-split evaluation of one window on real plans: (a) two plans on ONE GPU with the all-reduce done by hand,
(b) two ranks on two GPUs over NCCL (skipped with fewer than 2 devices)."""
import os
import socket

import numpy as np
import pytest

import eincm_b200.synth as S
from oracle import eincm_oracle as O

pytestmark = pytest.mark.gpu
HP = dict(alpha=20.0, beta=35.0, gamma=0.0025, delta=0.0)


def test_two_split_plans_on_one_gpu_match_oracle():
    import torch
    from eincm_b200 import parallel as PAR, plan as P
    w = S.make_workload('tiny', seed=2)
    th = S.theta_test_points(w, (4, 4))['perturbed']
    hp = P.make_hparams(**HP, cur_pyr_lvl=0)
    plans = []
    for r in range(2):
        p = P.Plan(w.sensor_size, max_events=len(w.xs), max_refs=3, flags=P.FLAG_EVENT_SPLIT)
        p.set_event_split(r, 2)
        p.set_window(*PAR.split_events(w.xs, w.ys, w.ts, 2, r), w.edges, w.edge_ts)
        plans.append(p)
    z = plans[0].zero_iwe() + plans[1].zero_iwe()
    m = torch.maximum(plans[0].event_mask(), plans[1].event_mask())
    for p in plans:
        p.zero_iwe().copy_(z); p.event_mask().copy_(m); p.window_finalize()
    th_d = torch.from_numpy(th).cuda()
    for p in plans:
        p.forward_events(th_d, hp)
    iwe = plans[0].iwe() + plans[1].iwe()
    losses, grads = [], []
    for p in plans:
        p.iwe().copy_(iwe)
        lo = torch.zeros(1, dtype=torch.float64, device='cuda'); g = torch.zeros_like(th_d)
        p.backward(hp, lo, g)
        losses.append(lo); grads.append(g)
    torch.cuda.synchronize()
    l_ref, g_ref = O.value_and_grad(th, *w.args(), **HP, cur_pyr_lvl=0, n_pyr_lvls=5, sensor_size=w.sensor_size)
    for lo in losses:
        assert abs(float(lo[0]) - l_ref) <= 1e-5 * abs(l_ref)
    g = (grads[0] + grads[1]).cpu().numpy()
    assert np.abs(g - g_ref).max() <= 1e-4 * np.abs(g_ref).max()
    # a split plan refuses the fused single-GPU call
    with pytest.raises(P.EincmError):
        plans[0].value_and_grad_host(th, hp)
    for p in plans:
        p.close()


def test_two_split_plans_with_peer_fused_splat_on_one_gpu():
    """Event split with peer access, in-process form: both plans live on one GPU and add their votes to BOTH fixed-point image
    buffers (what the ranks do over NVLink), so no all-reduce of the images is needed and the fused image pass runs on the
    complete image.  Losses are bit-identical on both "ranks" and equal to the unsplit evaluation."""
    import torch
    from eincm_b200 import parallel as PAR, plan as P
    w = S.make_workload('tiny', seed=2)
    th = S.theta_test_points(w, (4, 4))['perturbed']
    hp0 = dict(HP, gamma=0.0)
    hp = P.make_hparams(**hp0, cur_pyr_lvl=0)
    plans = []
    for r in range(2):
        p = P.Plan(w.sensor_size, max_events=len(w.xs), max_refs=3, flags=P.FLAG_EVENT_SPLIT)
        p.set_event_split(r, 2)
        plans.append(p)
    for p in plans:
        p.set_peer_pointers(plans)
    for p in plans:
        p.split_prepare()
    torch.cuda.synchronize()                                    # "barrier"
    for r, p in enumerate(plans):
        p.set_window(*PAR.split_events(w.xs, w.ys, w.ts, 2, r), w.edges, w.edge_ts)
    torch.cuda.synchronize()
    for p in plans:
        p.split_window_images(); p.window_finalize()
    th_d = torch.from_numpy(th).cuda()
    with pytest.raises(P.EincmError):                            # the image buffer still holds the zero-warp votes:
        plans[0].forward_events(th_d, hp)                        # a splat without split_prepare (+ barrier) is refused
    for p in plans:
        p.split_prepare()
    torch.cuda.synchronize()
    for p in plans:
        p.forward_events(th_d, hp)
    torch.cuda.synchronize()
    losses, grads = [], []
    for p in plans:
        lo = torch.zeros(1, dtype=torch.float64, device='cuda'); g = torch.zeros_like(th_d)
        p.backward(hp, lo, g)
        losses.append(lo); grads.append(g)
    torch.cuda.synchronize()
    l_ref, g_ref = O.value_and_grad(th, *w.args(), **hp0, cur_pyr_lvl=0, n_pyr_lvls=5, sensor_size=w.sensor_size)
    assert float(losses[0][0]) == float(losses[1][0])            # identical complete images on every rank
    assert abs(float(losses[0][0]) - l_ref) <= 1e-5 * abs(l_ref)
    g = (grads[0] + grads[1]).cpu().numpy()
    assert np.abs(g - g_ref).max() <= 1e-4 * np.abs(g_ref).max()
    # the unsplit plan sees the same fixed-point image, hence the same objective to the last bit
    single = P.Plan(w.sensor_size, max_events=len(w.xs), max_refs=3)
    single.set_window(*w.args())
    l1, _ = single.value_and_grad_host(th, hp)
    assert l1 == float(losses[0][0])
    np.testing.assert_array_equal(plans[0].iwe().cpu().numpy(), single.iwe().cpu().numpy())
    single.close()
    for p in plans:
        p.close()


def test_two_split_plans_with_fixed_point_allreduce_on_one_gpu():
    """Event split, fixed-point form of the collective (eincm_plan_set_split_fixed_point): the ranks' int64 images are summed (here by
    hand, what dist.all_reduce does over NCCL) and the fused image pass runs on the sum.  Integer sums: the objective is bit-identical
    on both "ranks" and equal to the unsplit evaluation."""
    import torch
    from eincm_b200 import parallel as PAR, plan as P
    w = S.make_workload('tiny', seed=2)
    th = S.theta_test_points(w, (4, 4))['perturbed']
    hp = P.make_hparams(**HP, cur_pyr_lvl=0)                     # gamma != 0: the TV gradient is added by rank 0 alone
    plans = []
    for r in range(2):
        p = P.Plan(w.sensor_size, max_events=len(w.xs), max_refs=3, flags=P.FLAG_EVENT_SPLIT)
        p.set_event_split(r, 2)
        p.set_split_fixed_point(True)
        p.set_window(*PAR.split_events(w.xs, w.ys, w.ts, 2, r), w.edges, w.edge_ts)
        plans.append(p)
    z = plans[0].zero_iwe() + plans[1].zero_iwe()
    m = torch.maximum(plans[0].event_mask(), plans[1].event_mask())
    for p in plans:
        p.zero_iwe().copy_(z); p.event_mask().copy_(m); p.window_finalize()
    th_d = torch.from_numpy(th).cuda()
    losses, grads = [], []
    for rep_ in range(2):                                        # twice: the image buffers are cleared and recycled correctly
        for p in plans:
            p.forward_events(th_d, hp)
        fix = plans[0].iwe_fix() + plans[1].iwe_fix()
        assert fix.dtype == torch.int64
        losses, grads = [], []
        for p in plans:
            p.iwe_fix().copy_(fix)
            lo = torch.zeros(1, dtype=torch.float64, device='cuda'); g = torch.zeros_like(th_d)
            p.backward(hp, lo, g)
            losses.append(lo); grads.append(g)
        torch.cuda.synchronize()
    l_ref, g_ref = O.value_and_grad(th, *w.args(), **HP, cur_pyr_lvl=0, n_pyr_lvls=5, sensor_size=w.sensor_size)
    assert float(losses[0][0]) == float(losses[1][0])
    assert abs(float(losses[0][0]) - l_ref) <= 1e-5 * abs(l_ref)
    g = (grads[0] + grads[1]).cpu().numpy()
    assert np.abs(g - g_ref).max() <= 1e-4 * np.abs(g_ref).max()
    single = P.Plan(w.sensor_size, max_events=len(w.xs), max_refs=3)
    single.set_window(*w.args())
    l1, _ = single.value_and_grad_host(th, hp)
    assert l1 == float(losses[0][0])                             # the same fixed-point image, hence the same objective to the last bit
    single.close()
    for p in plans:
        p.close()


def _worker(rank, world, port, q, mode='nccl'):
    import torch
    import torch.distributed as dist
    from eincm_b200 import parallel as PAR, plan as P
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        w = S.make_workload('mvsec_dt4', seed=5)
        th = S.theta_test_points(w, (16, 16))['perturbed']
        xs, ys, ts = PAR.split_events(w.xs, w.ys, w.ts, world, rank)
        p = P.Plan(w.sensor_size, max_events=len(w.xs), max_refs=5, flags=P.FLAG_EVENT_SPLIT)
        obj = PAR.EventSplitObjective(p, lambda lvl: P.make_hparams(**HP, cur_pyr_lvl=lvl), p2p=mode == 'p2p', fixed_point=mode == 'fixed')
        obj.set_datasample(xs, ys, ts, w.edges, w.edge_ts)
        loss, grad = obj.value_and_grad(torch.from_numpy(th).cuda(), 0)
        loss, grad = obj.value_and_grad(torch.from_numpy(th).cuda(), 0)          # twice: buffers are recycled correctly
        torch.cuda.synchronize()
        q.put((rank, float(loss[0]), grad.cpu().numpy().copy()))
        dist.barrier()
        p.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('mode', ['nccl', 'fixed', 'p2p'])
def test_event_split_two_gpus_nccl(mode):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (gpurun --gpus 2)')
    world = 2
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    w = S.make_workload('mvsec_dt4', seed=5)
    th = S.theta_test_points(w, (16, 16))['perturbed']
    l_ref, g_ref = O.value_and_grad(th, *w.args(), **HP, cur_pyr_lvl=0, n_pyr_lvls=5, sensor_size=w.sensor_size)
    for rank, loss, grad in res:
        assert abs(loss - l_ref) <= 1e-5 * abs(l_ref)
        assert np.abs(grad - g_ref).max() <= 1e-4 * np.abs(g_ref).max()
    if mode in ('p2p', 'fixed'):
        assert res[0][1] == res[1][1]                            # integer all-reduce: bit-identical objective on all ranks
