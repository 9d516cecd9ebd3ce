"""Host-side multi-GPU logic on CPU: sharding helpers, and the event-split collective sequence of
eincm_b200.parallel.EventSplitObjective at world size 2 over gloo (oracle-backed stand-in plan)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import eincm_b200.synth as S
from eincm_b200 import parallel as PAR
from oracle import eincm_oracle as O
from tests._oracle_split_plan import OracleSplitPlan

HP = dict(alpha=20.0, beta=35.0, gamma=0.0025, delta=0.3)


def test_shard_windows_covers_everything_once():
    for world in (1, 2, 4, 8):
        got = sorted(i for r in range(world) for i in PAR.shard_windows(2087, world, r))
        assert got == list(range(2087))


def test_shard_sequences_lpt_dsec():
    counts = [286, 541, 91, 376, 376, 56, 361]       # docs/assets/dsec_extended_evals/*.csv
    for world in (1, 2, 4, 8):
        bins = PAR.shard_sequences_lpt(counts, world)
        assert sorted(i for b in bins for i in b) == list(range(7))
        loads = [sum(counts[i] for i in b) for b in bins]
        assert max(loads) >= max(counts) and max(loads) <= max(max(counts), 2 * sum(counts) / world)
    assert max(sum(counts[i] for i in b) for b in PAR.shard_sequences_lpt(counts, 2)) <= 1050


def test_split_events_is_a_partition():
    w = S.make_workload('tiny', seed=1)
    for world in (2, 3, 8):
        parts = [PAR.split_events(w.xs, w.ys, w.ts, world, r) for r in range(world)]
        np.testing.assert_array_equal(np.concatenate([p[0] for p in parts]), w.xs)
        np.testing.assert_array_equal(np.concatenate([p[2] for p in parts]), w.ts)


@pytest.mark.parametrize('lvl,shape', [(0, (4, 4)), (2, (1, 1))])
def test_split_restatement_equals_value_and_grad(lvl, shape):
    """single process: summing the two halves' partial images / gradients reproduces the un-split oracle"""
    w = S.make_workload('tiny', seed=2)
    th = S.theta_test_points(w, shape)['perturbed']
    kw = dict(HP, cur_pyr_lvl=lvl, n_pyr_lvls=5, sensor_size=w.sensor_size)
    l_ref, g_ref = O.value_and_grad(th, *w.args(), **kw)
    halves = [PAR.split_events(w.xs, w.ys, w.ts, 2, r) for r in range(2)]
    iwes = sum(O.partial_images(th, *h, w.edge_ts, w.sensor_size) for h in halves)
    zero = sum(O.events_to_pdf_frame(h[0], h[1], w.sensor_size) for h in halves)
    mask = O.make_event_mask(w.xs, w.ys, w.sensor_size)
    tot = 0
    for r, h in enumerate(halves):
        l, g = O.split_value_and_grad(th, iwes, zero, mask, *h, w.edges, w.edge_ts, **kw, include_replicated_grad=(r == 0))
        assert l == pytest.approx(l_ref, rel=1e-12)
        tot = tot + g
    assert np.abs(tot - g_ref).max() <= 1e-10 * np.abs(g_ref).max()


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q, fixed=False):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        w = S.make_workload('tiny', seed=2)
        th = S.theta_test_points(w, (4, 4))['perturbed']
        obj = PAR.EventSplitObjective(OracleSplitPlan(w.sensor_size), lambda lvl: dict(HP, cur_pyr_lvl=lvl), fixed_point=fixed)
        obj.set_datasample(*PAR.split_events(w.xs, w.ys, w.ts, world, rank), w.edges, w.edge_ts)
        loss, grad = obj.value_and_grad(torch.from_numpy(th), 0)
        q.put((rank, float(loss[0]), grad.numpy().copy(), obj.n_collectives))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('fixed', [False, True])
def test_event_split_over_gloo_world2(fixed):
    """fixed: the per-evaluation collective runs on the int64 fixed-point images (eincm_plan_set_split_fixed_point): integer sums, the
    same bits on every rank whatever the reduction order."""
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, fixed)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w = S.make_workload('tiny', seed=2)
    th = S.theta_test_points(w, (4, 4))['perturbed']
    l_ref, g_ref = O.value_and_grad(th, *w.args(), **HP, cur_pyr_lvl=0, n_pyr_lvls=5, sensor_size=w.sensor_size)
    for rank, loss, grad, ncoll in res:
        # fixed: the stand-in quantises every cell of the partial images to 2^-21 / 2 pi
        assert loss == pytest.approx(l_ref, rel=1e-6 if fixed else 1e-12)
        assert np.abs(grad - g_ref).max() <= (1e-5 if fixed else 1e-10) * np.abs(g_ref).max()
        assert ncoll == 4          # zero-IWE, mask, IWE, gradient
    if fixed:
        assert res[0][1] == res[1][1]
