"""The objective on windows staged from a shard (eincm_b200/shards.py): same bits as the window staged from its arrays (the image of
warped events is a sum of exactly rounded fixed-point votes, so the pre-tiled order changes nothing), parity against the oracle, the
device ingest chain feeding a shard, and the event split cut at tile boundaries."""
import numpy as np
import pytest

import eincm_b200.synth as S
from oracle import eincm_oracle as O
from oracle import ingest_oracle as G

pytestmark = pytest.mark.gpu
HP = dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.0)


def _kw(win, lvl=1):
    return dict(**HP, cur_pyr_lvl=lvl, n_pyr_lvls=5, sensor_size=win.sensor_size)


@pytest.mark.parametrize('name,n,grid', [('tiny', None, (4, 4)), ('mvsec_dt4', None, (16, 16)), ('dsec', 300_000, (16, 16))])
def test_shard_staged_window_evaluates_to_the_same_bits(tmp_path, name, n, grid):
    from eincm_b200 import plan as P, shards as SH
    wins = [S.make_workload(name, seed=s, n_events=n) for s in (0, 1)]
    path = str(tmp_path / 'w.eshard')
    with SH.ShardWriter(path, wins[0].sensor_size) as wr:
        for w in wins:
            wr.add_window(*w.args())
    rd = SH.ShardReader(path, verify=True)
    hp = P.make_hparams(**HP, cur_pyr_lvl=1)
    p = P.Plan(wins[0].sensor_size, max_events=max(len(w.xs) for w in wins), max_refs=len(wins[0].edge_ts))
    try:
        for k, w in enumerate(wins):
            th = S.theta_test_points(w, grid)['perturbed']
            p.set_window(*w.args())
            la, ga = p.value_and_grad_host(th, hp)
            iwe_a = p.iwe().cpu().numpy().copy()
            sw = rd.stage(p, k)
            assert not np.array_equal(sw.xs, w.xs)                      # the stored order IS another one
            lb, gb = p.value_and_grad_host(th, hp)
            assert la == lb                                             # bit-identical loss
            np.testing.assert_array_equal(iwe_a, p.iwe().cpu().numpy())
            np.testing.assert_allclose(ga, gb, rtol=1e-9, atol=1e-12 * np.abs(ga).max())      # float64 reductions in another order
            if len(w.xs) <= 300_000:
                l_ref, g_ref = O.value_and_grad(th, *w.args(), **_kw(w))
                assert abs(lb - l_ref) <= 1e-5 * abs(l_ref)
                assert np.abs(gb - g_ref).max() <= 1e-4 * np.abs(g_ref).max()
    finally:
        p.close()


def test_device_ingest_to_shard_to_plan(tmp_path):
    """raw stream -> rectify -> fixed-N window -> normalised times (all on the device, bit-exact with the loader restatement) -> shard ->
    plan: the loss equals the oracle's on the loader restatement's arrays."""
    import torch
    from eincm_b200 import dataloaders as D, plan as P, shards as SH
    H, W, n = 48, 64, 30_000
    rng = np.random.default_rng(11)
    x = rng.integers(0, W, n).astype(np.int16); y = rng.integers(0, H, n).astype(np.int16)
    t = np.sort(rng.integers(50_000_000_000, 50_000_000_000 + 200_000, n)).astype(np.int64)
    pol = rng.integers(0, 2, n).astype(bool)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    rm = np.stack([xx + 1.5 * np.sin(yy / 9.0), yy + 1.5 * np.cos(xx / 11.0)], axis=-1).astype(np.float32)
    rx, ry, rt, _ = D.rectify_events(x, y, t, pol, rm, H, W)
    ex, ey, et, _ = G.rectify_events(x, y, t, pol, rm, H, W)
    edges = np.random.default_rng(12).random((3, H, W))
    edge_ts = np.array([0.0, 0.5, 1.0])
    path = str(tmp_path / 'seq.eshard')
    cuts = [(int(rx.numel() * a), int(rx.numel() * b)) for a, b in ((0.1, 0.3), (0.3, 0.5), (0.5, 0.7))]
    with SH.ShardWriter(path, (H, W)) as wr:
        for i0, i1 in cuts:
            ev = (int(rt[i0].item()), int(rt[i1 - 1].item()))
            xs, ys, ts, deficiency = D.stage_window_events(rx, ry, rt, i0, i1, ev, des_n_events=7000)
            wr.add_window(xs.cpu().numpy(), ys.cpu().numpy(), ts.cpu().numpy(), edges, edge_ts, t_start_us=ev[0], t_end_us=ev[1],
                          n_event_deficiency=deficiency)
    rd = SH.ShardReader(path, verify=True)
    assert len(rd) == 3
    p = P.Plan((H, W), max_events=7000, max_refs=3)
    try:
        th = np.random.default_rng(13).normal(0.0, 2.0, (2, 2, 2))
        for k, (i0, i1) in enumerate(cuts):
            w = rd.stage(p, k)
            a, b, d = G.window_event_range(i0, i1, len(ex), 7000)
            assert rd.n_events(k) == 7000 and w.n_event_deficiency == d and (w.t_start_us, w.t_end_us) == (int(et[i0]), int(et[i1 - 1]))
            ts_ref = G.normalize_times(et[a:b].astype(np.uint64), int(et[i0]), int(et[i1 - 1]))
            loss, grad = p.value_and_grad_host(th, P.make_hparams(**HP, cur_pyr_lvl=1))
            l_ref, g_ref = O.value_and_grad(th, ex[a:b], ey[a:b], ts_ref, edges, edge_ts, **HP, cur_pyr_lvl=1, n_pyr_lvls=5, sensor_size=(H, W))
            assert abs(loss - l_ref) <= 1e-5 * abs(l_ref)
            assert np.abs(grad - g_ref).max() <= 1e-4 * np.abs(g_ref).max()
    finally:
        p.close()


@pytest.mark.parametrize('world', [2, 3])
def test_event_split_from_tile_ranges_matches_the_unsplit_window(tmp_path, world):
    """Ranks of the event split read [a, b) of a pre-tiled window (whole source tiles each): summed fixed-point images are bit-identical
    to the unsplit evaluation, the loss is the same on every rank, the rank gradients add up."""
    import torch
    from eincm_b200 import plan as P, shards as SH
    w = S.make_workload('mvsec_dt1', seed=4)
    path = str(tmp_path / 'e.eshard')
    with SH.ShardWriter(path, w.sensor_size) as wr:
        wr.add_window(*w.args())
    rd = SH.ShardReader(path)
    th = S.theta_test_points(w, (4, 4))['perturbed']
    hp = P.make_hparams(**HP, cur_pyr_lvl=0)
    whole = P.Plan(w.sensor_size, max_events=len(w.xs), max_refs=2)
    whole.set_window(*w.args())
    l_whole, g_whole = whole.value_and_grad_host(th, hp)
    plans = []
    for r in range(world):
        p = P.Plan(w.sensor_size, max_events=len(w.xs), max_refs=2, flags=P.FLAG_EVENT_SPLIT)
        p.set_event_split(r, world)
        p.set_split_fixed_point(True)
        a, b = rd.event_range_for_rank(0, r, world)
        assert b > a
        rd.stage(p, 0, event_range=(a, b))
        plans.append(p)
    z = sum(p.zero_iwe() for p in plans)
    m = plans[0].event_mask().clone()
    for p in plans[1:]:
        m = torch.maximum(m, p.event_mask())
    for p in plans:
        p.zero_iwe().copy_(z); p.event_mask().copy_(m); p.window_finalize()
    th_d = torch.from_numpy(th).cuda()
    for p in plans:
        p.forward_events(th_d, hp)
    fix = sum(p.iwe_fix() for p in plans)
    losses, grads = [], []
    for p in plans:
        p.iwe_fix().copy_(fix)
        lo = torch.zeros(1, dtype=torch.float64, device='cuda'); g = torch.zeros_like(th_d)
        p.backward(hp, lo, g)
        losses.append(lo); grads.append(g)
    torch.cuda.synchronize()
    assert all(float(lo[0]) == float(losses[0][0]) for lo in losses)     # every rank holds the same summed image
    assert abs(float(losses[0][0]) - l_whole) <= 1e-12 * abs(l_whole)    # integer sums: the split does not change the images (the float64
    if world == 2:                                                       # sum of the ranks' zero-warp images is exact for two ranks)
        assert float(losses[0][0]) == l_whole
    g = sum(grads).cpu().numpy()
    np.testing.assert_allclose(g, g_whole, rtol=1e-9, atol=1e-12 * np.abs(g_whole).max())
    for p in plans + [whole]:
        p.close()
