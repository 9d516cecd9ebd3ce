"""Parity of the CUDA event-ingest steps (through the C-ABI) against the NumPy restatement of the reference's loader code
(oracle/ingest_oracle.py).  Index / byte work: bit-exact; the float64 time normalisation is one subtraction and one division: exact."""
import numpy as np
import pytest

from oracle import ingest_oracle as G

pytestmark = pytest.mark.gpu


def _stream(n, H, W, seed, distortion=6.0):
    rng = np.random.default_rng(seed)
    x = rng.integers(0, W, n).astype(np.int16); y = rng.integers(0, H, n).astype(np.int16)
    t = np.sort(rng.integers(50_000_000_000, 50_000_000_000 + 3_000_000, n)).astype(np.int64)
    p = rng.integers(0, 2, n).astype(bool)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    rm = np.stack([xx + distortion * np.sin(yy / 37.0) - 2.0, yy + distortion * np.cos(xx / 53.0) - 1.5], axis=-1).astype(np.float32)
    rm[::7, ::5] = np.round(rm[::7, ::5]) + 0.5              # exact ties: np.round goes to even
    return x, y, t, p, rm


@pytest.mark.parametrize('n,H,W', [(1, 8, 8), (4095, 40, 56), (4096, 40, 56), (4097, 40, 56), (300_000, 480, 640), (2_500_000, 480, 640)])
def test_rectify_events_matches_loader_code(n, H, W):
    from eincm_b200 import dataloaders as D
    x, y, t, p, rm = _stream(n, H, W, seed=n % 97)
    rx, ry, rt, rp = D.rectify_events(x, y, t, p, rm, H, W)
    ex, ey, et, ep = G.rectify_events(x, y, t, p, rm, H, W)
    if n > 1000:
        assert 0 < len(ex) < n                                     # some events leave the sensor
    assert rx.numel() == len(ex)
    assert np.array_equal(rx.cpu().numpy(), ex) and np.array_equal(ry.cpu().numpy(), ey)
    assert np.array_equal(rt.cpu().numpy(), et) and np.array_equal(rp.cpu().numpy(), ep)


def test_rectify_empty_and_all_dropped():
    from eincm_b200 import dataloaders as D
    H, W = 6, 7
    rm = np.full((H, W, 2), -5.0, np.float32)
    z = np.zeros(0, np.int16)
    out = D.rectify_events(z, z, np.zeros(0, np.int64), np.zeros(0, bool), rm, H, W)
    assert all(o.numel() == 0 for o in out)
    x = np.array([1, 2, 3], np.int16)
    out = D.rectify_events(x, x, np.arange(3), np.ones(3, bool), rm, H, W)
    assert all(o.numel() == 0 for o in out)


def test_normalize_times_exact():
    from eincm_b200 import dataloaders as D
    rng = np.random.default_rng(3)
    t = np.sort(rng.integers(51_234_567_890, 51_234_567_890 + 101_700, 100_000)).astype(np.int64)
    start, end = 51_234_567_890 + 500, 51_234_567_890 + 100_900
    got = D.normalize_times(t, start, end).cpu().numpy()
    np.testing.assert_array_equal(got, G.normalize_times(t.astype(np.uint64), start, end))
    assert got.min() < 0.0 and got.max() > 1.0                      # the fixed-N window is a superset of the evaluation interval


def test_stage_window_feeds_set_window():
    """rectify -> fixed-N window -> normalised times -> Plan.set_window, everything on the device."""
    import torch
    from eincm_b200 import dataloaders as D, plan as P
    H, W, n = 48, 64, 20_000
    x, y, t, p, rm = _stream(n, H, W, seed=5, distortion=2.0)
    rx, ry, rt, _ = D.rectify_events(x, y, t, p, rm, H, W)
    i0, i1 = int(rx.numel() * 0.4), int(rx.numel() * 0.6)
    ev = (int(rt[i0].item()), int(rt[i1 - 1].item()))
    xs, ys, ts, deficiency = D.stage_window_events(rx, ry, rt, i0, i1, ev, des_n_events=6000)
    assert xs.numel() == 6000 and deficiency == 6000 - (i1 - i0)
    ex, ey, et, _ = G.rectify_events(x, y, t, p, rm, H, W)
    a, b, d = G.window_event_range(i0, i1, len(ex), 6000)
    assert np.array_equal(xs.cpu().numpy(), ex[a:b]) and np.array_equal(ys.cpu().numpy(), ey[a:b])
    np.testing.assert_array_equal(ts.cpu().numpy(), G.normalize_times(et[a:b].astype(np.uint64), *ev))
    pl = P.Plan((H, W), max_events=6000, max_refs=3)
    edges = torch.rand((3, H, W), dtype=torch.float64, device='cuda')
    pl.set_window(xs, ys, ts, edges, np.array([0.0, 0.5, 1.0]))
    pl.close()


def test_mvsec_crop_matches_loader_code():
    from eincm_b200 import dataloaders as D
    rng = np.random.default_rng(9)
    n = 50_000
    xs = rng.integers(0, 346, n).astype(np.int16); ys = rng.integers(0, 260, n).astype(np.int16)
    ts = np.sort(rng.uniform(1.5e9, 1.5e9 + 30.0, n)); ps = rng.integers(0, 2, n).astype(bool)
    x, y, t, p = D.crop_events(xs, ys, ts, ps)
    ex, ey, et, ep = G.crop_events(xs, ys, ts, ps)
    assert 0 < len(ex) < n
    assert np.array_equal(x.cpu().numpy(), ex) and np.array_equal(y.cpu().numpy(), ey)
    assert np.array_equal(t.cpu().numpy(), et) and np.array_equal(p.cpu().numpy(), ep)       # float64 timestamps pass through bit for bit
    assert int(x.max()) <= 335 and int(y.max()) <= 255
