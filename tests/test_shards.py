"""Window shards (eincm_b200/shards.py): byte-exact round trips, the pre-tiled order and its tile counts, ragged
and empty windows, corruption / truncation / unclosed files, and the rank cuts of both sharding modes.  CPU only (file plumbing); the
objective on a shard-staged window is tests/test_gpu_shards.py."""
import os

import numpy as np
import pytest

from eincm_b200 import shards as SH
from eincm_b200 import synth as S
from oracle import eincm_oracle as O


H, W = 40, 56


def _win(n, R, seed):
    rng = np.random.default_rng(seed)
    xs = rng.integers(0, W, n).astype(np.int16); ys = rng.integers(0, H, n).astype(np.int16)
    ts = np.sort(rng.uniform(-0.1, 1.1, n))
    edges = rng.random((R, H, W))
    return xs, ys, ts, edges, np.linspace(0.0, 1.0, R)


def _write(path, wins, **kw):
    with SH.ShardWriter(path, (H, W), **kw) as wr:
        for k, (xs, ys, ts, edges, ets) in enumerate(wins):
            assert wr.add_window(xs, ys, ts, edges if kw.get('store_edges', True) else None, ets, t_start_us=1000 * k, t_end_us=1000 * k + 999,
                                 n_event_deficiency=k) == k


@pytest.mark.parametrize('tile_major', [True, False])
def test_round_trip_ragged_windows(tmp_path, tile_major):
    wins = [_win(5000, 3, 0), _win(0, 1, 1), _win(1, 2, 2), _win(4097, 5, 3)]
    path = str(tmp_path / 'a.eshard')
    _write(path, wins, tile_major=tile_major)
    rd = SH.ShardReader(path, verify=True)
    assert len(rd) == 4 and rd.sensor_size == (H, W) and rd.tile_major == tile_major
    assert os.path.getsize(path) % 8 == 0
    for k, (xs, ys, ts, edges, ets) in enumerate(wins):
        w = rd.window(k)
        order = SH.tile_major_order(xs, ys, (H, W))[0] if tile_major else np.arange(len(xs))
        np.testing.assert_array_equal(w.xs, xs[order]); np.testing.assert_array_equal(w.ys, ys[order])
        np.testing.assert_array_equal(w.ts, ts[order])                       # bit-exact: float64 stored as is
        np.testing.assert_array_equal(w.edges, edges); np.testing.assert_array_equal(w.edge_ts, ets)
        assert w.xs.dtype == np.int16 and w.ts.dtype == np.float64 and w.edges.shape == (len(ets), H, W)
        assert (w.t_start_us, w.t_end_us, w.n_event_deficiency) == (1000 * k, 1000 * k + 999, k)
        assert rd.n_events(k) == len(xs)
        off, nbytes = rd.payload_range(k)
        assert off % 64 == 0 and nbytes % 64 == 0
        if tile_major:
            assert w.tile_counts.dtype == np.uint32 and int(w.tile_counts.sum()) == len(xs)
        else:
            assert w.tile_counts is None
    with pytest.raises(IndexError):
        rd.window(4)


def test_pre_tiled_order_and_counts(tmp_path):
    xs, ys, ts, edges, ets = _win(20000, 1, 7)
    path = str(tmp_path / 't.eshard')
    _write(path, [(xs, ys, ts, edges, ets)])
    rd = SH.ShardReader(path)
    w = rd.window(0)
    ty, tx = rd.tiles
    assert (ty, tx) == (3, 4)                                               # 40 x 56 in 16 x 16 tiles, ragged last row / column
    key = (w.ys.astype(int) // 16) * tx + w.xs.astype(int) // 16
    assert np.all(np.diff(key) >= 0)                                        # tile-major
    np.testing.assert_array_equal(w.tile_counts, np.bincount(key, minlength=ty * tx))
    ends = np.cumsum(w.tile_counts)
    for t in range(ty * tx):                                                # the time order survives inside a tile (stable)
        seg = w.ts[ends[t] - w.tile_counts[t]:ends[t]]
        assert np.all(np.diff(seg) >= 0)
    # same multiset of events
    a = np.lexsort((ts, ys, xs)); b = np.lexsort((w.ts, w.ys, w.xs))
    np.testing.assert_array_equal(xs[a], w.xs[b]); np.testing.assert_array_equal(ys[a], w.ys[b]); np.testing.assert_array_equal(ts[a], w.ts[b])


def test_objective_does_not_depend_on_the_stored_order(tmp_path):
    """The pre-tiled order is a permutation of the events: the reference objective is a sum over events (oracle, float64: to rounding)."""
    win = S.make_window(H, W, 6000, edge_ts=(0.0, 0.5, 1.0), seed=3, n_segments=12, flow_mag=4.0)
    path = str(tmp_path / 'o.eshard')
    with SH.ShardWriter(path, (H, W)) as wr:
        wr.add_window(*win.args())
    w = SH.ShardReader(path).window(0)
    th = S.theta_test_points(win, (2, 2))['perturbed']
    kw = dict(alpha=20.0, beta=35.0, gamma=0.0, delta=0.0, cur_pyr_lvl=1, n_pyr_lvls=5, sensor_size=(H, W))
    la, ga = O.value_and_grad(th, *win.args(), **kw)
    lb, gb = O.value_and_grad(th, np.asarray(w.xs), np.asarray(w.ys), np.asarray(w.ts), np.asarray(w.edges), w.edge_ts, **kw)
    assert abs(la - lb) <= 1e-12 * abs(la)
    np.testing.assert_allclose(ga, gb, rtol=1e-9, atol=1e-12 * np.abs(ga).max())


def test_no_edges_shard(tmp_path):
    xs, ys, ts, edges, ets = _win(300, 2, 5)
    path = str(tmp_path / 'n.eshard')
    _write(path, [(xs, ys, ts, edges, ets)], store_edges=False)
    rd = SH.ShardReader(path, verify=True)
    w = rd.window(0)
    assert w.edges is None and len(w.xs) == 300
    with pytest.raises(SH.ShardError):
        rd.stage(object(), 0)


def test_corruption_truncation_and_unclosed_files_are_refused(tmp_path):
    wins = [_win(1000, 2, 0), _win(700, 2, 1)]
    path = str(tmp_path / 'c.eshard')
    _write(path, wins)
    raw = bytearray(open(path, 'rb').read())
    # a flipped payload byte: found by verify, by the window it belongs to
    bad = str(tmp_path / 'bad.eshard')
    off, nbytes = SH.ShardReader(path).payload_range(1)
    flipped = bytearray(raw); flipped[off + 17] ^= 0x40
    open(bad, 'wb').write(flipped)
    rd = SH.ShardReader(bad)
    rd.verify(0)
    with pytest.raises(SH.ShardError, match='checksum'):
        rd.verify(1)
    with pytest.raises(SH.ShardError, match='checksum'):
        SH.ShardReader(bad, verify=True)
    # truncated
    open(bad, 'wb').write(raw[:-8])
    with pytest.raises(SH.ShardError, match='truncated'):
        SH.ShardReader(bad)
    open(bad, 'wb').write(raw[:40])
    with pytest.raises(SH.ShardError):
        SH.ShardReader(bad)
    # not a shard
    open(bad, 'wb').write(b'NOTASHRD' + bytes(raw[8:]))
    with pytest.raises(SH.ShardError, match='magic'):
        SH.ShardReader(bad)
    # a writer that died before close leaves a header without index
    with pytest.raises(RuntimeError):
        with SH.ShardWriter(bad, (H, W)) as wr:
            wr.add_window(*wins[0])
            raise RuntimeError('writer interrupted')
    with pytest.raises(SH.ShardError, match='not closed'):
        SH.ShardReader(bad)
    # an index record that points outside the payload area
    rec0 = len(raw) - 2 * (56 + 8 * 8)
    broken = bytearray(raw); broken[rec0:rec0 + 8] = (len(raw)).to_bytes(8, 'little')
    open(bad, 'wb').write(broken)
    with pytest.raises(SH.ShardError, match='index record 0'):
        SH.ShardReader(bad)


def test_writer_refuses_windows_it_cannot_store(tmp_path):
    xs, ys, ts, edges, ets = _win(10, 2, 0)
    wr = SH.ShardWriter(str(tmp_path / 'w.eshard'), (H, W), r_max=2)
    with pytest.raises(SH.ShardError, match='outside the sensor'):
        wr.add_window(xs + W, ys, ts, edges, ets)
    with pytest.raises(SH.ShardError, match='outside the sensor'):
        wr.add_window(xs, ys - H, ts, edges, ets)
    with pytest.raises(SH.ShardError, match='same length'):
        wr.add_window(xs, ys[:-1], ts, edges, ets)
    with pytest.raises(SH.ShardError, match='shape'):
        wr.add_window(xs, ys, ts, edges[:, :-1], ets)
    with pytest.raises(SH.ShardError, match='reference times'):
        wr.add_window(xs, ys, ts, np.zeros((3, H, W)), [0.0, 0.5, 1.0])
    with pytest.raises(SH.ShardError, match='reference times'):
        wr.add_window(xs, ys, ts, np.zeros((0, H, W)), [])
    with pytest.raises(SH.ShardError, match='required'):
        wr.add_window(xs, ys, ts, None, ets)
    assert wr.add_window(xs, ys, ts, edges, ets) == 0                       # the refusals wrote nothing
    wr.close(); wr.close()
    with pytest.raises(SH.ShardError, match='closed'):
        wr.add_window(xs, ys, ts, edges, ets)
    assert len(SH.ShardReader(wr.path, verify=True)) == 1
    with pytest.raises(SH.ShardError):
        SH.ShardWriter(str(tmp_path / 'x.eshard'), (40000, 10))
    with pytest.raises(SH.ShardError):
        SH.ShardWriter(str(tmp_path / 'x.eshard'), (H, W), r_max=9)


@pytest.mark.parametrize('world', [1, 2, 3, 8, 64])
def test_rank_cuts(tmp_path, world):
    wins = [_win(9000, 1, 0), _win(0, 1, 1), _win(5, 1, 2)] + [_win(100, 1, 3 + k) for k in range(8)]
    path = str(tmp_path / 'r.eshard')
    _write(path, wins)
    rd = SH.ShardReader(path)
    # window sharding: contiguous blocks (the handover chains consecutive windows), every window exactly once
    got = [i for r in range(world) for i in rd.windows_for_rank(r, world)]
    assert got == list(range(len(wins)))
    # event split: the ranges partition the window, cut at tile boundaries, balanced to within the largest tile
    for i in range(len(wins)):
        n = rd.n_events(i)
        counts = rd.window(i).tile_counts
        bounds = set(np.concatenate([[0], np.cumsum(counts)]).tolist())
        prev_b = 0
        for r in range(world):
            a, b = rd.event_range_for_rank(i, r, world)
            assert a == prev_b and a <= b and a in bounds and b in bounds
            assert b - a <= n // world + 1 + 2 * int(counts.max())
            prev_b = b
        assert prev_b == n


def test_event_split_without_tile_order_is_the_slab_split(tmp_path):
    from eincm_b200 import parallel as PAR
    xs, ys, ts, edges, ets = _win(1001, 1, 0)
    path = str(tmp_path / 's.eshard')
    _write(path, [(xs, ys, ts, edges, ets)], tile_major=False)
    rd = SH.ShardReader(path)
    for r in range(3):
        a, b = rd.event_range_for_rank(0, r, 3)
        np.testing.assert_array_equal(rd.window(0).xs[a:b], PAR.split_events(xs, ys, ts, 3, r)[0])
