"""Known answers for the NumPy restatement of the reference's event ingest (oracle/ingest_oracle.py) and, without a GPU, the host
arithmetic of the library (eincm_window_event_range needs no device)."""
import numpy as np
import pytest

from oracle import ingest_oracle as G


def test_rectify_known_answer():
    H, W = 4, 5
    rm = np.zeros((H, W, 2), np.float32)
    rm[..., 0] = np.arange(W)[None, :] + 0.5          # x + 0.5: ties round to even
    rm[..., 1] = np.arange(H)[:, None] - 1.2          # y - 1.2: row 0 leaves the sensor
    x = np.array([0, 1, 2, 3, 4, 2], np.int16); y = np.array([1, 1, 2, 3, 3, 0], np.int16)
    t = np.arange(6, dtype=np.int64) * 10; p = np.array([1, 0, 1, 0, 1, 1], bool)
    rx, ry, rt, rp = G.rectify_events(x, y, t, p, rm, H, W)
    # x + 0.5 -> 0.5->0, 1.5->2, 2.5->2, 3.5->4, 4.5->4 ; y - 1.2 -> -0.2->-0 (kept, 0), 0.8->1, 1.8->2; last event: y=0 -> -1.2 -> -1 dropped
    assert rx.tolist() == [0, 2, 2, 4, 4] and ry.tolist() == [0, 0, 1, 2, 2]
    assert rt.tolist() == [0, 10, 20, 30, 40] and rp.tolist() == [True, False, True, False, True]
    assert rx.dtype == np.int16


@pytest.mark.parametrize('a,b,n,des,latest,want', [
    (100, 200, 1000, None, False, (100, 200, 0)),
    (100, 200, 1000, 100, False, (100, 200, 0)),
    (100, 200, 1000, 151, False, (74, 225, 51)),        # ceil(25.5) = 26 before, floor = 25 after
    (10, 20, 25, 100, False, (0, 25, 90)),              # clamped at both ends
    (100, 200, 1000, 60, False, (100, 160, -40)),
    (100, 200, 1000, 60, True, (140, 200, -40)),
])
def test_window_event_range(a, b, n, des, latest, want):
    assert G.window_event_range(a, b, n, des, latest) == want
    from eincm_b200 import dataloaders
    assert dataloaders.window_event_range(a, b, n, des, latest) == want          # host arithmetic of the library: no device needed


def test_normalize_times_known_answer():
    ts = np.array([1_000_000, 1_050_000, 1_100_000], np.uint64)
    out = G.normalize_times(ts, 1_000_000, 1_100_000)
    assert out.dtype == np.float64
    np.testing.assert_array_equal(out, np.array([0.0, 50000.0, 100000.0]) / (100000.0 + 2.220446049250313e-16))


def test_mvsec_crop_known_answer():
    xs = np.array([4, 5, 340, 341, 100], np.int64); ys = np.array([10, 1, 2, 257, 258], np.int64)
    ts = np.arange(5, dtype=np.float64); ps = np.array([1, 1, 0, 0, 1])
    x, y, t, p = G.crop_events(xs, ys, ts, ps)
    # x - 5 in [0, 336) and y - 2 in [0, 256): only (340, 2) -> (335, 0) and (341, 257) -> x = 336 is out; (100, 258) -> y = 256 is out
    assert x.tolist() == [335] and y.tolist() == [0] and t.tolist() == [2.0] and p.tolist() == [False]


def test_window_event_range_randomised_against_library():
    """eincm_window_event_range is host arithmetic (no device): compared with the loader's lines on random index ranges."""
    from hypothesis import given, settings, strategies as st
    from eincm_b200 import dataloaders

    @settings(max_examples=300, deadline=None, derandomize=True)
    @given(n=st.integers(0, 5_000_000), data=st.data())
    def check(n, data):
        a = data.draw(st.integers(0, n))
        b = data.draw(st.integers(a, n))
        des = data.draw(st.one_of(st.none(), st.integers(1, 6_000_000)))
        latest = data.draw(st.booleans())
        assert dataloaders.window_event_range(a, b, n, des, latest) == G.window_event_range(a, b, n, des, latest)

    check()
