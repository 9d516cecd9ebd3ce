"""The XLA FFI handler bodies (integration/xla_ffi/eincm_xla_ffi.cc, compiled unchanged against a stand-in for the FFI header) driven
the way XLA drives them: an eager set_window call per window with every operand - edge_ts included - a DEVICE buffer, then
evaluations that carry only theta and the token.  Two consecutive windows of IDENTICAL shape (the reference pads every window to
a fixed N, so the jitted executable and the plan are re-used): the second window's loss and gradient must be the second
window's (the first wrapper froze the window identity into the executable)."""
import ctypes as C

import numpy as np
import pytest

import eincm_b200.synth as S
from tests import _ffi_harness

pytestmark = pytest.mark.gpu


def _dev(t, a):
    return t.from_numpy(np.ascontiguousarray(a)).cuda()


def test_two_windows_of_identical_shape_through_the_handlers():
    import torch
    from eincm_b200 import plan as P
    lib = _ffi_harness.load()
    assert lib.ffi_binders_ok() == 1
    torch.cuda.set_device(0)
    wins = [S.make_workload('mvsec_dt4', seed=s) for s in (1, 2)]
    assert len(wins[0].xs) == len(wins[1].xs)
    H, W = wins[0].sensor_size
    hp = wins[0].hparams
    shape = (4, 4)
    msg = C.create_string_buffer(512)
    stream = torch.cuda.current_stream().cuda_stream
    token = torch.zeros(2, dtype=torch.int64, device='cuda')
    loss = torch.zeros((), dtype=torch.float64, device='cuda')
    grad = torch.zeros(shape + (2,), dtype=torch.float64, device='cuda')
    # before any window: an error, not a crash
    th0 = _dev(torch, np.zeros(shape + (2,)))
    rc = lib.ffi_value_and_grad(stream, 0, th0.data_ptr(), shape[0], shape[1], token.data_ptr(), hp['alpha'], hp['beta'], 0.0, 0.0, 1, 5, 7,
                                loss.data_ptr(), grad.data_ptr(), msg, 512)
    assert rc != 0 and b'before eincm_set_window' in msg.value
    got, want = [], []
    for k, win in enumerate(wins):
        xs, ys, ts, edges, edge_ts = win.args()
        ops = [_dev(torch, np.asarray(xs, np.int16)), _dev(torch, np.asarray(ys, np.int16)), _dev(torch, np.asarray(ts, np.float64)),
               _dev(torch, np.asarray(edges, np.float64)), _dev(torch, np.asarray(edge_ts, np.float64))]
        R = len(edge_ts)
        rc = lib.ffi_set_window(stream, 0, ops[0].data_ptr(), ops[1].data_ptr(), ops[2].data_ptr(), len(xs), ops[3].data_ptr(), R, H, W, ops[4].data_ptr(),
                                7, token.data_ptr(), msg, 512)
        assert rc == 0, msg.value
        assert token.cpu().tolist() == [7, k + 1]                       # {slot, generation}: a run-time value
        theta = S.theta_test_points(win, shape)['perturbed']
        th = _dev(torch, theta)
        for _ in range(2):                                              # the cached executable is called again and again
            rc = lib.ffi_value_and_grad(stream, 0, th.data_ptr(), shape[0], shape[1], token.data_ptr(), hp['alpha'], hp['beta'], 0.0, 0.0, 1, 5, 7,
                                        loss.data_ptr(), grad.data_ptr(), msg, 512)
            assert rc == 0, msg.value
        torch.cuda.synchronize()
        got.append((float(loss.item()), grad.cpu().numpy().copy()))
        p = P.Plan((H, W), max_events=len(xs), max_refs=max(3, R))
        try:
            p.set_window(*win.args())
            want.append(p.value_and_grad_host(theta, P.make_hparams(hp['alpha'], hp['beta'], 0.0, 0.0, 1)))
        finally:
            p.close()
    for (lg, gg), (lw, gw) in zip(got, want):
        assert lg == lw                                                 # same kernels, integer votes: bit-identical objective
        assert np.abs(gg - gw).max() <= 1e-9 * np.abs(gw).max()
    assert got[0][0] != got[1][0]                                       # and the two windows really differ


def test_handler_reports_bad_operands():
    import torch
    lib = _ffi_harness.load()
    torch.cuda.set_device(0)
    msg = C.create_string_buffer(512)
    d = torch.zeros(16, dtype=torch.float64, device='cuda')
    tok = torch.zeros(2, dtype=torch.int64, device='cuda')
    # theta with a last dimension that is not 2 cannot be expressed through the harness; a failing staging call can: event outside the sensor
    xs = torch.full((8,), 1000, dtype=torch.int16, device='cuda')
    ys = torch.zeros(8, dtype=torch.int16, device='cuda')
    ts = torch.linspace(0, 1, 8, dtype=torch.float64, device='cuda')
    edges = torch.zeros((1, 32, 32), dtype=torch.float64, device='cuda')
    ets = torch.zeros(1, dtype=torch.float64, device='cuda')
    rc = lib.ffi_set_window(torch.cuda.current_stream().cuda_stream, 0, xs.data_ptr(), ys.data_ptr(), ts.data_ptr(), 8, edges.data_ptr(), 1, 32, 32,
                            ets.data_ptr(), 9, tok.data_ptr(), msg, 512)
    assert rc != 0 and b'eincm_plan_set_window' in msg.value
    del d
