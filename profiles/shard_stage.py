"""Staging cost of one DSEC-shaped window (2 M events, 3 edge maps) from a window shard: file -> mapping -> device -> eincm_plan_set_window,
and the device time of set_window alone for events in time order against the shard's pre-tiled order.  Prints one JSON line."""
import json, os, sys, tempfile, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eincm_b200 import plan as P, shards as SH, synth as S

name = sys.argv[1] if len(sys.argv) > 1 else 'dsec'
wins = [S.make_workload(name, seed=s) for s in range(3)]
d = tempfile.mkdtemp()
path = os.path.join(d, 'seq.eshard')
t0 = time.perf_counter()
with SH.ShardWriter(path, wins[0].sensor_size) as wr:
    for w in wins:
        wr.add_window(*w.args())
t_write = (time.perf_counter() - t0) / len(wins)
rd = SH.ShardReader(path, verify=True)
p = P.Plan(wins[0].sensor_size, max_events=len(wins[0].xs), max_refs=len(wins[0].edge_ts))

def dev_time(xs, ys, ts, edges, ets, reps=10):
    xs, ys = torch.from_numpy(np.ascontiguousarray(xs)).cuda(), torch.from_numpy(np.ascontiguousarray(ys)).cuda()
    ts, edges = torch.from_numpy(np.ascontiguousarray(ts)).cuda(), torch.from_numpy(np.ascontiguousarray(edges)).cuda()
    for _ in range(3):
        p.set_window(xs, ys, ts, edges, ets)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        p.set_window(xs, ys, ts, edges, ets)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

w0 = wins[0]
sw = rd.window(0)
ms_time_order = dev_time(w0.xs, w0.ys, w0.ts, w0.edges, w0.edge_ts)
ms_tile_order = dev_time(sw.xs, sw.ys, sw.ts, sw.edges, sw.edge_ts)
# whole path from the file (page cache warm), host clock around a synchronised call
best = 1e9
for rep in range(5):
    for k in range(len(wins)):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rd.stage(p, k); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
hp = P.make_hparams(w0.hparams['alpha'], w0.hparams['beta'], 0.0, 0.0, 1)
th = S.theta_test_points(w0, (16, 16))['perturbed']
p.set_window(*w0.args()); la, _ = p.value_and_grad_host(th, hp)
rd.stage(p, 0); lb, _ = p.value_and_grad_host(th, hp)
off, nbytes = rd.payload_range(0)
print(json.dumps({'workload': name, 'events': len(w0.xs), 'payload_mb': round(nbytes / 1e6, 2), 'write_ms_per_window': round(1e3 * t_write, 1),
                  'set_window_device_ms': {'time_order': round(ms_time_order, 4), 'tile_major': round(ms_tile_order, 4)},
                  'stage_from_shard_ms': round(1e3 * best, 3), 'loss_identical': la == lb}))
