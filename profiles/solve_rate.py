"""windows/s of complete multi-level solves with the native optimizers: n_threads sequences per GPU, one host thread and one CUDA
stream each (the same loop as bench.py's windows_per_s), repeated to show the run-to-run spread.
usage: python profiles/solve_rate.py [--threads 4] [--windows 2] [--repeat 3] [--blocking]"""
import argparse
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from eincm_b200 import losses, plan as P, solver as SV, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='dsec')
ap.add_argument('--threads', type=int, default=4)
ap.add_argument('--windows', type=int, default=2)
ap.add_argument('--repeat', type=int, default=3)
ap.add_argument('--blocking', action='store_true')
ap.add_argument('--group', action='store_true', help='rendezvous the evaluations of the sequences (eincm_group)')
ap.add_argument('--burst', type=int, default=100)
ap.add_argument('--main-thread', action='store_true', help='threads == 1: run the sequence in the main thread')
a = ap.parse_args()
dev = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(dev)
seqs = [synth.make_sequence(a.workload, 1 + a.windows * a.repeat, seed=t) for t in range(a.threads)]       # the bench's drifting sequences
wins = seqs[0]
H, W = wins[0].sensor_size
hpd = wins[0].hparams
N, R = len(wins[0].xs), len(wins[0].edge_ts)
objs = [losses.WindowObjective((H, W), hpd['alpha'], hpd['beta'], hpd['gamma'], hpd['delta'], max_events=N, max_refs=max(R, 3),
                               flags=P.FLAG_BLOCKING_SYNC if a.blocking else 0) for _ in range(a.threads)]
grp = P.Group(a.burst) if a.group else None
if grp is not None:
    for o in objs:
        o.plan.set_group(grp)
sols = [SV.MultipleLevelEINCMSolver(o, backend='native', own_stream=True) for o in objs]
for t, sol in enumerate(sols):
    sol.set_datasample(*seqs[t][0].args())
    sol.solve()
cursor = [1] * a.threads
per_thread = [0.0] * a.threads


def work(t):
    torch.cuda.set_device(dev)
    t0 = time.perf_counter()
    for k in range(a.windows):
        sols[t].set_datasample(*seqs[t][cursor[t]].args())
        cursor[t] += 1
        sols[t].solve()
    per_thread[t] = time.perf_counter() - t0


def probes():
    """GPU probe: 20 device-resident evaluations on one stream (CUDA events); CPU probe: a fixed pure-Python loop."""
    pl = objs[0].plan
    th = torch.from_numpy(synth.theta_test_points(wins[0], (16, 16))['perturbed']).cuda()
    loss = torch.zeros(1, dtype=torch.float64, device='cuda'); grad = torch.zeros((16, 16, 2), dtype=torch.float64, device='cuda')
    hp = objs[0].hparams(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        pl.value_and_grad_device(th, hp, loss, grad)
    e1.record(); torch.cuda.synchronize()
    t0 = time.perf_counter(); x = 0
    for i in range(300000):
        x += i * i
    return e0.elapsed_time(e1) / 20 * 1e3, (time.perf_counter() - t0) * 1e3


for rep in range(a.repeat):
    n0 = sum(o.n_evals for o in objs)
    t0 = time.perf_counter()
    if a.main_thread and a.threads == 1:
        work(0)
    else:
        ths = [threading.Thread(target=work, args=(t,)) for t in range(a.threads)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n_ev = sum(o.n_evals for o in objs) - n0
    print(f'rank {os.environ.get("RANK", "-")} OMP={os.environ.get("OMP_NUM_THREADS", "-")} threads {a.threads} blocking {a.blocking} group {a.group} burst {a.burst}: '
          f'{a.threads * a.windows / dt:.2f} windows/s, {n_ev / (a.threads * a.windows):.0f} evals/window, {dt / n_ev * 1e6:.0f} us wall per eval, '
          f'per-thread s {[round(x, 2) for x in per_thread]}', flush=True)
    ht = [o.plan.host_times(reset=True) for o in objs]
    print('   host per eval (us): launch ' + ' '.join(f'{x[0] / max(x[2], 1) * 1e6:.0f}' for x in ht) + ' | wait ' + ' '.join(f'{x[1] / max(x[2], 1) * 1e6:.0f}' for x in ht), flush=True)
    g, c = probes()
    print(f'   probes: GPU {g:.0f} us per device-resident eval, CPU loop {c:.1f} ms', flush=True)
