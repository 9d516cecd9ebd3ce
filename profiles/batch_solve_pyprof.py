"""cProfile of the host side of one lockstep solve of B MVSEC-shaped windows (where the Python time between the batched calls goes).
usage: python profiles/batch_solve_pyprof.py [--batch 256]"""
import argparse, cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eincm_b200 import losses, solver as SV, synth
ap = argparse.ArgumentParser(); ap.add_argument('--batch', type=int, default=256); a = ap.parse_args()
torch.cuda.set_device(0)
seqs = [synth.make_sequence('mvsec_dt4', 3, seed=t) for t in range(a.batch)]
w0 = seqs[0][0]; H, W = w0.sensor_size; hpd = w0.hparams
objs = [losses.WindowObjective((H, W), hpd['alpha'], hpd['beta'], 0.0, hpd['delta'], max_events=len(w0.xs), max_refs=5) for _ in seqs]
lock = SV.BatchedMultipleLevelEINCMSolver(objs)
for k in range(2):
    lock.set_datasamples([s[k].args() for s in seqs]); lock.solve()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
lock.set_datasamples([s[2].args() for s in seqs]); lock.solve()
pr.disable()
print(f'one batch of {a.batch} windows: {time.perf_counter() - t0:.3f} s')
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28); print(s.getvalue()[:6000])
