"""Does a complete multi-level solve find the flow?  Per pyramid level: status, iterations, evaluations, loss, max|theta| and the
mean end-point error of the up-scaled theta against the truth flow of the synthetic window (both backends).
usage: python profiles/solve_quality.py [--workload dsec] [--windows 2]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from eincm_b200 import losses, solver as SV, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='dsec')
ap.add_argument('--windows', type=int, default=2)
ap.add_argument('--segments', type=int, default=None)
ap.add_argument('--backends', default='scipy,native')
a = ap.parse_args()
torch.cuda.set_device(0)
seq = synth.make_sequence(a.workload, a.windows, seed=0, **({'n_segments': a.segments} if a.segments else {}))
H, W = seq[0].sensor_size
hpd = seq[0].hparams


def dense(theta):
    return synth._bilinear_field(np.asarray(theta), H, W)


for backend in a.backends.split(','):
    obj = losses.WindowObjective((H, W), hpd['alpha'], hpd['beta'], hpd['gamma'], hpd['delta'], max_events=len(seq[0].xs), max_refs=3)
    sol = SV.MultipleLevelEINCMSolver(obj, backend=backend)
    for k, win in enumerate(seq):
        truth = dense(win.truth_theta)
        sol.set_datasample(*win.args())
        n0 = obj.n_evals
        out = sol.solve()
        print(f'[{backend}] window {k}: {obj.n_evals - n0} evaluations; truth max|flow| {np.abs(truth).max():.1f} px')
        for lvl in reversed(range(5)):
            key = f'pyr_lvl_{lvl}'
            st = out['theta_opt_state_pyr'][key]
            th = out['final_theta_pyr'][key]
            aee = float(np.sqrt(((dense(th) - truth) ** 2).sum(-1)).mean())
            print(f'    level {lvl}: status {st.status} nit {st.iter_num:3d} nfev {st.n_evals:4d} loss {st.fun_val:12.4f}  max|theta| {np.abs(th).max():9.2f} px  '
                  f'AEE vs truth {aee:8.3f} px  handover weight {out["final_handover_weight_pyr"].get(key)}')
    obj.close()
