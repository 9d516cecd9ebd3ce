"""Tiny driver for ncu captures: stages DSEC-shaped windows and runs objective+gradient evaluations (no timing).

    ncu --set full --import-source on --clock-control none --cache-control none -k regex:'k_splat9|k_backward9' \
        --launch-skip 8 --launch-count 2 -o gpurun_out/prof python profiles/prof_eval.py --evals 8
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from eincm_b200 import plan as P, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='dsec')
    ap.add_argument('--events', type=int, default=None)
    ap.add_argument('--theta', type=int, default=16)
    ap.add_argument('--evals', type=int, default=8)
    ap.add_argument('--windows', type=int, default=2)
    ap.add_argument('--gamma', type=float, default=None)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    wins = [synth.make_workload(args.workload, seed=k, n_events=args.events) for k in range(args.windows)]
    H, W = wins[0].sensor_size
    hpd = wins[0].hparams
    hp = P.make_hparams(hpd['alpha'], hpd['beta'], hpd['gamma'] if args.gamma is None else args.gamma, hpd['delta'], 0)
    plans, thetas = [], []
    for w in wins:
        p = P.Plan((H, W), max_events=len(w.xs), max_refs=max(len(w.edge_ts), 3))
        p.set_window(*w.args())
        plans.append(p)
        thetas.append(torch.from_numpy(synth.theta_test_points(w, (args.theta, args.theta))['perturbed']).cuda())
    loss = torch.zeros(1, dtype=torch.float64, device='cuda')
    grad = torch.zeros((args.theta, args.theta, 2), dtype=torch.float64, device='cuda')
    for i in range(args.evals):
        k = i % len(plans)
        plans[k].value_and_grad_device(thetas[k], hp, loss, grad)
    torch.cuda.synchronize()
    print('loss', float(loss.item()))


if __name__ == '__main__':
    main()
