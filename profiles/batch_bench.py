"""Device-resident timing of the batched evaluation (eincm_batch) on one GPU.

    python profiles/batch_bench.py --workload mvsec_dt4 --batch 64 512 [--dense] [--steps 20]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from eincm_b200 import plan as P, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='mvsec_dt4')
    ap.add_argument('--batch', type=int, nargs='+', default=[64])
    ap.add_argument('--theta', type=int, default=16)
    ap.add_argument('--dense', action='store_true')
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--distinct', type=int, default=16, help='distinct synthetic windows (cycled over the batch)')
    args = ap.parse_args()
    torch.cuda.set_device(0)
    wins = [synth.make_workload(args.workload, seed=k) for k in range(min(args.distinct, max(args.batch)))]
    H, W = wins[0].sensor_size
    R = len(wins[0].edge_ts)
    N = len(wins[0].xs)
    hpd = wins[0].hparams
    hp = P.make_hparams(hpd['alpha'], hpd['beta'], 0.0, 0.0, 1)
    shape = (H, W) if args.dense else (args.theta, args.theta)
    for B in args.batch:
        plans, th, lo, gr = [], [], [], []
        for k in range(B):
            w = wins[k % len(wins)]
            p = P.Plan((H, W), max_events=N, max_refs=max(R, 3))
            p.set_window(*w.args())
            plans.append(p)
            th.append(torch.from_numpy(synth.theta_test_points(w, shape, seed=k)['perturbed']).cuda())
            lo.append(torch.zeros(1, dtype=torch.float64, device='cuda'))
            gr.append(torch.zeros(shape + (2,), dtype=torch.float64, device='cuda'))
        batch = P.Batch(plans)
        for _ in range(5):
            batch.value_and_grad_device(th, hp, lo, gr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            batch.value_and_grad_device(th, hp, lo, gr)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        # one window at a time through the per-plan call, same stream
        n1 = min(B, 32)
        for k in range(n1):
            plans[k].value_and_grad_device(th[k], hp, lo[k], gr[k])
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(3):
            for k in range(n1):
                plans[k].value_and_grad_device(th[k], hp, lo[k], gr[k])
        s1.record()
        torch.cuda.synchronize()
        ms1 = s0.elapsed_time(s1) / (3 * n1)
        print(json.dumps({'workload': args.workload, 'sensor': [H, W], 'N': N, 'R': R, 'theta': list(shape), 'batch': B,
                          'ms_per_batch': ms, 'us_per_window': ms / B * 1e3, 'Gevents_per_s': B * N / (ms * 1e-3) / 1e9,
                          'per_plan_us_per_window': ms1 * 1e3, 'per_plan_Gevents_per_s': N / (ms1 * 1e-3) / 1e9,
                          'loss0': float(lo[0].item())}))
        batch.close()
        for p in plans:
            p.close()
        del plans, th, lo, gr
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
