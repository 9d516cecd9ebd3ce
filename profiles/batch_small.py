"""Batched evaluation of B MVSEC-shaped windows per launch against B: how the per-window cost falls with B (launch latency vs throughput)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, torch
torch.cuda.set_device(0)
for B in (1, 2, 4, 8, 32, 128):
    r = bench.batched_measure('mvsec_dt4', B, 16, False, 20, 5, 0.3, cpu=False, n_distinct=min(B, 16))
    print(B, f"{r['ms_per_step'] * 1e3:8.1f} us per step, {r['us_per_window']:6.1f} us per window", r['kernels_ms_per_launch'])
