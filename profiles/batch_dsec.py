import sys; sys.path.insert(0,'/root/repo')
import bench, json, torch
torch.cuda.set_device(0)
for B in (4, 8, 16):
    r = bench.batched_measure('dsec', B, 16, False, 20, 5, 0.5, cpu=False, n_distinct=B)
    print(B, round(r['value'],2), 'Gev/s', round(r['us_per_window'],1), 'us/window', r['kernels_ms_per_launch'], 'e2e', round(r['e2e']['value'],2))
