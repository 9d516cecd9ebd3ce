import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from eincm_b200 import plan as P, synth
w = synth.make_workload('dsec', seed=0)
H, W = w.sensor_size
hp = P.make_hparams(2000.0, 4000.0, 0.0, 0.0, 0)
p = P.Plan((H, W), max_events=len(w.xs), max_refs=3)
p.set_window(*w.args())
th = synth.theta_test_points(w, (16, 16))['perturbed']
thd = torch.from_numpy(th).cuda()
loss = torch.zeros(1, dtype=torch.float64, device='cuda'); grad = torch.zeros((16, 16, 2), dtype=torch.float64, device='cuda')
def t(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
def dev(): p.value_and_grad_device(thd, hp, loss, grad)
def dev_sync(): p.value_and_grad_device(thd, hp, loss, grad); torch.cuda.synchronize()
def host(): p.value_and_grad_host(th, hp)
def dev_sync_copy():
    p.value_and_grad_device(thd, hp, loss, grad); g = grad.cpu()
print('device async      us/eval', t(dev))
print('device + sync     us/eval', t(dev_sync))
print('device + D2H      us/eval', t(dev_sync_copy))
print('host call         us/eval', t(host))
s = torch.cuda.current_stream()
def empty_sync(): s.synchronize()
print('empty stream sync us', t(empty_sync))
