import sys, os, time, threading
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from eincm_b200 import plan as P, synth
torch.cuda.set_device(0)
w = synth.make_workload('dsec', seed=0)
H, W = w.sensor_size
hp = P.make_hparams(w.hparams['alpha'], w.hparams['beta'], 0.0, 0.0, 0)
p = P.Plan((H, W), max_events=len(w.xs), max_refs=3)
p.set_window(*w.args())
th = synth.theta_test_points(w, (16, 16))['perturbed']
def run(n, tag):
    for _ in range(10): p.value_and_grad_host(th, hp)
    t0 = time.perf_counter()
    for _ in range(n): p.value_and_grad_host(th, hp)
    dt = time.perf_counter() - t0
    print(tag, f'{dt / n * 1e6:.1f} us per host eval', p.host_times(reset=True))
run(300, 'main thread   ')
t = threading.Thread(target=run, args=(300, 'worker thread ')); t.start(); t.join()
run(300, 'main again    ')
# native bfgs in thread vs main: level 0 solve
from eincm_b200 import losses, solver as SV
