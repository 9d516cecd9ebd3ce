"""Phase breakdown of the cooperative image pass (CTA 0, %globaltimer).  EINCM_IMAGE_PASS_STAMPS=1 python profiles/image_pass_stamps.py"""
import ctypes as C
import os
import sys

os.environ['EINCM_IMAGE_PASS_STAMPS'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from eincm_b200 import plan as P, synth  # noqa: E402

for name in (sys.argv[1:] or ['dsec', 'mvsec_dt4']):
    w = synth.make_workload(name, seed=0)
    hp = P.make_hparams(w.hparams['alpha'], w.hparams['beta'], 0.0, 0.0, 0)
    p = P.Plan(w.sensor_size, max_events=len(w.xs), max_refs=5)
    p.set_window(*w.args())
    th = synth.theta_test_points(w, (16, 16))['perturbed']
    acc = np.zeros(5)
    n = 20
    for i in range(n + 5):
        p.value_and_grad_host(th, hp)
        out = (C.c_uint64 * 16)()
        p._check(p.lib.eincm_debug_image_pass_stamps(p._h, out))
        t = np.array(list(out), dtype=np.float64)
        if i >= 5:
            acc += np.diff(t[:6]) / 1e3
            sub = (t[6:10] - t[3]) / 1e3
            if i == n + 4:
                print('   phase 2 detail (us after barrier): loads+merge, shuffle tree, stats, sync', np.round(sub, 2))
    names = ['setup', 'phase 1 (stats)', 'grid barrier', 'phase 2 (reduce)', 'phase 3 (dL/dI)']
    print(name, {k: round(v / n, 2) for k, v in zip(names, acc)}, 'total us', round(acc.sum() / n, 2))
    p.close()
