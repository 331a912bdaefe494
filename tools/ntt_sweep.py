"""Timing of the Fr transforms over sizes and kinds (CUDA events on the context's stream): python tools/ntt_sweep.py [log_m ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zcash_gpu_thesis_b200 as zk
from tools.synthetic import random_scalars

sizes = [int(a) for a in sys.argv[1:]] or [12, 14, 16, 17, 18, 20, 22, 24]
w = zk.Worker(0)
rng = np.random.default_rng(3)
for log_m in sizes:
    d = w.to_device(random_scalars(rng, 1 << log_m))
    row = []
    for kind in range(4):
        for _ in range(3):
            assert w.lib.b200zk_ntt_dev(w.ctx, d.ptr, log_m, kind) == 0
        w.sync()
        reps = 10 if log_m >= 22 else 50
        w.timer_start()
        for _ in range(reps):
            assert w.lib.b200zk_ntt_dev(w.ctx, d.ptr, log_m, kind) == 0
        row.append(w.timer_stop() / reps)
    d.free()
    print("2^%d  fft %.4f  ifft %.4f  coset %.4f  icoset %.4f ms" % (log_m, *row), flush=True)
