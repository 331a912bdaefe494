"""A few 2^log_m forward NTTs (for ncu): python tools/ntt_one.py [log_m] [kind]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import zcash_gpu_thesis_b200 as zk

log_m = int(sys.argv[1]) if len(sys.argv) > 1 else 24
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 0
w = zk.Worker(0)
rng = np.random.default_rng(3)
d = w.to_device(bench.random_scalars(rng, 1 << log_m))
for _ in range(3):
    assert w.lib.b200zk_ntt_dev(w.ctx, d.ptr, log_m, kind) == 0
w.sync()
w.timer_start()
for _ in range(5):
    assert w.lib.b200zk_ntt_dev(w.ctx, d.ptr, log_m, kind) == 0
print("ms per NTT:", w.timer_stop() / 5)
