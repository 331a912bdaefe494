"""A few multiexps of one size (for ncu launch lists / timing): python tools/msm_one.py g1|g2 log_n [reps] [--plain] [--witness]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import zcash_gpu_thesis_b200 as zk
from tools import synthetic

group = zk.G1 if sys.argv[1] == "g1" else zk.G2
n = int(sys.argv[2]) if int(sys.argv[2]) > 40 else 1 << int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 5
w = zk.Worker(0)
rng = np.random.default_rng(3)
gen = bench.gen_g1_limbs() if group == zk.G1 else bench.gen_g2_limbs()
dxy, dinf, _ = zk.fixed_base_mul(w, group, gen, synthetic.base_multipliers(rng, n), 64)
bases = zk.Bases.from_device(w, group, dxy, n)
if "--plain" not in sys.argv:
    bases.precompute(0)
sc = synthetic.witness_scalars(rng, n) if "--witness" in sys.argv else synthetic.random_scalars(rng, n)
ds = w.to_device(sc)
out = w.alloc(320)
for _ in range(3):
    assert w.lib.b200zk_multiexp_dev(w.ctx, bases.handle, 0, ds.ptr, n, None, out.ptr, None) == 0
w.sync()
w.timer_start()
for _ in range(reps):
    assert w.lib.b200zk_multiexp_dev(w.ctx, bases.handle, 0, ds.ptr, n, None, out.ptr, None) == 0
print("ms per multiexp:", w.timer_stop() / reps, "points/s: %.4g" % (n / (w.timer_stop() / reps * 1e-3)) if False else "")
