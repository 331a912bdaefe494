"""One precomputed G2 MSM (for ncu): python tools/g2_one.py [log_n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import zcash_gpu_thesis_b200 as zk
from zcash_gpu_thesis_b200 import _lib as L

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << log_n
w = zk.Worker(0)
rng = np.random.default_rng(3)
k = np.zeros((n, 4), dtype=np.uint64)
k[:, 0] = rng.integers(1, 1 << 64, size=n, dtype=np.uint64)
dxy, dinf, _ = zk.fixed_base_mul(w, L.G2, bench.gen_g2_limbs(), k, 64)
bases = zk.Bases.from_device(w, L.G2, dxy, n)
bases.precompute(0)
ds = w.to_device(bench.random_scalars(rng, n))
out = w.alloc(288)
for _ in range(3):
    assert w.lib.b200zk_multiexp_dev(w.ctx, bases.handle, 0, ds.ptr, n, None, out.ptr, None) == 0
w.sync()
print("ok")
