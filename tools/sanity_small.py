"""Small end-to-end pass over every kernel family (for compute-sanitizer memcheck): MSM G1/G2 (plain, precomputed, density,
split buckets), NTT all kinds at a few sizes, H pipeline, Groth16-shaped prove on a synthetic CRS."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import zcash_gpu_thesis_b200 as zk
from zcash_gpu_thesis_b200 import _lib as L

w = zk.Worker(0)
rng = np.random.default_rng(1)
for group, gen in ((L.G1, bench.gen_g1_limbs()), (L.G2, bench.gen_g2_limbs())):
    for n in (1, 37, 700, 5000):
        k = np.zeros((n + 10, 4), dtype=np.uint64)
        k[:, 0] = rng.integers(1, 1 << 64, size=n + 10, dtype=np.uint64)
        dxy, dinf, _ = zk.fixed_base_mul(w, group, gen, k, 64)
        bases = zk.Bases.from_device(w, group, dxy, n + 10)
        s = bench.random_scalars(rng, n)
        s[rng.random(n) < 0.4] = (1, 0, 0, 0)
        d = zk.DensityTracker(rng.random(n) < 0.7)
        a = zk.multiexp(w, (bases, 0), zk.FullDensity(), s)
        b = zk.multiexp(w, (bases, 3), d, s)
        bases.precompute(0)
        a2 = zk.multiexp(w, (bases, 0), zk.FullDensity(), s)
        b2 = zk.multiexp(w, (bases, 3), d, s)
        assert np.array_equal(zk.into_affine(w, group, a)[0], zk.into_affine(w, group, a2)[0])
        assert np.array_equal(zk.into_affine(w, group, b)[0], zk.into_affine(w, group, b2)[0])
        f = zk.multiexp_async(w, (bases, 0), zk.FullDensity(), s)
        f.wait()
for lg in (1, 5, 10, 11, 13, 17):
    v = bench.random_scalars(rng, 1 << lg)
    for kind in range(4):
        zk.ntt_host(w, v, kind)
zk.h_poly(w, bench.random_scalars(rng, 777), bench.random_scalars(rng, 777), bench.random_scalars(rng, 777))
print("sanity pass OK")
