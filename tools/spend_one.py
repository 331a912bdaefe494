"""One Spend-shaped proof (for ncu launch lists): python tools/spend_one.py [streams]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import zcash_gpu_thesis_b200 as zk

w = zk.Worker(0)
rng = np.random.default_rng(5)
streams = int(sys.argv[1]) if len(sys.argv) > 1 else None
if streams:
    import types
    src = open(bench.__file__).read()
out = bench.bench_spend_proofs(bench.single_gpu_env(w, zk), rng)
print(json.dumps(out))
