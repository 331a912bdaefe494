"""One lock-step group of 8 Spend-shaped proofs on one context (for ncu launch lists): python tools/batch_one.py"""
import os, sys
os.environ["B200ZK_SPEND_STREAMS"] = "1"
os.environ["B200ZK_SPEND_PER_STREAM"] = "1"
os.environ["B200ZK_SPEND_BATCH_STREAMS"] = "1"
os.environ["B200ZK_SPEND_GROUPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import zcash_gpu_thesis_b200 as zk

w = zk.Worker(0)
print(bench.bench_spend_proofs(bench.single_gpu_env(w, zk), np.random.default_rng(5))["batched"])
