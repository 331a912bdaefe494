"""Longer randomised parity run of the multiexp pipeline against the CPU oracle: python tools/fuzz.py [cases] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import util
from tests.test_gpu_fuzz import check_case, random_case
import zcash_gpu_thesis_b200 as zk

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 400
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
w = zk.Worker(0)
r = util.rng(seed)
for i in range(cases):
    check_case(w, r, random_case(r, max_n=20000))
print(f"{cases} random multiexps equal the oracle (seed {seed})")
