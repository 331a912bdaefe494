"""CPU checks of bench.py's host-side pieces: the reference arm prints one valid JSON line, the exact dot product used for
the size-independent MSM check is right, the scalar sampler stays below r."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_a_json_line():
    env = dict(os.environ, B200ZK_REF_LOG_N="10")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "points/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0


def test_dot_mod_r_and_sampler():
    sys.path.insert(0, ROOT)
    import bench

    rng = np.random.default_rng(5)
    n = 5000
    s = bench.random_scalars(rng, n)
    k = rng.integers(1, 1 << 64, size=n, dtype=np.uint64)
    ints = [sum(int(v) << (64 * i) for i, v in enumerate(row)) for row in s]
    assert all(x < bench.FR_MODULUS for x in ints)
    want = sum(int(a) * b for a, b in zip(k, ints)) % bench.FR_MODULUS
    assert bench.dot_mod_r(k, s) == want


def test_generator_constants_match_the_oracle():
    sys.path.insert(0, ROOT)
    import bench
    from oracle.curve import G1, G2

    assert list(bench.gen_g1_limbs()) == G1.affine_to_limbs(G1.gen)
    assert list(bench.gen_g2_limbs()) == G2.affine_to_limbs(G2.gen)
