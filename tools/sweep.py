"""Size sweep of the hot path on one GPU (BASELINE.json configs 2-4): G1 / G2 MSM and Fr NTT, every MSM result checked
through sum s_i [k_i]G == [sum s_i k_i]G.  Writes gpurun_out/sweep.json.
usage: python tools/sweep.py [--g1 16,20,22,24,26] [--g2 16,20,22] [--ntt 16,20,24]"""
import argparse, ctypes, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import zcash_gpu_thesis_b200 as zk
from zcash_gpu_thesis_b200 import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--g1", default="16,18,20,22,24,26")
ap.add_argument("--g2", default="16,20,22")
ap.add_argument("--ntt", default="16,17,20,22,24")
ap.add_argument("--no-precompute", action="store_true")
args = ap.parse_args()
w = zk.Worker(0)
lib = w.lib
rng = np.random.default_rng(11)
out = {"g1_msm": {}, "g2_msm": {}, "ntt": {}}


def msm(group, log_n, gen):
    n = 1 << log_n
    k = np.zeros((n, 4), dtype=np.uint64)
    k[:, 0] = rng.integers(1, 1 << 64, size=n, dtype=np.uint64)
    dxy, dinf, _ = zk.fixed_base_mul(w, group, gen, k, 64)
    bases = zk.Bases.from_device(w, group, dxy, n)
    dxy.free(); dinf.free()
    res = {}
    scalars = bench.random_scalars(rng, n)
    ds = w.to_device(scalars)
    dout = w.alloc(288)
    total = bench.dot_mod_r(k[:, 0], scalars)
    tl = np.array([[(total >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]], dtype=np.uint64)
    exp_xy, exp_inf, _ = zk.fixed_base_mul(w, group, gen, tl, 255)
    want = exp_xy.download(np.uint64, 12 if group == L.G1 else 24)
    for mode in (["plain"] if args.no_precompute else ["plain", "precomputed"]):
        if mode == "precomputed":
            t0 = time.perf_counter()
            bases.precompute(0)
            res["precompute_s"] = time.perf_counter() - t0
        def run():
            st = lib.b200zk_multiexp_dev(w.ctx, bases.handle, 0, ds.ptr, n, None, dout.ptr, None)
            assert st == 0, w.last_error()
        run(); w.sync()
        got = dout.download(np.uint64, 18 if group == L.G1 else 36)
        aff, inf = zk.into_affine(w, group, got)
        assert not inf[0] and np.array_equal(aff[0], want), f"{group} 2^{log_n} {mode}: result check failed"
        for _ in range(2):
            run()
        steps = 3 if log_n >= 22 else 10
        w.sync(); w.timer_start()
        for _ in range(steps):
            run()
        ms = w.timer_stop() / steps
        res[mode] = {"ms": ms, "points_per_s": n / (ms * 1e-3)}
    ds.free(); dout.free(); bases.free()
    return res


for lg in [int(x) for x in args.g1.split(",") if x]:
    out["g1_msm"][f"2^{lg}"] = msm(L.G1, lg, bench.gen_g1_limbs())
    print("g1", lg, out["g1_msm"][f"2^{lg}"], flush=True)
for lg in [int(x) for x in args.g2.split(",") if x]:
    out["g2_msm"][f"2^{lg}"] = msm(L.G2, lg, bench.gen_g2_limbs())
    print("g2", lg, out["g2_msm"][f"2^{lg}"], flush=True)
for lg in [int(x) for x in args.ntt.split(",") if x]:
    m = 1 << lg
    d = w.to_device(bench.random_scalars(rng, m))
    res = {}
    for kind, name in ((L.FFT, "fft"), (L.IFFT, "ifft"), (L.COSET_FFT, "coset_fft"), (L.ICOSET_FFT, "icoset_fft")):
        for _ in range(3):
            assert lib.b200zk_ntt_dev(w.ctx, d.ptr, lg, kind) == 0
        steps = 10 if lg >= 22 else 50
        w.sync(); w.timer_start()
        for _ in range(steps):
            assert lib.b200zk_ntt_dev(w.ctx, d.ptr, lg, kind) == 0
        ms = w.timer_stop() / steps
        res[name] = {"ms": ms, "ntt_per_s": 1e3 / ms, "algorithmic_GBps": 64.0 * m / 1e9 / (ms * 1e-3)}
    d.free()
    out["ntt"][f"2^{lg}"] = res
    print("ntt", lg, res, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep.json", "w"), indent=1)
