#!/usr/bin/env python
"""bench.py -- the hot-path benchmark (contract: `python bench.py --gpus N --steps K --warmup W`).

Metric (BASELINE.json): BLS12-381 G1 MSM points/s at 2^24 (1-8 B200); also Fr NTT/s and Spend-shaped H-pipeline
rates as `extra`.  One "step" = one pass of bellman::multiexp over the rank's 2^24-point shard of synthetic
random bases / scalars (weak scaling: N ranks = one N*2^24-point MSM sharded by base range, the per-rank partial
sums combined by an NCCL all-gather + point adds inside the step).

  value   points/s with scalars and bases resident in HBM (device-timed, max over ranks)
  e2e     the same through the host-buffer C-ABI call b200zk_multiexp: scalars H2D from pinned memory and the
          144-byte result D2H inside the timed region (bases stay resident: they are the CRS, uploaded once)
  roofline  dominant kernel k_msm_accumulate against the measured integer-pipe peak (MSM is
          bound by the INT32 multiply-add pipe, not HBM or tensor cores); HBM figures alongside
  cpu_baseline  the C++ port of the reference's multiexp (oracle/csrc/cref.cpp) on the host cores, bounded sample

`--impl reference` times that CPU port alone (the reference is Rust and cannot be built in this image).
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FR_MODULUS = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
METRIC = "BLS12-381 G1 MSM points/s at 2^24"
UNIT = "points/s"
SEED = 0x5DBE62598D313D76


def env_int(name, default):
    return int(os.environ.get(name, default))


def random_scalars(rng, n):
    """uniform canonical Fr scalars (n, 4) uint64 -- 255-bit rejection sampling like fr.rs:255-268"""
    mod = [(FR_MODULUS >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
    out = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    out[:, 3] &= np.uint64((1 << 63) - 1)
    while True:
        ge = np.zeros(n, dtype=bool)
        eq = np.ones(n, dtype=bool)
        for l in (3, 2, 1, 0):
            ge |= eq & (out[:, l] > np.uint64(mod[l]))
            eq &= out[:, l] == np.uint64(mod[l])
        bad = ge | eq
        k = int(bad.sum())
        if k == 0:
            return out
        fresh = rng.integers(0, 1 << 64, size=(k, 4), dtype=np.uint64)
        fresh[:, 3] &= np.uint64((1 << 63) - 1)
        out[bad] = fresh


def dot_mod_r(k64, scalars):
    """sum_i k_i * s_i mod r with k_i 64-bit, s_i 256-bit, via 16-bit digit dot products in uint64 (exact)."""
    n = k64.shape[0]
    total = 0
    kd = [((k64 >> np.uint64(16 * a)) & np.uint64(0xFFFF)) for a in range(4)]
    for limb in range(4):
        for b in range(4):
            sd = (scalars[:, limb] >> np.uint64(16 * b)) & np.uint64(0xFFFF)
            for a in range(4):
                # each product < 2^32, n <= 2^26 terms -> < 2^58: no overflow
                total += int(np.dot(kd[a], sd)) << (16 * a + 64 * limb + 16 * b)
    return total % FR_MODULUS


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def windows_reference(n):
    c = 3 if n < 32 else int(math.ceil(math.log(float(n))))
    return c, (255 + c - 1) // c


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """The reference's own CPU multiexp (C++ port of multiexp.rs:140-335, all host threads) on a bounded sample."""
    if rank != 0:
        return
    from oracle import cref
    from oracle.curve import G1

    log_n = env_int("B200ZK_REF_LOG_N", 20)
    n = 1 << log_n
    rng = np.random.default_rng([SEED & 0xFFFFFFFF, SEED >> 32, 99])
    k = np.zeros((n, 4), dtype=np.uint64)
    k[:, 0] = rng.integers(1, 1 << 64, size=n, dtype=np.uint64)
    gen = np.array(G1.affine_to_limbs(G1.gen), dtype=np.uint64)
    bases, _ = cref.scalar_muls("g1", gen, k)
    scalars = random_scalars(rng, n)
    cores = cref.hardware_threads()
    for _ in range(args.warmup):
        cref.multiexp("g1", bases[: n // 8], scalars[: n // 8])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st, _ = cref.multiexp("g1", bases, scalars)
        assert st == 0
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = f"2^{log_n}-point G1 multiexp per step (same distribution as the 2^24 workload), C++ port of bellman multiexp, {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (u64 on CPU)",
        "data": "synthetic", "config": {"workload": "G1 MSM 2^24 points per GPU (reference arm: bounded 2^%d sample on host cores)" % log_n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200zk")
    ap.add_argument("--log-n", type=int, default=env_int("B200ZK_BENCH_LOG_N", 24))
    ap.add_argument("--no-extra", action="store_true", help="skip the NTT / H-pipeline extra lines")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-precompute", action="store_true", help="do not build the 2^(cw)P base table (plain windows)")
    args = ap.parse_args()
    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist

    import zcash_gpu_thesis_b200 as zk
    from zcash_gpu_thesis_b200 import _lib as L

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    w = zk.Worker(local_rank)
    lib = w.lib
    # the library's own NCCL communicator (gather of the per-shard partial sums): id from rank 0 via torch.distributed
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (ctypes.c_uint8 * 128)()
            assert lib.b200zk_nccl_unique_id(buf) == 0
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        ub = (ctypes.c_uint8 * 128)(*uid.cpu().tolist())
        st = lib.b200zk_comm_init(w.ctx, ub, rank, world)
        assert st == 0, w.last_error()

    n = 1 << args.log_n
    rng = np.random.default_rng([SEED & 0xFFFFFFFF, SEED >> 32, rank])
    # ---- synthetic inputs: bases [k_i]G generated on the device, uniform scalars from the host
    k = np.zeros((n, 4), dtype=np.uint64)
    k[:, 0] = rng.integers(1, 1 << 64, size=n, dtype=np.uint64)
    gen = gen_g1_limbs()
    dxy, dinf, _ = zk.fixed_base_mul(w, zk.G1, gen, k, 64)
    bases = zk.Bases.from_device(w, zk.G1, dxy, n)
    cpu_log = min(env_int("B200ZK_CPU_LOG_N", 23), args.log_n)  # 2^23 pairs: about 12 s of CPU work on the box (the 10-30 s the contract asks for)
    cpu_bases = dxy.download(np.uint64, 12 << cpu_log).reshape(-1, 12) if (rank == 0 and not args.no_cpu) else None
    dxy.free(); dinf.free()
    t_pre = None
    if not args.no_precompute:
        t0 = time.perf_counter()
        bases.precompute(0)  # one-time, part of loading the bases (like Parameters::read); not in the timed region
        t_pre = time.perf_counter() - t0
    scalars = random_scalars(rng, n)
    # pinned host copy for the e2e path
    hp = ctypes.c_void_p()
    assert lib.b200zk_host_alloc_pinned(scalars.nbytes, ctypes.byref(hp)) == 0
    pinned = np.ctypeslib.as_array(ctypes.cast(hp, ctypes.POINTER(ctypes.c_uint64)), shape=(n, 4))
    pinned[:] = scalars
    d_scalars = w.to_device(scalars)
    d_part = w.alloc(288)
    d_out = w.alloc(288)

    def step_resident():
        st = lib.b200zk_multiexp_dev(w.ctx, bases.handle, 0, d_scalars.ptr, n, None, d_part.ptr, None)
        assert st == 0, w.last_error()
        if world > 1:
            st = lib.b200zk_allgather_sum_dev(w.ctx, L.G1, d_part.ptr, d_out.ptr)
            assert st == 0, w.last_error()

    out_host = np.zeros(18, dtype=np.uint64)

    def step_e2e():
        if world == 1:
            # the call a bellman user makes: host scalars in, Jacobian result out (H2D + MSM + D2H inside)
            st = lib.b200zk_multiexp(w.ctx, bases.handle, 0, pinned.ctypes.data_as(ctypes.c_void_p), n, None, out_host.ctypes.data_as(ctypes.c_void_p))
            assert st == 0, w.last_error()
        else:
            # sharded form: H2D of this rank's scalars, shard MSM, NCCL gather + sum, D2H of the total
            st = lib.b200zk_h2d(w.ctx, d_scalars.ptr, pinned.ctypes.data_as(ctypes.c_void_p), pinned.nbytes)
            assert st == 0, w.last_error()
            step_resident()
            st = lib.b200zk_d2h(w.ctx, out_host.ctypes.data_as(ctypes.c_void_p), d_out.ptr, 144)
            assert st == 0, w.last_error()

    def barrier():
        w.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, steps):
        barrier()
        w.timer_start()
        for _ in range(steps):
            fn()
        ms = w.timer_stop()
        barrier()
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- correctness gate (untimed): sum s_i [k_i]G == [sum s_i k_i]G through an independent device path
    step_resident()
    w.sync()
    got = (d_out if world > 1 else d_part).download(np.uint64, 18)
    got_aff, got_inf = zk.into_affine(w, zk.G1, got)
    total = dot_mod_r(k[:, 0], scalars)
    if world > 1:
        tt = torch.tensor([(total >> (32 * i)) & 0xFFFFFFFF for i in range(8)], dtype=torch.int64, device="cuda")
        gathered = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(gathered, tt)
        total = sum(sum(int(v) << (32 * i) for i, v in enumerate(g.cpu().tolist())) for g in gathered) % FR_MODULUS
    tl = np.array([[(total >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]], dtype=np.uint64)
    exp_xy, exp_inf, _ = zk.fixed_base_mul(w, zk.G1, gen, tl, 255)
    want = exp_xy.download(np.uint64, 12)
    ok = (not got_inf[0]) and np.array_equal(got_aff[0], want)
    if not ok:
        print(json.dumps({"error": "MSM result failed the sum(s_i k_i) identity check", "rank": rank}))
        sys.exit(1)

    # ---- timed: resident
    for _ in range(max(args.warmup, 3)):
        step_resident()
    lib.b200zk_profile_enable(w.ctx, 1)
    lib.b200zk_launch_count(w.ctx, 1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches_timed = int(lib.b200zk_launch_count(w.ctx, 1))
    acc_ms = ctypes.c_double()
    acc_n = ctypes.c_int()
    lib.b200zk_profile_read(w.ctx, ctypes.byref(acc_ms), ctypes.byref(acc_n))
    lib.b200zk_profile_enable(w.ctx, 0)
    ms_step = ms_total / args.steps
    value = world * n / (ms_step * 1e-3)

    # ---- timed: e2e through the host-buffer entry point: the reference's future-returning multiexp() (sharded form for N > 1): step k + 1 is
    # submitted before step k is waited for (the prover keeps its multiexps in flight, prover.rs:289-354), so the H2D copy of
    # the next exponents overlaps the running multiexp; every step still copies its 512 MiB of exponents from pinned host
    # memory and reads its result back inside the timed region.
    def run_e2e(steps):
        submit = lib.b200zk_multiexp_async if world == 1 else lib.b200zk_multiexp_sharded_async
        pend = None
        for _ in range(steps):
            job = ctypes.c_void_p()
            st = submit(w.ctx, bases.handle, 0, pinned.ctypes.data_as(ctypes.c_void_p), n, None, ctypes.byref(job))
            assert st == 0, w.last_error()
            if pend is not None:
                assert lib.b200zk_job_wait(pend, out_host.ctypes.data_as(ctypes.c_void_p)) == 0, w.last_error()
            pend = job
        assert lib.b200zk_job_wait(pend, out_host.ctypes.data_as(ctypes.c_void_p)) == 0, w.last_error()

    step_e2e()
    ref_out = out_host.copy()
    run_e2e(2)
    assert np.array_equal(zk.into_affine(w, zk.G1, out_host)[0], zk.into_affine(w, zk.G1, ref_out)[0])
    barrier()
    t0 = time.perf_counter()
    w.timer_start()
    run_e2e(args.steps)
    e2e_dev_ms = w.timer_stop()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    tt = torch.tensor([max(e2e_dev_ms, e2e_wall_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_ms = float(tt.item()) / args.steps
    e2e_value = world * n / (e2e_ms * 1e-3)
    # for transparency: the same through the blocking entry point (no overlap between steps)
    sync_ms = timed(step_e2e, args.steps) / args.steps

    # ---- roofline for the dominant kernel (bucket accumulation), integer pipe
    # The multiplier pipe (fmaheavy) issues one 32x32->64 IMAD.WIDE per 4 cycles per SM sub-partition = 32/clk/SM (ncu:
    # 94.8 % pipe-busy at 9.24 T wide-IMAD/s, profiles/r01_mulbench_ncu_pipes.csv); plain 32-bit IMAD is 64/clk/SM.
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    int_peak = 32.0 * w.sm_count() * sm_mhz * 1e6
    wide_measured = w.microbench(1, 4000)  # dependent mad.wide chains, measured live (a lower bound of the pipe rate)
    c_ref, w_ref = windows_reference(n)
    alg_imad = n * 3300.0 * w_ref  # W(n) mixed adds x 11 Fq-mul-equiv x 300 IMAD per point
    acc_launch_ms = acc_ms.value / max(acc_n.value, 1)
    achieved = alg_imad / (acc_launch_ms * 1e-3) if acc_launch_ms > 0 else 0.0
    c_used = 22 if t_pre is not None else 16
    w_used = (256 + c_used - 1) // c_used
    # mixed adds actually executed x (6 products x 300 + 2 squarings x 234 + the fused r(q-x3) - y ppp: 2 x 144 + 156)
    true_imad = n * w_used * (6 * 300.0 + 2 * 234.0 + 444.0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = n * (96 + 32)
    roofline = {
        "kernel": "k_msm_accumulate<fq_t>", "bound": "int32",
        "achieved": achieved / 1e12, "peak": int_peak / 1e12, "unit": "TIMAD/s", "frac": achieved / int_peak if int_peak else None,
        "traffic": traffic_from_profile(),
        "note": "MSM is bound by the INT32 multiplier pipe (no dense contraction, HBM-light): achieved = algorithmic 32x32->64 multiply-adds of "
                "the reference algorithm (3300*W(n) per point, W(2^24)=15 windows of c=17; SURVEY.md 8d) / measured accumulate-kernel time; peak = "
                "32 IMAD.WIDE/clk/SM x SMs x sampled SM clock (the fmaheavy pipe rate established with ncu); the 64/clk/SM figure of plain 32-bit "
                "IMAD does not apply to 64-bit products.",
        "imad_wide_measured_TIMADps": wide_measured / 1e12,
        "achieved_true": true_imad / (acc_launch_ms * 1e-3) / 1e12 if acc_launch_ms else None,
        "frac_true": true_imad / (acc_launch_ms * 1e-3) / int_peak if acc_launch_ms else None,
        "true_note": "multiply-adds the kernel really executes: n x %d windows (c = %d, signed digits) x (6 products x 300 + 2 squarings x 234 + one fused product difference x 444) (madd-2008-s, XYZZ); "
                     "frac above 1 against the reference formula means the schedule needs fewer adds than the reference's c = 17, W = 15 Jacobian one" % (w_used, c_used),
        "kernel_ms_per_launch": acc_launch_ms, "kernel_share_of_step": acc_launch_ms / ms_step if ms_step else None,
        "hbm": {"algorithmic_GB": alg_bytes / 1e9, "achieved_GBps": alg_bytes / 1e9 / (acc_launch_ms * 1e-3) if acc_launch_ms else None,
                "peak_GBps": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
    }

    extra = {}
    if not args.no_extra:
        extra = bench_extra(w, zk, lib, rng, timed, world)

    cpu_baseline = None
    if rank == 0 and not args.no_cpu:
        cpu_baseline = bench_cpu_baseline(cpu_bases, scalars, cpu_log)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (381-bit Fq / 255-bit Fr Montgomery)",
            "data": "synthetic",
            "config": {"workload": f"G1 MSM 2^{args.log_n} points per GPU (BASELINE configs[2]); bases [k_i]G resident in HBM, uniform Fr scalars",
                       "points_per_gpu": n, "sharding": "base range per rank, NCCL all-gather of 144-byte partials + point adds" if world > 1 else "single GPU",
                       "l2": "inputs (512 MiB scalars + 1.5 GiB bases per GPU) exceed the 126 MB L2; no flush needed",
                       "result_check": "sum s_i [k_i]G == [sum s_i k_i]G verified before timing",
                       "bases_precomputed": None if t_pre is None else {"table": "2^(c w) P_i for all windows resident in HBM (b200zk_bases_precompute, c = 22, 12 x 1.5 GiB)",
                                                                        "one_time_setup_s": t_pre}},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "blocking_call_value": world * n / (sync_ms * 1e-3), "blocking_call_ms_per_step": sync_ms, "h2d_bytes_per_step": int(scalars.nbytes), "d2h_bytes_per_step": 148,
                    "note": "b200zk_multiexp_async / b200zk_job_wait (the future-returning multiexp of the reference, depth-2 pipeline) with pinned host "
                            "scalars; bases (the CRS) stay resident; time = max(device events, host wall clock) over the steps"},
            "gpu_launches": launches_timed + (2 * args.steps if world > 1 else 0),
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "extra": extra,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def traffic_from_profile():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full capture
    (profiles/r01_traffic.json); None when the capture does not match the precomputed-table configuration."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        return {"dram_bytes_per_launch": t["dram_bytes_per_launch"], "source": t["source"], "workload": t["workload"]}
    except Exception:
        return None


def bench_extra(w, zk, lib, rng, timed, world):
    """Fr NTT/s at 2^24 and 2^17, and the Spend-shaped H pipeline (7 NTTs at 2^17), per GPU replicas."""
    from zcash_gpu_thesis_b200 import _lib as L

    out = {}
    peaks_hbm = 6539.9
    for log_m in (24, 17):
        m = 1 << log_m
        coeffs = random_scalars(rng, m)  # canonical values < r are valid Montgomery residues
        d = w.to_device(coeffs)
        for kind, name in ((L.FFT, "fft"), (L.IFFT, "ifft"), (L.COSET_FFT, "coset_fft")):
            def run():
                st = lib.b200zk_ntt_dev(w.ctx, d.ptr, log_m, kind)
                assert st == 0
            for _ in range(3):
                run()
            steps = 10 if log_m == 24 else 50
            ms = timed(run, steps) / steps
            # the binding roofline is the 64-bit multiplier pipe (32 IMAD.WIDE/clk/SM at the maximum SM clock): (m/2) log2 m
            # butterflies x 114 wide multiply-adds for the Fr product; the HBM view (64 m bytes per transform) is reported beside it
            int_peak = 32.0 * w.sm_count() * 1965e6
            out[f"fr_{name}_2^{log_m}"] = {"ntt_per_s": world * 1e3 / ms, "ms": ms, "hbm_GBps_algorithmic": 64.0 * m / 1e9 / (ms * 1e-3),
                                           "hbm_frac_of_measured_peak": 64.0 * m / 1e9 / (ms * 1e-3) / peaks_hbm,
                                           "multiplier_frac": (m / 2) * log_m * 114.0 / (ms * 1e-3) / int_peak, "bound": "int32 multiplier pipe"}
        d.free()
    # Spend-shaped H block: m = 2^17 (98 785 constraints)
    m = 1 << 17
    a, b, c = (w.to_device(random_scalars(rng, m)) for _ in range(3))
    o = w.alloc(m * 32)

    def run_h():
        st = lib.b200zk_h_poly_dev(w.ctx, a.ptr, b.ptr, c.ptr, 17, o.ptr)
        assert st == 0
    for _ in range(3):
        run_h()
    ms = timed(run_h, 20) / 20
    out["spend_h_poly_2^17"] = {"per_s": world * 1e3 / ms, "ms": ms}
    for x in (a, b, c, o):
        x.free()
    out["sapling_spend_proofs"] = bench_spend_proofs(w, zk, rng, world)
    # the reference's largest circuit: Sprout JoinSplit on Groth16, 1 989 085 constraints -> m = 2^21 (SURVEY.md section 8;
    # sapling-crypto/src/circuit/sprout/mod.rs:465); variable counts and densities are not published: taken proportional to Spend's
    out["sprout_joinsplit_shaped_proofs"] = bench_spend_proofs(w, zk, rng, world, shape=(1989085, 10, 1986000, 1719000, 1, 1234000))
    return out


def _slowest_rank(seconds, world):
    """max over ranks of a wall-clock interval (every rank calls it at the same point), so that a multi-GPU proofs/s number is
    all the proofs of the job over the time of the slowest rank"""
    if world == 1:
        return seconds
    import torch
    import torch.distributed as dist

    t = torch.tensor([seconds], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _line_up(world):
    if world > 1:
        import torch.distributed as dist

        dist.barrier()


def bench_spend_proofs(w, zk, rng, world, shape=None):
    """Sapling-Spend-shaped create_proof: 98 785 constraints -> m = 2^17; H 131 071, L 98 638, A 8 + 85 382,
    B 1 + 61 299 (G1 and G2) bases; witness-like scalars (half of them 0/1).  The CRS is synthetic ([k]G points generated on
    the device), so the proofs are not meaningful -- the arithmetic is identical to a real Spend proof.  Proofs of a batch
    are independent: each rank proves its own share on `streams` concurrent contexts."""
    import threading

    from zcash_gpu_thesis_b200 import _lib as L

    n_con, n_in, n_aux = 98785, 8, 98638
    a_dense, b_in_dense, b_aux_dense = 85382, 1, 61299
    light = shape is not None  # another circuit shape (the Sprout JoinSplit one): single-call latency and a short batch only
    if light:
        n_con, n_in, n_aux, a_dense, b_in_dense, b_aux_dense = shape
    log_m = (n_con - 1).bit_length()
    g1 = gen_g1_limbs()
    g2 = gen_g2_limbs()

    def bases(group, n, gen):
        k = np.zeros((n, 4), dtype=np.uint64)
        k[:, 0] = rng.integers(1, 1 << 64, size=n, dtype=np.uint64)
        dxy, dinf, _ = zk.fixed_base_mul(w, group, gen, k, 64)
        b = zk.Bases.from_device(w, group, dxy, n)
        dxy.free(); dinf.free()
        return b

    h, l = bases(L.G1, (1 << log_m) - 1, g1), bases(L.G1, n_aux, g1)
    a, b1, b2 = bases(L.G1, n_in + a_dense, g1), bases(L.G1, b_in_dense + b_aux_dense, g1), bases(L.G2, b_in_dense + b_aux_dense, g2)
    head1 = np.zeros((3, 12), dtype=np.uint64)
    st = w.lib.b200zk_d2h(w.ctx, head1.ctypes.data_as(ctypes.c_void_p), bases_ptr(w, zk, L.G1, g1, rng, 3), 3 * 96)
    head2 = np.zeros((2, 24), dtype=np.uint64)
    st |= w.lib.b200zk_d2h(w.ctx, head2.ctypes.data_as(ctypes.c_void_p), bases_ptr(w, zk, L.G2, g2, rng, 2), 2 * 192)
    assert st == 0
    t0 = time.perf_counter()
    for q in (h, l, a, b1, b2):
        q.precompute(0)  # one-time, at CRS load
    t_pre = time.perf_counter() - t0
    params = zk.Parameters(w, h, l, a, b1, b2, head1[0], head1[1], head2[0], head1[2], head2[1])

    def witness(n):
        v = random_scalars(rng, n)
        small = rng.random(n) < 0.5
        v[small] = 0
        v[small, 0] = rng.integers(0, 2, size=int(small.sum()), dtype=np.uint64)
        return v

    def density(n, total):
        d = np.zeros(n, dtype=np.uint8)
        d[rng.choice(n, size=total, replace=False)] = 1
        return d

    # the assignment a prover hands over lives in page-locked memory (the e2e rule: H2D from pinned host memory)
    ev = [w.pinned_copy(random_scalars(rng, n_con)) for _ in range(3)]
    inputs, aux = witness(n_in), witness(n_aux)
    inputs[0] = (1, 0, 0, 0)
    inputs, aux = w.pinned_copy(inputs), w.pinned_copy(aux)
    da, dbi, dba = (w.pinned_copy(x) for x in (density(n_aux, a_dense), density(n_in, b_in_dense), density(n_aux, b_aux_dense)))
    r, s = 0x1234567890ABCDEF1234567890ABCDEF, 0x0FEDCBA0987654321FEDCBA098765432

    def prove(worker):
        return zk.create_proof_from_assignment(worker, params, ev[0], ev[1], ev[2], inputs, aux, da, dbi, dba, r, s)

    p0 = prove(w)
    p1 = prove(w)
    assert np.array_equal(p0.a, p1.a) and np.array_equal(p0.c, p1.c)  # deterministic
    t0 = time.perf_counter()
    reps = 3 if light else 5
    for _ in range(reps):
        prove(w)
    single_ms = (time.perf_counter() - t0) / reps * 1e3
    if light:
        one = (ev[0], ev[1], ev[2], inputs, aux, da, dbi, dba, r, s)
        ref = p0.write(w)
        got = zk.create_proofs_from_assignments(w, params, [one] * 4, 4)
        assert all(g.write(w) == ref for g in got)
        _line_up(world)
        t0 = time.perf_counter()
        zk.create_proofs_from_assignments(w, params, [one] * 8, 4)
        bdt = _slowest_rank(time.perf_counter() - t0, world)
        for q in (h, l, a, b1, b2):
            q.free()
        return {"single_call_ms_per_proof": single_ms, "proofs_per_s": world * 8 / bdt, "lockstep": 4, "batch": world * 8,
                "shape": f"m=2^{log_m} ({n_con} constraints), multiexps {(1 << log_m) - 1}/{n_aux}/{n_in}+{a_dense}/{b_in_dense}+{b_aux_dense} (G1) and "
                         f"{b_in_dense}+{b_aux_dense} (G2), synthetic CRS and densities", "crs_precompute_s": t_pre}
    streams = env_int("B200ZK_SPEND_STREAMS", 8)
    workers = [zk.Worker(w.device) for _ in range(streams)]
    for x in workers:
        prove(x)
    per_thread = env_int("B200ZK_SPEND_PER_STREAM", 6)

    def loop(x):
        for _ in range(per_thread):
            prove(x)

    ths = [threading.Thread(target=loop, args=(x,)) for x in workers]
    _line_up(world)
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = _slowest_rank(time.perf_counter() - t0, world)
    for x in workers:
        x.close()
    # the batch API: `lockstep` proofs share five batched multiexps (b200zk_groth16_prove_batch); two contexts alternate so
    # that the uploads and the serial tails of one group overlap the multiexps of the other
    lockstep = env_int("B200ZK_SPEND_LOCKSTEP", 8)
    bstreams = env_int("B200ZK_SPEND_BATCH_STREAMS", 2)
    groups = env_int("B200ZK_SPEND_GROUPS", 4)
    one = (ev[0], ev[1], ev[2], inputs, aux, da, dbi, dba, r, s)
    bworkers = [zk.Worker(w.device) for _ in range(bstreams)]
    ref = p0.write(w)
    for x in bworkers:
        got = zk.create_proofs_from_assignments(x, params, [one] * lockstep, lockstep)
        assert all(g.write(x) == ref for g in got)  # the batch gives the single-call proof

    def bloop(x):
        zk.create_proofs_from_assignments(x, params, [one] * (lockstep * groups), lockstep)

    ths = [threading.Thread(target=bloop, args=(x,)) for x in bworkers]
    _line_up(world)
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    bdt = _slowest_rank(time.perf_counter() - t0, world)
    for x in bworkers:
        x.close()
    batched = {"proofs_per_s": world * bstreams * lockstep * groups / bdt, "lockstep": lockstep, "contexts_per_gpu": bstreams,
               "batch": world * bstreams * lockstep * groups, "api": "b200zk_groth16_prove_batch"}
    return {"proofs_per_s": max(world * streams * per_thread / dt, batched["proofs_per_s"]), "batched": batched,
            "independent_calls_proofs_per_s": world * streams * per_thread / dt, "single_stream_ms_per_proof": single_ms, "streams_per_gpu": streams,
            "batch": world * streams * per_thread, "timing": "host wall clock around the prove calls incl. H2D of a/b/c/assignments and D2H of the proofs; with N GPUs every rank proves its own share and the time is that of the slowest rank",
            "shape": "m=2^17, MSM sizes 131071/98638/8+85382/1+61299 (G1) and 1+61299 (G2), synthetic CRS",
            "crs_precompute_s": t_pre}


def gen_g1_limbs():
    gx = 0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb
    gy = 0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1
    return fq_mont_limbs([gx, gy])


def gen_g2_limbs():
    x0 = 0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8
    x1 = 0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e
    y0 = 0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801
    y1 = 0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be
    return fq_mont_limbs([x0, x1, y0, y1])


def fq_mont_limbs(vals):
    q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    R = (1 << 384) % q
    return np.array([((v * R % q) >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for v in vals for i in range(6)], dtype=np.uint64)


_KEEP = []


def bases_ptr(w, zk, group, gen, rng, n):
    """n device-generated points [k]G as a raw device pointer (kept alive for the duration of the bench)"""
    k = np.zeros((n, 4), dtype=np.uint64)
    k[:, 0] = rng.integers(1, 1 << 64, size=n, dtype=np.uint64)
    dxy, dinf, _ = zk.fixed_base_mul(w, group, gen, k, 64)
    _KEEP.append((dxy, dinf))
    return dxy.ptr


def bench_cpu_baseline(bases, scalars, log_s):
    """The oracle's C++ port of bellman multiexp on the host cores: bounded sample = the first 2^log_s (base, scalar) pairs of
    rank 0's own workload (bases downloaded from the GPU), one multiexp with the reference's c = ceil(ln n) windows."""
    from oracle import cref

    s = 1 << log_s
    cores = cref.hardware_threads()
    cref.multiexp("g1", bases[: s // 16], scalars[: s // 16])
    t0 = time.perf_counter()
    st, _ = cref.multiexp("g1", bases[:s], scalars[:s])
    dt = time.perf_counter() - t0
    assert st == 0
    return {"value": s / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first 2^{log_s} (base, scalar) pairs of rank 0's workload, one multiexp, one pool task per window (c = ceil(ln n)), wall {dt:.2f} s"}


if __name__ == "__main__":
    main()
