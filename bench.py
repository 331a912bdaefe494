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

`--impl reference` times that CPU port alone (the reference is Rust and cannot be built in this image) on the SAME
workload: one 2^24-point G1 multiexp per step on the host cores, c = ceil(ln n) = 17, 15 window tasks.

Strong scaling (BASELINE configs[2]: "2^20-2^26 points sharded across 1/2/4/8"): every run also times fixed-size
multiexps (2^24 and 2^26 points in total, n / N per rank) and reports them in `roofline.strong_*` / `extra.strong_scaling`;
`--scaling strong` makes the 2^log_n-total split the headline `value` instead of the 2^log_n-per-rank one.
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
from tools import synthetic  # noqa: E402

FR_MODULUS = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
METRIC = "BLS12-381 G1 MSM points/s at 2^24"
UNIT = "points/s"
SEED = 0x5DBE62598D313D76


def env_int(name, default):
    return int(os.environ.get(name, default))


from tools.synthetic import random_scalars  # noqa: E402  (seeded synthetic inputs shared with the parity tests)


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
    """The reference's own CPU multiexp (C++ port of multiexp.rs:140-335: one pool task per window, c = ceil(ln n)) on the host
    cores, on the headline configuration: one 2^log_n-point G1 multiexp per step (log_n = 24 unless --log-n says otherwise).
    The bases are a walk P_i = G + i * [7]G generated on the host (input preparation, untimed); scalars uniform in Fr.
    A step takes ~10-20 s, so the run stops early when B200ZK_REF_BUDGET_S (default 240 s) is used up and prints the
    number of steps it really timed."""
    if rank != 0:
        return
    from oracle import cref

    log_n = env_int("B200ZK_REF_LOG_N", args.log_n)
    n = 1 << log_n
    budget = float(os.environ.get("B200ZK_REF_BUDGET_S", 240))
    rng = np.random.default_rng([SEED & 0xFFFFFFFF, SEED >> 32, 99])
    gen = gen_g1_limbs()
    step, _ = cref.scalar_muls("g1", gen, np.array([[7, 0, 0, 0]], dtype=np.uint64))
    t_setup = time.perf_counter()
    bases, _ = cref.ResidentBases.walk_g1(gen, step[0], n)
    scalars = random_scalars(rng, n)
    t_setup = time.perf_counter() - t_setup
    cores = cref.hardware_threads()
    c_ref, w_ref = windows_reference(n)
    for _ in range(min(args.warmup, 1)):  # one short warm-up (thread pool, page faults of the scalar array); a full step costs ~15 s
        bases.multiexp(scalars[: n // 16])
    done, t0 = 0, time.perf_counter()
    while done < args.steps:
        st, _ = bases.multiexp(scalars)
        assert st == 0
        done += 1
        if time.perf_counter() - t0 + t_setup > budget:
            break
    dt = time.perf_counter() - t0
    value = n * done / dt
    sample = (f"{done} x one 2^{log_n}-point G1 multiexp (the headline workload itself), C++ port of bellman multiexp, c = {c_ref}, {w_ref} window tasks "
              f"on {cores} host threads; bases = walk G + i*[7]G generated on the host in {t_setup:.1f} s (untimed)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done, "steps_requested": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / done * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 limbs (381-bit Fq / 255-bit Fr Montgomery, u128 products)", "data": "synthetic",
        "config": workload_config(log_n, 1, None),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(log_n, world, t_pre, scaling="weak"):
    """the `config` object: identical for both arms (the driver compares it)"""
    n = 1 << log_n
    return {"workload": f"G1 MSM 2^{log_n} points per GPU (BASELINE configs[2]); bases resident, uniform Fr scalars", "points_per_gpu": n,
            "log_n": log_n, "l2": "inputs (32 B x n scalars + 96 B x n bases per GPU) exceed the 126 MB L2 at the headline size; no flush needed"}


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200zk")
    ap.add_argument("--log-n", type=int, default=env_int("B200ZK_BENCH_LOG_N", 24))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: 2^log_n points per rank (default); strong: 2^log_n points in total")
    ap.add_argument("--no-extra", action="store_true", help="skip the NTT / G2 / proof / strong-scaling extra measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
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

    env = BenchEnv(args, rank, world, local_rank, w, lib, zk, L, torch, dist, barrier, timed)
    # --scaling strong: the rank's shard is 2^log_n / world points of ONE 2^log_n-point multiexp
    shard_log = args.log_n
    if args.scaling == "strong":
        assert world & (world - 1) == 0, "strong scaling needs a power-of-two number of ranks"
        shard_log = args.log_n - (world.bit_length() - 1)
    n = 1 << shard_log
    rng = np.random.default_rng([SEED & 0xFFFFFFFF, SEED >> 32, rank])
    # ---- synthetic inputs: bases [k_i]G generated on the device, uniform scalars from the host
    wl = MsmWorkload(env, rng, n, precompute=not args.no_precompute, keep_cpu_sample=(rank == 0 and not args.no_cpu))
    wl.check()  # correctness gate (untimed): sum s_i [k_i]G == [sum s_i k_i]G through an independent device path

    # ---- timed: resident
    for _ in range(max(args.warmup, 3)):
        wl.step_resident()
    lib.b200zk_profile_enable(w.ctx, 1)
    lib.b200zk_launch_count(w.ctx, 1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(wl.step_resident, args.steps)
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
    # the next exponents overlaps the running multiexp; every step still copies its exponents from pinned host memory and reads
    # its result back inside the timed region.
    wl.step_e2e_blocking()
    ref_out = wl.out_host.copy()
    wl.run_e2e(2)
    assert np.array_equal(zk.into_affine(w, zk.G1, wl.out_host)[0], zk.into_affine(w, zk.G1, ref_out)[0])
    barrier()
    t0 = time.perf_counter()
    w.timer_start()
    wl.run_e2e(args.steps)
    e2e_dev_ms = w.timer_stop()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    tt = torch.tensor([max(e2e_dev_ms, e2e_wall_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_ms = float(tt.item()) / args.steps
    e2e_value = world * n / (e2e_ms * 1e-3)
    # for transparency: the same through the blocking entry point (no overlap between steps)
    sync_ms = timed(wl.step_e2e_blocking, args.steps) / args.steps

    # ---- the same multiexp with plain windows (no precomputed 2^(cw) P table): what a caller gets who cannot spend W x the base memory
    plain_ms = None
    if not args.no_precompute and not args.no_extra:
        lib.b200zk_set_msm_window(w.ctx, 16 if shard_log >= 23 else max(8, shard_log - 7))  # a window other than the table's: the table is bypassed
        for _ in range(2):
            wl.step_resident()
        plain_ms = timed(wl.step_resident, min(args.steps, 5)) / min(args.steps, 5)
        lib.b200zk_set_msm_window(w.ctx, 0)

    # ---- roofline for the dominant kernel (bucket accumulation), integer multiplier pipe
    # The multiplier pipe (fmaheavy) issues one 32x32->64 IMAD.WIDE per 4 cycles per SM sub-partition = 32/clk/SM (ncu:
    # 94.8 % pipe-busy at 9.24 T wide-IMAD/s, profiles/r01_mulbench_ncu_pipes.csv); plain 32-bit IMAD is 64/clk/SM.
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    int_peak = 32.0 * w.sm_count() * sm_mhz * 1e6
    c_ref, w_ref = windows_reference(n)
    alg_imad = n * 3300.0 * w_ref  # W(n) mixed adds x 11 Fq-mul-equiv x 300 IMAD per point
    acc_launch_ms = acc_ms.value / max(acc_n.value, 1)
    achieved = alg_imad / (acc_launch_ms * 1e-3) if acc_launch_ms > 0 else 0.0
    c_used, w_used = wl.window()
    # mixed adds actually executed x (6 products x 300 + 2 squarings x 234 + the fused r(q-x3) - y ppp: 2 x 144 + 156)
    true_imad = n * w_used * (6 * 300.0 + 2 * 234.0 + 444.0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = n * (96 + 32)
    traffic = traffic_from_profile()
    roofline = {
        "kernel": "k_msm_accumulate<fq_t>", "bound": "int32",
        "achieved": achieved / 1e12, "peak": int_peak / 1e12, "unit": "TIMAD/s", "frac": achieved / int_peak if int_peak else None,
        "traffic": traffic["dram_bytes_per_launch"] if traffic else None, "traffic_source": traffic["source"] if traffic else None,
        "note": "MSM is bound by the INT32 multiplier pipe (no dense contraction, HBM-light): achieved = algorithmic 32x32->64 multiply-adds of "
                "the reference algorithm (3300*W(n) per point, W(2^24)=15 windows of c=17; SURVEY.md 8d) / measured accumulate-kernel time; peak = "
                "32 IMAD.WIDE/clk/SM x SMs x sampled SM clock (the fmaheavy pipe rate established with ncu); the 64/clk/SM figure of plain 32-bit "
                "IMAD does not apply to 64-bit products.",
        "achieved_true": true_imad / (acc_launch_ms * 1e-3) / 1e12 if acc_launch_ms else None,
        "frac_true": true_imad / (acc_launch_ms * 1e-3) / int_peak if acc_launch_ms else None,
        "true_note": "multiply-adds the kernel really executes: n x %d windows (c = %d, signed digits) x (6 products x 300 + 2 squarings x 234 + one fused product difference x 444) (madd-2008-s, XYZZ); "
                     "frac above 1 against the reference formula means the schedule needs fewer adds than the reference's c = 17, W = 15 Jacobian one" % (w_used, c_used),
        "kernel_ms_per_launch": acc_launch_ms, "kernel_share_of_step": acc_launch_ms / ms_step if ms_step else None,
        "hbm_algorithmic_GB": alg_bytes / 1e9, "hbm_achieved_GBps": alg_bytes / 1e9 / (acc_launch_ms * 1e-3) if acc_launch_ms else None,
        "hbm_peak_GBps": hbm_peak, "hbm_peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
        "g1_plain_window_points_per_s": world * n / (plain_ms * 1e-3) if plain_ms else None, "g1_plain_window_ms": plain_ms,
    }
    cpu_bases, scalars = wl.cpu_bases, wl.scalars
    wl.free()

    extra = {}
    if not args.no_extra:
        extra = bench_extra(env, rng)
        # flat copies of the secondary metrics in keys the driver's parser keeps
        for k, v in flat_secondary(extra).items():
            roofline[k] = v

    cpu_baseline = None
    if rank == 0 and not args.no_cpu:
        cpu_baseline = bench_cpu_baseline(cpu_bases, scalars, min(env_int("B200ZK_CPU_LOG_N", 23), shard_log), full=not args.no_extra)

    if rank == 0 and cpu_baseline is not None and (extra.get("sapling_spend_proofs") or {}).get("cpu_baseline"):
        sp = extra["sapling_spend_proofs"]["cpu_baseline"]
        cpu_baseline.update({"spend_proofs_per_s": sp["proofs_per_s"], "spend_sample": sp["sample"]})
    if rank == 0:
        cfg = workload_config(args.log_n, world, wl.t_pre)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u32 limbs (381-bit Fq / 255-bit Fr Montgomery)",
            "data": "synthetic",
            "config": cfg,
            "setup": {"points_per_rank": n, "sharding": "base range per rank, NCCL all-gather of 320-byte records (partial point + status) + point adds" if world > 1 else "single GPU",
                      "result_check": "sum s_i [k_i]G == [sum s_i k_i]G verified before timing; Spend-shaped proof == CPU oracle's 192 bytes before its timing",
                      "bases_precomputed": None if wl.t_pre is None else {"table": "2^(c w) P_i for all windows resident in HBM (b200zk_bases_precompute, c = %d, %d x the base vector)" % (c_used, w_used),
                                                                           "one_time_setup_s": wl.t_pre}},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "blocking_call_value": world * n / (sync_ms * 1e-3), "blocking_call_ms_per_step": sync_ms, "h2d_bytes_per_step": int(scalars.nbytes), "d2h_bytes_per_step": 320,
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


class BenchEnv:
    def __init__(self, args, rank, world, local_rank, w, lib, zk, L, torch, dist, barrier, timed):
        self.args, self.rank, self.world, self.local_rank = args, rank, world, local_rank
        self.w, self.lib, self.zk, self.L, self.torch, self.dist, self.barrier, self.timed = w, lib, zk, L, torch, dist, barrier, timed


def single_gpu_env(w, zk, no_cpu=True):
    """a BenchEnv for the tools/ scripts: one rank, no torch.distributed"""
    import types

    from zcash_gpu_thesis_b200 import _lib as L

    args = types.SimpleNamespace(no_cpu=no_cpu, steps=5, warmup=3, log_n=24)

    def timed(fn, steps):
        w.sync()
        w.timer_start()
        for _ in range(steps):
            fn()
        return w.timer_stop()

    return BenchEnv(args, 0, 1, 0, w, w.lib, zk, L, None, None, w.sync, timed)


class MsmWorkload:
    """One rank's shard of a G1 (or G2) multiexp: n bases [k_i] * generator generated on the device and kept resident (with the
    precomputed table), n uniform scalars in pinned host memory and in HBM."""

    def __init__(self, env, rng, n, precompute=True, keep_cpu_sample=False, group=None, tile_scalars_from=None):
        self.env, self.n = env, n
        w, zk, lib = env.w, env.zk, env.lib
        self.group = group or zk.G1
        g1 = self.group == zk.G1
        self.words = 18 if g1 else 36
        self.k = np.zeros((n, 4), dtype=np.uint64)
        self.k[:, 0] = rng.integers(1, 1 << 64, size=n, dtype=np.uint64)
        self.gen = gen_g1_limbs() if g1 else gen_g2_limbs()
        dxy, dinf, _ = zk.fixed_base_mul(w, self.group, self.gen, self.k, 64)
        self.bases = zk.Bases.from_device(w, self.group, dxy, n)
        cpu_log = min(env_int("B200ZK_CPU_LOG_N", 23), n.bit_length() - 1)
        self.cpu_bases = dxy.download(np.uint64, 12 << cpu_log).reshape(-1, 12) if (keep_cpu_sample and g1) else None
        dxy.free(); dinf.free()
        self.t_pre = None
        if precompute:
            t0 = time.perf_counter()
            self.bases.precompute(0)  # one-time, part of loading the bases (like Parameters::read); not in the timed region
            self.t_pre = time.perf_counter() - t0
        if tile_scalars_from is not None:  # a larger workload reuses a smaller scalar set (repeated): host generation of 2 GiB of scalars takes longer than the bench
            reps = (n + tile_scalars_from.shape[0] - 1) // tile_scalars_from.shape[0]
            self.scalars = np.tile(tile_scalars_from, (reps, 1))[:n]
        else:
            self.scalars = random_scalars(rng, n)
        hp = ctypes.c_void_p()
        assert lib.b200zk_host_alloc_pinned(self.scalars.nbytes, ctypes.byref(hp)) == 0
        self.hp = hp
        self.pinned = np.ctypeslib.as_array(ctypes.cast(hp, ctypes.POINTER(ctypes.c_uint64)), shape=(n, 4))
        self.pinned[:] = self.scalars
        self.d_scalars = w.to_device(self.scalars)
        self.d_part = w.alloc(320)
        self.d_out = w.alloc(320)
        self.out_host = np.zeros(self.words, dtype=np.uint64)

    def window(self):
        """(c, W) of the schedule the resident step runs"""
        lg = self.n.bit_length() - 1
        if self.t_pre is not None:
            c = min(22, max(8, lg if lg < 19 else lg - 1))
            c += 1 if 255 % c == 0 else 0
        else:
            c = 16 if lg >= 23 else 15 if lg >= 22 else 14
        return c, (256 + c - 1) // c

    def step_resident(self):
        env, w, lib = self.env, self.env.w, self.env.lib
        st = lib.b200zk_multiexp_dev(w.ctx, self.bases.handle, 0, self.d_scalars.ptr, self.n, None, self.d_part.ptr, None)
        assert st == 0, w.last_error()
        if env.world > 1:
            st = lib.b200zk_allgather_sum_dev(w.ctx, self.group, self.d_part.ptr, self.d_out.ptr)
            assert st == 0, w.last_error()

    def step_e2e_blocking(self):
        env, w, lib = self.env, self.env.w, self.env.lib
        if env.world == 1:
            # the call a bellman user makes: host scalars in, Jacobian result out (H2D + MSM + D2H inside)
            st = lib.b200zk_multiexp(w.ctx, self.bases.handle, 0, self.pinned.ctypes.data_as(ctypes.c_void_p), self.n, None, self.out_host.ctypes.data_as(ctypes.c_void_p))
            assert st == 0, w.last_error()
        else:
            self.run_e2e(1)

    def run_e2e(self, steps):
        env, w, lib = self.env, self.env.w, self.env.lib
        submit = lib.b200zk_multiexp_async if env.world == 1 else lib.b200zk_multiexp_sharded_async
        pend = None
        for _ in range(steps):
            job = ctypes.c_void_p()
            st = submit(w.ctx, self.bases.handle, 0, self.pinned.ctypes.data_as(ctypes.c_void_p), self.n, None, ctypes.byref(job))
            assert st == 0, w.last_error()
            if pend is not None:
                assert lib.b200zk_job_wait(pend, self.out_host.ctypes.data_as(ctypes.c_void_p)) == 0, w.last_error()
            pend = job
        assert lib.b200zk_job_wait(pend, self.out_host.ctypes.data_as(ctypes.c_void_p)) == 0, w.last_error()

    def check(self):
        """sum s_i [k_i]G == [sum s_i k_i]G, the right-hand side through the fixed-base path (all ranks' shards together)"""
        env, w, zk = self.env, self.env.w, self.env.zk
        self.step_resident()
        w.sync()
        got = (self.d_out if env.world > 1 else self.d_part).download(np.uint64, self.words)
        got_aff, got_inf = zk.into_affine(w, self.group, got)
        total = dot_mod_r(self.k[:, 0], self.scalars)
        if env.world > 1:
            torch, dist = env.torch, env.dist
            tt = torch.tensor([(total >> (32 * i)) & 0xFFFFFFFF for i in range(8)], dtype=torch.int64, device="cuda")
            gathered = [torch.zeros_like(tt) for _ in range(env.world)]
            dist.all_gather(gathered, tt)
            total = sum(sum(int(v) << (32 * i) for i, v in enumerate(g.cpu().tolist())) for g in gathered) % FR_MODULUS
        tl = np.array([[(total >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]], dtype=np.uint64)
        exp_xy, exp_inf, _ = zk.fixed_base_mul(w, self.group, self.gen, tl, 255)
        want = exp_xy.download(np.uint64, 12 if self.group == zk.G1 else 24)
        if got_inf[0] or not np.array_equal(got_aff[0], want):
            print(json.dumps({"error": "MSM result failed the sum(s_i k_i) identity check", "rank": env.rank, "n": self.n}))
            sys.exit(1)

    def free(self):
        self.bases.free()
        for b in (self.d_scalars, self.d_part, self.d_out):
            b.free()
        self.env.lib.b200zk_host_free_pinned(self.hp)
        self.pinned = None


def flat_secondary(extra):
    """the secondary metrics as flat scalars (they ride in `roofline`, which the driver's parser keeps)"""
    out = {}
    for name in ("fr_fft_2^24", "fr_ifft_2^24", "fr_coset_fft_2^24", "fr_icoset_fft_2^24"):
        if name in extra:
            out[f"{name}_ms"] = extra[name]["ms"]
    if "fr_fft_2^24" in extra:
        e = extra["fr_fft_2^24"]
        out.update({"ntt_2^24_per_s": e["ntt_per_s"], "ntt_multiplier_frac": e["multiplier_frac"], "ntt_hbm_frac": e["hbm_frac_of_measured_peak"]})
    for name in ("g2_msm_2^22", "g2_msm_2^24"):
        if name in extra:
            out[f"{name}_points_per_s"] = extra[name]["points_per_s"]
    sp = extra.get("sapling_spend_proofs")
    if sp:
        out.update({"spend_proofs_per_s": sp["proofs_per_s"], "spend_single_proof_ms": sp["single_stream_ms_per_proof"]})
    if extra.get("verify_proofs"):
        out["verify_proofs_per_s"] = extra["verify_proofs"]["proofs_per_s"]
    for k, v in (extra.get("strong_scaling") or {}).items():
        out[f"strong_{k}_points_per_s"] = v["points_per_s"]
        out[f"strong_{k}_ms"] = v["ms_per_step"]
    return out


def traffic_from_profile():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full capture
    (profiles/r02_traffic.json); None when the capture does not match the precomputed-table configuration."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        return {"dram_bytes_per_launch": t["dram_bytes_per_launch"], "source": t["source"], "workload": t["workload"]}
    except Exception:
        return None


def bench_extra(env, rng):
    """The other two metrics of BASELINE.json (Fr NTT/s, Sapling Spend proofs/s) and the configs the headline line does not cover:
    G2 multiexps at 2^22 / 2^24 (configs[3]), the Sprout-JoinSplit-shaped proof, and strong scaling of fixed-size G1 multiexps."""
    w, lib, L, world, timed = env.w, env.lib, env.L, env.world, env.timed
    out = {}
    peaks_hbm = 6539.9
    try:
        peaks_hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for log_m in (24, 17):
        m = 1 << log_m
        coeffs = random_scalars(rng, m)  # canonical values < r are valid Montgomery residues
        d = w.to_device(coeffs)
        for kind, name in ((L.FFT, "fft"), (L.IFFT, "ifft"), (L.COSET_FFT, "coset_fft"), (L.ICOSET_FFT, "icoset_fft")):
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
    # G2 multiexps (BASELINE configs[3]: the B query of a proof), per-rank replicas, result checked through the dlog identity
    for log_n in (22, 24):
        g2 = MsmWorkload(env, rng, 1 << log_n, precompute=True, group=env.zk.G2)
        one = env.world
        env.world = 1  # replicas: no gather
        g2.check()
        for _ in range(2):
            g2.step_resident()
        steps = 5 if log_n == 22 else 3
        ms = timed(g2.step_resident, steps) / steps
        env.world = one
        out[f"g2_msm_2^{log_n}"] = {"points_per_s": world * (1 << log_n) / (ms * 1e-3), "ms": ms, "table_setup_s": g2.t_pre, "replicas": world}
        g2.free()
    out["sapling_spend_proofs"] = bench_spend_proofs(env, rng)
    # the reference's largest circuit: Sprout JoinSplit on Groth16, 1 989 085 constraints -> m = 2^21
    out["sprout_joinsplit_shaped_proofs"] = bench_spend_proofs(env, rng, shape=synthetic.SPROUT_SHAPE)
    out["strong_scaling"] = bench_strong_scaling(env, rng)
    out["verify_proofs"] = bench_verify(env)
    return out


class _SquareChain:
    """a small real circuit for the verifier measurement: x_(i+1) = x_i^2 + i over `rounds` constraints, the last value public"""

    def __init__(self, x0, rounds):
        self.x0, self.rounds = x0, rounds

    def synthesize(self, cs):
        p = FR_MODULUS
        v = self.x0 % p
        x = cs.alloc(lambda: v)
        for i in range(self.rounds):
            nv = (v * v + i) % p
            y = cs.alloc_input(lambda nv=nv: nv) if i == self.rounds - 1 else cs.alloc(lambda nv=nv: nv)
            cs.enforce([(x, 1)], [(x, 1)], [(y, 1), (("in", 0), -i)])
            x, v = y, nv


def bench_verify(env):
    """groth16::verify_proof on the device (b200zk_verify_proofs): a real (small) circuit -- CRS generated on the GPU, proof made on
    the GPU, accepted by the GPU verifier and rejected for a wrong public input -- then a batch of copies for the rate."""
    w, zk, world = env.w, env.zk, env.world
    circ = _SquareChain(3, 16)
    asm = zk.KeypairAssembly()
    asm.alloc_input()
    _SquareChain(0, 16).synthesize(asm)
    for i in range(asm.num_inputs):
        asm.enforce([(("in", i), 1)], [], [])
    gen = zk.generate_parameters(w, asm, gen_g1_limbs(), gen_g2_limbs(), 0x1111, 0x2222, 0x3333, 0x4444, 0x5555)
    params = gen.to_device(w)
    proof = zk.create_proof(w, circ, params, 0xABCDEF, 0x123456)
    public = zk.synthesize(circ).input_assignment[1:]
    pvk = zk.PreparedVerifyingKey(w, gen.alpha_g1, gen.beta_g2, gen.gamma_g2, gen.delta_g2, gen.ic)
    if zk.verify_proofs(w, pvk, [proof, proof], [public, [(public[0] + 1) % FR_MODULUS]]) != [True, False]:
        print(json.dumps({"error": "device verify_proof: accept / reject check failed"}))
        sys.exit(1)
    t0 = time.perf_counter()
    zk.verify_proof(w, pvk, proof, public)
    single_ms = (time.perf_counter() - t0) * 1e3
    n = env_int("B200ZK_VERIFY_BATCH", 2048)
    zk.verify_proofs(w, pvk, [proof] * n, [public] * n)
    _line_up(world)
    t0 = time.perf_counter()
    ok = zk.verify_proofs(w, pvk, [proof] * n, [public] * n)
    dt = _slowest_rank(time.perf_counter() - t0, world)
    assert all(ok)
    return {"proofs_per_s": world * n / dt, "batch": world * n, "single_proof_ms": single_ms, "public_inputs": len(public),
            "note": "one thread block per proof: 3 Miller loops on 3 threads + the final exponentiation on one; host wall clock incl. copies"}


def bench_strong_scaling(env, rng):
    """BASELINE configs[2] read literally: ONE multiexp of 2^24 (and 2^26) points split by base range over the N ranks -- every rank
    holds n / N bases (and their table), computes its partial and the partials are gathered and summed (NCCL) inside the step.
    value = total points / max-over-ranks device time.  The driver's 1/2/4/8 runs give the strong-scaling curve."""
    out = {}
    world = env.world
    if world & (world - 1):
        return out
    base_scalars = None
    for total_log in (24, 26):
        shard_log = total_log - (world.bit_length() - 1)
        wl = MsmWorkload(env, rng, 1 << shard_log, precompute=True, tile_scalars_from=base_scalars)
        if base_scalars is None:
            base_scalars = wl.scalars
        wl.check()
        for _ in range(3):
            wl.step_resident()
        steps = 5 if total_log == 24 else 3
        ms = env.timed(wl.step_resident, steps) / steps
        # e2e: this rank's n / N exponents from pinned host memory, result back, depth-2 futures
        wl.run_e2e(2)
        env.barrier()
        t0 = time.perf_counter()
        wl.run_e2e(steps)
        e2e = _slowest_rank(time.perf_counter() - t0, world) / steps * 1e3
        out[f"2^{total_log}"] = {"points_per_s": (1 << total_log) / (ms * 1e-3), "ms_per_step": ms, "e2e_points_per_s": (1 << total_log) / (e2e * 1e-3),
                                 "e2e_ms_per_step": e2e, "points_per_rank": 1 << shard_log, "ranks": world, "table_setup_s": wl.t_pre}
        wl.free()
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


def bench_spend_proofs(env, rng, shape=None):
    """Sapling-Spend-shaped create_proof: 98 785 constraints -> m = 2^17; H 131 071, L 98 638, A 8 + 85 382,
    B 1 + 61 299 (G1 and G2) bases; witness-like scalars (half of them 0/1).  The CRS is synthetic ([k]G points generated on
    the device), so the proofs are not meaningful -- the arithmetic is identical to a real Spend proof.  Before anything is
    timed, rank 0 proves the same assignment with the CPU oracle (C++ H block + the reference's eight multiexps + assembly) and
    the GPU proof must have the same 192 bytes; that CPU run is also the proofs/s baseline.  Proofs of a batch are
    independent: each rank proves its own share on `streams` concurrent contexts."""
    import threading

    w, zk, L, world = env.w, env.zk, env.L, env.world
    light = shape is not None  # another circuit shape (the Sprout JoinSplit one): single-call latency and a short batch only
    shape = shape or synthetic.SPEND_SHAPE
    sizes = synthetic.crs_sizes(shape)
    log_m = (shape["n_con"] - 1).bit_length()
    want_cpu = env.rank == 0 and not env.args.no_cpu and not light
    dev, host = {}, {}
    for name, n in sizes.items():
        group, gen = (L.G2, gen_g2_limbs()) if name == "b_g2" else (L.G1, gen_g1_limbs())
        dxy, dinf, _ = zk.fixed_base_mul(w, group, gen, synthetic.base_multipliers(rng, n), 64)
        if want_cpu:
            host[name] = dxy.download(np.uint64, n * (24 if group == L.G2 else 12)).reshape(n, -1)
        dev[name] = zk.Bases.from_device(w, group, dxy, n)
        dxy.free(); dinf.free()
    vk1d, _, _ = zk.fixed_base_mul(w, L.G1, gen_g1_limbs(), synthetic.base_multipliers(rng, 3), 64)
    vk2d, _, _ = zk.fixed_base_mul(w, L.G2, gen_g2_limbs(), synthetic.base_multipliers(rng, 2), 64)
    head1, head2 = vk1d.download(np.uint64, 36).reshape(3, 12), vk2d.download(np.uint64, 48).reshape(2, 24)
    t0 = time.perf_counter()
    for q in dev.values():
        q.precompute(0)  # one-time, at CRS load
    t_pre = time.perf_counter() - t0
    params = zk.Parameters(w, dev["h"], dev["l"], dev["a"], dev["b_g1"], dev["b_g2"], head1[0], head1[1], head2[0], head1[2], head2[1])
    asg = synthetic.spend_assignment(rng, shape)
    order = ("a", "b", "c", "inputs", "aux", "a_aux_density", "b_input_density", "b_aux_density")
    # the assignment a prover hands over lives in page-locked memory (the e2e rule: H2D from pinned host memory)
    pinned = [w.pinned_copy(asg[k]) for k in order]
    r, s = 0x1234567890ABCDEF1234567890ABCDEF, 0x0FEDCBA0987654321FEDCBA098765432

    def prove(worker):
        return zk.create_proof_from_assignment(worker, params, *pinned, r, s)

    p0 = prove(w)
    p1 = prove(w)
    assert np.array_equal(p0.a, p1.a) and np.array_equal(p0.c, p1.c)  # deterministic
    cpu = None
    if want_cpu:  # parity gate + CPU baseline: the reference's path on the host cores, same CRS, same assignment, same (r, s)
        from oracle import cref, spend

        crs = spend.HostCrs(host["h"], host["l"], host["a"], host["b_g1"], host["b_g2"], head1[0], head1[1], head2[0], head1[2], head2[1])
        want, t_msm, t_asm = spend.prove(crs, asg, r, s)
        if p0.write(w) != want:
            print(json.dumps({"error": "Spend-shaped GPU proof differs from the CPU oracle's 192 bytes"}))
            sys.exit(1)
        cpu = {"proofs_per_s": 1.0 / t_msm, "seconds_h_block_and_multiexps": t_msm, "seconds_assembly_python": t_asm, "cores": cref.hardware_threads(), "kind": "port",
               "sample": "ONE Spend-shaped create_proof: C++ port of the H block (7 FFTs, best_fft) and of the reference's eight multiexps on the host threads; "
                         "the assembly (5 scalar multiplications, prover.rs:326-363) runs in the Python restatement and is excluded from proofs_per_s",
               "parity": "GPU proof bytes == CPU oracle proof bytes (192 B)"}
        del crs, host
    t0 = time.perf_counter()
    reps = 3 if light else 5
    for _ in range(reps):
        prove(w)
    single_ms = (time.perf_counter() - t0) / reps * 1e3
    one = tuple(pinned) + (r, s)
    shape_str = (f"m=2^{log_m} ({shape['n_con']} constraints), multiexps {sizes['h']}/{sizes['l']}/{shape['n_in']}+{shape['a_dense']}/"
                 f"{shape['b_in_dense']}+{shape['b_aux_dense']} (G1) and {shape['b_in_dense']}+{shape['b_aux_dense']} (G2), synthetic CRS and densities")
    if light:
        ref = p0.write(w)
        got = zk.create_proofs_from_assignments(w, params, [one] * 4, 4)
        assert all(g.write(w) == ref for g in got)
        _line_up(world)
        t0 = time.perf_counter()
        zk.create_proofs_from_assignments(w, params, [one] * 8, 4)
        bdt = _slowest_rank(time.perf_counter() - t0, world)
        for q in dev.values():
            q.free()
        return {"single_call_ms_per_proof": single_ms, "proofs_per_s": world * 8 / bdt, "lockstep": 4, "batch": world * 8, "shape": shape_str, "crs_precompute_s": t_pre}
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
    # the batch API: `lockstep` proofs share five batched multiexps (b200zk_groth16_prove_batch); three contexts (host threads)
    # alternate so that the uploads and the serial tails of one group overlap the multiexps of the others
    # (measured: lock-step 8 x 2 contexts 387, 8 x 3 398, 16 x 2 393, 16 x 3 398, 4 x 4 392 proofs/s)
    lockstep = env_int("B200ZK_SPEND_LOCKSTEP", 8)
    bstreams = env_int("B200ZK_SPEND_BATCH_STREAMS", 3)
    groups = env_int("B200ZK_SPEND_GROUPS", 4)
    bworkers = [zk.Worker(w.device) for _ in range(bstreams)]
    ref = p0.write(w)
    for x in bworkers:
        got = zk.create_proofs_from_assignments(x, params, [one] * lockstep, lockstep)
        assert all(g.write(x) == ref for g in got)  # the batch gives the single-call proof (== the oracle's bytes, checked above)

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
    for q in dev.values():
        q.free()
    batched = {"proofs_per_s": world * bstreams * lockstep * groups / bdt, "lockstep": lockstep, "contexts_per_gpu": bstreams,
               "batch": world * bstreams * lockstep * groups, "api": "b200zk_groth16_prove_batch"}
    return {"proofs_per_s": max(world * streams * per_thread / dt, batched["proofs_per_s"]), "batched": batched,
            "independent_calls_proofs_per_s": world * streams * per_thread / dt, "single_stream_ms_per_proof": single_ms, "streams_per_gpu": streams,
            "batch": world * streams * per_thread, "timing": "host wall clock around the prove calls incl. H2D of a/b/c/assignments and D2H of the proofs; with N GPUs every rank proves its own share and the time is that of the slowest rank",
            "shape": shape_str, "crs_precompute_s": t_pre, "cpu_baseline": cpu}


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


def bench_cpu_baseline(bases, scalars, log_s, full=True):
    """The oracle's C++ port of the reference on the host cores, each on a bounded sample:
      * multiexp (the headline metric): the first 2^log_s (base, scalar) pairs of rank 0's own workload (bases downloaded from the
        GPU), one multiexp with the reference's c = ceil(ln n) windows, one pool task per window;
      * Fr NTT: one fft of 2^24 elements with best_fft (domain.rs:261-270: 2^floor(log2 cores) sub-FFTs on the host threads);
      * the Spend-shaped proof is timed inside bench_spend_proofs (it doubles as the parity gate) and copied here by the caller."""
    from oracle import cref

    s = 1 << log_s
    cores = cref.hardware_threads()
    rb = cref.ResidentBases.load("g1", bases[:s])
    rb.multiexp(scalars[: s // 16])
    t0 = time.perf_counter()
    st, _ = rb.multiexp(scalars[:s])
    dt = time.perf_counter() - t0
    assert st == 0
    rb.free()
    out = {"value": s / dt, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"first 2^{log_s} (base, scalar) pairs of rank 0's workload, one multiexp, one pool task per window (c = ceil(ln n)), wall {dt:.2f} s"}
    if full:
        log_m = env_int("B200ZK_CPU_NTT_LOG_M", 24)
        coeffs = np.ascontiguousarray(scalars[: 1 << log_m]) if scalars.shape[0] >= (1 << log_m) else random_scalars(np.random.default_rng(5), 1 << log_m)
        cref.fft(coeffs[: 1 << 16])
        t0 = time.perf_counter()
        cref.fft(coeffs, cref.FFT)
        dt = time.perf_counter() - t0 - 0.0
        out.update({"ntt_per_s": 1.0 / dt, "ntt_sample": f"one fft of 2^{log_m} Fr elements, best_fft on {cores} host threads, wall {dt:.2f} s (includes one copy of the 512 MiB vector)"})
    return out


if __name__ == "__main__":
    main()
