"""Integer-pipe calibration on the GPU box: IMAD / IMAD.WIDE / IMAD.HI / IADD3 rates and the Fq / Fr Montgomery
multiply rate over all SMs (the integer roofline must be measured, not assumed)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zcash_gpu_thesis_b200 as zk  # noqa: E402

NAMES = {0: "imad_lo", 1: "imad_wide", 2: "imad_hi", 3: "iadd3_carry", 4: "fq_mul", 5: "fr_mul"}


def main():
    w = zk.Worker(0)
    out = {"sm_count": w.sm_count()}
    for kind, name in NAMES.items():
        iters = 4000 if kind <= 3 else 2000
        best = max(w.microbench(kind, iters) for _ in range(3))
        out[name + "_per_s"] = best
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/microbench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
