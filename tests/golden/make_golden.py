"""Extracts the reference's own golden vectors for the hot path into tests/golden/kat.json.

Run in the build container only (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden.py
Sources (all under /root/reference/librustzcash):
  pairing/src/bls12_381/fr.rs:1240-1262, 1306-1325     Fr mul / square KATs
  pairing/src/bls12_381/fq.rs:2558-2584, 2630-2651     Fq mul / square KATs
  pairing/src/bls12_381/ec.rs:1060-1175                G1 add / double KATs (canonical coordinates)
  pairing/src/bls12_381/fq2.rs:273-681                 Fq2 square / mul / inverse / add / sub / negate / double KATs
  pairing/src/bls12_381/tests/*.dat                    1000 multiples of the generators, 4 encodings
  pairing/src/bls12_381/tests/mod.rs:5-53              pairing KAT e(G1, G2) (RELIC)
  bellman/src/groth16/tests/mod.rs:98-400              test_xordemo constants (DummyEngine, Fr = Z/64513)
Only literal test constants are extracted (no source code is copied).
"""
import hashlib
import json
import os
import re

REF = "/root/reference/librustzcash"
HERE = os.path.dirname(os.path.abspath(__file__))


def hexes(path, lo, hi):
    lines = open(os.path.join(REF, path)).read().split("\n")[lo - 1:hi]
    keep = []
    for ln in lines:
        if "XorShiftRng" in ln:  # the property-test part of the function starts here
            break
        keep.append(ln)
    return re.findall(r"0x[0-9a-fA-F]+", "\n".join(keep))


def main():
    out = {}
    h = hexes("pairing/src/bls12_381/fr.rs", 1240, 1270)
    assert len(h) == 12
    out["fr_mul"] = {"a": h[0:4], "b": h[4:8], "out": h[8:12], "src": "fr.rs:1240-1262 (raw Montgomery limbs)"}
    h = hexes("pairing/src/bls12_381/fr.rs", 1306, 1330)
    assert len(h) == 8
    out["fr_square"] = {"a": h[0:4], "out_repr": h[4:8], "src": "fr.rs:1306-1325 (a raw limbs; out via from_repr)"}
    h = hexes("pairing/src/bls12_381/fq.rs", 2558, 2590)
    assert len(h) == 18
    out["fq_mul"] = {"a": h[0:6], "b": h[6:12], "out": h[12:18], "src": "fq.rs:2558-2584"}
    h = hexes("pairing/src/bls12_381/fq.rs", 2630, 2660)
    assert len(h) == 12
    out["fq_square"] = {"a": h[0:6], "out_repr": h[6:12], "src": "fq.rs:2630-2651"}
    h = hexes("pairing/src/bls12_381/ec.rs", 1060, 1126)
    assert len(h) == 36
    out["g1_add"] = {"p": [h[0:6], h[6:12]], "q": [h[12:18], h[18:24]], "sum": [h[24:30], h[30:36]],
                     "src": "ec.rs:1060-1125 (canonical affine coordinates, z = 1)"}
    h = hexes("pairing/src/bls12_381/ec.rs", 1128, 1176)
    assert len(h) == 24
    out["g1_double"] = {"p": [h[0:6], h[6:12]], "dbl": [h[12:18], h[18:24]], "src": "ec.rs:1128-1175"}
    # Fq2 KATs: canonical (from_repr) limbs, c0 then c1 of every element in source order
    fq2 = {}
    for name, lo, hi, parts in (("square", 273, 346, ("a", "out")), ("mul", 347, 410, ("a", "b", "out")), ("inverse", 411, 459, ("a", "out")),
                                ("add", 460, 523, ("a", "b", "out")), ("sub", 524, 587, ("a", "b", "out")), ("negate", 588, 634, ("a", "out")),
                                ("double", 635, 681, ("a", "out"))):
        h = hexes("pairing/src/bls12_381/fq2.rs", lo, hi)
        assert len(h) == 12 * len(parts), (name, len(h))
        fq2[name] = {part: [h[12 * i:12 * i + 6], h[12 * i + 6:12 * i + 12]] for i, part in enumerate(parts)}
        fq2[name]["src"] = f"fq2.rs:{lo}-{hi} (canonical limbs via from_repr; [c0, c1])"
    out["fq2"] = fq2
    dat = {}
    for name, sz in (("g1_compressed", 48), ("g1_uncompressed", 96), ("g2_compressed", 96), ("g2_uncompressed", 192)):
        b = open(os.path.join(REF, f"pairing/src/bls12_381/tests/{name}_valid_test_vectors.dat"), "rb").read()
        assert len(b) == 1000 * sz
        dat[name] = {"entry_bytes": sz, "entries": 1000, "sha256": hashlib.sha256(b).hexdigest(),
                     "first": [b[i * sz:(i + 1) * sz].hex() for i in range(4)], "last": b[999 * sz:].hex()}
    out["dat"] = dat
    src = open(os.path.join(REF, "pairing/src/bls12_381/tests/mod.rs")).read().split("\n")[20:55]
    vals = re.findall(r'from_str\("(\d+)"\)', "\n".join(src))
    assert len(vals) == 12
    out["pairing_g1_g2"] = {"src": "pairing/src/bls12_381/tests/mod.rs:5-53 (RELIC): e(G1::one(), G2::one()) as c0.c0.c0, c0.c0.c1, c0.c1.c0, ... decimal",
                            "fq12": vals}
    out["xordemo"] = {
        "src": "bellman/src/groth16/tests/mod.rs:98-400 + tests/dummy_engine.rs (Fr = Z/64513, generator 5, S = 10)",
        "modulus": 64513, "generator": 5, "s": 10, "root_2_10": 57751, "root_2_3": 20201,
        "alpha": 48577, "beta": 22580, "gamma": 53332, "delta": 5481, "tau": 3673, "r": 27134, "s_rand": 17146,
        "u_i": [59158, 48317, 21767, 10402], "v_i": [0, 0, 60619, 30791], "w_i": [0, 23320, 41193, 41193],
        "h_coeffs": [5040, 11763, 10755, 63633, 128, 9747, 8739],
    }
    json.dump(out, open(os.path.join(HERE, "kat.json"), "w"), indent=1)
    print("wrote", os.path.join(HERE, "kat.json"))


if __name__ == "__main__":
    main()
