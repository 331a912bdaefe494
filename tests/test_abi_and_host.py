"""CPU-only checks of the boundary: libb200zk.so loads and exports every symbol include/b200zk.h declares,
the product has no CPU fallback, and the host-side mirror keeps the reference's argument / error behaviour."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "b200zk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200zk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import zcash_gpu_thesis_b200 as zk

    lib = ctypes.CDLL(zk._lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200zk.h but not exported"
    # the ctypes table covers exactly the header
    assert sorted(zk._lib.SIGNATURES) == syms
    # ... and so does the generated Rust -sys crate (bindings/rust/b200zk-sys/src/lib.rs, tools/gen_rust_bindings.py)
    rust = open(os.path.join(ROOT, "bindings", "rust", "b200zk-sys", "src", "lib.rs")).read()
    assert sorted(re.findall(r"pub fn (b200zk_[a-z0-9_]+)\(", rust)) == syms
    # ... and the glue crate only calls entry points that exist
    glue = open(os.path.join(ROOT, "bindings", "rust", "bellman-b200zk", "src", "lib.rs")).read()
    assert set(re.findall(r"\b(b200zk_[a-z0-9_]+)\(", glue)) <= set(syms)


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path refuses to run (status ERR_CUDA), it never computes on the CPU."""
    import zcash_gpu_thesis_b200 as zk

    lib = zk._lib.load()
    if lib.b200zk_device_count() > 0:
        pytest.skip("a GPU is present")
    ctx = ctypes.c_void_p()
    assert lib.b200zk_init(0, ctypes.byref(ctx)) == zk._lib.ERR_CUDA
    with pytest.raises(zk.CudaError):
        zk.Worker(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "zcash-gpu-thesis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("never imports `oracle/`", ""), f"{f} references the oracle"
                assert "cref" not in src, f"{f} references the CPU port"


def test_density_tracker_and_window_formula():
    """multiexp.rs:99-138 DensityTracker; multiexp.rs:296-300 window size."""
    import zcash_gpu_thesis_b200 as zk
    from oracle.multiexp import window_size

    d = zk.DensityTracker()
    for _ in range(5):
        d.add_element()
    d.inc(1); d.inc(1); d.inc(4)
    assert d.get_total_density() == 2 and d.get_query_size() == 5
    assert list(d.as_bytes()) == [0, 1, 0, 0, 1]
    assert zk.FullDensity().get_query_size() is None
    for n, c in ((1, 3), (31, 3), (32, 4), (1 << 14, 10), (1 << 16, 12), (1 << 20, 14), (1 << 24, 17), (1 << 26, 19)):
        assert zk.bellman.window_size_reference(n) == c == window_size(n)


def test_fr_limb_helpers_roundtrip():
    import zcash_gpu_thesis_b200 as zk
    from oracle.fields import Fr

    for v in (0, 1, 7, Fr.p - 1, 0x1234567890ABCDEF1234567890ABCDEF):
        l = zk.bellman.fr_to_mont_limbs(v)
        assert [int(x) for x in l] == Fr.to_mont_limbs(v)
        assert zk.bellman.fr_from_mont_limbs(l) == v


def test_from_coeffs_degree_check_is_host_side():
    """domain.rs:59-61: exp >= Fr::S fails before anything touches the device (checked on the padded length loop)."""
    import zcash_gpu_thesis_b200 as zk

    class FakeCoeffs:
        shape = ((1 << 32) + 1, 4)

    # emulate the loop of from_coeffs without allocating 2^32 elements
    m, exp = 1, 0
    with pytest.raises(zk.PolynomialDegreeTooLarge):
        while m < FakeCoeffs.shape[0]:
            m *= 2
            exp += 1
            if exp >= zk.bellman.FR_S:
                raise zk.PolynomialDegreeTooLarge()


def test_product_synthesis_matches_the_oracle():
    """ProvingAssignment / KeypairAssembly of the product (host logic, prover.rs:84-234, generator.rs:57-212) against the oracle's"""
    import zcash_gpu_thesis_b200 as zk
    from oracle.fields import Fr
    from oracle.groth16 import KeypairAssembly as OracleAssembly
    from oracle.groth16 import synthesize_assignment
    from oracle.pairing import Bls12
    from tests.test_gpu_groth16 import MiMCLike

    consts = [(i * 7919 + 3) ** 5 % Fr.p for i in range(9)]
    circ = MiMCLike(12345, Fr.p - 5, consts)
    want = synthesize_assignment(Bls12, circ)
    got = zk.synthesize(circ)
    assert got.a == want.a and got.b == want.b and got.c == want.c
    assert got.input_assignment == want.input_assignment and got.aux_assignment == want.aux_assignment
    assert [bool(x) for x in got.a_aux_density] == want.a_aux_density and [bool(x) for x in got.b_aux_density] == want.b_aux_density
    assert [bool(x) for x in got.b_input_density] == want.b_input_density
    a, b, c, inputs, aux, da, dbi, dba, r, s = got.as_tuple(Fr.p + 3, 9)
    assert a.shape == (len(want.a), 4) and inputs.shape == (len(want.input_assignment), 4) and da.dtype.name == "uint8" and r == 3
    assert list(map(int, a[1])) == Fr.to_mont_limbs(want.a[1])
    asm, oasm = zk.KeypairAssembly(), OracleAssembly(Fr)
    for x in (asm, oasm):
        x.alloc_input()
        circ.synthesize(x)
    assert asm.at_aux == oasm.at_aux and asm.bt_aux == oasm.bt_aux and asm.ct_inputs == oasm.ct_inputs and asm.num_constraints == oasm.num_constraints


def test_every_transform_size_maps_to_built_kernels():
    """b200zk_ntt_plan (host logic, no GPU): for every log_m the library accepts, with the default threshold, with the one the
    tests use to force the large kernels and with the large kernels off, for single transforms and for the batches of the H
    blocks -- the passes add up to log_m, each (stages, columns) pair is a kernel shape ntt.cu instantiates, a later pass never
    has more columns than stage bits below it, and every pass has at least one tile."""
    import zcash_gpu_thesis_b200 as zk

    small = {(b, 2) for b in (5, 6, 7, 8)} | {(b, 0) for b in (5, 6, 7, 8, 9)}
    small_first_only = {(3, 0), (4, 0)}
    large = {(b, 1) for b in (6, 7, 8)} | {(b, 2) for b in (6, 7, 8, 9)}
    for large_from in (20, 12, 31):
        for batch in (1, 3, 24, 192):
            for sm in (148, 132):
                for log_m in range(3, 31):
                    passes, radix4 = zk.ntt_plan(log_m, large_from, sm, batch)
                    assert passes, (log_m, large_from)
                    assert sum(b for b, _ in passes) == log_m
                    assert radix4 == (log_m >= max(large_from, 12))
                    s0 = 0
                    for i, (b, q) in enumerate(passes):
                        shapes = large if radix4 else (small | small_first_only if i == 0 else small)
                        assert (b, q) in shapes, f"log_m {log_m}, large_from {large_from}: pass {i} = ({b}, {q}) is not a built kernel"
                        assert i == 0 or s0 >= q
                        assert log_m >= b + q
                        s0 += b
    assert zk.ntt_plan(2)[0] == [] and zk.ntt_plan(31)[0] == []
