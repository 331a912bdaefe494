"""bellman::multiexp on the GPU vs the CPU oracle (multiexp.rs:337-376 test_with_bls12 re-targeted, plus the
edge cases the reference's Source semantics imply, multiexp.rs:42-68, 174-196).  Projective results are
compared after into_affine (the reference compares with a Jacobian-aware PartialEq, ec.rs:45-85)."""
import numpy as np
import pytest

from oracle import cref
from oracle.fields import Fr, int_to_limbs
from tests import util

pytestmark = pytest.mark.gpu


def _check(worker, group, bases_xy, exps, density=None, offset=0, inf=None, bases=None):
    import zcash_gpu_thesis_b200 as zk

    code = zk.G1 if group == "g1" else zk.G2
    if bases is None:
        bases = zk.Bases(worker, code, bases_xy, inf)
    dm = zk.FullDensity() if density is None else zk.DensityTracker(density)
    got = zk.multiexp(worker, (bases, offset), dm, exps)
    st, want = cref.multiexp(group, bases_xy, exps, density=density, base_offset=offset, inf=inf)
    assert st == 0
    got_aff, got_inf = zk.into_affine(worker, code, got)
    want_aff, want_inf = cref.into_affine(group, want)
    assert bool(got_inf[0]) == want_inf
    assert np.array_equal(got_aff[0], want_aff)
    return got_aff[0], bool(got_inf[0])


@pytest.mark.parametrize("n", [0, 1, 2, 3, 31, 32, 33, 100, 1 << 10, (1 << 12) + 7, 1 << 14])
def test_g1_multiexp_vs_oracle(worker, n):
    r = util.rng(500 + n)
    xy, ks = util.random_bases("g1", r, max(n, 1))
    exps = util.random_fr_repr(r, n)
    aff, inf = _check(worker, "g1", xy[:n] if n else xy[:0], exps)
    # independent check through the discrete logs: sum s_i [k_i]G = [sum s_i k_i]G
    want_xy, want_inf = util.affine_of_scalar("g1", util.expected_scalar(ks, exps))
    assert inf == want_inf and (inf or np.array_equal(aff, want_xy))


def test_g1_multiexp_2_16_naive_sum(worker):
    """BASELINE config 1: 2^16 random bases / scalars vs the naive sum (here: oracle multiexp and the dlog identity)."""
    n = 1 << 16
    r = util.rng(516)
    xy, ks = util.random_bases("g1", r, n)
    exps = util.random_fr_repr(r, n)
    aff, inf = _check(worker, "g1", xy, exps)
    want_xy, want_inf = util.affine_of_scalar("g1", util.expected_scalar(ks, exps))
    assert not inf and np.array_equal(aff, want_xy)


@pytest.mark.parametrize("c", [2, 3, 5, 8, 13, 16, 19])
def test_g1_window_override(worker, c):
    """Every window width gives the same group element (signed-digit recoding, carries across windows)."""
    n = 777
    r = util.rng(600 + c)
    xy, ks = util.random_bases("g1", r, n)
    exps = util.random_fr_repr(r, n)
    exps[0] = int_to_limbs(Fr.p - 1, 4)  # top bits set, exercises the carry into the last window
    exps[1] = int_to_limbs((1 << 254) + (1 << 253), 4)
    worker.set_msm_window(c)
    try:
        _check(worker, "g1", xy, exps)
    finally:
        worker.set_msm_window(0)


def test_special_scalars(worker):
    """all-zero, all-one, r-1, powers of two, duplicated bases (doubling branch), P + (-P) -> identity."""
    n = 300
    r = util.rng(700)
    xy, ks = util.random_bases("g1", r, n)
    zero = np.zeros((n, 4), dtype=np.uint64)
    one = zero.copy()
    one[:, 0] = 1
    aff, inf = _check(worker, "g1", xy, zero)
    assert inf
    _check(worker, "g1", xy, one)
    rm1 = np.tile(np.array(int_to_limbs(Fr.p - 1, 4), dtype=np.uint64), (n, 1))
    _check(worker, "g1", xy, rm1)
    pw = zero.copy()
    for i in range(n):
        b = (i * 7) % 255
        pw[i, b // 64] = np.uint64(1) << np.uint64(b % 64)
    _check(worker, "g1", xy, pw)
    # duplicated bases with equal scalars hit the doubling branch of the mixed add (ec.rs:473-475)
    dup = np.repeat(xy[:1], n, axis=0)
    _check(worker, "g1", dup, one)
    same = np.tile(util.random_fr_repr(r, 1), (n, 1))
    _check(worker, "g1", dup, same)
    # s*P + (r-s)*P = identity
    s = util.rows_to_ints(util.random_fr_repr(r, 1))[0]
    pair = np.array([int_to_limbs(s, 4), int_to_limbs(Fr.p - s, 4)], dtype=np.uint64)
    aff, inf = _check(worker, "g1", dup[:2], pair)
    assert inf
    # small (witness-like) scalars: mostly 0/1/small values
    small = zero.copy()
    small[:, 0] = r.integers(0, 4, size=n, dtype=np.uint64)
    _check(worker, "g1", xy, small)


def test_density_maps_and_offsets(worker):
    """DensityTracker semantics (multiexp.rs:174-196): a base is consumed only where the density bit is set;
    the source starts at `offset` (groth16/mod.rs:456-481)."""
    import zcash_gpu_thesis_b200 as zk

    n = 5000
    r = util.rng(800)
    xy, ks = util.random_bases("g1", r, 4000)
    exps = util.random_fr_repr(r, n)
    exps[r.integers(0, n, size=500)] = 0
    exps[r.integers(0, n, size=500)] = (1, 0, 0, 0)
    bases = zk.Bases(worker, zk.G1, xy)
    for p, offset in ((0.5, 0), (0.5, 1000), (0.0, 0), (0.7, 300), (0.01, 3990)):
        density = (r.random(n) < p).astype(np.uint8)
        if density.sum() + offset > 4000:
            density[np.nonzero(density)[0][4000 - offset:]] = 0
        aff, inf = _check(worker, "g1", xy, exps, density=density, offset=offset, bases=bases)
        want_xy, want_inf = util.affine_of_scalar("g1", util.expected_scalar(ks, exps, density, offset))
        assert inf == want_inf and (inf or np.array_equal(aff, want_xy))
    # query size mismatch is an assert in the reference (multiexp.rs:306)
    with pytest.raises(AssertionError):
        zk.multiexp(worker, (bases, 0), zk.DensityTracker([True] * 10), exps)


def test_error_semantics(worker):
    """UnexpectedIdentity / IoError(UnexpectedEof) exactly where the reference's Source raises them."""
    import zcash_gpu_thesis_b200 as zk

    n = 200
    r = util.rng(900)
    xy, ks = util.random_bases("g1", r, n)
    exps = util.random_fr_repr(r, n)
    # too few bases -> UnexpectedEof, also when the missing base would only be skipped (zero scalar)
    bases = zk.Bases(worker, zk.G1, xy[:150])
    with pytest.raises(zk.IoError):
        zk.multiexp(worker, (bases, 0), zk.FullDensity(), exps)
    z = exps.copy()
    z[150:] = 0
    with pytest.raises(zk.IoError):
        zk.multiexp(worker, (bases, 0), zk.FullDensity(), z)
    assert cref.multiexp("g1", xy[:150], z)[0] == cref.UNEXPECTED_EOF
    with pytest.raises(zk.IoError):
        zk.multiexp(worker, (bases, 60), zk.FullDensity(), exps[:100])
    # enough bases once the density map drops exponents
    density = np.zeros(n, dtype=np.uint8)
    density[:150] = 1
    _check(worker, "g1", xy[:150], exps, density=density, bases=bases)
    # identity base: consumed with a non-zero scalar -> UnexpectedIdentity; with a zero scalar or density 0 -> fine
    inf = np.zeros(n, dtype=np.uint8)
    inf[17] = 1
    ib = zk.Bases(worker, zk.G1, xy, inf)
    with pytest.raises(zk.UnexpectedIdentity):
        zk.multiexp(worker, (ib, 0), zk.FullDensity(), exps)
    assert cref.multiexp("g1", xy, exps, inf=inf)[0] == cref.UNEXPECTED_IDENTITY
    one = exps.copy()
    one[17] = (1, 0, 0, 0)
    with pytest.raises(zk.UnexpectedIdentity):
        zk.multiexp(worker, (ib, 0), zk.FullDensity(), one)
    z = exps.copy()
    z[17] = 0
    _check(worker, "g1", xy, z, inf=inf, bases=ib)
    d = np.ones(n, dtype=np.uint8)
    d[17] = 0
    # with density[17] = 0 exponent 18 consumes base 17 (the identity) -> error; shift the identity out of reach instead
    with pytest.raises(zk.UnexpectedIdentity):
        zk.multiexp(worker, (ib, 0), zk.DensityTracker(d), exps)
    # first offending exponent decides: EOF position before identity position and vice versa
    short = zk.Bases(worker, zk.G1, xy[:10], inf[:10])
    with pytest.raises(zk.IoError):
        zk.multiexp(worker, (short, 0), zk.FullDensity(), exps)


@pytest.mark.parametrize("n", [0, 1, 33, 500, 1 << 12])
def test_g2_multiexp_vs_oracle(worker, n):
    r = util.rng(1000 + n)
    xy, ks = util.random_bases("g2", r, max(n, 1))
    exps = util.random_fr_repr(r, n)
    if n > 10:
        exps[3] = 0
        exps[4] = (1, 0, 0, 0)
    aff, inf = _check(worker, "g2", xy[:n] if n else xy[:0], exps)
    want_xy, want_inf = util.affine_of_scalar("g2", util.expected_scalar(ks, exps))
    assert inf == want_inf and (inf or np.array_equal(aff, want_xy))


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 64, 128, 131])
def test_sum_points(worker, n):
    """multiexp.rs:942-1200: the final point reduction at the reference's awkward lengths."""
    import zcash_gpu_thesis_b200 as zk

    r = util.rng(1100 + n)
    xy, ks = util.random_bases("g1", r, n)
    one = np.array(int_to_limbs((1 << 384) % 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab, 6), dtype=np.uint64)
    jac = np.concatenate([xy, np.tile(one, (n, 1))], axis=1)
    din = worker.to_device(jac)
    dout = worker.alloc(144)
    st = worker.lib.b200zk_sum_points_dev(worker.ctx, zk.G1, din.ptr, n, dout.ptr)
    assert st == 0
    got = dout.download(np.uint64, 18)
    aff, inf = zk.into_affine(worker, zk.G1, got)
    total = sum(util.rows_to_ints(ks)) % Fr.p
    want_xy, want_inf = util.affine_of_scalar("g1", total)
    assert bool(inf[0]) == want_inf and np.array_equal(aff[0], want_xy)


def test_fixed_base_and_large_identity(worker):
    """2^20 bases generated on the device as [k_i]G, MSM checked through sum s_i k_i (size-independent property)."""
    import zcash_gpu_thesis_b200 as zk

    n = 1 << 20
    r = util.rng(1200)
    k = np.zeros((n, 4), dtype=np.uint64)
    k[:, 0] = r.integers(1, 1 << 64, size=n, dtype=np.uint64)
    dxy, dinf, _ = zk.fixed_base_mul(worker, zk.G1, util.g1_gen_limbs(), k, 64)
    # spot check the generator against the oracle
    head = dxy.download(np.uint64, 12 * 8).reshape(8, 12)
    want, _ = cref.scalar_muls("g1", util.g1_gen_limbs(), k[:8])
    assert np.array_equal(head, want)
    bases = zk.Bases.from_device(worker, zk.G1, dxy, n)
    exps = util.random_fr_repr(r, n)
    got = zk.multiexp(worker, (bases, 0), zk.FullDensity(), exps)
    aff, inf = zk.into_affine(worker, zk.G1, got)
    kk = k[:, 0].astype(object)
    ee = [int(a) | (int(b) << 64) | (int(c) << 128) | (int(d) << 192) for a, b, c, d in exps]
    total = sum(int(x) * y for x, y in zip(kk, ee)) % Fr.p
    want_xy, want_inf = util.affine_of_scalar("g1", total)
    assert bool(inf[0]) == want_inf and np.array_equal(aff[0], want_xy)


@pytest.mark.parametrize("group,n,c", [("g1", 1, 8), ("g1", 33, 9), ("g1", 3000, 0), ("g1", 3000, 13), ("g1", 1 << 14, 0), ("g2", 700, 0), ("g2", 700, 11)])
def test_precomputed_bases_give_the_same_result(worker, group, n, c):
    """b200zk_bases_precompute: the 2^(c w) * P table changes the schedule (one shared bucket set), not the result."""
    import zcash_gpu_thesis_b200 as zk

    code = zk.G1 if group == "g1" else zk.G2
    r = util.rng(1300 + n + c)
    xy, ks = util.random_bases(group, r, n + 50)
    exps = util.random_fr_repr(r, n)
    exps[0] = int_to_limbs(Fr.p - 1, 4)
    if n > 10:
        exps[3] = 0
        exps[4] = (1, 0, 0, 0)
    bases = zk.Bases(worker, code, xy).precompute(c)
    _check(worker, group, xy, exps, bases=bases)
    density = (r.random(n) < 0.6).astype(np.uint8)
    aff, inf = _check(worker, group, xy, exps, density=density, offset=17, bases=bases)
    want_xy, want_inf = util.affine_of_scalar(group, util.expected_scalar(ks, exps, density, 17))
    assert inf == want_inf and (inf or np.array_equal(aff, want_xy))
    # witness-like scalars: a large share of 0 / 1 / small values (splits the oversized buckets)
    small = exps.copy()
    mask = r.random(n) < 0.7
    small[mask] = 0
    small[mask, 0] = r.integers(0, 3, size=int(mask.sum()), dtype=np.uint64)
    _check(worker, group, xy, small, bases=bases)
    # too few bases still reports UnexpectedEof
    with pytest.raises(zk.IoError):
        zk.multiexp(worker, (bases, 60), zk.FullDensity(), np.concatenate([exps, exps])[: n + 10])


def test_witness_like_scalars_split_buckets(worker):
    """Half of the exponents equal to one (a Sapling witness is full of boolean wires): the oversized bucket is split into
    tasks and folded by a warp; the result must not change."""
    n = 20000
    r = util.rng(1400)
    xy, ks = util.random_bases("g1", r, n)
    exps = util.random_fr_repr(r, n)
    ones = r.random(n) < 0.5
    exps[ones] = (1, 0, 0, 0)
    exps[r.random(n) < 0.2] = 0
    aff, inf = _check(worker, "g1", xy, exps)
    want_xy, want_inf = util.affine_of_scalar("g1", util.expected_scalar(ks, exps))
    assert inf == want_inf and np.array_equal(aff, want_xy)


def test_multiexp_futures_in_flight(worker):
    """multiexp() returns a future and the prover keeps several in flight (prover.rs:289-318, 339-354): results equal the
    synchronous call, errors surface at wait()."""
    import zcash_gpu_thesis_b200 as zk

    n = 3000
    r = util.rng(1500)
    xy, ks = util.random_bases("g1", r, n)
    bases = zk.Bases(worker, zk.G1, xy)
    exps = [util.random_fr_repr(r, n) for _ in range(4)]
    density = zk.DensityTracker((r.random(n) < 0.5))
    futs = [zk.multiexp_async(worker, (bases, 0), zk.FullDensity() if i % 2 == 0 else density, e) for i, e in enumerate(exps)]
    for i, (f, e) in enumerate(zip(futs, exps)):
        got = f.wait()
        want = zk.multiexp(worker, (bases, 0), zk.FullDensity() if i % 2 == 0 else density, e)
        assert np.array_equal(zk.into_affine(worker, zk.G1, got)[0], zk.into_affine(worker, zk.G1, want)[0])
    short = zk.Bases(worker, zk.G1, xy[:100])
    f_ok = zk.multiexp_async(worker, (bases, 0), zk.FullDensity(), exps[0])
    f_bad = zk.multiexp_async(worker, (short, 0), zk.FullDensity(), exps[1])
    f_ok.wait()
    with pytest.raises(zk.IoError):
        f_bad.wait()
    # a fifth job while four are pending is refused, not queued silently
    pend = [zk.multiexp_async(worker, (bases, 0), zk.FullDensity(), exps[0]) for _ in range(4)]
    with pytest.raises(ValueError):
        zk.multiexp_async(worker, (bases, 0), zk.FullDensity(), exps[0])
    for f in pend:
        f.wait()


def test_duplicates_and_cancelling_points_with_table(worker):
    """Many duplicate bases with equal scalars (doubling branch), opposite points (identity) and oversized buckets (a third of the
    exponents are 1) through the precomputed-table schedule, whose small-multiexp form cuts every chain into tasks and folds the
    partial sums with the lane-pair kernels: == the oracle, with and without the table."""
    import zcash_gpu_thesis_b200 as zk

    n = 6000
    r = util.rng(1600)
    xy, ks = util.random_bases("g1", r, n)
    xy[1::7] = xy[0]  # many duplicates: doubling / cancelling pairs inside one bucket
    neg = xy[2].copy()
    neg[6:] = cref.field_vec("fq", "negate", neg[6:].reshape(1, 6)).reshape(-1)
    xy[3::11] = neg   # and their opposites
    exps = util.random_fr_repr(r, n)
    exps[1::7] = exps[0]
    exps[3::11] = exps[2]
    exps[r.random(n) < 0.3] = (1, 0, 0, 0)
    st, ref = cref.multiexp("g1", xy, exps)
    assert st == 0
    want = cref.into_affine("g1", ref)[0]
    bases = zk.Bases(worker, zk.G1, xy)
    for rounds in range(2):
        got = zk.multiexp(worker, (bases, 0), zk.FullDensity(), exps)
        assert np.array_equal(zk.into_affine(worker, zk.G1, got)[0][0], want)
        bases.precompute(0)


@pytest.mark.parametrize("group,n,K,pre", [("g1", 0, 3, None), ("g1", 1, 2, None), ("g1", 777, 5, None), ("g1", 3000, 4, 0), ("g1", 1 << 13, 3, 12),
                                           ("g2", 500, 3, None), ("g2", 900, 4, 0)])
def test_batched_multiexps_equal_separate_ones(worker, group, n, K, pre):
    """b200zk_multiexp_batch_dev: K multiexps over the same bases as bucket sets of one pipeline == K separate multiexps
    (different exponents, different density maps, witness-like and uniform scalars mixed)."""
    import zcash_gpu_thesis_b200 as zk

    code = zk.G1 if group == "g1" else zk.G2
    r = util.rng(4100 + n + K)
    xy, _ = util.random_bases(group, r, n + 20)
    bases = zk.Bases(worker, code, xy)
    if pre is not None:
        bases.precompute(pre)
    exps = np.stack([util.random_fr_repr(r, n) for _ in range(K)]) if n else np.zeros((K, 0, 4), np.uint64)
    if n > 10:
        exps[0, 0] = int_to_limbs(Fr.p - 1, 4)
        exps[1, 5] = 0
        mask = r.random(n) < 0.6  # proof 1 of the batch has a witness-like vector
        exps[1, mask] = 0
        exps[1, mask, 0] = r.integers(0, 2, size=int(mask.sum()), dtype=np.uint64)
    for dens in (None, [(r.random(n) < 0.3 + 0.1 * k).astype(np.uint8) for k in range(K)]):
        got = zk.multiexp_batch(worker, (bases, 7), dens, exps)
        for k in range(K):
            st, want = cref.multiexp(group, xy, exps[k], density=None if dens is None else dens[k], base_offset=7)
            assert st == 0
            got_aff, got_inf = zk.into_affine(worker, code, got[k])
            want_aff, want_inf = cref.into_affine(group, want)
            assert bool(got_inf[0]) == want_inf and np.array_equal(got_aff[0], want_aff), (k, dens is None)


def test_batched_multiexp_reports_the_failing_member(worker):
    import zcash_gpu_thesis_b200 as zk

    r = util.rng(4200)
    xy, _ = util.random_bases("g1", r, 100)
    inf = np.zeros(100, np.uint8)
    inf[40] = 1
    bases = zk.Bases(worker, zk.G1, xy, inf)
    exps = np.stack([util.random_fr_repr(r, 60) for _ in range(3)])
    exps[:, 40] = 0  # nobody consumes the identity base: fine
    assert zk.multiexp_batch(worker, bases, None, exps).shape == (3, 18)
    exps[2, 40] = (5, 0, 0, 0)
    with pytest.raises(zk.UnexpectedIdentity, match="multiexp 2"):
        zk.multiexp_batch(worker, bases, None, exps)
    exps[2, 40] = 0
    with pytest.raises(zk.IoError, match="multiexp 0"):
        zk.multiexp_batch(worker, (bases, 50), None, exps)  # 60 exponents, 50 bases left


@pytest.mark.parametrize("group,log_n", [("g1", 24), ("g2", 22)])
def test_baseline_full_sizes_through_the_discrete_log_identity(worker, group, log_n):
    """BASELINE.json configs 3 / 4 at their full sizes (G1 2^24, G2 2^22), far beyond what the CPU oracle finishes: bases are
    [k_i]G made on the device, so sum s_i P_i must equal [sum s_i k_i]G; the plain-window schedule, the precomputed-table
    schedule and the sum of two half-range shards (the multi-GPU decomposition) must all give that same point."""
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    import zcash_gpu_thesis_b200 as zk

    code = zk.G1 if group == "g1" else zk.G2
    gen = util.g1_gen_limbs() if group == "g1" else bench.gen_g2_limbs()
    n = 1 << log_n
    r = np.random.default_rng(77 + log_n)
    k = np.zeros((n, 4), dtype=np.uint64)
    k[:, 0] = r.integers(1, 1 << 64, size=n, dtype=np.uint64)
    dxy, dinf, _ = zk.fixed_base_mul(worker, code, gen, k, 64)
    bases = zk.Bases.from_device(worker, code, dxy, n)
    dxy.free(); dinf.free()
    exps = bench.random_scalars(r, n)
    exps[5] = 0
    exps[6] = (1, 0, 0, 0)
    exps[7] = int_to_limbs(Fr.p - 1, 4)
    want_xy, want_inf = util.affine_of_scalar(group, bench.dot_mod_r(k[:, 0], exps))
    assert not want_inf

    def affine(jac):
        aff, inf = zk.into_affine(worker, code, jac)
        assert not inf[0]
        return aff[0]

    plain = zk.multiexp(worker, (bases, 0), zk.FullDensity(), exps)
    assert np.array_equal(affine(plain), want_xy)
    half = n // 2
    lo = zk.multiexp(worker, (bases, 0), zk.FullDensity(), exps[:half])
    hi = zk.multiexp(worker, (bases, half), zk.FullDensity(), exps[half:])
    both = zk.point_op(worker, code, zk._lib.POINT_ADD, lo.reshape(1, -1), hi.reshape(1, -1))
    assert np.array_equal(affine(both[0]), want_xy)
    bases.precompute(0)
    assert np.array_equal(affine(zk.multiexp(worker, (bases, 0), zk.FullDensity(), exps)), want_xy)
    bases.free()


def test_one_context_shared_by_several_threads(worker):
    """`Worker` is Clone in the reference and multiexp may be called from any thread (multicore.rs:19-49): calls on one
    context from concurrent host threads are serialised by the library and all give the right answer."""
    import threading

    import zcash_gpu_thesis_b200 as zk

    r = util.rng(4300)
    n = 3000
    xy, _ = util.random_bases("g1", r, n)
    bases = zk.Bases(worker, zk.G1, xy)
    jobs = [util.random_fr_repr(r, n - 100 * t) for t in range(6)]
    got, errs = [None] * len(jobs), []

    def run(t):
        try:
            for _ in range(3):
                got[t] = zk.multiexp(worker, (bases, t), zk.FullDensity(), jobs[t])
                d = zk.EvaluationDomain.from_coeffs(worker, jobs[t][:256] >> np.uint64(2))
                d.fft(worker)
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ths = [threading.Thread(target=run, args=(t,)) for t in range(len(jobs))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs
    for t, exps in enumerate(jobs):
        st, want = cref.multiexp("g1", xy, exps, base_offset=t)
        assert st == 0
        got_aff, got_inf = zk.into_affine(worker, zk.G1, got[t])
        want_aff, want_inf = cref.into_affine("g1", want)
        assert bool(got_inf[0]) == want_inf and np.array_equal(got_aff[0], want_aff)


def test_two_devices_in_one_process(worker):
    """A single host process (the reference's FFI caller) driving two GPUs: one context per device, used from two threads.
    Skipped on a one-GPU box."""
    import threading

    import zcash_gpu_thesis_b200 as zk

    if worker.lib.b200zk_device_count() < 2:
        pytest.skip("needs two GPUs")
    r = util.rng(4400)
    n = 5000
    xy, _ = util.random_bases("g1", r, n)
    exps = [util.random_fr_repr(r, n) for _ in range(2)]
    got, errs = [None, None], []

    def run(dev):
        try:
            w = worker if dev == 0 else zk.Worker(dev)
            bases = zk.Bases(w, zk.G1, xy).precompute(0)
            jac = zk.multiexp(w, (bases, 0), zk.FullDensity(), exps[dev])
            got[dev] = zk.into_affine(w, zk.G1, jac)
            d = zk.EvaluationDomain.from_coeffs(w, exps[dev][:1024] >> np.uint64(2))
            d.fft(w); d.ifft(w)
            bases.free()
            if dev:
                w.close()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ths = [threading.Thread(target=run, args=(d,)) for d in range(2)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs
    for dev in range(2):
        st, want = cref.multiexp("g1", xy, exps[dev])
        want_aff, want_inf = cref.into_affine("g1", want)
        assert st == 0 and bool(got[dev][1][0]) == want_inf and np.array_equal(got[dev][0][0], want_aff)


def test_non_canonical_scalar_is_reported(worker):
    """The ABI takes any 4 x u64 as an exponent.  A canonical FrRepr (< r < 2^255) never carries out of the top window; a value with
    bit 255 set can, when the window width divides 256: reported as a bad argument instead of silently dropping 2^256 P."""
    import zcash_gpu_thesis_b200 as zk

    r = util.rng(4700)
    n = 300
    xy, _ = util.random_bases("g1", r, n)
    exps = util.random_fr_repr(r, n)
    bases = zk.Bases(worker, zk.G1, xy)
    bad = exps.copy()
    bad[5] = (0xFFFFFFFFFFFFFFFF,) * 4
    for c in (4, 8, 16):
        worker.set_msm_window(c)
        zk.multiexp(worker, (bases, 0), zk.FullDensity(), exps)
        with pytest.raises(ValueError):
            zk.multiexp(worker, (bases, 0), zk.FullDensity(), bad)
    worker.set_msm_window(13)  # 20 windows of 13 bits = 260 bits: every 256-bit value fits
    zk.multiexp(worker, (bases, 0), zk.FullDensity(), bad)
    worker.set_msm_window(0)


def _devices_for_group(worker, shards):
    """as many distinct GPUs as the box has, repeating device ids when it has fewer than `shards` (two shards on one GPU run the
    same code path: own contexts, streams and records; only the NVLink hop is then a local store)"""
    n = worker.lib.b200zk_device_count()
    return [d % n for d in range(shards)]


@pytest.mark.parametrize("group,shards", [("g1", 2), ("g1", 3), ("g2", 2)])
def test_one_process_multi_gpu_multiexp(worker, group, shards):
    """b200zk_init_multi / b200zk_multi_bases_upload / b200zk_multi_multiexp: one host process, bases sharded by contiguous range,
    exponents cut at the shard boundaries (density-aware), partials summed on the first device == the oracle's unsharded multiexp.
    Offsets, density maps, precomputed tables, futures in flight, and the Source's error semantics across shards."""
    import zcash_gpu_thesis_b200 as zk

    code = zk.G1 if group == "g1" else zk.G2
    r = util.rng(4500 + shards)
    n = 3001 if group == "g1" else 1201
    xy, ks = util.random_bases(group, r, n)
    mw = zk.MultiWorker(_devices_for_group(worker, shards))
    try:
        assert len(mw) == shards and all(mw.peer_access(i) in (True, False) for i in range(shards))
        bases = zk.ShardedBases(mw, code, xy)
        assert len(bases) == n

        def check(exps, density=None, offset=0):
            dm = zk.FullDensity() if density is None else zk.DensityTracker(density)
            got = zk.multiexp(mw, (bases, offset), dm, exps)
            st, want = cref.multiexp(group, xy, exps, density=density, base_offset=offset)
            assert st == 0
            ga, gi = zk.into_affine(worker, code, got)
            wa, wi = cref.into_affine(group, want)
            assert bool(gi[0]) == wi and np.array_equal(ga[0], wa)

        exps = util.random_fr_repr(r, n)
        for rounds in range(2):  # plain windows, then precomputed tables per shard
            check(exps)
            check(exps[: n - 700], offset=700)
            check(exps[:5])                      # everything in the first shard
            check(exps[:9], offset=n - 9)        # everything in the last shard
            d = (r.random(n) < 0.55).astype(np.uint8)
            check(exps, density=d)
            d2 = (r.random(n) < 0.3).astype(np.uint8)
            check(exps, density=d2, offset=n // 2)
            check(exps[:0])
            bases.precompute(0)
        # futures in flight (prover.rs:289-318 keeps several)
        jobs = [util.random_fr_repr(r, n - 10 * t) for t in range(3)]
        futs = [zk.multiexp_async(mw, (bases, 3 * t), zk.FullDensity(), jobs[t]) for t in range(3)]
        for t, f in enumerate(futs):
            st, want = cref.multiexp(group, xy, jobs[t], base_offset=3 * t)
            ga, gi = zk.into_affine(worker, code, f.wait())
            wa, wi = cref.into_affine(group, want)
            assert st == 0 and bool(gi[0]) == wi and np.array_equal(ga[0], wa)
        # UnexpectedEof: only the last shard can run out of bases, whichever shard the exponents start in
        with pytest.raises(zk.IoError):
            zk.multiexp(mw, (bases, 1), zk.FullDensity(), exps)
        with pytest.raises(zk.IoError):
            zk.multiexp(mw, (bases, n + 5), zk.FullDensity(), exps[:3])
        z = exps.copy()
        z[n - 1] = 0
        with pytest.raises(zk.IoError):
            zk.multiexp(mw, (bases, 1), zk.FullDensity(), z)  # the missing base would only have been skipped: still EOF (multiexp.rs:60-62)
        bases.free()
        # UnexpectedIdentity in a middle shard fails the whole multiexp; a zero scalar on it does not
        inf = np.zeros(n, dtype=np.uint8)
        bad = n // shards + 7
        inf[bad] = 1
        ib = zk.ShardedBases(mw, code, xy, inf)
        with pytest.raises(zk.UnexpectedIdentity):
            zk.multiexp(mw, (ib, 0), zk.FullDensity(), exps)
        z = exps.copy()
        z[bad] = 0
        got = zk.multiexp(mw, (ib, 0), zk.FullDensity(), z)
        st, want = cref.multiexp(group, xy, z, inf=inf)
        assert st == 0 and np.array_equal(zk.into_affine(worker, code, got)[0][0], cref.into_affine(group, want)[0])
        # which error wins: the first offending exponent in iteration order (identity in shard 1 before the EOF of the last shard)
        with pytest.raises(zk.UnexpectedIdentity):
            zk.multiexp(mw, (ib, 1), zk.FullDensity(), exps)
        ib.free()
    finally:
        mw.close()


@pytest.mark.parametrize("group", ["g1", "g2"])
def test_nccl_sharded_multiexp_two_gpus_in_process(worker, group):
    """The process-per-GPU path without torchrun: two contexts on two GPUs, one host thread each, a communicator from
    b200zk_nccl_unique_id / b200zk_comm_init, then b200zk_multiexp_sharded_async on the two base-range shards == the oracle on
    the unsharded input; an identity in ONE shard fails the multiexp on BOTH ranks.  Needs two GPUs (NCCL refuses two ranks on one)."""
    import ctypes
    import threading

    import zcash_gpu_thesis_b200 as zk
    from zcash_gpu_thesis_b200.sharding import shard_density, shard_range

    if worker.lib.b200zk_device_count() < 2:
        pytest.skip("needs two GPUs")
    code = zk.G1 if group == "g1" else zk.G2
    r = util.rng(4600)
    n = 4000 if group == "g1" else 1500
    xy, _ = util.random_bases(group, r, n)
    exps = util.random_fr_repr(r, n)
    density = (r.random(n) < 0.7).astype(np.uint8)
    inf = np.zeros(n, dtype=np.uint8)
    inf[n - 5] = 1  # in rank 1's shard
    uid = (ctypes.c_uint8 * 128)()
    assert worker.lib.b200zk_nccl_unique_id(uid) == 0
    results, errs = [None, None], []

    def run(rank):
        try:
            w = zk.Worker(rank)
            st = w.lib.b200zk_comm_init(w.ctx, uid, rank, 2)
            assert st == 0, w.last_error()
            lo, hi = shard_range(n, rank, 2)
            out = {}
            for name, dens, use_inf in (("full", None, False), ("dense", density, False), ("identity", None, True)):
                d_shard, off = shard_density(dens, lo, hi)
                # the shard owns bases [lo, hi): its cursor starts at off - lo inside its own slice when the density map is full
                if dens is None:
                    b = zk.Bases(w, code, xy[lo:hi], inf[lo:hi] if use_inf else None)
                    local_off = 0
                else:
                    b = zk.Bases(w, code, xy)  # density-mapped: every rank keeps the whole vector and starts at its cursor
                    local_off = off
                job = ctypes.c_void_p()
                e = np.ascontiguousarray(exps[lo:hi])
                st = w.lib.b200zk_multiexp_sharded_async(w.ctx, b.handle, local_off, e.ctypes.data_as(ctypes.c_void_p), hi - lo,
                                                         None if d_shard is None else np.ascontiguousarray(d_shard).ctypes.data_as(ctypes.c_void_p), ctypes.byref(job))
                assert st == 0, w.last_error()
                res = np.zeros(18 if group == "g1" else 36, dtype=np.uint64)
                out[name] = (w.lib.b200zk_job_wait(job, res.ctypes.data_as(ctypes.c_void_p)), res)
                b.free()
            results[rank] = out
            w.close()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ths = [threading.Thread(target=run, args=(k,)) for k in range(2)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs
    for name, dens in (("full", None), ("dense", density)):
        st, want = cref.multiexp(group, xy, exps, density=dens)
        wa, wi = cref.into_affine(group, want)
        for rank in range(2):
            code_st, res = results[rank][name]
            ga, gi = zk.into_affine(worker, code, res)
            assert code_st == 0 and st == 0 and bool(gi[0]) == wi and np.array_equal(ga[0], wa), (name, rank)
    assert [results[k]["identity"][0] for k in range(2)] == [zk._lib.ERR_UNEXPECTED_IDENTITY] * 2
