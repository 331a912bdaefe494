"""EvaluationDomain on the GPU vs the CPU oracle, limb-exact (bellman/src/domain.rs:379-494: polynomial_arith,
fft_composition, parallel_fft_consistency re-targeted at the CUDA path)."""
import numpy as np
import pytest

from oracle import cref
from oracle import domain as odomain
from oracle.fields import Fr
from tests import util

pytestmark = pytest.mark.gpu

KINDS = [(0, "fft"), (1, "ifft"), (2, "coset_fft"), (3, "icoset_fft")]


@pytest.mark.parametrize("log_m", list(range(0, 21)))
def test_ntt_matches_serial_fft(worker, log_m):
    """Every transform kind at 2^0..2^20: device output == oracle serial_fft output, every limb."""
    import zcash_gpu_thesis_b200 as zk

    m = 1 << log_m
    coeffs = util.random_fr_mont(util.rng(100 + log_m), m)
    for kind, name in KINDS:
        got = zk.ntt_host(worker, coeffs, kind)
        want = cref.fft(coeffs, kind)
        assert np.array_equal(got, want), f"{name} at 2^{log_m}"


@pytest.mark.parametrize("log_m", [1, 5, 9, 13, 19, 22])
def test_fft_composition(worker, log_m):
    """domain.rs:426-461: ifft(fft(v)) == v, fft(ifft(v)) == v, icoset(coset(v)) == v, coset(icoset(v)) == v."""
    import zcash_gpu_thesis_b200 as zk

    m = 1 << log_m
    v = util.random_fr_mont(util.rng(200 + log_m), m)
    for first, second in (("ifft", "fft"), ("fft", "ifft"), ("icoset_fft", "coset_fft"), ("coset_fft", "icoset_fft")):
        d = zk.EvaluationDomain.from_coeffs(worker, v)
        getattr(d, first)(worker)
        getattr(d, second)(worker)
        assert np.array_equal(d.into_coeffs(), v), f"{second}({first}(v)) != v at 2^{log_m}"


def test_ntt_2_24_round_trip_and_spot_check(worker):
    """The headline size: round trip is the identity, and 32 outputs equal the direct O(n) evaluation
    X[k] = sum_j x[j] w^(jk) computed on a structured input (x[j] = a * b^j has a closed form)."""
    import zcash_gpu_thesis_b200 as zk

    log_m = 24
    m = 1 << log_m
    # x[j] = a * b^j  =>  X[k] = a * ((b w^k)^m - 1) / (b w^k - 1)
    a, b = 0x1234567890ABCDEF % Fr.p, 0x0FEDCBA987654321 % Fr.p
    ones = np.tile(zk.bellman.fr_to_mont_limbs(a), (m, 1))
    d = zk.EvaluationDomain.from_coeffs(worker, ones)
    d.distribute_powers(worker, b)
    x = d.into_coeffs()
    d.fft(worker)
    X = d.into_coeffs()
    omega = Fr.root_of_unity
    for _ in range(log_m, 32):
        omega = Fr.sqr(omega)
    r = util.rng(24)
    for k in [0, 1, m - 1, m // 2] + [int(v) for v in r.integers(0, m, size=28)]:
        q = b * pow(omega, k, Fr.p) % Fr.p
        want = a * (pow(q, m, Fr.p) - 1) * pow(q - 1, -1, Fr.p) % Fr.p
        assert zk.bellman.fr_from_mont_limbs(X[k]) == want, f"X[{k}]"
    d.ifft(worker)
    assert np.array_equal(d.into_coeffs(), x)


@pytest.mark.parametrize("log_m", [21, 22, 23, 24])
def test_ntt_large_all_kinds_full_compare(worker, log_m):
    """2^21..2^24 (the headline size): all four transform kinds, every limb of every output against the oracle's best_fft
    (domain.rs:261-270: parallel_fft on the host threads; == serial_fft by tests/test_oracle_golden.py) -- this covers the
    coset tables (g^i, g^-i / m) over the full index range."""
    import zcash_gpu_thesis_b200 as zk

    m = 1 << log_m
    coeffs = util.random_fr_mont(util.rng(500 + log_m), m)
    for kind, name in KINDS:
        got = zk.ntt_host(worker, coeffs, kind)
        want = cref.fft(coeffs, kind, serial=False)
        assert np.array_equal(got, want), f"{name} at 2^{log_m}"
        del got, want


@pytest.mark.parametrize("log_m", list(range(12, 20)) + [25])
def test_large_transform_kernels_on_every_tile_shape(log_m, monkeypatch):
    """The radix-4 large-transform kernels (k_ntt_pass4: values in [0, 2r) between stages, in-place middle passes) on every pass
    shape they are built for -- B200ZK_NTT_LARGE_FROM=12 sends 2^12..2^19 through them (passes of 6 / 7 / 8 / 9 stages, the
    one-stage first round of an odd pass), 2^25 is the four-pass split 7 + 6 + 6 + 6 -- all kinds, every limb against the oracle."""
    import zcash_gpu_thesis_b200 as zk

    monkeypatch.setenv("B200ZK_NTT_LARGE_FROM", "12")
    w = zk.Worker(0)  # the threshold is read when the context is created
    try:
        m = 1 << log_m
        coeffs = util.random_fr_mont(util.rng(700 + log_m), m)
        for kind, name in KINDS:
            got = zk.ntt_host(w, coeffs, kind)
            want = cref.fft(coeffs, kind, serial=log_m < 20)
            assert np.array_equal(got, want), f"{name} at 2^{log_m}"
            del got, want
    finally:
        w.close()


@pytest.mark.parametrize("log_m", [12, 16, 20])
def test_large_transform_kernels_extreme_values(log_m, monkeypatch):
    """The [0, 2r) value range of the large-transform kernels at its edges: vectors made of r - 1, 0 and 1 (as Montgomery residues,
    i.e. the limbs themselves are r - 1 / 0 / 1) in long runs and alternating patterns, so that sums reach 2r - 2 + (2r - 1) and
    differences -(2r - 1) at every stage; all kinds, every limb against the oracle."""
    import zcash_gpu_thesis_b200 as zk

    monkeypatch.setenv("B200ZK_NTT_LARGE_FROM", "12")
    w = zk.Worker(0)
    try:
        m = 1 << log_m
        top = np.array([(Fr.p - 1 >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
        one = np.array([1, 0, 0, 0], dtype=np.uint64)
        r = util.rng(900 + log_m)
        patterns = {
            "all r-1": np.tile(top, (m, 1)),
            "alternating r-1 / 0": np.where((np.arange(m) % 2 == 0)[:, None], top[None, :], np.zeros((1, 4), dtype=np.uint64)),
            "random of {r-1, 0, 1}": np.stack([top, np.zeros(4, dtype=np.uint64), one])[r.integers(0, 3, size=m)],
        }
        for label, coeffs in patterns.items():
            coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64)
            for kind, name in KINDS:
                got = zk.ntt_host(w, coeffs, kind)
                want = cref.fft(coeffs, kind, serial=log_m < 20)
                assert np.array_equal(got, want), f"{name} of '{label}' at 2^{log_m}"
    finally:
        w.close()


def test_polynomial_arith(worker):
    """domain.rs:379-423: fft * fft -> ifft equals the schoolbook product (a sample of degree pairs < 70)."""
    import zcash_gpu_thesis_b200 as zk

    r = util.rng(300)
    for (da, db) in [(1, 1), (2, 3), (5, 64), (33, 31), (69, 69), (17, 50)]:
        pa = util.rows_to_ints(util.random_fr_repr(r, da))
        pb = util.rows_to_ints(util.random_fr_repr(r, db))
        naive = [0] * (da + db)
        for i, x in enumerate(pa):
            for j, y in enumerate(pb):
                naive[i + j] = (naive[i + j] + x * y) % Fr.p
        to_m = lambda v: np.array([Fr.to_mont_limbs(x) for x in v], dtype=np.uint64)
        A = zk.EvaluationDomain.from_coeffs(worker, to_m(pa + [0] * db))
        B = zk.EvaluationDomain.from_coeffs(worker, to_m(pb + [0] * da))
        A.fft(worker)
        B.fft(worker)
        A.mul_assign(worker, B)
        A.ifft(worker)
        got = [Fr.from_mont_limbs(row) for row in A.into_coeffs()]
        assert got[: da + db] == naive
        assert all(v == 0 for v in got[da + db:])


def test_elementwise_domain_ops(worker):
    """distribute_powers / divide_by_z_on_coset / sub_assign (domain.rs:105-189) vs the Python oracle."""
    import zcash_gpu_thesis_b200 as zk

    m = 1 << 9
    r = util.rng(301)
    a = util.random_fr_mont(r, m)
    b = util.random_fr_mont(r, m)
    ai = [Fr.from_mont_limbs(x) for x in a]
    bi = [Fr.from_mont_limbs(x) for x in b]
    od = odomain.EvaluationDomain(Fr, ai)
    ob = odomain.EvaluationDomain(Fr, bi)
    od.distribute_powers(12345)
    od.sub_assign(ob)
    od.divide_by_z_on_coset()
    d = zk.EvaluationDomain.from_coeffs(worker, a)
    e = zk.EvaluationDomain.from_coeffs(worker, b)
    d.distribute_powers(worker, 12345)
    d.sub_assign(worker, e)
    d.divide_by_z_on_coset(worker)
    assert [Fr.from_mont_limbs(x) for x in d.into_coeffs()] == od.coeffs
    assert d.z(7) == od.z(7)


def test_degree_too_large(worker):
    """domain.rs:59-61: exp >= Fr::S -> PolynomialDegreeTooLarge (no allocation of 2^32 elements needed)."""
    import zcash_gpu_thesis_b200 as zk

    st = worker.lib.b200zk_ntt_dev(worker.ctx, None, 32, 0)
    assert st == zk._lib.ERR_DEGREE_TOO_LARGE


@pytest.mark.parametrize("n", [1, 5, 8, 1000, 98785 // 16])
def test_h_poly_matches_oracle(worker, n):
    """prover.rs:256-287: the fused device H pipeline vs the oracle's EvaluationDomain sequence."""
    import zcash_gpu_thesis_b200 as zk

    r = util.rng(400 + n)
    a, b, c = (util.random_fr_mont(r, n) for _ in range(3))
    m = 1
    while m < n:
        m *= 2
    pad = lambda v: np.concatenate([v, np.zeros((m - n, 4), dtype=np.uint64)])
    want = cref.h_poly(pad(a), pad(b), pad(c))
    got = zk.h_poly(worker, a, b, c)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n", [5000, 98785])
def test_h_poly_batched_on_the_large_transform_kernels(n, monkeypatch):
    """The H block runs its a / b / c transforms as ONE batch of launches (blockIdx.y = the vector); B200ZK_NTT_LARGE_FROM=12
    sends them through the radix-4 kernels, whose batch strides differ between the caller's vectors and the scratch."""
    import zcash_gpu_thesis_b200 as zk

    monkeypatch.setenv("B200ZK_NTT_LARGE_FROM", "12")
    w = zk.Worker(0)
    try:
        r = util.rng(450 + n)
        a, b, c = (util.random_fr_mont(r, n) for _ in range(3))
        m = 1
        while m < n:
            m *= 2
        pad = lambda v: np.concatenate([v, np.zeros((m - n, 4), dtype=np.uint64)])
        assert np.array_equal(zk.h_poly(w, a, b, c), cref.h_poly(pad(a), pad(b), pad(c)))
    finally:
        w.close()


@pytest.mark.parametrize("n,shards", [(1, 3), (1000, 3), (98785, 3), (98785, 2), (5000, 1)])
def test_h_poly_over_a_group_of_gpus(worker, n, shards):
    """b200zk_multi_h_poly: a, b, c transformed on separate GPUs of a one-process group (prover.rs:257-266 runs them as three
    scoped tasks), b and c copied peer to peer to the first device for the pointwise combine == the oracle's H coefficients.
    With fewer GPUs than shards the device ids repeat (separate contexts and streams on one GPU: the same code path)."""
    import zcash_gpu_thesis_b200 as zk

    r = util.rng(600 + n + shards)
    a, b, c = (util.random_fr_mont(r, n) for _ in range(3))
    m = 1
    while m < n:
        m *= 2
    pad = lambda v: np.concatenate([v, np.zeros((m - n, 4), dtype=np.uint64)])
    want = cref.h_poly(pad(a), pad(b), pad(c))
    ndev = worker.lib.b200zk_device_count()
    mw = zk.MultiWorker([d % ndev for d in range(shards)])
    try:
        assert np.array_equal(zk.h_poly(mw, a, b, c), want)
        assert np.array_equal(zk.h_poly(mw, a, b, c), want)  # workspaces reused
    finally:
        mw.close()
