"""Randomised parity of the multiexp pipeline against the CPU oracle: sizes, groups, density maps, base offsets, scalar
mixes (uniform / witness-like / tiny / one hot bucket), plain and precomputed schedules, window overrides, batches.
The case generator is shared with tools/fuzz.py, which runs many more cases."""
import numpy as np
import pytest

from oracle import cref
from oracle.fields import Fr, int_to_limbs
from tests import util

pytestmark = pytest.mark.gpu


def random_case(r, max_n=3000):
    group = "g1" if r.random() < 0.7 else "g2"
    n = int(r.choice([0, 1, 2, 3, 5, 17, 31, 32, 33, 64, 100, 257, 1000, int(r.integers(1, max_n))]))
    if group == "g2":
        n = min(n, 700)
    exps = util.random_fr_repr(r, n)
    kind = r.integers(0, 5)
    if n:
        if kind == 1:  # witness-like
            mask = r.random(n) < 0.6
            exps[mask] = 0
            exps[mask, 0] = r.integers(0, 2, size=int(mask.sum()), dtype=np.uint64)
        elif kind == 2:  # tiny scalars
            exps[:, 1:] = 0
            exps[:, 0] = r.integers(0, 1 << 12, size=n, dtype=np.uint64)
        elif kind == 3:  # one hot digit pattern: most exponents equal
            exps[r.random(n) < 0.8] = exps[0]
        elif kind == 4:  # extremes
            exps[r.random(n) < 0.3] = int_to_limbs(Fr.p - 1, 4)
    density = (r.random(n) < r.uniform(0.2, 1.0)).astype(np.uint8) if r.random() < 0.5 else None
    offset = int(r.integers(0, 9))
    pre = int(r.choice([-1, -1, 0, 8, 11, 13]))  # -1: plain windows
    window = int(r.choice([0, 0, 0, 3, 7, 12])) if pre < 0 else 0
    return dict(group=group, n=n, exps=exps, density=density, offset=offset, pre=pre, window=window)


def check_case(worker, r, case):
    import zcash_gpu_thesis_b200 as zk

    group, n = case["group"], case["n"]
    code = zk.G1 if group == "g1" else zk.G2
    xy, _ = util.random_bases(group, r, n + case["offset"] + 3)
    if n > 4 and r.random() < 0.3:
        xy[2] = xy[1]  # a repeated base: the doubling branch of the mixed addition
    bases = zk.Bases(worker, code, xy)
    if case["pre"] >= 0:
        bases.precompute(case["pre"])
    worker.set_msm_window(case["window"])
    try:
        dm = zk.FullDensity() if case["density"] is None else zk.DensityTracker(case["density"])
        got = zk.multiexp(worker, (bases, case["offset"]), dm, case["exps"])
    finally:
        worker.set_msm_window(0)
    if n and r.random() < 0.25:  # the same multiexp as member 1 of a batch of three over these bases
        others = [util.random_fr_repr(r, n) for _ in range(2)]
        dens = None if case["density"] is None else [np.ones(n, np.uint8), case["density"], (r.random(n) < 0.5).astype(np.uint8)]
        batch = zk.multiexp_batch(worker, (bases, case["offset"]), dens, np.stack([others[0], case["exps"], others[1]]))
        a1, i1 = zk.into_affine(worker, code, batch[1])
        a0, i0 = zk.into_affine(worker, code, got)
        assert bool(i1[0]) == bool(i0[0]) and np.array_equal(a1[0], a0[0])
    st, want = cref.multiexp(group, xy, case["exps"], density=case["density"], base_offset=case["offset"])
    assert st == 0
    got_aff, got_inf = zk.into_affine(worker, code, got)
    want_aff, want_inf = cref.into_affine(group, want)
    assert bool(got_inf[0]) == want_inf and np.array_equal(got_aff[0], want_aff), {k: v for k, v in case.items() if k not in ("exps", "density")}
    bases.free()


def test_random_multiexps(worker):
    r = util.rng(9001)
    for _ in range(60):
        check_case(worker, r, random_case(r))
