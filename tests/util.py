"""Shared helpers for the parity tests: seeded inputs in the reference's memory layout."""
import numpy as np

from oracle import cref
from oracle.curve import G1, G2
from oracle.fields import Fq, Fr, int_to_limbs, limbs_to_int

SEED = 0x5DBE62598D313D76  # the reference's XorShiftRng seed words (pairing tests), used as our base seed


def rng(tag=0):
    return np.random.default_rng([SEED & 0xFFFFFFFF, SEED >> 32, tag])


def random_field_canonical(r, modulus, n, nlimbs):
    """n uniform elements of [0, modulus) as (n, nlimbs) uint64 canonical limbs (rejection sampling like fr.rs:255-268)."""
    bits = modulus.bit_length()
    out = np.zeros((n, nlimbs), dtype=np.uint64)
    todo = np.arange(n)
    mod_limbs = np.array(int_to_limbs(modulus, nlimbs), dtype=np.uint64)
    while todo.size:
        cand = r.integers(0, 1 << 64, size=(todo.size, nlimbs), dtype=np.uint64)
        top_bits = bits - 64 * (nlimbs - 1)
        cand[:, -1] &= np.uint64((1 << top_bits) - 1)
        # lexicographic compare cand < modulus, most significant limb first
        lt = np.zeros(todo.size, dtype=bool)
        eq = np.ones(todo.size, dtype=bool)
        for l in range(nlimbs - 1, -1, -1):
            lt |= eq & (cand[:, l] < mod_limbs[l])
            eq &= cand[:, l] == mod_limbs[l]
        out[todo[lt]] = cand[lt]
        todo = todo[~lt]
    return out


def random_fr_repr(r, n):
    return random_field_canonical(r, Fr.p, n, 4)


def random_fr_mont(r, n):
    """random Fr elements in Montgomery form (a uniform canonical limb pattern < r is a uniform element)."""
    return random_field_canonical(r, Fr.p, n, 4)


def random_fq_mont(r, n):
    return random_field_canonical(r, Fq.p, n, 6)


def rows_to_ints(a):
    return [limbs_to_int(row) for row in a]


def g1_gen_limbs():
    return np.array(G1.affine_to_limbs(G1.gen), dtype=np.uint64)


def g2_gen_limbs():
    return np.array(G2.affine_to_limbs(G2.gen), dtype=np.uint64)


def random_bases(group, r, n, bits64=True):
    """n bases [k_i] * generator (k_i uniform 64-bit or full Fr), affine Montgomery limbs, via the CPU oracle.
    Returns (xy, k as (n,4) canonical limbs)."""
    if bits64:
        k = np.zeros((n, 4), dtype=np.uint64)
        k[:, 0] = r.integers(1, 1 << 64, size=n, dtype=np.uint64)
    else:
        k = random_fr_repr(r, n)
    gen = g1_gen_limbs() if group == "g1" else g2_gen_limbs()
    xy, inf = cref.scalar_muls(group, gen, k)
    assert not inf.any()
    return xy, k


def expected_scalar(ks, exps, density=None, offset=0):
    """sum_i exps[i] * k[offset + rank(i)] mod r for bases [k]G: the discrete log of the MSM result."""
    kk = rows_to_ints(ks)
    ee = rows_to_ints(exps)
    acc, j = 0, offset
    for i, e in enumerate(ee):
        if density is not None and not density[i]:
            continue
        acc += e * kk[j]
        j += 1
    return acc % Fr.p


def affine_of_scalar(group, s):
    """[s] * generator as affine limbs via the oracle: (xy, is_infinity)."""
    gen = g1_gen_limbs() if group == "g1" else g2_gen_limbs()
    xy, inf = cref.scalar_muls(group, gen, np.array([int_to_limbs(s % Fr.p, 4)], dtype=np.uint64))
    return xy[0], bool(inf[0])
