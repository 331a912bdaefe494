"""bellman::multiexp restated (bellman/src/multiexp.rs:19-335).

ORACLE -- test infrastructure only (see oracle/__init__.py).

`multiexp` follows the reference window by window: c = 3 if n < 32 else
ceil(ln n) (multiexp.rs:296-300); one region per window (multiexp.rs:140-233)
with the zero / one special cases, bucket fill by mixed addition, the
running-sum reduction and the c doublings per join.  The group G is any object
with zero/add/add_mixed/double/affine_is_zero (oracle.curve.Curve or the dummy
group of oracle.dummy_engine), scalars are canonical ints (FrRepr).
"""
from __future__ import annotations

import math


class SynthesisError(Exception):
    """bellman/src/lib.rs:171-188"""


class UnexpectedIdentity(SynthesisError):
    pass


class UnexpectedEof(SynthesisError):
    """IoError(UnexpectedEof), multiexp.rs:44-46, 60-62"""


class PolynomialDegreeTooLarge(SynthesisError):
    pass


class Source:
    """impl Source for (Arc<Vec<G>>, usize), multiexp.rs:42-68"""

    def __init__(self, G, bases, idx=0):
        self.G = G
        self.bases = bases
        self.idx = idx

    def add_assign_mixed(self, to):
        if len(self.bases) <= self.idx:
            raise UnexpectedEof("expected more bases from source")
        if self.G.affine_is_zero(self.bases[self.idx]):
            raise UnexpectedIdentity()
        r = self.G.add_mixed(to, self.bases[self.idx])
        self.idx += 1
        return r

    def skip(self, amt):
        if len(self.bases) <= self.idx:
            raise UnexpectedEof("expected more bases from source")
        self.idx += amt


def window_size(n: int) -> int:
    """multiexp.rs:296-300"""
    if n < 32:
        return 3
    return int(math.ceil(math.log(float(n))))


def multiexp_region(G, bases, base_offset, density, exponents, skip, c, handle_trivial):
    """The closure body of multiexp_inner, multiexp.rs:160-209."""
    acc = G.zero()
    src = Source(G, bases, base_offset)
    buckets = [G.zero() for _ in range((1 << c) - 1)]
    mask = (1 << c) - 1
    for i, exp in enumerate(exponents):
        d = True if density is None else bool(density[i])
        if not d:
            continue
        if exp == 0:
            src.skip(1)
        elif exp == 1:
            if handle_trivial:
                acc = src.add_assign_mixed(acc)
            else:
                src.skip(1)
        else:
            digit = (exp >> skip) & mask
            if digit != 0:
                buckets[digit - 1] = src.add_assign_mixed(buckets[digit - 1])
            else:
                src.skip(1)
    running_sum = G.zero()
    for b in reversed(buckets):
        running_sum = G.add(running_sum, b)
        acc = G.add(acc, running_sum)
    return acc


def multiexp_inner(G, bases, base_offset, density, exponents, skip, c, handle_trivial, num_bits=255):
    """multiexp.rs:140-233 (recursion unrolled; errors of the lowest region win, like Join polling `this` first)."""
    this = multiexp_region(G, bases, base_offset, density, exponents, skip, c, handle_trivial)
    skip += c
    if skip >= num_bits:
        return this
    higher = multiexp_inner(G, bases, base_offset, density, exponents, skip, c, False, num_bits)
    for _ in range(c):
        higher = G.double(higher)
    return G.add(higher, this)


def multiexp(G, bases, exponents, density=None, base_offset=0, num_bits=255):
    """bellman::multiexp::multiexp, multiexp.rs:285-335.

    bases: list of affine points; (bases, base_offset) is the SourceBuilder.
    density: None (FullDensity) or a list of bools with len == len(exponents).
    """
    c = window_size(len(exponents))
    if density is not None:
        assert len(density) == len(exponents)  # multiexp.rs:306
    return multiexp_inner(G, bases, base_offset, density, exponents, 0, c, True, num_bits)


def naive_multiexp(G, bases, exponents):
    """The check of test_with_bls12, multiexp.rs:342-352."""
    acc = G.zero()
    for b, e in zip(bases, exponents):
        acc = G.add(acc, G.mul(b, e))
    return acc
