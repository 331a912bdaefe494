"""Field arithmetic of the reference's `pairing` crate, restated with Python ints.

ORACLE -- test infrastructure only (see oracle/__init__.py).

Follows:
  * pairing/src/bls12_381/fr.rs:4-55   (Fr constants), :276-303 (from_repr/into_repr),
    :341-375 (add/double/sub/negate), :438-500 (mul/square), :520-571 (mont_reduce)
  * pairing/src/bls12_381/fq.rs:5-66   (Fq constants), :813-1123
  * pairing/src/bls12_381/fq2.rs:84-140 (Fq2 mul/square), :21-30 (ordering)
  * pairing/src/lib.rs:306-324 (Field::pow), :645-679 (adc/sbb/mac_with_carry)

Field elements are canonical Python ints in [0, p).  Every reference field
operation fully reduces (fr.rs:513-517, fq.rs:1030), so a canonical int is in
one-to-one correspondence with the reference's Montgomery limbs; `to_mont_limbs`
/ `from_mont_limbs` convert to the exact in-memory representation.
"""
from __future__ import annotations

MASK64 = (1 << 64) - 1


# --------------------------------------------------------------------------- limb primitives
def adc(a: int, b: int, carry: int):
    """pairing/src/lib.rs:662-668"""
    t = a + b + carry
    return t & MASK64, t >> 64


def sbb(a: int, b: int, borrow: int):
    """pairing/src/lib.rs:650-656"""
    t = (1 << 64) + a - b - borrow
    return t & MASK64, (1 if (t >> 64) == 0 else 0)


def mac_with_carry(a: int, b: int, c: int, carry: int):
    """pairing/src/lib.rs:673-679"""
    t = a + b * c + carry
    return t & MASK64, t >> 64


def int_to_limbs(x: int, n: int):
    return [(x >> (64 * i)) & MASK64 for i in range(n)]


def limbs_to_int(limbs):
    x = 0
    for i, l in enumerate(limbs):
        x |= int(l) << (64 * i)
    return x


class PrimeField:
    """A Montgomery prime field as the reference lays it out (n x u64 limbs)."""

    def __init__(self, name, modulus, nlimbs, generator=None, s=None):
        self.name = name
        self.p = modulus
        self.nlimbs = nlimbs
        self.num_bits = modulus.bit_length()
        self.R = (1 << (64 * nlimbs)) % modulus
        self.R2 = (self.R * self.R) % modulus
        self.Rinv = pow(self.R, -1, modulus)
        self.INV = (-pow(modulus, -1, 1 << 64)) % (1 << 64)
        self.generator = generator
        self.S = s
        if generator is not None:
            t = (modulus - 1) >> s
            assert t & 1 and (t << s) == modulus - 1
            self.root_of_unity = pow(generator, t, modulus)  # fr.rs:49-55
        self.zero = 0
        self.one = 1

    # -- canonical-int arithmetic
    def add(self, a, b):
        r = a + b
        return r - self.p if r >= self.p else r

    def sub(self, a, b):
        r = a - b
        return r + self.p if r < 0 else r

    def neg(self, a):
        return self.p - a if a else 0

    def dbl(self, a):
        return self.add(a, a)

    def mul(self, a, b):
        return (a * b) % self.p

    def sqr(self, a):
        return (a * a) % self.p

    def inv(self, a):
        if a == 0:
            return None
        return pow(a, -1, self.p)

    def pow(self, a, e):
        return pow(a, e, self.p)

    def is_zero(self, a):
        return a == 0

    def eq(self, a, b):
        return a == b

    # -- exact memory representation
    def to_mont(self, a):
        return (a * self.R) % self.p

    def from_mont(self, m):
        return (m * self.Rinv) % self.p

    def to_mont_limbs(self, a):
        return int_to_limbs(self.to_mont(a), self.nlimbs)

    def from_mont_limbs(self, limbs):
        m = limbs_to_int(limbs)
        assert m < self.p, "non-canonical Montgomery limbs"
        return self.from_mont(m)

    def repr_limbs(self, a):
        """into_repr(): canonical (non-Montgomery) limbs, fr.rs:290-303"""
        return int_to_limbs(a, self.nlimbs)

    # -- limb-level Montgomery multiply, literal restatement (fq.rs:910-1017 / fr.rs:438-500 + mont_reduce)
    def mont_mul_limbs(self, a_limbs, b_limbs):
        n = self.nlimbs
        r = [0] * (2 * n)
        for i in range(n):
            carry = 0
            for j in range(n):
                r[i + j], carry = mac_with_carry(r[i + j], a_limbs[i], b_limbs[j], carry)
            r[i + n] = carry
        return self.mont_reduce_limbs(r)

    def mont_reduce_limbs(self, r):
        """fq.rs:1040-1123 / fr.rs:520-571 (HAC 14.32), then the conditional subtract (`reduce`)."""
        n = self.nlimbs
        r = list(r)
        mod = int_to_limbs(self.p, n)
        carry2 = 0
        for i in range(n):
            k = (r[i] * self.INV) & MASK64
            carry = 0
            _, carry = mac_with_carry(r[i], k, mod[0], carry)
            for j in range(1, n):
                r[i + j], carry = mac_with_carry(r[i + j], k, mod[j], carry)
            r[i + n], carry2 = adc(r[i + n], carry2, carry)
        out = limbs_to_int(r[n:])
        # carry2 is always 0 for p < 2^(64n-1); the reference ignores it too
        if out >= self.p:
            out -= self.p
        return int_to_limbs(out, n)


# BLS12-381 scalar field, fr.rs:4-55
FR_MODULUS = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
# BLS12-381 base field, fq.rs:5-42
FQ_MODULUS = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab

Fr = PrimeField("Fr", FR_MODULUS, 4, generator=7, s=32)
Fq = PrimeField("Fq", FQ_MODULUS, 6, generator=2, s=1)


def fq_sqrt(a):
    """fq.rs:1147-1172: q = 3 mod 4, a^((q+1)/4) with a check."""
    r = pow(a, (FQ_MODULUS + 1) // 4, FQ_MODULUS)
    return r if (r * r) % FQ_MODULUS == a else None


class Fq2Field:
    """Fq[u]/(u^2+1), elements are (c0, c1) tuples of canonical ints. fq2.rs"""

    name = "Fq2"
    base = Fq
    nlimbs = 12
    zero = (0, 0)
    one = (1, 0)

    def add(self, a, b):
        return (Fq.add(a[0], b[0]), Fq.add(a[1], b[1]))

    def sub(self, a, b):
        return (Fq.sub(a[0], b[0]), Fq.sub(a[1], b[1]))

    def neg(self, a):
        return (Fq.neg(a[0]), Fq.neg(a[1]))

    def dbl(self, a):
        return self.add(a, a)

    def mul(self, a, b):
        """fq2.rs:118-132 (Karatsuba: 3 Fq muls)"""
        aa = Fq.mul(a[0], b[0])
        bb = Fq.mul(a[1], b[1])
        o = Fq.add(b[0], b[1])
        c1 = Fq.mul(Fq.add(a[1], a[0]), o)
        c1 = Fq.sub(Fq.sub(c1, aa), bb)
        c0 = Fq.sub(aa, bb)
        return (c0, c1)

    def sqr(self, a):
        """fq2.rs:84-98 (2 Fq muls)"""
        ab = Fq.mul(a[0], a[1])
        c0c1 = Fq.add(a[0], a[1])
        c0 = Fq.mul(Fq.add(Fq.neg(a[1]), a[0]), c0c1)
        c0 = Fq.sub(c0, ab)
        c1 = Fq.add(ab, ab)
        c0 = Fq.add(c0, ab)
        return (c0, c1)

    def inv(self, a):
        """fq2.rs:134-153"""
        t = Fq.add(Fq.sqr(a[0]), Fq.sqr(a[1]))
        ti = Fq.inv(t)
        if ti is None:
            return None
        return (Fq.mul(a[0], ti), Fq.neg(Fq.mul(a[1], ti)))

    def pow(self, a, e):
        r = self.one
        for i in reversed(range(e.bit_length())):
            r = self.sqr(r)
            if (e >> i) & 1:
                r = self.mul(r, a)
        return r

    def is_zero(self, a):
        return a[0] == 0 and a[1] == 0

    def eq(self, a, b):
        return a == b

    def mul_by_nonresidue(self, a):
        """fq2.rs:51-58: multiply by (u + 1)"""
        return (Fq.sub(a[0], a[1]), Fq.add(a[0], a[1]))

    def to_mont_limbs(self, a):
        return Fq.to_mont_limbs(a[0]) + Fq.to_mont_limbs(a[1])

    def from_mont_limbs(self, limbs):
        return (Fq.from_mont_limbs(limbs[:6]), Fq.from_mont_limbs(limbs[6:12]))

    def sqrt(self, a):
        """Square root in Fq2 (q = 3 mod 4), algorithm 9 of eprint 2012/685 as in fq2.rs:160-204."""
        if self.is_zero(a):
            return a
        q = FQ_MODULUS
        a1 = self.pow(a, (q - 3) // 4)
        alpha = self.mul(self.sqr(a1), a)
        a0 = self.mul(self.frobenius(alpha), alpha)
        neg1 = (Fq.neg(1), 0)
        if a0 == neg1:
            return None
        a1 = self.mul(a1, a)
        if alpha == neg1:
            return self.mul(a1, (0, 1))
        alpha = self.add(alpha, self.one)
        alpha = self.pow(alpha, (q - 1) // 2)
        return self.mul(alpha, a1)

    def frobenius(self, a):
        return (a[0], Fq.neg(a[1]))


Fq2 = Fq2Field()
