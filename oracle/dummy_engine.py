"""The reference's DummyEngine (groth16/tests/dummy_engine.rs): Fr = integers mod 64513
(S = 10, generator 5, :20, 231-261), G1 = G2 = Fr as an additive group, "pairing" = product.

ORACLE -- test infrastructure only.  Used to replay test_xordemo (groth16/tests/mod.rs:98-400).
"""
from .fields import PrimeField

DUMMY_FR = PrimeField("DummyFr", 64513, 1, generator=5, s=10)


class DummyGroup:
    """impl CurveProjective/CurveAffine for Fr, dummy_engine.rs:266-450"""
    F = DUMMY_FR

    def zero(self):
        return 0

    def is_zero(self, p):
        return p == 0

    def affine_is_zero(self, a):
        return a == 0

    def add(self, p, q):
        return (p + q) % 64513

    def add_mixed(self, p, a):
        return (p + a) % 64513

    def double(self, p):
        return (2 * p) % 64513

    def mul(self, a, k):
        return (a * k) % 64513

    def mul_proj(self, p, k):
        return (p * k) % 64513

    def into_affine(self, p):
        return p

    def into_projective(self, a):
        return a

    def affine_negate(self, a):
        return (-a) % 64513

    def eq(self, p, q):
        return p == q


class DummyEngine:
    Fr = DUMMY_FR
    G1 = DummyGroup()
    G2 = DummyGroup()

    @staticmethod
    def pairing_product_is_one(pairs):
        # "Fqk" is Fr written additively: the product of pairings is a sum of products
        return sum(a * b for a, b in pairs) % 64513 == 0
