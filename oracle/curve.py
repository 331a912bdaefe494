"""BLS12-381 G1 / G2 group law, restated from pairing/src/bls12_381/ec.rs.

ORACLE -- test infrastructure only (see oracle/__init__.py).

Jacobian points are (x, y, z) tuples, identity <=> z == 0 (ec.rs:224-240);
affine points are (x, y, infinity) with zero = (0, 1, True) (ec.rs:158-164).
"""
from __future__ import annotations

from .fields import Fq, Fq2, Fr, FQ_MODULUS, fq_sqrt


class Curve:
    def __init__(self, name, F, b, gen_xy, coord_bytes):
        self.name = name
        self.F = F
        self.b = b
        self.gen = (gen_xy[0], gen_xy[1], False)
        self.coord_bytes = coord_bytes  # 48 for G1, 96 for G2

    # ---- identities
    def zero(self):
        return (self.F.zero, self.F.one, self.F.zero)

    def affine_zero(self):
        return (self.F.zero, self.F.one, True)

    def is_zero(self, p):
        return self.F.is_zero(p[2])

    def affine_is_zero(self, a):
        return a[2]

    # ---- ec.rs:296-354 (dbl-2009-l)
    def double(self, p):
        F = self.F
        if self.is_zero(p):
            return p
        x, y, z = p
        a = F.sqr(x)
        b = F.sqr(y)
        c = F.sqr(b)
        d = F.sqr(F.add(x, b))
        d = F.sub(F.sub(d, a), c)
        d = F.dbl(d)
        e = F.add(F.dbl(a), a)
        f = F.sqr(e)
        z3 = F.dbl(F.mul(z, y))
        x3 = F.sub(F.sub(f, d), d)
        y3 = F.mul(F.sub(d, x3), e)
        c8 = F.dbl(F.dbl(F.dbl(c)))
        y3 = F.sub(y3, c8)
        return (x3, y3, z3)

    # ---- ec.rs:356-444 (add-2007-bl)
    def add(self, p, q):
        F = self.F
        if self.is_zero(p):
            return q
        if self.is_zero(q):
            return p
        x1, y1, z1 = p
        x2, y2, z2 = q
        z1z1 = F.sqr(z1)
        z2z2 = F.sqr(z2)
        u1 = F.mul(x1, z2z2)
        u2 = F.mul(x2, z1z1)
        s1 = F.mul(F.mul(y1, z2), z2z2)
        s2 = F.mul(F.mul(y2, z1), z1z1)
        if u1 == u2 and s1 == s2:
            return self.double(p)
        h = F.sub(u2, u1)
        i = F.sqr(F.dbl(h))
        j = F.mul(h, i)
        r = F.dbl(F.sub(s2, s1))
        v = F.mul(u1, i)
        x3 = F.sub(F.sub(F.sub(F.sqr(r), j), v), v)
        y3 = F.mul(F.sub(v, x3), r)
        s1j = F.dbl(F.mul(s1, j))
        y3 = F.sub(y3, s1j)
        z3 = F.sqr(F.add(z1, z2))
        z3 = F.mul(F.sub(F.sub(z3, z1z1), z2z2), h)
        return (x3, y3, z3)

    # ---- ec.rs:446-526 (madd-2007-bl)
    def add_mixed(self, p, a):
        F = self.F
        if a[2]:
            return p
        if self.is_zero(p):
            return (a[0], a[1], F.one)
        x1, y1, z1 = p
        z1z1 = F.sqr(z1)
        u2 = F.mul(a[0], z1z1)
        s2 = F.mul(F.mul(a[1], z1), z1z1)
        if x1 == u2 and y1 == s2:
            return self.double(p)
        h = F.sub(u2, x1)
        hh = F.sqr(h)
        i = F.dbl(F.dbl(hh))
        j = F.mul(h, i)
        r = F.dbl(F.sub(s2, y1))
        v = F.mul(x1, i)
        x3 = F.sub(F.sub(F.sub(F.sqr(r), j), v), v)
        j2 = F.dbl(F.mul(j, y1))
        y3 = F.sub(F.mul(F.sub(v, x3), r), j2)
        z3 = F.sub(F.sub(F.sqr(F.add(z1, h)), z1z1), hh)
        return (x3, y3, z3)

    def negate(self, p):
        if self.is_zero(p):
            return p
        return (p[0], self.F.neg(p[1]), p[2])

    def affine_negate(self, a):
        if a[2]:
            return a
        return (a[0], self.F.neg(a[1]), False)

    # ---- ec.rs:45-85
    def eq(self, p, q):
        F = self.F
        if self.is_zero(p):
            return self.is_zero(q)
        if self.is_zero(q):
            return False
        z1 = F.sqr(p[2])
        z2 = F.sqr(q[2])
        if F.mul(p[0], z2) != F.mul(q[0], z1):
            return False
        z1 = F.mul(z1, p[2])
        z2 = F.mul(z2, q[2])
        return F.mul(z2, p[1]) == F.mul(z1, q[1])

    # ---- ec.rs:586-619
    def into_affine(self, p):
        F = self.F
        if self.is_zero(p):
            return self.affine_zero()
        if p[2] == F.one:
            return (p[0], p[1], False)
        zinv = F.inv(p[2])
        zinv2 = F.sqr(zinv)
        x = F.mul(p[0], zinv2)
        y = F.mul(p[1], F.mul(zinv2, zinv))
        return (x, y, False)

    def into_projective(self, a):
        if a[2]:
            return self.zero()
        return (a[0], a[1], self.F.one)

    # ---- ec.rs:87-99 mul_bits (MSB-first double and add), scalar is an int
    def mul(self, a, k):
        res = self.zero()
        found = False
        for i in reversed(range(max(k.bit_length(), 1))):
            if found:
                res = self.double(res)
            if (k >> i) & 1:
                found = True
                res = self.add_mixed(res, a)
        return res

    def mul_proj(self, p, k):
        """$projective::mul_assign, ec.rs:528-552"""
        res = self.zero()
        found = False
        for i in reversed(range(max(k.bit_length(), 1))):
            if found:
                res = self.double(res)
            if (k >> i) & 1:
                found = True
                res = self.add(res, p)
        return res

    def is_on_curve(self, a):
        """ec.rs:125-139"""
        if a[2]:
            return True
        F = self.F
        return F.sqr(a[1]) == F.add(F.mul(F.sqr(a[0]), a[0]), self.b)

    # ---- encodings (pairing/src/bls12_381/README.md:59-75; ec.rs:686-868, 2624-2830)
    def _coord_to_bytes(self, c):
        if self.F is Fq:
            return c.to_bytes(48, "big")
        return c[1].to_bytes(48, "big") + c[0].to_bytes(48, "big")  # Fq2: c1 then c0

    def _coord_from_bytes(self, b):
        if self.F is Fq:
            v = int.from_bytes(b, "big")
            assert v < FQ_MODULUS
            return v
        c1 = int.from_bytes(b[:48], "big")
        c0 = int.from_bytes(b[48:], "big")
        assert c0 < FQ_MODULUS and c1 < FQ_MODULUS
        return (c0, c1)

    def _lex_largest(self, y):
        """y > -y: Fq compares canonical ints (fq.rs:703-708); Fq2 compares c1 then c0 (fq2.rs:21-30)"""
        ny = self.F.neg(y)
        if self.F is Fq:
            return y > ny
        return (y[1], y[0]) > (ny[1], ny[0])

    def encode_uncompressed(self, a):
        n = self.coord_bytes
        if a[2]:
            out = bytearray(2 * n)
            out[0] |= 0x40
            return bytes(out)
        return self._coord_to_bytes(a[0]) + self._coord_to_bytes(a[1])

    def encode_compressed(self, a):
        n = self.coord_bytes
        if a[2]:
            out = bytearray(n)
            out[0] |= 0x40 | 0x80
            return bytes(out)
        out = bytearray(self._coord_to_bytes(a[0]))
        if self._lex_largest(a[1]):
            out[0] |= 0x20
        out[0] |= 0x80
        return bytes(out)

    def decode_uncompressed(self, b):
        n = self.coord_bytes
        assert len(b) == 2 * n
        if b[0] & 0x80:
            raise ValueError("unexpected compression flag")
        if b[0] & 0x40:
            assert all(v == 0 for v in bytes([b[0] & 0x3F]) + b[1:])
            return self.affine_zero()
        assert not (b[0] & 0x20)
        x = self._coord_from_bytes(b[:n])
        y = self._coord_from_bytes(b[n:])
        return (x, y, False)

    def decode_compressed(self, b):
        n = self.coord_bytes
        assert len(b) == n and (b[0] & 0x80)
        if b[0] & 0x40:
            return self.affine_zero()
        greatest = bool(b[0] & 0x20)
        xb = bytes([b[0] & 0x1F]) + bytes(b[1:])
        x = self._coord_from_bytes(xb)
        F = self.F
        rhs = F.add(F.mul(F.sqr(x), x), self.b)
        y = fq_sqrt(rhs) if F is Fq else F.sqrt(rhs)
        if y is None:
            raise ValueError("not on curve")
        if self._lex_largest(y) != greatest:
            y = F.neg(y)
        return (x, y, False)

    # ---- memory layout helpers (little-endian u64 Montgomery limbs as in Rust memory)
    def affine_to_limbs(self, a):
        """x || y Montgomery limbs (12 u64 for G1, 24 for G2); infinity carried separately."""
        F = self.F
        return F.to_mont_limbs(a[0]) + F.to_mont_limbs(a[1])

    def jacobian_from_limbs(self, limbs):
        F = self.F
        k = F.nlimbs if F is not Fq2 else 12
        return (F.from_mont_limbs(limbs[0:k]), F.from_mont_limbs(limbs[k:2 * k]), F.from_mont_limbs(limbs[2 * k:3 * k]))

    def jacobian_to_limbs(self, p):
        F = self.F
        return F.to_mont_limbs(p[0]) + F.to_mont_limbs(p[1]) + F.to_mont_limbs(p[2])


# Generators, fq.rs:77-136 (decimal values quoted in the reference's comments)
G1_GEN_X = 3685416753713387016781088315183077757961620795782546409894578378688607592378376318836054947676345821548104185464507
G1_GEN_Y = 1339506544944476473020471379941921221584933875938349620426543736416511423956333506472724655353366534992391756441569
G2_GEN_X = (
    352701069587466618187139116011060144890029952792775240219908644239793785735715026873347600343865175952761926303160,
    3059144344244213709971259814753781636986470325476647558659373206291635324768958432433509563104347017837885763365758,
)
G2_GEN_Y = (
    1985150602287291935568054521177171638300868978215655730859378665066344726373823718423869104263333984641494340347905,
    927553665492332455747201965776037880757740193453592970025027978793976877002675564980949289727957565575433344219582,
)

G1 = Curve("G1", Fq, 4, (G1_GEN_X, G1_GEN_Y), 48)
G2 = Curve("G2", Fq2, (4, 4), (G2_GEN_X, G2_GEN_Y), 96)  # b = 4(u+1), ec.rs:2848-2853

assert G1.is_on_curve(G1.gen) and G2.is_on_curve(G2.gen)

__all__ = ["Curve", "G1", "G2", "Fr", "Fq", "Fq2"]
