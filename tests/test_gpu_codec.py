"""Wire format on the device vs the reference's golden byte vectors (pairing/src/bls12_381/tests/mod.rs:55-97 and the four
tests/*.dat files: 1000 multiples of each generator in both encodings).  The oracle regenerates the bytes (pinned by sha256 in
tests/golden/kat.json); the device decodes / encodes them and must agree byte for byte."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle.curve import G1, G2
from tests import util

pytestmark = pytest.mark.gpu
KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))


def _multiples(G, count):
    pts, p = [], G.zero()
    for _ in range(count):
        pts.append(G.into_affine(p))
        p = G.add_mixed(p, G.gen)
    return pts


@pytest.mark.parametrize("name", ["g1", "g2"])
def test_dat_vectors_roundtrip_on_device(worker, name):
    import zcash_gpu_thesis_b200 as zk

    G, code = (G1, zk.G1) if name == "g1" else (G2, zk.G2)
    pts = _multiples(G, 1000)
    unc = b"".join(G.encode_uncompressed(p) for p in pts)
    comp = b"".join(G.encode_compressed(p) for p in pts)
    assert hashlib.sha256(unc).hexdigest() == KAT["dat"][f"{name}_uncompressed"]["sha256"]
    assert hashlib.sha256(comp).hexdigest() == KAT["dat"][f"{name}_compressed"]["sha256"]
    # decode (entry 0 is the point at infinity, flag 0x40), with the curve + subgroup checks
    xy, inf = zk.decode_points(worker, code, unc, checked=True)
    assert inf[0] == 1 and not inf[1:].any()
    want = np.array([G.affine_to_limbs(p) for p in pts[1:]], dtype=np.uint64)
    assert np.array_equal(xy[1:], want)
    # encode both ways: byte-for-byte the reference's files
    assert zk.encode_points(worker, code, xy, inf, compressed=False) == unc
    got_c = zk.encode_points(worker, code, xy, inf, compressed=True)
    assert got_c == comp
    assert hashlib.sha256(got_c).hexdigest() == KAT["dat"][f"{name}_compressed"]["sha256"]


def test_parameters_read_path_and_multiexp(worker):
    """Bases.read (the point vectors of Parameters::read): decoded bases give the same multiexp as limb-uploaded ones; the
    identity is rejected like groth16/mod.rs:300-304 unless allowed."""
    import zcash_gpu_thesis_b200 as zk
    from oracle import cref

    pts = _multiples(G1, 300)
    data = b"".join(G1.encode_uncompressed(p) for p in pts[1:])
    bases = zk.Bases.read(worker, zk.G1, data, checked=True)
    assert len(bases) == 299
    r = util.rng(1700)
    exps = util.random_fr_repr(r, 299)
    got = zk.multiexp(worker, (bases, 0), zk.FullDensity(), exps)
    xy = np.array([G1.affine_to_limbs(p) for p in pts[1:]], dtype=np.uint64)
    st, want = cref.multiexp("g1", xy, exps)
    assert st == 0 and np.array_equal(zk.into_affine(worker, zk.G1, got)[0][0], cref.into_affine("g1", want)[0])
    with pytest.raises(zk.GroupDecodingError, match="point 0: point at infinity"):
        zk.Bases.read(worker, zk.G1, G1.encode_uncompressed(pts[0]) + data)
    ok = zk.Bases.read(worker, zk.G1, G1.encode_uncompressed(pts[0]) + data, allow_infinity=True)
    with pytest.raises(zk.UnexpectedIdentity):  # the identity base is consumed by a non-zero exponent
        zk.multiexp(worker, (ok, 0), zk.FullDensity(), exps)


def test_decoding_errors(worker):
    """GroupDecodingError cases of ec.rs:686-736 / 125-144."""
    import zcash_gpu_thesis_b200 as zk
    from oracle.fields import FQ_MODULUS

    good = G1.encode_uncompressed(G1.gen)
    bad_flag = bytes([good[0] | 0x80]) + good[1:]
    with pytest.raises(zk.GroupDecodingError, match="compression"):
        zk.decode_points(worker, zk.G1, bad_flag)
    sign_flag = bytes([good[0] | 0x20]) + good[1:]
    with pytest.raises(zk.GroupDecodingError, match="unexpected information"):
        zk.decode_points(worker, zk.G1, sign_flag)
    not_in_field = FQ_MODULUS.to_bytes(48, "big") + good[48:]
    with pytest.raises(zk.GroupDecodingError, match="not in the field"):
        zk.decode_points(worker, zk.G1, good + not_in_field)
    off_curve = good[:48] + (int.from_bytes(good[48:], "big") ^ 1).to_bytes(48, "big")
    zk.decode_points(worker, zk.G1, off_curve, checked=False)  # unchecked accepts it (into_affine_unchecked)
    with pytest.raises(zk.GroupDecodingError, match="point 1: point is not on the curve"):
        zk.decode_points(worker, zk.G1, good + off_curve, checked=True)
    # on the curve but outside the prime-order subgroup: x = 4 gives a point of E(Fq) whose order does not divide r (ec.rs:1040-1056 style)
    from oracle.fields import fq_sqrt, Fq
    x = 0
    while True:
        x += 1
        y = fq_sqrt((x * x * x + 4) % FQ_MODULUS)
        if y is not None and not G1.is_zero(G1.mul((x, y, False), 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001)):
            break
    rogue = x.to_bytes(48, "big") + y.to_bytes(48, "big")
    with pytest.raises(zk.GroupDecodingError, match="subgroup"):
        zk.decode_points(worker, zk.G1, rogue, checked=True)
    bad_inf = bytes([0x40]) + bytes(94) + b"\x01"
    with pytest.raises(zk.GroupDecodingError, match="unexpected information"):
        zk.decode_points(worker, zk.G1, bad_inf)
