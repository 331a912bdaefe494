"""groth16::verify_proof and the BLS12-381 pairing on the GPU vs the oracle (bellman/src/groth16/verifier.rs:18-66;
pairing/src/bls12_381/tests/mod.rs:5-53 RELIC known-answer test; pairing/src/tests/engine.rs bilinearity re-targeted)."""
import json
import os

import numpy as np
import pytest

from oracle import pairing as op
from oracle.curve import G1, G2
from oracle.fields import Fq, Fr, int_to_limbs
from oracle.groth16 import ONE, generate_parameters, parameters_bytes, prove_from_assignment, synthesize_assignment, verify_proof
from tests import util
from tests.test_gpu_groth16 import BlsEngine, MiMCLike, _gpu_prove

pytestmark = pytest.mark.gpu


def _fq12_ints(row):
    """72 u64 Montgomery limbs -> 12 canonical ints in the order c0.c0.c0, c0.c0.c1, c0.c1.c0, ..."""
    return [Fq.from_mont_limbs([int(v) for v in row[6 * i:6 * i + 6]]) for i in range(12)]


def test_pairing_known_answer_and_oracle(worker):
    import zcash_gpu_thesis_b200 as zk

    kat = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))["pairing_g1_g2"]
    g1, g2 = util.g1_gen_limbs(), util.g2_gen_limbs()
    got = zk.pairing(worker, g1[None, :], g2[None, :])
    assert _fq12_ints(got[0]) == [int(v) for v in kat["fq12"]]  # e(G1::one(), G2::one()) of RELIC
    # random points: the device value == the oracle's Fq12 value; identity operands give one
    r = util.rng(3100)
    p_xy, pk = util.random_bases("g1", r, 3, bits64=False)
    q_xy, qk = util.random_bases("g2", r, 3, bits64=False)
    got = zk.pairing(worker, p_xy, q_xy)
    aff = lambda G, limbs: (G.F.from_mont_limbs([int(v) for v in limbs[: len(limbs) // 2]]), G.F.from_mont_limbs([int(v) for v in limbs[len(limbs) // 2:]]), False)
    for i in range(3):
        assert _fq12_ints(got[i]) == op.f12_flat(op.pairing(aff(G1, p_xy[i]), aff(G2, q_xy[i]))), i
    one = zk.pairing(worker, p_xy[:2], q_xy[:2], g1_inf=[1, 0], g2_inf=[0, 1])
    assert all(_fq12_ints(row) == [1] + [0] * 11 for row in one)
    # bilinearity (pairing/src/tests/engine.rs): e([a]G1, [b]G2) == e([ab]G1, G2) == e(G1, [ab]G2)
    a, b = util.rows_to_ints(util.random_fr_repr(r, 2))
    ab = a * b % Fr.p
    sa, _ = util.affine_of_scalar("g1", a)
    sb, _ = util.affine_of_scalar("g2", b)
    sab1, _ = util.affine_of_scalar("g1", ab)
    sab2, _ = util.affine_of_scalar("g2", ab)
    vals = zk.pairing(worker, np.stack([sa, sab1, g1]), np.stack([sb, g2, sab2]))
    assert np.array_equal(vals[0], vals[1]) and np.array_equal(vals[0], vals[2])


def test_verify_proofs_matches_oracle(worker):
    """Proofs made by the GPU prover over a CRS read from Parameters::write bytes: verify_proof on the device accepts them and
    rejects what the oracle's verify_proof rejects (wrong public input, swapped / altered proof elements, the identity)."""
    import zcash_gpu_thesis_b200 as zk

    E = BlsEngine
    r0 = util.rng(3200)
    rnd = lambda: util.rows_to_ints(util.random_fr_repr(r0, 1))[0]
    consts = [rnd() for _ in range(5)]
    params, _ = generate_parameters(E, MiMCLike(0, 0, consts), G1.gen, G2.gen, *[rnd() for _ in range(5)])
    dev = zk.Parameters.read(worker, parameters_bytes(params), checked=True)
    pvk = zk.prepare_verifying_key(worker, dev.verifying_key())
    proofs, publics, oracle_proofs = [], [], []
    for _ in range(3):
        asg = synthesize_assignment(E, MiMCLike(rnd(), rnd(), consts))
        r, s = rnd(), rnd()
        proofs.append(_gpu_prove(worker, dev, asg, r, s))
        publics.append(asg.input_assignment[1:])
        oracle_proofs.append(prove_from_assignment(E, asg, params, r, s))
    assert zk.verify_proofs(worker, pvk, proofs, publics) == [True, True, True]
    assert all(verify_proof(E, params.vk, p, pub) for p, pub in zip(oracle_proofs, publics))
    # wrong public input / inputs of another proof
    bad = [[(publics[0][0] + 1) % Fr.p] + publics[0][1:], publics[2], publics[1]]
    assert zk.verify_proofs(worker, pvk, proofs, bad) == [False, False, False]
    assert not verify_proof(E, params.vk, oracle_proofs[0], bad[0])
    # altered proofs: c of another proof; a negated; b at infinity
    mixed = zk.Proof(proofs[0].a, proofs[0].b, proofs[1].c, proofs[0].inf)
    neg_y = np.array(Fq.to_mont_limbs((Fq.p - Fq.from_mont_limbs([int(v) for v in proofs[0].a[6:]])) % Fq.p), dtype=np.uint64)
    negated = zk.Proof(np.concatenate([proofs[0].a[:6], neg_y]), proofs[0].b, proofs[0].c, proofs[0].inf)
    b_inf = zk.Proof(proofs[0].a, proofs[0].b, proofs[0].c, [False, True, False])
    assert zk.verify_proofs(worker, pvk, [mixed, negated, b_inf, proofs[0]], [publics[0]] * 4) == [False, False, False, True]
    assert zk.verify_proof(worker, pvk, proofs[1], publics[1]) is True
    # MalformedVerifyingKey (verifier.rs:41-43)
    with pytest.raises(ValueError):
        zk.verify_proofs(worker, pvk, proofs[:1], [publics[0] + [5]])
    assert zk.verify_proofs(worker, pvk, [], []) == []
