"""groth16::create_proof on the GPU vs the oracle: identical proof points and identical 192 proof bytes for fixed (r, s)
(groth16/mod.rs:493-575 and bellman/tests/mimc.rs re-targeted; the CRS comes from the oracle's generate_parameters)."""
import numpy as np
import pytest

from oracle.curve import G1, G2
from oracle.fields import Fq, Fq2, Fr, int_to_limbs
from oracle.groth16 import ONE, Circuit, generate_parameters, proof_bytes, prove_from_assignment, synthesize_assignment
from tests import util

pytestmark = pytest.mark.gpu


from oracle.pairing import Bls12 as BlsEngine  # Fr, G1, G2 + the pairing check of verify_proof
from oracle.groth16 import Proof as OracleProof, verify_proof


class MiMCLike(Circuit):
    """A small MiMC-style circuit (bellman/tests/mimc.rs:86-167): x_{i+1} = (x_i + k + c_i)^3 + y_i, swapped each round."""

    def __init__(self, xl, xr, constants):
        self.xl, self.xr, self.constants = xl, xr, constants

    def synthesize(self, cs):
        p = Fr.p
        xl_v, xr_v = self.xl, self.xr
        xl = cs.alloc(lambda: xl_v)
        xr = cs.alloc(lambda: xr_v)
        n = len(self.constants)
        for i, c in enumerate(self.constants):
            t = (xl_v + c) % p
            tmp_v = t * t % p
            tmp = cs.alloc(lambda v=tmp_v: v)
            cs.enforce([(xl, 1), (ONE, c)], [(xl, 1), (ONE, c)], [(tmp, 1)])
            new_v = (tmp_v * t + xr_v) % p
            if i == n - 1:
                new = cs.alloc_input(lambda v=new_v: v)
            else:
                new = cs.alloc(lambda v=new_v: v)
            cs.enforce([(tmp, 1)], [(xl, 1), (ONE, c)], [(new, 1), (xr, -1)])
            xr, xr_v = xl, xl_v
            xl, xl_v = new, new_v


class Silly(Circuit):
    """groth16/mod.rs:493-535 MySillyCircuit: a * b = c with c public."""

    def __init__(self, a, b):
        self.a, self.b = a, b

    def synthesize(self, cs):
        a = cs.alloc(lambda: self.a)
        b = cs.alloc(lambda: self.b)
        c = cs.alloc_input(lambda: self.a * self.b % Fr.p)
        cs.enforce([(a, 1)], [(b, 1)], [(c, 1)])


def _aff_limbs(G, p):
    return np.array(G.affine_to_limbs(p), dtype=np.uint64)


def _upload(worker, params):
    import zcash_gpu_thesis_b200 as zk

    pack = lambda G, v: np.array([G.affine_to_limbs(p) for p in v], dtype=np.uint64).reshape(len(v), -1)
    vk = params.vk
    return zk.Parameters(worker, pack(G1, params.h), pack(G1, params.l), pack(G1, params.a), pack(G1, params.b_g1), pack(G2, params.b_g2),
                         _aff_limbs(G1, vk.alpha_g1), _aff_limbs(G1, vk.beta_g1), _aff_limbs(G2, vk.beta_g2), _aff_limbs(G1, vk.delta_g1),
                         _aff_limbs(G2, vk.delta_g2))


def _gpu_prove(worker, dev_params, asg, r, s):
    import zcash_gpu_thesis_b200 as zk

    mont = lambda v: np.array([Fr.to_mont_limbs(x) for x in v], dtype=np.uint64).reshape(len(v), 4)
    rep = lambda v: np.array([int_to_limbs(x, 4) for x in v], dtype=np.uint64).reshape(len(v), 4)
    return zk.create_proof_from_assignment(worker, dev_params, mont(asg.a), mont(asg.b), mont(asg.c), rep(asg.input_assignment),
                                           rep(asg.aux_assignment), asg.a_aux_density, asg.b_input_density, asg.b_aux_density, r, s)


def _same_point(G, limbs, inf, want):
    if want[2]:
        return bool(inf)
    return (not inf) and list(map(int, limbs)) == G.affine_to_limbs(want)


@pytest.mark.parametrize("which", ["silly", "mimc"])
def test_create_proof_matches_oracle(worker, which):
    E = BlsEngine
    r0 = util.rng(2000)
    rnd = lambda: util.rows_to_ints(util.random_fr_repr(r0, 1))[0]
    if which == "silly":
        blank, circ = Silly(0, 0), Silly(rnd(), rnd())
    else:
        consts = [rnd() for _ in range(24)]
        blank, circ = MiMCLike(0, 0, consts), MiMCLike(rnd(), rnd(), consts)
    toxic = [rnd() for _ in range(5)]
    params, _ = generate_parameters(E, blank, G1.gen, G2.gen, *toxic)
    dev = _upload(worker, params)
    for trial in range(2):
        r, s = rnd(), rnd()
        asg = synthesize_assignment(E, circ)
        want = prove_from_assignment(E, asg, params, r, s)
        got = _gpu_prove(worker, dev, asg, r, s)
        assert _same_point(G1, got.a, got.inf[0], want.a)
        assert _same_point(G2, got.b, got.inf[1], want.b)
        assert _same_point(G1, got.c, got.inf[2], want.c)
        assert got.write(worker) == proof_bytes(want)
        assert len(got.write(worker)) == 192  # groth16/mod.rs:567
    # the GPU proof verifies under the real pairing check (verifier.rs:35-66) and fails for a wrong public input
    aff = lambda G, limbs, inf: G.affine_zero() if inf else tuple(
        [G.F.from_mont_limbs(list(map(int, limbs[: len(limbs) // 2]))), G.F.from_mont_limbs(list(map(int, limbs[len(limbs) // 2:]))), False])
    gp = OracleProof(a=aff(G1, got.a, got.inf[0]), b=aff(G2, got.b, got.inf[1]), c=aff(G1, got.c, got.inf[2]))
    public = asg.input_assignment[1:]
    assert verify_proof(E, params.vk, gp, public)
    assert not verify_proof(E, params.vk, gp, [(public[0] + 1) % Fr.p] + public[1:])


def test_batched_proofs_equal_single_proofs(worker):
    """b200zk_groth16_prove_batch: proofs of one circuit proved in lock-step (groups of 3 here: 3 + 2) are the proofs
    create_proof gives one at a time, for different witnesses and different (r, s)."""
    import zcash_gpu_thesis_b200 as zk

    E = BlsEngine
    r0 = util.rng(2100)
    rnd = lambda: util.rows_to_ints(util.random_fr_repr(r0, 1))[0]
    consts = [rnd() for _ in range(16)]
    params, _ = generate_parameters(E, MiMCLike(0, 0, consts), G1.gen, G2.gen, *[rnd() for _ in range(5)])
    dev = _upload(worker, params)
    mont = lambda v: np.array([Fr.to_mont_limbs(x) for x in v], dtype=np.uint64).reshape(len(v), 4)
    rep = lambda v: np.array([int_to_limbs(x, 4) for x in v], dtype=np.uint64).reshape(len(v), 4)
    batch, wants = [], []
    for _ in range(5):
        asg = synthesize_assignment(E, MiMCLike(rnd(), rnd(), consts))
        r, s = rnd(), rnd()
        wants.append(prove_from_assignment(E, asg, params, r, s))
        batch.append((mont(asg.a), mont(asg.b), mont(asg.c), rep(asg.input_assignment), rep(asg.aux_assignment), asg.a_aux_density,
                      asg.b_input_density, asg.b_aux_density, r, s))
    proofs = zk.create_proofs_from_assignments(worker, dev, batch, lockstep=3)
    assert len(proofs) == 5
    for got, want in zip(proofs, wants):
        assert got.write(worker) == proof_bytes(want)
    one = zk.create_proof_from_assignment(worker, dev, *batch[3])
    assert one.write(worker) == proofs[3].write(worker)
    # the outer-FFI form: N assignments in, N x 192 bytes out (b200zk_groth16_prove_batch_bytes)
    assert zk.create_proof_bytes_from_assignments(worker, dev, batch, lockstep=2) == [proof_bytes(w_) for w_ in wants]
    assert zk.create_proofs_from_assignments(worker, dev, []) == []
    with pytest.raises(ValueError):
        zk.create_proofs_from_assignments(worker, dev, [batch[0], tuple(x[:-1] for x in batch[1][:3]) + batch[1][3:]])  # fewer constraints


def test_circuit_facing_api(worker):
    """create_proof / create_random_proof / create_proofs (prover.rs:192-211): synthesis by the product's own
    ProvingAssignment, then the GPU prover; fixed (r, s) gives the oracle's bytes, random (r, s) verifies."""
    import random

    import zcash_gpu_thesis_b200 as zk

    E = BlsEngine
    r0 = util.rng(2200)
    rnd = lambda: util.rows_to_ints(util.random_fr_repr(r0, 1))[0]
    consts = [rnd() for _ in range(12)]
    params, _ = generate_parameters(E, MiMCLike(0, 0, consts), G1.gen, G2.gen, *[rnd() for _ in range(5)])
    dev = _upload(worker, params)
    circ = MiMCLike(rnd(), rnd(), consts)
    r, s = rnd(), rnd()
    asg = synthesize_assignment(E, circ)
    assert zk.create_proof(worker, circ, dev, r, s).write(worker) == proof_bytes(prove_from_assignment(E, asg, params, r, s))
    aff = lambda G, limbs: tuple([G.F.from_mont_limbs(list(map(int, limbs[: len(limbs) // 2]))), G.F.from_mont_limbs(list(map(int, limbs[len(limbs) // 2:]))), False])
    p = zk.create_random_proof(worker, circ, dev, random.Random(7))
    assert verify_proof(E, params.vk, OracleProof(a=aff(G1, p.a), b=aff(G2, p.b), c=aff(G1, p.c)), asg.input_assignment[1:])
    circs = [MiMCLike(rnd(), rnd(), consts) for _ in range(3)]
    rs = [(rnd(), rnd()) for _ in range(3)]
    for c, (ri, si), got in zip(circs, rs, zk.create_proofs(worker, circs, dev, rs)):
        assert got.write(worker) == proof_bytes(prove_from_assignment(E, synthesize_assignment(E, c), params, ri, si))


def test_subversion_check(worker):
    """prover.rs:320-324: delta at infinity -> UnexpectedIdentity."""
    import zcash_gpu_thesis_b200 as zk

    E = BlsEngine
    params, _ = generate_parameters(E, Silly(0, 0), G1.gen, G2.gen, 3, 5, 7, 11, 13)
    pack = lambda G, v: np.array([G.affine_to_limbs(p) for p in v], dtype=np.uint64).reshape(len(v), -1)
    vk = params.vk
    dev = zk.Parameters(worker, pack(G1, params.h), pack(G1, params.l), pack(G1, params.a), pack(G1, params.b_g1), pack(G2, params.b_g2),
                        _aff_limbs(G1, vk.alpha_g1), _aff_limbs(G1, vk.beta_g1), _aff_limbs(G2, vk.beta_g2), _aff_limbs(G1, vk.delta_g1),
                        _aff_limbs(G2, vk.delta_g2), vk_infinity=[0, 0, 0, 1, 0])
    asg = synthesize_assignment(E, Silly(2, 3))
    with pytest.raises(zk.UnexpectedIdentity):
        _gpu_prove(worker, dev, asg, 5, 6)


def test_batch_round_robin_over_contexts(worker):
    """sharding.prove_on_devices: lock-step groups of one batch go round-robin over several contexts (two GPUs when the box has
    them, else two contexts of the one GPU); the proofs come back in order and equal the oracle's."""
    import zcash_gpu_thesis_b200 as zk
    from zcash_gpu_thesis_b200.sharding import lockstep_groups, prove_on_devices

    assert lockstep_groups(7, 3) == [[0, 1, 2], [3, 4, 5], [6]] and lockstep_groups(0, 4) == []
    E = BlsEngine
    r0 = util.rng(2300)
    rnd = lambda: util.rows_to_ints(util.random_fr_repr(r0, 1))[0]
    consts = [rnd() for _ in range(10)]
    params, _ = generate_parameters(E, MiMCLike(0, 0, consts), G1.gen, G2.gen, *[rnd() for _ in range(5)])
    second = zk.Worker(1 if worker.lib.b200zk_device_count() > 1 else 0)
    workers = [worker, second]
    devs = [_upload(w, params) for w in workers]
    batch, wants = [], []
    for _ in range(7):
        circ = MiMCLike(rnd(), rnd(), consts)
        r, s = rnd(), rnd()
        wants.append(proof_bytes(prove_from_assignment(E, synthesize_assignment(E, circ), params, r, s)))
        batch.append(zk.synthesize(circ).as_tuple(r, s))
    proofs = prove_on_devices(workers, devs, batch, lockstep=2)
    assert [p.write(worker) for p in proofs] == wants
    del devs
    second.close()


def _spend_crs(worker, rng, shape):
    """Synthetic CRS of a circuit shape: query vectors [k_i]G generated on the device (checked against the oracle on a sample)
    and downloaded, so that the GPU prover and the CPU oracle consume the same arrays."""
    import zcash_gpu_thesis_b200 as zk
    from oracle import cref, spend
    from tools import synthetic as syn

    sizes = syn.crs_sizes(shape)
    host, dev = {}, {}
    for name, n in sizes.items():
        group, gen, gname, w = (zk.G2, util.g2_gen_limbs(), "g2", 24) if name == "b_g2" else (zk.G1, util.g1_gen_limbs(), "g1", 12)
        k = syn.base_multipliers(rng, n)
        dxy, dinf, _ = zk.fixed_base_mul(worker, group, gen, k, 64)
        xy = dxy.download(np.uint64, n * w).reshape(n, w)
        pick = rng.choice(n, size=8, replace=False)
        want, _ = cref.scalar_muls(gname, gen, k[pick])
        assert np.array_equal(xy[pick], want)
        host[name] = xy
        dev[name] = zk.Bases.from_device(worker, group, dxy, n)
        dxy.free(); dinf.free()
    vk1, _ = cref.scalar_muls("g1", util.g1_gen_limbs(), syn.base_multipliers(rng, 3))   # alpha_g1, beta_g1, delta_g1
    vk2, _ = cref.scalar_muls("g2", util.g2_gen_limbs(), syn.base_multipliers(rng, 2))   # beta_g2, delta_g2
    params = zk.Parameters(worker, dev["h"], dev["l"], dev["a"], dev["b_g1"], dev["b_g2"], vk1[0], vk1[1], vk2[0], vk1[2], vk2[1])
    crs = spend.HostCrs(host["h"], host["l"], host["a"], host["b_g1"], host["b_g2"], vk1[0], vk1[1], vk2[0], vk1[2], vk2[1])
    return params, crs, dev


@pytest.mark.parametrize("precompute", [False, True])
def test_spend_shaped_proof_matches_oracle(worker, precompute):
    """BASELINE config 5 at its real size (m = 2^17, multiexps of 131 071 / 98 638 / 8 + 85 382 / 1 + 61 299 bases, witness-like
    scalars, random density maps): the 192 proof bytes of the GPU prover -- single call and lock-step 8 -- equal the CPU oracle's
    (C++ port of the H block and of the reference's eight multiexps + the restated assembly of prover.rs:326-363)."""
    import zcash_gpu_thesis_b200 as zk
    from oracle import spend
    from tools import synthetic as syn

    r0 = util.rng(2400)
    params, crs, dev = _spend_crs(worker, r0, syn.SPEND_SHAPE)
    if precompute:
        for b in dev.values():
            b.precompute(0)
    asgs = [syn.spend_assignment(r0) for _ in range(3)]
    answers = [spend.multiexp_phase(crs, a) for a in asgs]
    rnd = lambda: util.rows_to_ints(util.random_fr_repr(r0, 1))[0]
    order = ("a", "b", "c", "inputs", "aux", "a_aux_density", "b_input_density", "b_aux_density")
    r, s = rnd(), rnd()
    got = zk.create_proof_from_assignment(worker, params, *[asgs[0][k] for k in order], r, s)
    assert got.write(worker) == spend.assemble(crs, answers[0], r, s)
    batch, wants = [], []
    for k in range(8):
        r, s = rnd(), rnd()
        batch.append(tuple(asgs[k % 3][f] for f in order) + (r, s))
        wants.append(spend.assemble(crs, answers[k % 3], r, s))
    proofs = zk.create_proofs_from_assignments(worker, params, batch, lockstep=8)
    assert [p.write(worker) for p in proofs] == wants


def test_parameters_read_write_round_trip(worker):
    """groth16/mod.rs:536-545 (serialization test of the reference) re-targeted: Parameters::write bytes made by the oracle ->
    Parameters::read on the device (checked and unchecked) -> the proofs equal the oracle's, Parameters::write gives the same bytes
    back, the VerifyingKey comes back element by element; truncated input, a point at infinity in a query vector and (checked) a
    point off the curve are decoding errors."""
    import zcash_gpu_thesis_b200 as zk
    from oracle.groth16 import parameters_bytes

    E = BlsEngine
    r0 = util.rng(2500)
    rnd = lambda: util.rows_to_ints(util.random_fr_repr(r0, 1))[0]
    consts = [rnd() for _ in range(6)]
    params, _ = generate_parameters(E, MiMCLike(0, 0, consts), G1.gen, G2.gen, *[rnd() for _ in range(5)])
    data = parameters_bytes(params)
    asg = synthesize_assignment(E, MiMCLike(rnd(), rnd(), consts))
    r, s = rnd(), rnd()
    want = proof_bytes(prove_from_assignment(E, asg, params, r, s))
    for checked in (True, False):
        for pre in (False, True):
            dev = zk.Parameters.read(worker, data, checked=checked, precompute=pre)
            assert dev.query_sizes() == dict(h=len(params.h), l=len(params.l), a=len(params.a), b_g1=len(params.b_g1), b_g2=len(params.b_g2))
            assert _gpu_prove(worker, dev, asg, r, s).write(worker) == want
            assert dev.write() == data
            vk = dev.verifying_key()
            assert list(map(int, vk["gamma_g2"][0])) == G2.affine_to_limbs(params.vk.gamma_g2) and not vk["gamma_g2"][1]
            assert [list(map(int, row)) for row in vk["ic"]] == [G1.affine_to_limbs(p) for p in params.vk.ic]
            dev.free()
    for cut in (10, 96 * 3 + 192 * 3 + 2, len(data) - 1):
        with pytest.raises(zk.GroupDecodingError):
            zk.Parameters.read(worker, data[:cut])
    # a point at infinity inside the h vector (flag byte 0x40, all else zero): rejected even when unchecked (mod.rs:300-304)
    h_off = 96 * 3 + 192 * 3 + 4 + 96 * len(params.vk.ic) + 4
    bad = bytearray(data)
    bad[h_off:h_off + 96] = bytes([0x40]) + bytes(95)
    with pytest.raises(zk.GroupDecodingError):
        zk.Parameters.read(worker, bytes(bad), checked=False)
    # a coordinate changed: off the curve -> only the checked read notices
    bad = bytearray(data)
    bad[h_off + 95] ^= 1
    with pytest.raises(zk.GroupDecodingError):
        zk.Parameters.read(worker, bytes(bad), checked=True)
    zk.Parameters.read(worker, bytes(bad), checked=False).free()
    # parameters assembled from separate vectors carry no gamma_g2 / ic: write refuses
    with pytest.raises(ValueError):
        _upload(worker, params).write()


def test_identity_vk_elements_are_skipped(worker):
    """add_assign_mixed skips an identity operand (ec.rs:447-449): alpha_g1 / beta_g1 / beta_g2 at infinity must not be added
    as finite points by the assembly (prover.rs:326-345).  Oracle: the same parameters with those elements at infinity."""
    import zcash_gpu_thesis_b200 as zk

    E = BlsEngine
    r0 = util.rng(2600)
    rnd = lambda: util.rows_to_ints(util.random_fr_repr(r0, 1))[0]
    consts = [rnd() for _ in range(4)]
    params, _ = generate_parameters(E, MiMCLike(0, 0, consts), G1.gen, G2.gen, *[rnd() for _ in range(5)])
    params.vk.alpha_g1 = G1.affine_zero()
    params.vk.beta_g1 = G1.affine_zero()
    params.vk.beta_g2 = G2.affine_zero()
    pack = lambda G, v: np.array([G.affine_to_limbs(p) for p in v], dtype=np.uint64).reshape(len(v), -1)
    vk = params.vk
    zero1, zero2 = np.zeros(12, np.uint64), np.zeros(24, np.uint64)
    dev = zk.Parameters(worker, pack(G1, params.h), pack(G1, params.l), pack(G1, params.a), pack(G1, params.b_g1), pack(G2, params.b_g2),
                        zero1, zero1, zero2, _aff_limbs(G1, vk.delta_g1), _aff_limbs(G2, vk.delta_g2), vk_infinity=[1, 1, 1, 0, 0])
    asg = synthesize_assignment(E, MiMCLike(rnd(), rnd(), consts))
    r, s = rnd(), rnd()
    assert _gpu_prove(worker, dev, asg, r, s).write(worker) == proof_bytes(prove_from_assignment(E, asg, params, r, s))
