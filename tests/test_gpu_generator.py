"""groth16::generate_parameters with the numeric part on the GPU (generator.rs:173-482) vs the oracle's restatement:
every element of Parameters / VerifyingKey must be the same affine point; a proof made with the GPU-generated CRS by the
GPU prover verifies under the oracle's pairing check."""
import numpy as np
import pytest

from oracle.curve import G1, G2
from oracle.fields import Fr, int_to_limbs
from oracle.groth16 import Proof as OracleProof
from oracle.groth16 import generate_parameters as oracle_generate
from oracle.groth16 import proof_bytes, prove_from_assignment, synthesize_assignment, verify_proof
from oracle.pairing import Bls12 as BlsEngine
from tests import util
from tests.test_gpu_groth16 import MiMCLike, Silly

pytestmark = pytest.mark.gpu


def _limbs(G, pts):
    return np.array([G.affine_to_limbs(p) for p in pts], dtype=np.uint64).reshape(len(pts), -1)


def _product_assembly(circuit):
    """synthesize into the product's own KeypairAssembly the way generator.rs:187-212 does"""
    import zcash_gpu_thesis_b200 as zk

    asm = zk.KeypairAssembly()
    asm.alloc_input()  # the constant ONE
    circuit.synthesize(asm)
    for i in range(asm.num_inputs):
        asm.enforce([(("in", i), 1)], [], [])
    return asm


@pytest.mark.parametrize("which", ["silly", "mimc"])
def test_generate_parameters_matches_oracle(worker, which):
    import zcash_gpu_thesis_b200 as zk

    r0 = util.rng(5000)
    rnd = lambda: util.rows_to_ints(util.random_fr_repr(r0, 1))[0]
    if which == "silly":
        blank, circ = Silly(0, 0), Silly(rnd(), rnd())
    else:
        consts = [rnd() for _ in range(40)]
        blank, circ = MiMCLike(0, 0, consts), MiMCLike(rnd(), rnd(), consts)
    toxic = [rnd() for _ in range(5)]
    want, oracle_asm = oracle_generate(BlsEngine, blank, G1.gen, G2.gen, *toxic)
    asm = _product_assembly(blank)
    assert (asm.num_inputs, asm.num_aux, asm.num_constraints) == (oracle_asm.num_inputs, oracle_asm.num_aux, oracle_asm.num_constraints)
    assert asm.at_aux == oracle_asm.at_aux and asm.bt_inputs == oracle_asm.bt_inputs and asm.ct_aux == oracle_asm.ct_aux
    g1, g2 = np.array(G1.affine_to_limbs(G1.gen), np.uint64), np.array(G2.affine_to_limbs(G2.gen), np.uint64)
    got = zk.generate_parameters(worker, asm, g1, g2, *toxic)
    for name, G in (("h", G1), ("l", G1), ("a", G1), ("b_g1", G1), ("b_g2", G2)):
        w = _limbs(G, getattr(want, name))
        g = getattr(got, name)
        assert g.shape == w.shape and np.array_equal(g, w), name
    assert np.array_equal(got.ic, _limbs(G1, want.vk.ic))
    for name, G in (("alpha_g1", G1), ("beta_g1", G1), ("delta_g1", G1), ("beta_g2", G2), ("gamma_g2", G2), ("delta_g2", G2)):
        assert list(map(int, getattr(got, name))) == G.affine_to_limbs(getattr(want.vk, name)), name
    # prove with the GPU-generated CRS on the GPU, verify with the oracle's pairing check
    dev = got.to_device(worker)
    asg = synthesize_assignment(BlsEngine, circ)
    mont = lambda v: np.array([Fr.to_mont_limbs(x) for x in v], dtype=np.uint64).reshape(len(v), 4)
    rep = lambda v: np.array([int_to_limbs(x, 4) for x in v], dtype=np.uint64).reshape(len(v), 4)
    r, s = rnd(), rnd()
    proof = zk.create_proof_from_assignment(worker, dev, mont(asg.a), mont(asg.b), mont(asg.c), rep(asg.input_assignment), rep(asg.aux_assignment),
                                            asg.a_aux_density, asg.b_input_density, asg.b_aux_density, r, s)
    assert proof.write(worker) == proof_bytes(prove_from_assignment(BlsEngine, asg, want, r, s))
    aff = lambda G, limbs: tuple([G.F.from_mont_limbs(list(map(int, limbs[: len(limbs) // 2]))), G.F.from_mont_limbs(list(map(int, limbs[len(limbs) // 2:]))), False])
    gp = OracleProof(a=aff(G1, proof.a), b=aff(G2, proof.b), c=aff(G1, proof.c))
    assert verify_proof(BlsEngine, want.vk, gp, asg.input_assignment[1:])


def test_unconstrained_variable_and_bad_toxic_waste(worker):
    import zcash_gpu_thesis_b200 as zk

    class Loose(Silly):
        def synthesize(self, cs):
            super().synthesize(cs)
            cs.alloc(lambda: 5)  # never used in a constraint (generator.rs:452-456)

    g1, g2 = np.array(G1.affine_to_limbs(G1.gen), np.uint64), np.array(G2.affine_to_limbs(G2.gen), np.uint64)
    with pytest.raises(zk.UnconstrainedVariable):
        zk.generate_parameters(worker, _product_assembly(Loose(0, 0)), g1, g2, 3, 5, 7, 11, 13)
    with pytest.raises(zk.UnexpectedIdentity):
        zk.generate_parameters(worker, _product_assembly(Silly(0, 0)), g1, g2, 3, 5, 7, 0, 13)  # delta = 0 (generator.rs:198-199)


def test_fr_spmv_matches_python(worker):
    """b200zk_fr_spmv_dev on a ragged matrix: empty rows, one long row (the constant ONE), rows crossing warp strides"""
    import zcash_gpu_thesis_b200 as zk
    from zcash_gpu_thesis_b200 import _lib as L
    from zcash_gpu_thesis_b200.generator import _csr

    r = util.rng(5100)
    n_cols, p = 300, Fr.p
    x = util.rows_to_ints(util.random_fr_repr(r, n_cols))
    lens = [0, 1, 2, 31, 32, 33, 300, 0, 65, 7]
    rows = [[(int(r.integers(1, 1 << 62)) * 977 % p, int(r.integers(0, n_cols))) for _ in range(k)] for k in lens]
    ptr, col, val = _csr(rows)
    xm = np.array([Fr.to_mont_limbs(v) for v in x], dtype=np.uint64).reshape(n_cols, 4)
    bufs = [worker.to_device(a) for a in (ptr, col, val, xm)]
    y = worker.alloc(len(rows) * 32)
    st = worker.lib.b200zk_fr_spmv_dev(worker.ctx, bufs[0].ptr, bufs[1].ptr, bufs[2].ptr, bufs[3].ptr, len(rows), y.ptr)
    assert st == 0
    got = y.download(np.uint64, len(rows) * 4).reshape(len(rows), 4)
    for i, row in enumerate(rows):
        want = sum(c * x[j] for c, j in row) % p
        assert list(map(int, got[i])) == Fr.to_mont_limbs(want), i
    for b in bufs + [y]:
        b.free()
