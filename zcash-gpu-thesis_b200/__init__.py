"""zcash-gpu-thesis_b200: B200-native (sm_100a CUDA) Groth16 prover numerics behind bellman's API.

  csrc/         hand-written CUDA kernels + the C ABI (-> libb200zk.so, declared in include/b200zk.h)
  bellman.py    host-side mirror of the reference interface (Worker, multiexp, EvaluationDomain, ...)

The directory name has hyphens (it is the reference's name); import it as `zcash_gpu_thesis_b200`
(the shim at the repo root) or through importlib.
"""
from . import _lib  # noqa: F401
from .bellman import (  # noqa: F401
    Bases,
    CudaError,
    DensityTracker,
    DeviceBuffer,
    EvaluationDomain,
    FullDensity,
    GroupDecodingError,
    IoError,
    PolynomialDegreeTooLarge,
    SynthesisError,
    UnexpectedIdentity,
    Parameters,
    Proof,
    Worker,
    create_proof_from_assignment,
    create_proofs_from_assignments,
    create_proof_bytes_from_assignments,
    create_proof,
    create_random_proof,
    create_proofs,
    ProvingAssignment,
    synthesize,
    ONE,
    decode_points,
    encode_points,
    field_vec,
    fixed_base_mul,
    h_poly,
    into_affine,
    multiexp,
    multiexp_async,
    multiexp_batch,
    MultiWorker,
    PreparedVerifyingKey,
    pairing,
    prepare_verifying_key,
    verify_proof,
    verify_proofs,
    ShardedBases,
    multi_plan,
    ntt_plan,
    ntt_host,
    point_op,
)
from .generator import GeneratedParameters, KeypairAssembly, UnconstrainedVariable, generate_parameters  # noqa: F401
from ._lib import G1, G2, FR, FQ, FQ2, FQ2_PAIR, FFT, IFFT, COSET_FFT, ICOSET_FFT  # noqa: F401
