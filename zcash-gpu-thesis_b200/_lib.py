"""ctypes binding of libb200zk.so (include/b200zk.h).  There is no fallback: if the CUDA library is not
built, importing the product path raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200ZK_LIB") or os.path.join(_HERE, "libb200zk.so")  # B200ZK_LIB: another build of the same library (A/B measurements)

OK, ERR_UNEXPECTED_IDENTITY, ERR_UNEXPECTED_EOF, ERR_DEGREE_TOO_LARGE, ERR_BAD_ARG, ERR_CUDA, ERR_NCCL, ERR_DECODE = range(8)
G1, G2 = 1, 2
FR, FQ, FQ2, FQ2_PAIR = 0, 1, 2, 3
FFT, IFFT, COSET_FFT, ICOSET_FFT = 0, 1, 2, 3
OP_ADD, OP_SUB, OP_MUL, OP_SQUARE, OP_DOUBLE, OP_NEGATE, OP_INTO_REPR, OP_FROM_REPR, OP_INVERSE, OP_INVERSE_BINARY, OP_MULSUB = range(11)
POINT_DOUBLE, POINT_ADD, POINT_ADD_MIXED = 0, 1, 2

_vp, _sz, _i, _u32 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32

# every symbol include/b200zk.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "b200zk_version": (C.c_char_p, []),
    "b200zk_device_count": (_i, []),
    "b200zk_init": (_i, [_i, C.POINTER(_vp)]),
    "b200zk_destroy": (None, [_vp]),
    "b200zk_last_error": (C.c_char_p, [_vp]),
    "b200zk_sync": (_i, [_vp]),
    "b200zk_set_stream": (_i, [_vp, _vp]),
    "b200zk_sm_count": (_i, [_vp]),
    "b200zk_dev_alloc": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "b200zk_dev_free": (_i, [_vp, _vp]),
    "b200zk_h2d": (_i, [_vp, _vp, _vp, _sz]),
    "b200zk_d2h": (_i, [_vp, _vp, _vp, _sz]),
    "b200zk_d2d": (_i, [_vp, _vp, _vp, _sz]),
    "b200zk_host_alloc_pinned": (_i, [_sz, C.POINTER(_vp)]),
    "b200zk_host_free_pinned": (_i, [_vp]),
    "b200zk_timer_start": (_i, [_vp]),
    "b200zk_timer_stop": (_i, [_vp, C.POINTER(C.c_float)]),
    "b200zk_bases_upload": (_i, [_vp, _i, _vp, _sz, _sz, _vp, _sz, C.POINTER(_vp)]),
    "b200zk_bases_from_device": (_i, [_vp, _i, _vp, _sz, _vp, C.POINTER(_vp)]),
    "b200zk_bases_precompute": (_i, [_vp, _vp, _i]),
    "b200zk_bases_upload_encoded": (_i, [_vp, _i, _vp, _sz, _i, _i, C.POINTER(_vp)]),
    "b200zk_decode_points": (_i, [_vp, _i, _vp, _sz, _i, _vp, _vp]),
    "b200zk_encode_points": (_i, [_vp, _i, _vp, _vp, _sz, _i, _vp]),
    "b200zk_bases_len": (_sz, [_vp]),
    "b200zk_bases_free": (None, [_vp]),
    "b200zk_multiexp": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _vp]),
    "b200zk_multiexp_dev": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _vp, _vp]),
    "b200zk_multiexp_batch_dev": (_i, [_vp, _vp, _sz, _vp, _sz, _sz, _vp, _sz, C.c_uint32, _vp, _vp]),
    "b200zk_multiexp_async": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, C.POINTER(_vp)]),
    "b200zk_multiexp_sharded_async": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, C.POINTER(_vp)]),
    "b200zk_job_wait": (_i, [_vp, _vp]),
    "b200zk_set_msm_window": (_i, [_vp, _i]),
    "b200zk_sum_points_dev": (_i, [_vp, _i, _vp, _sz, _vp]),
    "b200zk_into_affine": (_i, [_vp, _i, _vp, _sz, _vp, _vp]),
    "b200zk_fixed_base_mul_dev": (_i, [_vp, _i, _vp, _vp, _sz, _u32, _vp, _vp]),
    "b200zk_nccl_unique_id": (_i, [_vp]),
    "b200zk_comm_init": (_i, [_vp, _vp, _i, _i]),
    "b200zk_allgather_sum_dev": (_i, [_vp, _i, _vp, _vp]),
    "b200zk_init_multi": (_i, [_vp, _i, C.POINTER(_vp)]),
    "b200zk_group_destroy": (None, [_vp]),
    "b200zk_group_last_error": (C.c_char_p, [_vp]),
    "b200zk_group_size": (_i, [_vp]),
    "b200zk_group_ctx": (_vp, [_vp, _i]),
    "b200zk_group_peer_access": (_i, [_vp, _i]),
    "b200zk_multi_bases_upload": (_i, [_vp, _i, _vp, _sz, _sz, _vp, _sz, C.POINTER(_vp)]),
    "b200zk_multi_bases_precompute": (_i, [_vp, _vp, _i]),
    "b200zk_multi_bases_len": (_sz, [_vp]),
    "b200zk_multi_bases_free": (None, [_vp]),
    "b200zk_multi_multiexp": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, _vp]),
    "b200zk_multi_multiexp_async": (_i, [_vp, _vp, _sz, _vp, _sz, _vp, C.POINTER(_vp)]),
    "b200zk_multi_job_wait": (_i, [_vp, _vp]),
    "b200zk_multi_h_poly": (_i, [_vp, _vp, _vp, _vp, _u32, _vp]),
    "b200zk_multi_plan": (_i, [_vp, _i, _sz, _vp, _sz, _vp, _vp]),
    "b200zk_ntt_plan": (_i, [_u32, _i, _i, _u32, _vp, _vp, _vp]),
    "b200zk_ntt": (_i, [_vp, _vp, _u32, _i]),
    "b200zk_ntt_dev": (_i, [_vp, _vp, _u32, _i]),
    "b200zk_distribute_powers_dev": (_i, [_vp, _vp, _sz, _vp]),
    "b200zk_fr_spmv_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200zk_field_vec_dev": (_i, [_vp, _i, _i, _vp, _vp, _vp, _sz]),
    "b200zk_field_vec": (_i, [_vp, _i, _i, _vp, _vp, _vp, _sz]),
    "b200zk_divide_by_z_on_coset_dev": (_i, [_vp, _vp, _u32]),
    "b200zk_domain_z": (_i, [_vp, _vp, _u32, _vp]),
    "b200zk_fr_scale_dev": (_i, [_vp, _vp, _sz, _vp]),
    "b200zk_point_op": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _sz]),
    "b200zk_h_poly": (_i, [_vp, _vp, _vp, _vp, _u32, _vp]),
    "b200zk_h_poly_dev": (_i, [_vp, _vp, _vp, _vp, _u32, _vp]),
    "b200zk_crs_create": (_i, [_vp] * 12 + [C.POINTER(_vp)]),
    "b200zk_crs_free": (None, [_vp]),
    "b200zk_parameters_read": (_i, [_vp, _vp, _sz, _i, C.POINTER(_vp)]),
    "b200zk_parameters_size": (_sz, [_vp]),
    "b200zk_parameters_write": (_i, [_vp, _vp, _vp, _sz]),
    "b200zk_crs_verifying_key": (_i, [_vp, _vp, _vp, _vp, C.POINTER(_sz)]),
    "b200zk_crs_query_sizes": (_i, [_vp, _vp]),
    "b200zk_crs_precompute": (_i, [_vp, _vp, _i]),
    "b200zk_groth16_prove": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200zk_groth16_prove_batch": (_i, [_vp, _vp, _vp, _sz, _sz, _sz, _sz, _i, _vp, _vp, _vp, _vp]),
    "b200zk_groth16_prove_batch_bytes": (_i, [_vp, _vp, _vp, _sz, _sz, _sz, _sz, _i, _vp]),
    "b200zk_pairing": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200zk_prepare_verifying_key": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, C.POINTER(_vp)]),
    "b200zk_pvk_free": (None, [_vp]),
    "b200zk_verify_proofs": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _sz, _vp]),
    "b200zk_launch_count": (C.c_ulonglong, [_vp, _i]),
    "b200zk_profile_enable": (_i, [_vp, _i]),
    "b200zk_profile_read": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(_i)]),
    "b200zk_microbench": (_i, [_vp, _i, _i, C.POINTER(C.c_double)]),
}



class ProveInput(C.Structure):
    """b200zk_prove_input (include/b200zk.h): the per-proof pointers of b200zk_groth16_prove_batch"""
    _fields_ = [(name, _vp) for name in ("a", "b", "c", "inputs", "aux", "a_aux_density", "b_input_density", "b_aux_density", "r", "s")]


_lib = None


def load():
    """Load libb200zk.so; raises ImportError when it has not been built (run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build the CUDA library first (make -C zcash-gpu-thesis_b200/csrc). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
