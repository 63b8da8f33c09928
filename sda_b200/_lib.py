"""ctypes binding of libsda_b200.so (include/sda_b200.h).

The library is the product; this module only declares its prototypes.  There is no Python or
CPU implementation behind it: if the shared object is missing the import fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SDA_B200_LIB: developer override to A/B another build of the same library
LIB_PATH = os.environ.get("SDA_B200_LIB") or os.path.join(_HERE, "libsda_b200.so")

SDA_OK, SDA_ERR_INVALID, SDA_ERR_CUDA, SDA_ERR_NCCL, SDA_ERR_UNSUPPORTED, SDA_ERR_REJECTED = 0, 1, 2, 3, 4, 5
SHARING_ADDITIVE, SHARING_PACKED_SHAMIR = 0, 1
MASK_NONE, MASK_FULL, MASK_CHACHA = 0, 1, 2
PACKED_PATH_AUTO, PACKED_PATH_CUDA_CORES, PACKED_PATH_TENSOR_CORES, PACKED_PATH_TENSOR_CORES_V1 = 0, 1, 2, 3
PACKED_PATH_TENSOR_CORES_ANY_SHAPE = 4


class sda_sharing_scheme(C.Structure):
    _fields_ = [("kind", C.c_int32), ("share_count", C.c_uint64), ("secret_count", C.c_uint64),
                ("privacy_threshold", C.c_uint64), ("modulus", C.c_int64),
                ("omega_secrets", C.c_int64), ("omega_shares", C.c_int64)]


class sda_masking_scheme(C.Structure):
    _fields_ = [("kind", C.c_int32), ("modulus", C.c_int64), ("dimension", C.c_uint64),
                ("seed_bitsize", C.c_uint64)]


_vp, _sz, _i64, _u64, _int = C.c_void_p, C.c_size_t, C.c_int64, C.c_uint64, C.c_int
_ss, _ms = C.POINTER(sda_sharing_scheme), C.POINTER(sda_masking_scheme)
_psz = C.POINTER(C.c_size_t)

# name -> (restype, argtypes); must list every symbol include/sda_b200.h declares
PROTOTYPES = {
    "sda_abi_version": (_int, []),
    "sda_ctx_create": (_int, [_int, C.POINTER(_vp)]),
    "sda_ctx_destroy": (None, [_vp]),
    "sda_last_error": (C.c_char_p, [_vp]),
    "sda_ctx_set_rng_rounds": (_int, [_vp, _int]),
    "sda_ctx_get_rng_rounds": (_int, [_vp]),
    "sda_ctx_set_packed_path": (_int, [_vp, _int]),
    "sda_ctx_set_stream": (_int, [_vp, _vp]),
    "sda_ctx_get_stream": (_vp, [_vp]),
    "sda_ctx_synchronize": (_int, [_vp]),
    "sda_ctx_set_deferred_checks": (_int, [_vp, _int]),
    "sda_ctx_launch_count": (_u64, [_vp]),
    "sda_ctx_last_kernel": (C.c_char_p, [_vp]),
    "sda_host_alloc": (_int, [_vp, _sz, C.POINTER(_vp)]),
    "sda_host_free": (_int, [_vp, _vp]),
    "sda_input_size": (_sz, [_ss]),
    "sda_output_size": (_sz, [_ss]),
    "sda_privacy_threshold": (_sz, [_ss]),
    "sda_reconstruction_threshold": (_sz, [_ss]),
    "sda_share_batches": (_sz, [_ss, _sz]),
    "sda_mask_len": (_sz, [_ms, _sz]),
    "sda_sharing_scheme_validate": (_int, [_vp, _ss]),
    "sda_packed_share_matrix": (_int, [_vp, _ss, _vp]),
    "sda_packed_reconstruct_matrix": (_int, [_vp, _ss, _vp, _sz, _vp]),
    "sda_share_generate": (_int, [_vp, _ss, _vp, _sz, _vp, _vp]),
    "sda_share_combine": (_int, [_vp, _ss, _vp, _sz, _sz, _vp]),
    "sda_share_combine_rows": (_int, [_vp, _ss, _vp, _vp, _sz, _vp, _psz]),
    "sda_secret_reconstruct": (_int, [_vp, _ss, _sz, _vp, _vp, _sz, _sz, _vp, _psz]),
    "sda_secret_reconstruct_rows": (_int, [_vp, _ss, _sz, _vp, _vp, _vp, _sz, _vp, _psz]),
    "sda_mask": (_int, [_vp, _ms, _vp, _sz, _vp, _vp, _psz, _vp]),
    "sda_mask_combine": (_int, [_vp, _ms, _vp, _sz, _sz, _vp, _psz]),
    "sda_unmask": (_int, [_vp, _ms, _vp, _sz, _vp, _sz, _vp]),
    "sda_share_generate_dev": (_int, [_vp, _ss, _vp, _sz, _sz, _sz, _vp, _vp]),
    "sda_mask_share_generate": (_int, [_vp, _ms, _ss, _vp, _sz, _vp, _vp, _vp, _vp]),
    "sda_mask_share_generate_dev": (_int, [_vp, _ms, _ss, _vp, _sz, _sz, _sz, _vp, _vp, _vp, _vp]),
    "sda_share_combine_dev": (_int, [_vp, _ss, _vp, _sz, _sz, _sz, _vp, _vp]),
    "sda_share_generate_combine_dev": (_int, [_vp, _ss, _vp, _sz, _sz, _sz, _vp, _vp, _vp]),
    "sda_mod_reduce_dev": (_int, [_vp, _i64, _vp, _sz, _vp]),
    "sda_mod_reduce_u64_dev": (_int, [_vp, _i64, _vp, _sz, _vp]),
    "sda_secret_reconstruct_dev": (_int, [_vp, _ss, _sz, _vp, _vp, _sz, _sz, _sz, _vp]),
    "sda_mask_dev": (_int, [_vp, _ms, _vp, _sz, _vp, _vp, _vp]),
    "sda_mask_combine_dev": (_int, [_vp, _ms, _vp, _sz, _sz, _vp]),
    "sda_unmask_dev": (_int, [_vp, _ms, _vp, _vp, _sz, _vp]),
    "sda_nccl_unique_id": (_int, [_vp]),
    "sda_ctx_comm_init_rank": (_int, [_vp, _vp, _int, _int]),
    "sda_ctx_comm_rank": (_int, [_vp]),
    "sda_ctx_comm_size": (_int, [_vp]),
    "sda_partial_sums_reduce_dev": (_int, [_vp, _i64, _vp, _sz, _int]),
    "sda_share_combine_ranks_dev": (_int, [_vp, _ss, _vp, _sz, _sz, _sz, _vp, _vp, _int]),
    "sda_ctx_create_multi": (_int, [C.POINTER(_int), _int, C.POINTER(_vp)]),
    "sda_ctx_multi_count": (_int, [_vp]),
    "sda_ctx_multi_member": (_vp, [_vp, _int]),
    "sda_share_combine_multi_dev": (_int, [_vp, _ss, _vp, _sz, _vp, _sz, _vp, _vp]),
    "sda_share_combine_rows_multi": (_int, [_vp, _ss, _vp, _vp, _sz, _vp, _psz]),
    "sda_varint_max_bytes": (_sz, [_sz]),
    "sda_varint_encode": (_int, [_vp, _vp, _sz, _vp, _psz]),
    "sda_varint_decode": (_int, [_vp, _vp, _sz, _vp, _sz, _psz]),
    "sda_varint_encode_dev": (_int, [_vp, _vp, _sz, _vp, _psz]),
    "sda_varint_decode_dev": (_int, [_vp, _vp, _sz, _vp, _sz, _psz]),
    "sda_snapshot_transpose_dev": (_int, [_vp, _vp, _vp, _sz, _sz, _vp, _vp]),
    "sda_fixed_encode_dev": (_int, [_vp, _i64, _int, _vp, _sz, _vp]),
    "sda_fixed_decode_dev": (_int, [_vp, _i64, _int, _u64, _vp, _sz, _vp]),
    "sda_fixed_encode_mask_dev": (_int, [_vp, _ms, _i64, _int, _vp, _sz, _sz, _sz, _vp, _vp, _vp, _sz]),
    "sda_synth_fill_dev": (_int, [_vp, C.c_uint32, _i64, _u64, _sz, _vp]),
}

_lib = None


def load():
    """dlopen libsda_b200.so and type its entry points.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `make -C sda_b200/csrc` (or "
                "`python -c 'import __graft_entry__ as g; g.build()'`). sda_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)          # AttributeError if the .so does not export it
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib
