"""ctypes declarations for include/rsgpu.h (one entry per exported symbol)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


class RsgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rsgpu error {code}: {msg}")
        self.code = code


def lib_path():
    return os.path.join(_HERE, "librsgpu.so")


_vp, _sz, _u64, _int = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); mirrors include/rsgpu.h line by line
SIGNATURES = {
    "rsg_last_error": (C.c_char_p, []),
    "rsg_device_count": (_int, []),
    "rsg_context_create": (_int, [_pp, _sz, _sz, _vp, _sz, _sz, _vp, _int]),
    "rsg_context_destroy": (None, [_vp]),
    "rsg_context_sync": (_int, [_vp]),
    "rsg_context_set_stream": (_int, [_vp, _vp]),
    "rsg_context_launch_count": (_u64, [_vp]),
    "rsg_context_stat": (_u64, [_vp, C.c_char_p]),
    "rsg_crs_create": (_int, [_vp, _sz, _pp]),
    "rsg_crs_upload": (_int, [_vp, _sz, _sz, _vp]),
    "rsg_crs_download": (_int, [_vp, _sz, _sz, _vp]),
    "rsg_crs_fill_uniform": (_int, [_vp, _u64]),
    "rsg_crs_fill_uniform_at": (_int, [_vp, _sz, _sz, _u64, _u64]),
    "rsg_crs_device_ptr": (_vp, [_vp]),
    "rsg_crs_destroy": (None, [_vp]),
    "rsg_ringvec_create": (_int, [_vp, _sz, _pp]),
    "rsg_ringvec_upload": (_int, [_vp, _sz, _sz, _vp]),
    "rsg_ringvec_download": (_int, [_vp, _sz, _sz, _vp]),
    "rsg_ringvec_fill_uniform": (_int, [_vp, _u64]),
    "rsg_ringvec_device_ptr": (_vp, [_vp]),
    "rsg_ringvec_size": (_sz, [_vp]),
    "rsg_ringvec_destroy": (None, [_vp]),
    "rsg_ringvec_is_zero_prefix": (_int, [_vp, _sz, _sz, _vp]),
    "rsg_ring_binop": (_int, [_vp, _int, _vp, _sz, _vp, _sz, _vp, _sz, _sz]),
    "rsg_ring_scalar_op": (_int, [_vp, _int, _vp, _sz, _u64, _vp, _sz, _sz]),
    "rsg_ring_negate": (_int, [_vp, _vp, _sz, _vp, _sz, _sz]),
    "rsg_ring_invert": (_int, [_vp, _vp, _sz, _vp, _sz, _sz, _vp]),
    "rsg_inner_product": (_int, [_vp, _vp, _sz, _vp, _sz, _sz, _vp, _vp, _vp, C.POINTER(_sz)]),
    "rsg_inner_product_idx": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp, C.POINTER(_sz)]),
    "rsg_enc_add": (_int, [_vp, _vp, _vp]),
    "rsg_crs_copy": (_int, [_vp, _sz, _vp, _sz, _sz]),
    "rsg_enc_sum": (_int, [_vp, _vp, _sz, _sz, _vp]),
    "rsg_enc_sum_strided": (_int, [_vp, _vp, _sz, _sz, _sz, _vp]),
    "rsg_witness_map": (_int, [_vp, _sz, _vp, _vp, _vp]),
    "rsg_witness_map_zk": (_int, [_vp, _sz, _vp, _vp, _vp, _vp]),
    "rsg_witness_map_r1cs": (_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "rsg_witness_map_groth16": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "rsg_interpolate": (_int, [_vp, _sz, _sz, _vp, _sz, _vp, _sz]),
    "rsg_instance_map": (_int, [_vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "rsg_decode": (_int, [_vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "rsg_encode": (_int, [_vp, _vp, _vp, _sz, _sz, _vp, _vp, _sz]),
    "rsg_enc_file_info": (_int, [C.c_char_p, _vp, _vp, _vp]),
    "rsg_enc_file_write": (_int, [C.c_char_p, C.c_uint64, _sz, _sz, _vp, _sz, _sz, _vp, _sz, _vp]),
    "rsg_enc_file_read": (_int, [C.c_char_p, _sz, _sz, _vp, _sz, _sz, _vp, _sz, _vp, _vp, _vp]),
    "rsg_crs_save": (_int, [_vp, _sz, _sz, C.c_char_p]),
    "rsg_crs_load": (_int, [_vp, _sz, C.c_char_p, _vp]),
    "rsg_vanishing": (_int, [_vp, _sz, _vp]),
    "rsg_r1cs_create": (_int, [_vp, _sz, _sz, _sz, _vp, _vp, _vp, _pp]),
    "rsg_r1cs_destroy": (None, [_vp]),
    "rsg_r1cs_evaluate": (_int, [_vp, _vp, _vp, _vp]),
    "rsg_groth16_prove": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rsg_groth16_prove_refs": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rsg_rinocchio_prove": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rsg_groth16_lincombs": (_int, [_vp, _vp, _vp, _sz, _sz, _vp, _vp, _vp, _vp, _vp]),
    "rsg_groth16_shard_block_words": (_sz, [_sz, _sz]),
    "rsg_groth16_lincombs_shard": (_int, [_vp, _vp, _vp, _sz, _sz, _vp, _vp, _vp, _sz, _vp]),
    "rsg_groth16_shard_check": (_int, [_vp, _sz, _sz, _sz, _u64, C.POINTER(_int)]),
    "rsg_groth16_lincombs_chain": (_int, [_vp, _vp, _sz, _vp, _sz, _sz, _vp, _vp, _vp, _vp]),
    "rsg_groth16_chain_finish": (_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "rsg_exchange_p2p": (_int, [_vp, _vp, _sz, _sz, _sz, _sz, _vp]),
    "rsg_enc_sum_p2p": (_int, [_vp, _vp, _vp, _sz, _sz, _sz, _sz, _vp]),
    "rsg_ringvec_wrap": (_int, [_vp, _vp, _sz, _pp]),
    "rsg_batch_encode": (_int, [_vp, _vp, _sz, _vp]),
    "rsg_plain_to_ntt": (_int, [_vp, _vp, _sz, _vp]),
    "rsg_ntt": (_int, [_vp, _vp, _sz, _int, _sz, _int]),
    "rsg_crs_lincomb": (_int, [_vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "rsg_trace_report": (_sz, [C.c_char_p, _sz, _int]),
    "rsg_context_enable_timing": (_int, [_vp, _int]),
    "rsg_context_last_timing": (_int, [_vp, C.c_char_p, C.POINTER(C.c_float), C.POINTER(_u64)]),
}

_lib = None


def load_library():
    """Load librsgpu.so (built in-tree by __graft_entry__.build()); fail loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RsgError(-2, f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RsgError(rc, load_library().rsg_last_error().decode())
