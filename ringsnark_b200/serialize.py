"""Files of encodings (CRS / proving-key ranges, proofs): thin mirror of the rsg_enc_file_* / rsg_crs_save / rsg_crs_load
entry points of include/rsgpu.h.  The container is specified in csrc/serialize.inl; the reference itself declares
proving-key and proof stream operators (zk_proof_systems/r1cs_ppzksnark.hpp:43-47,142-146) without defining them."""
import ctypes as C

import numpy as np

from .capi import check, load_library

FILE_CRS, FILE_PROOF = 1, 2


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def file_info(path):
    """Header of a file: dict(kind, N_R, L_R, N_E, L_E, n_elems, q, Q).  Host only."""
    lib = load_library()
    info, q, Q = np.zeros(6, dtype=np.uint64), np.zeros(8, dtype=np.uint64), np.zeros(16, dtype=np.uint64)
    check(lib.rsg_enc_file_info(str(path).encode(), info.ctypes.data, q.ctypes.data, Q.ctypes.data))
    kind, N_R, L_R, N_E, L_E, n = (int(x) for x in info)
    return dict(kind=kind, N_R=N_R, L_R=L_R, N_E=N_E, L_E=L_E, n_elems=n, q=[int(x) for x in q[:L_R]], Q=[int(x) for x in Q[:L_E]])


def write_encodings(path, words, N_R, q, N_E, Q, kind=FILE_PROOF):
    """words: [n_elems][L_R*2*L_E*N_E] canonical residues (a proof is three encodings).  Host only."""
    lib = load_library()
    q, Q, w = _u64(q), _u64(Q), _u64(words)
    per = len(q) * 2 * len(Q) * N_E
    assert w.size % per == 0, "words do not hold whole encodings"
    check(lib.rsg_enc_file_write(str(path).encode(), kind, N_R, len(q), q.ctypes.data, N_E, len(Q), Q.ctypes.data, w.size // per,
                                 w.ctypes.data))


def read_encodings(path, N_R, q, N_E, Q):
    """Returns (words [n_elems][enc_words], kind); raises RsgError on foreign parameters, corruption or truncation."""
    lib = load_library()
    n = file_info(path)["n_elems"]
    q, Q = _u64(q), _u64(Q)
    per = len(q) * 2 * len(Q) * N_E
    out = np.zeros((max(n, 1), per), dtype=np.uint64)
    got, kind = C.c_size_t(0), C.c_uint64(0)
    check(lib.rsg_enc_file_read(str(path).encode(), N_R, len(q), q.ctypes.data, N_E, len(Q), Q.ctypes.data, n, out.ctypes.data,
                                C.byref(got), C.byref(kind)))
    return out[:got.value], int(kind.value)


def save_crs(crs, path, first=0, count=None):
    """HBM arena range -> file (streams through pinned memory)."""
    count = len(crs) - first if count is None else count
    check(crs.ctx.lib.rsg_crs_save(crs.h, first, count, str(path).encode()))


def load_crs(crs, path, first=0):
    """File -> HBM arena starting at element `first`; returns the number of encodings loaded."""
    n = C.c_size_t(0)
    check(crs.ctx.lib.rsg_crs_load(crs.h, first, str(path).encode(), C.byref(n)))
    return int(n.value)
