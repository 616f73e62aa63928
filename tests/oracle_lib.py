"""ctypes binding of oracle/librs_oracle.so (the CPU restatement). TEST INFRASTRUCTURE: import from tests only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def _load():
    so = os.path.join(_ORACLE_DIR, "librs_oracle.so")
    src = os.path.join(_ORACLE_DIR, "rs_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "librs_oracle.so"], stdout=subprocess.DEVNULL)
    lib = C.CDLL(so)
    sz, u64 = C.c_size_t, C.c_uint64
    lib.ro_minimal_primitive_root.restype = u64
    lib.ro_minimal_primitive_root.argtypes = [u64, u64]
    lib.ro_ntt_forward.argtypes = [_u64p, sz, u64]
    lib.ro_ntt_inverse.argtypes = [_u64p, sz, u64]
    lib.ro_batch_index_map.argtypes = [sz, _u64p]
    lib.ro_batch_encode.argtypes = [_u64p, sz, sz, u64, _u64p]
    lib.ro_plain_lift_ntt.argtypes = [_u64p, sz, u64, _u64p, sz, _u64p]
    lib.ro_is_zero_quirk.restype = C.c_int
    lib.ro_is_zero_quirk.argtypes = [_u64p, sz]
    lib.ro_is_equal_quirk.restype = C.c_int
    lib.ro_is_equal_quirk.argtypes = [_u64p, _u64p, sz]
    lib.ro_inner_product.restype = sz
    lib.ro_inner_product.argtypes = [_u64p, _u64p, _u8p, sz, sz, sz, _u64p, sz, sz, _u64p, _u64p]
    lib.ro_enc_add.argtypes = [_u64p, _u64p, sz, sz, sz, _u64p]
    lib.ro_vanishing.argtypes = [sz, u64, _u64p]
    lib.ro_interpolate.argtypes = [sz, _u64p, sz, sz, _u64p, _u64p]
    lib.ro_witness_H.argtypes = [sz, _u64p, _u64p, _u64p, sz, sz, sz, sz, sz, _u64p, _u64p, C.POINTER(sz)]
    return lib


lib = _load()
TAG_SKIP, TAG_ONE, TAG_GENERAL = 0, 1, 2


def c(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def minimal_primitive_root(degree, p):
    return int(lib.ro_minimal_primitive_root(degree, p))


def ntt_forward(a, p):
    a = c(a).copy(); lib.ro_ntt_forward(a, a.size, int(p)); return a


def ntt_inverse(a, p):
    a = c(a).copy(); lib.ro_ntt_inverse(a, a.size, int(p)); return a


def batch_index_map(N):
    m = np.zeros(N, dtype=np.uint64); lib.ro_batch_index_map(N, m); return m


def batch_encode(vals, N_E, t):
    out = np.zeros(N_E, dtype=np.uint64); vals = c(vals)
    lib.ro_batch_encode(vals, vals.size, N_E, int(t), out); return out


def plain_lift_ntt(plain, t, Q):
    plain, Q = c(plain), c(Q)
    out = np.zeros(Q.size * plain.size, dtype=np.uint64)
    lib.ro_plain_lift_ntt(plain, plain.size, int(t), Q, Q.size, out); return out.reshape(Q.size, plain.size)


def is_zero_quirk(words):
    words = c(words); return bool(lib.ro_is_zero_quirk(words, words.size))


def term_tags(words, tag, scalar):
    """Reference dispatch of one coefficient vector: seal_ring.tcc:390-396 (is_zero) and :525-528 (scalar one)."""
    out = np.zeros(len(tag), dtype=np.uint8)
    for i in range(len(tag)):
        if int(tag[i]) == 0:
            out[i] = TAG_SKIP if int(scalar[i]) == 0 else (TAG_ONE if int(scalar[i]) == 1 else TAG_GENERAL)
        else:
            out[i] = TAG_SKIP if is_zero_quirk(words[i]) else TAG_GENERAL
    return out


def inner_product(crs, coeff, tags, N_R, L_R, q, N_E, L_E, Q):
    crs, coeff, q, Q = c(crs), c(coeff), c(q), c(Q)
    tags = np.ascontiguousarray(tags, dtype=np.uint8)
    out = np.zeros(L_R * 2 * L_E * N_E, dtype=np.uint64)
    used = lib.ro_inner_product(crs, coeff, tags, len(tags), N_R, L_R, q, N_E, L_E, Q, out)
    return out, int(used)


def enc_add(acc, other, L_R, N_E, L_E, Q):
    acc = c(acc).copy(); lib.ro_enc_add(acc, c(other), L_R, N_E, L_E, c(Q)); return acc


def vanishing(n, p):
    Z = np.zeros(n + 1, dtype=np.uint64); lib.ro_vanishing(n, int(p), Z); return Z


def interpolate(y, N_R, L_R, q):
    y = c(y); n = y.shape[0]
    out = np.zeros_like(y); lib.ro_interpolate(n, y, N_R, L_R, c(q), out); return out


def witness_H(aA, aB, aC, N_R, L_R, q, lens=None):
    aA, aB, aC = c(aA), c(aB), c(aC); n = aA.shape[0]
    lens = lens or (n, n, n)
    H = np.zeros((max(n - 1, 0), N_R * L_R), dtype=np.uint64)
    hl = C.c_size_t(0)
    lib.ro_witness_H(n, aA, aB, aC, lens[0], lens[1], lens[2], N_R, L_R, c(q), H, C.byref(hl))
    return H, int(hl.value)


def instance_map(n, n_vars, row_ptr, col, coeff, t, N_R, L_R, q):
    """r1cs_to_qrp_instance_map_with_evaluation (r1cs_to_qrp.tcc:75-116): returns (ABCt [3][n_vars+1][W], Ht [n+1][W], Zt [W])."""
    nv1 = n_vars + 1
    W = N_R * L_R
    ABCt = np.zeros((3 * nv1, W), dtype=np.uint64); Ht = np.zeros((n + 1, W), dtype=np.uint64); Zt = np.zeros(W, dtype=np.uint64)
    rp = np.ascontiguousarray(row_ptr, dtype=np.uint32); cl = np.ascontiguousarray(col, dtype=np.uint32)
    fn = lib.ro_instance_map
    fn.restype = C.c_int
    fn.argtypes = [C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                   C.c_void_p, C.c_void_p, C.c_void_p]
    cf, tt, qq = c(coeff), c(t), c(q)
    rc = fn(n, nv1, rp.ctypes.data, cl.ctypes.data, cf.ctypes.data, tt.ctypes.data, N_R, L_R, qq.ctypes.data,
            ABCt.ctypes.data, Ht.ctypes.data, Zt.ctypes.data)
    if rc:
        raise ValueError("t hits a domain point in some slot")
    return ABCt, Ht, Zt


def decode(enc, sk, N_R, L_R, q, N_E, L_E, Q):
    """EncodingElem::decode (seal_ring.tcc:435-477) of one encoding [L_R][2][L_E][N_E] with the secret keys
    [L_R][L_E][N_E]: returns (ring words [L_R*N_R], budgets [L_R])."""
    enc, sk, Q = c(enc).reshape(L_R, 2 * L_E * N_E), c(sk).reshape(L_R, L_E * N_E), c(Q)
    fn = lib.ro_decode_limb
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_uint64, C.c_size_t, C.c_void_p]
    out = np.zeros((L_R, N_R), dtype=np.uint64)
    budgets = []
    for j in range(L_R):
        e, s = np.ascontiguousarray(enc[j]), np.ascontiguousarray(sk[j])
        budgets.append(int(fn(e.ctypes.data, s.ctypes.data, N_E, L_E, Q.ctypes.data, int(q[j]), N_R, out[j].ctypes.data)))
    return out.reshape(-1), budgets


def prng_bytes(seed, off, n):
    """Bytes [off, off+n) of SEAL's Blake2xbPRNG stream for the 8-word seed (randomgen.cpp:201-211, util/blake2xb.c)."""
    seed = c(seed)
    out = np.zeros(n, dtype=np.uint8)
    lib.ro_prng_bytes.argtypes = [C.c_void_p, C.c_uint64, C.c_size_t, C.c_void_p]
    lib.ro_prng_bytes(seed.ctypes.data, off, n, out.ctypes.data)
    return out


def encode(ring, sk, seeds, N_R, L_R, q, N_E, L_E, Q):
    """EncodingElem::encode (seal_ring.tcc:324-359) of one ring element [L_R*N_R] under the secret keys [L_R][L_E][N_E];
    seeds: [L_R][8] words, the seed every PRNG of ring limb j's context starts from.  Returns [L_R][2][L_E][N_E] words."""
    ring, sk, Q, seeds = c(ring).reshape(L_R, N_R), c(sk).reshape(L_R, L_E * N_E), c(Q), c(seeds).reshape(L_R, 8)
    fn = lib.ro_encrypt_limb
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    out = np.zeros((L_R, 2 * L_E * N_E), dtype=np.uint64)
    for j in range(L_R):
        r, s, sd = np.ascontiguousarray(ring[j]), np.ascontiguousarray(sk[j]), np.ascontiguousarray(seeds[j])
        fn(r.ctypes.data, N_R, int(q[j]), N_E, L_E, Q.ctypes.data, s.ctypes.data, sd.ctypes.data, out[j].ctypes.data)
    return out.reshape(-1)
