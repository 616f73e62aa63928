"""Python mirror of the backend objects (thin; all computation happens in librsgpu.so on the GPU).

Names follow the reference: a Context replaces the statics of RingElem/EncodingElem (ringsnark/seal/seal_ring.hpp:25,
218-223), a Crs is a vector<EncodingElem> (groth16.hpp:14-20), a RingVec is a vector<RingElem>.
"""
import ctypes as C

import numpy as np

from .capi import check, load_library

TERM_SKIP, TERM_ONE, TERM_GENERAL = 0, 1, 2


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


class Context:
    def __init__(self, N_R, q, N_E, Q, device=0):
        self.lib = load_library()
        self.N_R, self.N_E = int(N_R), int(N_E)
        self.q, self.Q = _u64(q).copy(), _u64(Q).copy()
        self.L_R, self.L_E = self.q.size, self.Q.size
        self.ring_words = self.L_R * self.N_R
        self.enc_words = self.L_R * 2 * self.L_E * self.N_E
        h = C.c_void_p()
        check(self.lib.rsg_context_create(C.byref(h), self.N_R, self.L_R, _ptr(self.q), self.N_E, self.L_E, _ptr(self.Q), device))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.rsg_context_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.lib.rsg_context_sync(self.h))

    def set_stream(self, cuda_stream):
        check(self.lib.rsg_context_set_stream(self.h, C.c_void_p(cuda_stream)))

    def launch_count(self):
        return int(self.lib.rsg_context_launch_count(self.h))

    def stat(self, name):
        return int(self.lib.rsg_context_stat(self.h, name.encode()))

    def enable_timing(self, on=True):
        check(self.lib.rsg_context_enable_timing(self.h, 1 if on else 0))

    def timing(self, kernel=None):
        ms, n = C.c_float(0), C.c_uint64(0)
        check(self.lib.rsg_context_last_timing(self.h, kernel.encode() if kernel else None, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    # ---- factories
    def crs(self, n):
        return Crs(self, n)

    def ringvec(self, n):
        return RingVec(self, n)

    def crs_from(self, words):
        words = _u64(words).reshape(-1, self.enc_words)
        c = Crs(self, words.shape[0])
        c.upload(words)
        return c

    def ringvec_from(self, words):
        words = _u64(words).reshape(-1, self.ring_words)
        v = RingVec(self, words.shape[0])
        v.upload(words)
        return v

    # ---- hot path (b)
    def term_tags(self, vec, tag=None, scalar=None, first=0, count=None):
        """The reference's per-term dispatch (seal_ring.tcc:390-396, 509-548). `tag`/`scalar` describe elements the
        caller holds as scalars (RingElem variant); poly elements get SealPoly::is_zero's prefix test on the GPU."""
        count = len(vec) - first if count is None else count
        flags = vec.is_zero_prefix(first, count)
        out = np.where(flags != 0, TERM_SKIP, TERM_GENERAL).astype(np.uint8)
        if tag is not None:
            for i in range(count):
                if int(tag[i]) == 0:
                    s = int(scalar[i])
                    out[i] = TERM_SKIP if s == 0 else (TERM_ONE if s == 1 else TERM_GENERAL)
        return out

    def inner_product(self, crs, coeffs, tags, crs_first=0, coeff_first=0, to_host=True, d_out=None):
        """EncodingElem::inner_product (seal_ring.tcc:361-433). Returns (words or None, n_used)."""
        tags = np.ascontiguousarray(tags, dtype=np.uint8)
        out = np.empty(self.enc_words, dtype=np.uint64) if to_host else None
        used = C.c_size_t(0)
        check(self.lib.rsg_inner_product(self.h, crs.h, crs_first, coeffs.h, coeff_first, tags.size, _ptr(tags),
                                         _ptr(out) if to_host else None, C.c_void_p(d_out) if d_out else None, C.byref(used)))
        return out, int(used.value)

    def groth16_lincombs(self, crs, layout, n, n_aux, d_vec, aux_kind=None, to_host=True, d_proof=None):
        """The six inner products + operator+= chain of groth16.tcc:89-112 over the term ranges of `layout` from
        device-resident coefficient vectors: d_vec = 6 device pointers (A_io, A_mid, B_io, B_mid, H, aux), each at the
        FIRST element of its range.  Returns (proof words [3][enc_words] or None, n_used[3])."""
        out = np.empty((3, self.enc_words), dtype=np.uint64) if to_host else None
        used = (C.c_size_t * 3)()
        ptrs = (C.c_void_p * 6)(*[int(p) for p in d_vec])
        if aux_kind is not None:
            aux_kind = np.ascontiguousarray(aux_kind, dtype=np.uint8)
        check(self.lib.rsg_groth16_lincombs(self.h, crs.h, C.byref(layout), n, n_aux, ptrs,
                                            _ptr(aux_kind) if aux_kind is not None else None,
                                            _ptr(out) if to_host else None, C.c_void_p(d_proof) if d_proof else None, used))
        return out, [int(u) for u in used]

    def enc_sum(self, d_parts, parts, n_enc, d_out):
        """Modular sum of `parts` blocks of n_enc encodings (device pointers): runs after the NCCL all-gather."""
        check(self.lib.rsg_enc_sum(self.h, C.c_void_p(d_parts), parts, n_enc, C.c_void_p(d_out)))

    # ---- hot path (a)
    def witness_map(self, n, evals, coeffs=None, H=None):
        coeffs = coeffs or RingVec(self, 6 * n)
        H = H or RingVec(self, n + 1)
        check(self.lib.rsg_witness_map(self.h, n, evals.h, coeffs.h, H.h))
        return coeffs, H

    def interpolate(self, n, y, batch=1, out=None, y_first=0, out_first=0):
        out = out or RingVec(self, batch * n)
        check(self.lib.rsg_interpolate(self.h, n, batch, y.h, y_first, out.h, out_first))
        return out

    def decode(self, sk, enc_words):
        """EncodingElem::decode (seal_ring.tcc:435-477) of host encodings [count][enc_words] under the secret keys
        sk [L_R][L_E][N_E] (NTT form).  Returns (ring words [count][L_R*N_R], noise budgets [count][L_R]); raises RsgError
        (code -6) if some ciphertext has no noise budget left -- the reference's decoding_error."""
        sk, enc = _u64(sk), _u64(enc_words)
        count = enc.size // self.enc_words
        ring = np.zeros((count, self.L_R * self.N_R), dtype=np.uint64)
        budget = np.zeros((count, self.L_R), dtype=np.int32)
        check(self.lib.rsg_decode(self.h, _ptr(sk), None, _ptr(enc), count, _ptr(ring), budget.ctypes.data_as(C.c_void_p)))
        return ring, budget

    def encode(self, sk, elems, seeds, out=None, first=0, count=None, out_first=0):
        """EncodingElem::encode (seal_ring.tcc:324-359) of device ring elements into a CRS arena.  sk: [L_R][L_E][N_E] (NTT
        form); seeds: [count][L_R][8] words (what SEAL's random generator factory would seed each encryption with)."""
        count = len(elems) - first if count is None else count
        out = out or Crs(self, count)
        sk, seeds = _u64(sk), _u64(seeds)
        assert seeds.size == count * self.L_R * 8
        check(self.lib.rsg_encode(self.h, _ptr(sk), elems.h, first, count, _ptr(seeds), out.h, out_first))
        return out

    def vanishing(self, n):
        Z = np.zeros((self.L_R, n + 1), dtype=np.uint64)
        check(self.lib.rsg_vanishing(self.h, n, _ptr(Z)))
        return Z


class Groth16Layout(C.Structure):
    """rsg_groth16_layout (include/rsgpu.h)."""
    _fields_ = [(k, C.c_size_t) for k in (
        "s_pows_off", "s_pows_lo", "s_pows_hi", "delta_ts_off", "delta_ts_lo", "delta_ts_hi",
        "delta_mid_off", "delta_mid_lo", "delta_mid_hi", "alpha_idx", "beta_idx")]


NONE = (1 << 64) - 1
AUX_POLY = 0xFF


class R1cs:
    """Constraint system in CSR form on the device (relations/constraint_satisfaction_problems/r1cs/r1cs.hpp:118-162)."""

    def __init__(self, ctx, n, n_io, n_aux, row_ptr, col, coeff):
        self.ctx, self.n, self.n_io, self.n_aux = ctx, int(n), int(n_io), int(n_aux)
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint32)
        col = np.ascontiguousarray(col, dtype=np.uint32)
        coeff = _u64(coeff)
        assert row_ptr.size == 3 * self.n + 1
        h = C.c_void_p()
        check(ctx.lib.rsg_r1cs_create(ctx.h, self.n, self.n_io, self.n_aux, _ptr(row_ptr), _ptr(col), _ptr(coeff), C.byref(h)))
        self.h = h

    def evaluate(self, assignment, evals=None):
        evals = evals or RingVec(self.ctx, 9 * self.n)
        check(self.ctx.lib.rsg_r1cs_evaluate(self.ctx.h, self.h, assignment.h, evals.h))
        return evals

    def instance_map(self, t, t_first=0):
        """r1cs_to_qrp_instance_map_with_evaluation (r1cs_to_qrp.tcc:75-116) at the ring element t[t_first].
        Returns (ABCt [3*(vars+1)], Ht [n+1], Zt [1]) as device ring vectors."""
        nv1 = self.n_io + self.n_aux + 1
        ABCt, Ht, Zt = RingVec(self.ctx, 3 * nv1), RingVec(self.ctx, self.n + 1), RingVec(self.ctx, 1)
        check(self.ctx.lib.rsg_instance_map(self.ctx.h, self.h, t.h, t_first, ABCt.h, Ht.h, Zt.h))
        return ABCt, Ht, Zt

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.rsg_r1cs_destroy(self.h)
            self.h = None
        except Exception:
            pass


def groth16_shard_layout(n, n_aux, rank=0, world=1):
    """Term sharding of the ringGroth16 CRS across `world` ranks (pure host logic, no GPU needed).
    s_pows[0..n], delta_ts[0..n] and delta_mid[0..n_aux) are each cut into `world` contiguous term ranges; rank r's
    arena is [ s_pows[lo:hi) | delta_ts[lo:hi) | delta_mid[lo:hi) | alpha | beta ]; only rank 0 adds alpha and beta
    (groth16.tcc:95,103).  Returns a dict with the rsg_groth16_layout fields plus n_elems."""
    def shard(total):
        per = (total + world - 1) // world
        return min(total, rank * per), min(total, (rank + 1) * per)

    s_lo, s_hi = shard(n + 1)
    t_lo, t_hi = shard(n + 1)
    m_lo, m_hi = shard(n_aux)
    d = dict(s_pows_off=0, s_pows_lo=s_lo, s_pows_hi=s_hi)
    d.update(delta_ts_off=s_hi - s_lo, delta_ts_lo=t_lo, delta_ts_hi=t_hi)
    d.update(delta_mid_off=d["delta_ts_off"] + (t_hi - t_lo), delta_mid_lo=m_lo, delta_mid_hi=m_hi)
    end = d["delta_mid_off"] + (m_hi - m_lo)
    d.update(alpha_idx=end if rank == 0 else NONE, beta_idx=end + 1 if rank == 0 else NONE, n_elems=end + 2)
    return d


class Groth16ProvingKey:
    """groth16::proving_key (zk_proof_systems/groth16/groth16.hpp:10-47) with every CRS vector in ONE arena:
    [ s_pows[lo:hi) | delta_ts[lo:hi) | delta_mid[lo:hi) | alpha | beta ].  rank/world shard each vector by term."""

    def __init__(self, ctx, r1cs, rank=0, world=1):
        self.ctx, self.r1cs = ctx, r1cs
        d = groth16_shard_layout(r1cs.n, r1cs.n_aux, rank, world)
        L = Groth16Layout()
        for k, _ in Groth16Layout._fields_:
            setattr(L, k, d[k])
        self.layout = L
        self.n_elems = d["n_elems"]
        self.crs = Crs(ctx, self.n_elems)
        self.assignment = RingVec(ctx, r1cs.n_io + r1cs.n_aux)

    def load(self, s_pows, delta_ts, delta_mid, alpha, beta):
        """Full (unsharded) CRS vectors as host words -> this rank's shard of the arena."""
        L = self.layout
        self.crs.upload(s_pows[L.s_pows_lo:L.s_pows_hi], L.s_pows_off)
        self.crs.upload(delta_ts[L.delta_ts_lo:L.delta_ts_hi], L.delta_ts_off)
        if L.delta_mid_hi > L.delta_mid_lo:
            self.crs.upload(delta_mid[L.delta_mid_lo:L.delta_mid_hi], L.delta_mid_off)
        if L.alpha_idx != NONE:
            self.crs.upload(alpha, L.alpha_idx)
            self.crs.upload(beta, L.beta_idx)

    def fill_synthetic(self, seed):
        """Synthetic CRS (uniform residues) that is the SAME key for every (rank, world): each shard range is filled with
        the words the unsharded arena [s_pows (n+1) | delta_ts (n+1) | delta_mid (n_aux) | alpha | beta] would hold."""
        L, n, aux = self.layout, self.r1cs.n, self.r1cs.n_aux
        self.crs.fill_uniform_at(L.s_pows_off, L.s_pows_hi - L.s_pows_lo, L.s_pows_lo, seed)
        self.crs.fill_uniform_at(L.delta_ts_off, L.delta_ts_hi - L.delta_ts_lo, n + 1 + L.delta_ts_lo, seed)
        self.crs.fill_uniform_at(L.delta_mid_off, L.delta_mid_hi - L.delta_mid_lo, 2 * n + 2 + L.delta_mid_lo, seed)
        if L.alpha_idx != NONE:
            self.crs.fill_uniform_at(L.alpha_idx, 1, 2 * n + 2 + aux, seed)
            self.crs.fill_uniform_at(L.beta_idx, 1, 2 * n + 3 + aux, seed)

    def save(self, path):
        """This rank's arena (its shard of every CRS vector, in layout order) -> one file (csrc/serialize.inl)."""
        from . import serialize
        serialize.save_crs(self.crs, path, 0, self.n_elems)

    def load_file(self, path):
        """Arena written by save() for the same (n, n_aux, rank, world) -> HBM."""
        from . import serialize
        got = serialize.load_crs(self.crs, path, 0)
        if got != self.n_elems:
            raise ValueError(f"{path} holds {got} encodings, this key shard needs {self.n_elems}")

    def prove(self, h_assignment=None, aux_kind=None, to_host=True, d_proof=None):
        """groth16::prover (groth16.tcc:69-115). h_assignment: host words [n_io+n_aux][L_R*N_R] (None = use what is
        already resident in self.assignment). Returns (proof words [3][enc_words] or None, n_used[3])."""
        ctx = self.ctx
        out = np.empty((3, ctx.enc_words), dtype=np.uint64) if to_host else None
        used = (C.c_size_t * 3)()
        if h_assignment is not None:
            h_assignment = _u64(h_assignment)
        if aux_kind is not None:
            aux_kind = np.ascontiguousarray(aux_kind, dtype=np.uint8)
        check(ctx.lib.rsg_groth16_prove(ctx.h, self.r1cs.h, self.crs.h, C.byref(self.layout), self.assignment.h,
                                        _ptr(h_assignment) if h_assignment is not None else None,
                                        _ptr(aux_kind) if aux_kind is not None else None,
                                        _ptr(out) if to_host else None, C.c_void_p(d_proof) if d_proof else None, used))
        return out, [int(u) for u in used]


class CrsRef(C.Structure):
    """rsg_crs_ref (include/rsgpu.h): an arena handle and the index of the first encoding of a key vector."""
    _fields_ = [("crs", C.c_void_p), ("first", C.c_size_t)]


def rinocchio_prove(ctx, r1cs, refs, assignment, h_d=None, aux_kind=None, h_assignment=None):
    """rinocchio::prover (rinocchio.tcc:74-190) in one call (rsg_rinocchio_prove).  refs: six (Crs, first) pairs -- s_pows,
    alpha_s_pows, beta_prods, beta_rv_ts, beta_rw_ts, beta_ry_ts; h_d: [3][L_R*N_R] words d1, d2, d3 or None (non-ZK).
    Returns (proof words [9][enc_words], n_used[9])."""
    arr = (CrsRef * 6)()
    for k, (crs, first) in enumerate(refs):
        arr[k].crs, arr[k].first = (crs.h if crs is not None else None), first
    out = np.empty((9, ctx.enc_words), dtype=np.uint64)
    used = (C.c_size_t * 9)()
    if h_d is not None:
        h_d = _u64(h_d)
    if aux_kind is not None:
        aux_kind = np.ascontiguousarray(aux_kind, dtype=np.uint8)
    if h_assignment is not None:
        h_assignment = _u64(h_assignment)
    check(ctx.lib.rsg_rinocchio_prove(ctx.h, r1cs.h, arr, assignment.h, _ptr(h_assignment) if h_assignment is not None else None,
                                      _ptr(aux_kind) if aux_kind is not None else None, _ptr(h_d) if h_d is not None else None,
                                      _ptr(out), None, used))
    return out, [int(u) for u in used]


def groth16_prove_refs(ctx, r1cs, refs, assignment, aux_kind=None, h_assignment=None):
    """groth16::prover over key vectors in different arenas (rsg_groth16_prove_refs): refs = five (Crs, first) pairs."""
    arr = (CrsRef * 5)()
    for k, (crs, first) in enumerate(refs):
        arr[k].crs, arr[k].first = (crs.h if crs is not None else None), first
    out = np.empty((3, ctx.enc_words), dtype=np.uint64)
    used = (C.c_size_t * 3)()
    if aux_kind is not None:
        aux_kind = np.ascontiguousarray(aux_kind, dtype=np.uint8)
    if h_assignment is not None:
        h_assignment = _u64(h_assignment)
    check(ctx.lib.rsg_groth16_prove_refs(ctx.h, r1cs.h, arr, assignment.h, _ptr(h_assignment) if h_assignment is not None else None,
                                         _ptr(aux_kind) if aux_kind is not None else None, _ptr(out), None, used))
    return out, [int(u) for u in used]


class _Arena:
    def __len__(self):
        return self.n


class Crs(_Arena):
    def __init__(self, ctx, n):
        self.ctx, self.n = ctx, int(n)
        h = C.c_void_p()
        check(ctx.lib.rsg_crs_create(ctx.h, self.n, C.byref(h)))
        self.h = h

    def upload(self, words, first=0):
        words = _u64(words).reshape(-1, self.ctx.enc_words)
        check(self.ctx.lib.rsg_crs_upload(self.h, first, words.shape[0], _ptr(words)))

    def download(self, first=0, count=None):
        count = self.n - first if count is None else count
        out = np.empty((count, self.ctx.enc_words), dtype=np.uint64)
        check(self.ctx.lib.rsg_crs_download(self.h, first, count, _ptr(out)))
        return out

    def fill_uniform(self, seed):
        check(self.ctx.lib.rsg_crs_fill_uniform(self.h, seed))

    def fill_uniform_at(self, first, count, virtual_first, seed):
        """Elements [first, first+count) := elements [virtual_first, ...) of an arena filled by fill_uniform(seed)."""
        check(self.ctx.lib.rsg_crs_fill_uniform_at(self.h, first, count, virtual_first, seed))

    def device_ptr(self):
        return int(self.ctx.lib.rsg_crs_device_ptr(self.h) or 0)

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.rsg_crs_destroy(self.h)
            self.h = None
        except Exception:
            pass


class RingVec(_Arena):
    def __init__(self, ctx, n):
        self.ctx, self.n = ctx, int(n)
        h = C.c_void_p()
        check(ctx.lib.rsg_ringvec_create(ctx.h, self.n, C.byref(h)))
        self.h = h

    def upload(self, words, first=0):
        words = _u64(words).reshape(-1, self.ctx.ring_words)
        check(self.ctx.lib.rsg_ringvec_upload(self.h, first, words.shape[0], _ptr(words)))

    def download(self, first=0, count=None):
        count = self.n - first if count is None else count
        out = np.empty((count, self.ctx.ring_words), dtype=np.uint64)
        check(self.ctx.lib.rsg_ringvec_download(self.h, first, count, _ptr(out)))
        return out

    def fill_uniform(self, seed):
        check(self.ctx.lib.rsg_ringvec_fill_uniform(self.h, seed))

    def device_ptr(self):
        return int(self.ctx.lib.rsg_ringvec_device_ptr(self.h) or 0)

    def is_zero_prefix(self, first=0, count=None):
        count = self.n - first if count is None else count
        flags = np.zeros(count, dtype=np.uint8)
        check(self.ctx.lib.rsg_ringvec_is_zero_prefix(self.h, first, count, _ptr(flags)))
        return flags

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.rsg_ringvec_destroy(self.h)
            self.h = None
        except Exception:
            pass
