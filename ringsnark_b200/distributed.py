"""Multi-GPU ringGroth16 prover: one process per GPU, torch.distributed (NCCL over NVLink) for the plumbing.

SURVEY.md section 8(e): the witness map shards by SLOT (every slot of Z_q^N is an independent copy of the computation),
the CRS linear combinations shard by TERM.  Between the two sits the path's one real exchange step: after the witness map
rank r holds slots [r*S, (r+1)*S) of EVERY coefficient, for the lincombs it needs ALL slots of the coefficients of ITS
terms -- an all-to-all (about 10 MB per rank at C4).  Each rank then produces a partial proof (3 encodings) over its term
range; one all-gather of the partials and the modular-add kernel (modular addition is not an NCCL reduction) finish.

    phase 1  rsg_r1cs_evaluate + rsg_witness_map_groth16  on a context with N_R/G slots      (sharded by slot)
    phase 2  all_to_all_single                                                             (slots <-> terms)
    phase 3  rsg_groth16_lincombs                        on the full context, term shard of the CRS
    phase 4  all_gather_into_tensor + rsg_enc_sum_strided, then rsg_groth16_shard_check on the gathered probe blocks

The reference's inner product DROPS a running sum whose c1 polynomial vanishes (seal_ring.tcc:493-504): an order-dependent
rule over the GLOBAL term order.  Every rank therefore ships, behind its partial proof, the running sums of its inner products
at one NTT slot; after the all-gather each rank shifts them by the totals of the ranks before it and checks that no global
prefix (and no operator+= of groth16.tcc:89-112) vanished there -- then the modular sum of the partials IS the reference's
proof.  Otherwise (a structured CRS, e.g. the tiny_transp golden case; never for real encryptions) the ranks run the exact
chain: rank 0 .. G-1 in order extend the six inner products (`chain_step`, the carry travels rank to rank), rank 0 -- which
holds alpha and beta -- applies the operator+= chain (`chain_finish`).

With peer memory (`set_peers`: every rank maps the others' buffers -- torch symmetric memory in bench.py) phases 2 and 4 are
single kernels over NVLink instead of NCCL collectives (csrc/p2p.cuh): `exchange_p2p` writes this rank's slot block of every
coefficient straight into the term owner's buffer, `combine_p2p` is reduce-scatter + all-gather of the modular sum in one
launch; device-side barriers order them.  The NCCL form stays as the fallback and as the reference the tests compare with.

The index bookkeeping (`send_rows`, `unpack`) is pure host logic and is exercised on CPU with gloo in
tests/test_multi_rank_cpu.py; tests/test_gpu_parity.py runs all phases for G simulated ranks on one GPU.
"""
import ctypes as C

import numpy as np

from .backend import Context, Groth16Layout, NONE, R1cs, groth16_shard_layout
from .capi import check

VEC_ROWS = ("A_io", "A_mid", "B_io", "B_mid", "H")


def rows_per_rank(n, world):
    """Terms of s_pows[0..n] / delta_ts[0..n] per rank (the last rank may hold fewer)."""
    return (n + 1 + world - 1) // world


def send_rows(n, world):
    """Row indices into the rank-local witness tensor [7n+2 rows] = [A_io|B_io|C_io|A_mid|B_mid|C_mid (n each) | H (n+1) |
    one zero row], in the order the all-to-all ships them: for every destination d, for every vector of VEC_ROWS, the
    `per` coefficients of d's term range (indices past the end of a vector point at the zero row)."""
    per = rows_per_rank(n, world)
    base = {"A_io": 0, "A_mid": 3 * n, "B_io": n, "B_mid": 4 * n, "H": 6 * n}
    length = {"A_io": n, "A_mid": n, "B_io": n, "B_mid": n, "H": n + 1}
    zero_row = 7 * n + 1
    idx = np.empty((world, len(VEC_ROWS), per), dtype=np.int64)
    for d in range(world):
        for v, name in enumerate(VEC_ROWS):
            for i in range(per):
                k = d * per + i
                idx[d, v, i] = base[name] + k if k < length[name] else zero_row
    return idx.reshape(-1), per


def unpack(recv, world, per, L_R, S):
    """recv: [world (source rank = slot block), 5, per, L_R, S] -> [5, per, L_R, world*S]: full ring elements of this
    rank's terms (element layout [L_R][N_R], poly_arith.h:81-102)."""
    return recv.reshape(world, len(VEC_ROWS), per, L_R, S).permute(1, 2, 3, 0, 4).reshape(len(VEC_ROWS), per, L_R * world * S)


def slot_shard(words, L_R, N_R, rank, world):
    """[k][L_R*N_R] host words -> the rank's slot block [k][L_R*S]."""
    S = N_R // world
    w = np.asarray(words).reshape(-1, L_R, N_R)
    return np.ascontiguousarray(w[:, :, rank * S:(rank + 1) * S]).reshape(w.shape[0], L_R * S)


class ShardedGroth16Prover:
    """One rank of the G-GPU prover.  `exchange(send) -> recv` and `gather(part) -> all_parts` are injected so that the
    same code runs under NCCL (bench.py), and rank-by-rank on one GPU in the tests."""

    def __init__(self, cfg, r1cs_csr, rank, world, device=0, stream=None, alloc=None):
        """alloc (optional): callable(numel) -> int64 CUDA tensor for the three buffers other ranks touch in the peer-memory
        form (coefficients of this rank's terms, partial-proof record, proof) -- e.g. torch symmetric memory."""
        import torch
        self.torch = torch
        self.rank, self.world = rank, world
        self.n, self.io, self.aux = cfg["n"], cfg["io"], cfg["aux"]
        self.N_R, self.L_R = cfg["N_R"], len(cfg["q"])
        assert self.N_R % world == 0, "slot sharding needs N_R divisible by the number of ranks"
        self.S = self.N_R // world
        n = self.n
        self.ctxP = Context(cfg["N_R"], cfg["q"], cfg["N_E"], cfg["Q"], device=device)      # lincombs: full ring elements
        self.ctxW = Context(self.S, cfg["q"], cfg["N_E"], cfg["Q"], device=device)          # witness map: this rank's slots
        if stream is not None:
            self.ctxP.set_stream(stream)
            self.ctxW.set_stream(stream)
        row_ptr, col, coeff = r1cs_csr
        self.r1csW = R1cs(self.ctxW, n, self.io, self.aux, row_ptr, col, coeff)
        d = groth16_shard_layout(n, self.aux, rank, world)
        self.layout = Groth16Layout()
        for k, _ in Groth16Layout._fields_:
            setattr(self.layout, k, d[k])
        self.carry_first = d["n_elems"]
        self.crs = self.ctxP.crs(d["n_elems"] + 6)   # six spare encodings: the chain's carry is staged there
        self.pstride = max(rows_per_rank(n, world), (self.aux + world - 1) // world, 1)
        self.block_words = int(self.ctxP.lib.rsg_groth16_shard_block_words(self.L_R, self.pstride))
        self.part_words = 3 * self.ctxP.enc_words + self.block_words   # what one rank contributes to the all-gather
        dev = torch.device("cuda", device)
        Ws = self.L_R * self.S
        # witness tensor of this slot block: [6n coefficients | n+1 H | zero row]; the library writes into it through wraps
        self.t_wit = torch.zeros(7 * n + 2, Ws, dtype=torch.int64, device=dev)
        self.t_assign = torch.zeros(self.io + self.aux, Ws, dtype=torch.int64, device=dev)
        self.t_evals = torch.zeros(9 * n, Ws, dtype=torch.int64, device=dev)
        idx, self.per = send_rows(n, world)
        self.idx = torch.from_numpy(idx).to(dev)
        self.m_lo, self.m_hi = d["delta_mid_lo"], d["delta_mid_hi"]
        self.t_aux = torch.zeros(max(self.m_hi - self.m_lo, 1), self.L_R * self.N_R, dtype=torch.int64, device=dev)
        alloc = alloc or (lambda numel: torch.zeros(numel, dtype=torch.int64, device=dev))
        self.t_part = alloc(self.part_words)
        self.t_final = alloc(3 * self.ctxP.enc_words)
        self.t_full_raw = alloc(len(VEC_ROWS) * self.per * self.L_R * self.N_R)
        self.t_full = self.t_full_raw.view(len(VEC_ROWS), self.per, self.L_R * self.N_R)
        self.t_blocks = torch.zeros(world * self.block_words, dtype=torch.int64, device=dev)
        self._peers = None
        self._wraps = []
        self._h_blocks = None
        self.rv_assign = self._wrap(self.ctxW, self.t_assign, self.io + self.aux)
        self.rv_evals = self._wrap(self.ctxW, self.t_evals, 9 * n)
        self.rv_coeffs = self._wrap(self.ctxW, self.t_wit, 6 * n)
        self.rv_H = self._wrap(self.ctxW, self.t_wit[6 * n:], n + 1)

    def _wrap(self, ctx, tensor, n_elems):
        h = C.c_void_p()
        check(ctx.lib.rsg_ringvec_wrap(ctx.h, C.c_void_p(tensor.data_ptr()), n_elems, C.byref(h)))
        self._wraps.append((ctx, h))
        return h

    def load_assignment(self, words, non_blocking=False):
        """Host words of the FULL assignment [io+aux][L_R*N_R] (numpy, or pre-sharded pinned torch tensors via
        load_assignment_shards)."""
        torch = self.torch
        shard = torch.from_numpy(slot_shard(words, self.L_R, self.N_R, self.rank, self.world).view(np.int64))
        aux = torch.from_numpy(np.ascontiguousarray(np.asarray(words)[self.io + self.m_lo:self.io + self.m_hi]).view(np.int64))
        self.load_assignment_shards(shard, aux, non_blocking)

    def load_assignment_shards(self, shard, aux, non_blocking=True):
        self.t_assign.copy_(shard, non_blocking=non_blocking)
        if self.m_hi > self.m_lo:
            self.t_aux[:self.m_hi - self.m_lo].copy_(aux, non_blocking=non_blocking)

    def witness_phase(self):
        """Phase 1 + the pack of phase 2: returns the send buffer [world * 5 * per, L_R*S]."""
        lib, ctx = self.ctxW.lib, self.ctxW
        check(lib.rsg_r1cs_evaluate(ctx.h, self.r1csW.h, self.rv_assign, self.rv_evals))
        check(lib.rsg_witness_map_groth16(ctx.h, self.r1csW.h, self.rv_evals, self.rv_coeffs, self.rv_H))
        return self.t_wit.index_select(0, self.idx)

    def fill_synthetic(self, seed):
        """Synthetic CRS that is the SAME key for every world size (backend.Groth16ProvingKey.fill_synthetic)."""
        L, n, aux = self.layout, self.n, self.aux
        self.crs.fill_uniform_at(L.s_pows_off, L.s_pows_hi - L.s_pows_lo, L.s_pows_lo, seed)
        self.crs.fill_uniform_at(L.delta_ts_off, L.delta_ts_hi - L.delta_ts_lo, n + 1 + L.delta_ts_lo, seed)
        self.crs.fill_uniform_at(L.delta_mid_off, L.delta_mid_hi - L.delta_mid_lo, 2 * n + 2 + L.delta_mid_lo, seed)
        if L.alpha_idx != NONE:
            self.crs.fill_uniform_at(L.alpha_idx, 1, 2 * n + 2 + aux, seed)
            self.crs.fill_uniform_at(L.beta_idx, 1, 2 * n + 3 + aux, seed)

    # ---- peer-memory form of phases 2 and 4
    def set_peers(self, full_ptrs, part_ptrs, final_ptrs, barrier):
        """Device pointers (rank order) of every rank's t_full / t_part / t_final as mapped into THIS process, and a callable
        that enqueues a device-side barrier over all ranks on the current stream."""
        G = self.world
        assert len(full_ptrs) == len(part_ptrs) == len(final_ptrs) == G
        self._peers = ((C.c_void_p * G)(*full_ptrs), (C.c_void_p * G)(*part_ptrs), (C.c_void_p * G)(*final_ptrs), barrier)

    def witness_phase_p2p(self):
        """Phase 1, then phase 2 as ONE kernel: this rank's slots of every coefficient go straight into the term owners' t_full."""
        lib, ctx = self.ctxW.lib, self.ctxW
        check(lib.rsg_r1cs_evaluate(ctx.h, self.r1csW.h, self.rv_assign, self.rv_evals))
        check(lib.rsg_witness_map_groth16(ctx.h, self.r1csW.h, self.rv_evals, self.rv_coeffs, self.rv_H))
        check(lib.rsg_exchange_p2p(ctx.h, C.c_void_p(self.t_wit.data_ptr()), self.n, self.world, self.rank, self.per, self._peers[0]))
        self._peers[3]()     # every rank's coefficients have arrived

    def combine_p2p(self):
        """Phase 4 as ONE kernel (after lincomb_phase): slice sums over peer loads, stored into every rank's t_final; the probe
        blocks come along for the global transparent-prefix check.  Returns the verdict (see combine)."""
        _, parts, finals, barrier = self._peers
        barrier()            # every rank's partial proof is complete
        check(self.ctxP.lib.rsg_enc_sum_p2p(self.ctxP.h, parts, finals, self.world, self.rank, 3, self.block_words,
                                            C.c_void_p(self.t_blocks.data_ptr())))
        barrier()            # every slice has landed everywhere (and nobody still reads the partials / coefficient buffers)
        if self._h_blocks is None:
            self._h_blocks = self.torch.empty(self.world, self.block_words, dtype=self.torch.int64).pin_memory()
        self._h_blocks.view(-1).copy_(self.t_blocks, non_blocking=True)
        self.torch.cuda.current_stream().synchronize()
        verdict = C.c_int(0)
        check(self.ctxP.lib.rsg_groth16_shard_check(self._h_blocks.numpy().view(np.uint64).ctypes.data_as(C.c_void_p), self.world,
                                                    self.L_R, self.pstride, int(self.ctxP.Q[0]), C.byref(verdict)))
        return verdict.value

    def lincomb_phase(self, recv=None, h_proof_ptr=None, aux_kind=None):
        """Phase 3 from the received buffer [world, 5, per, L_R, S] (NCCL form) or, with recv=None, from t_full (peer-memory
        form); leaves the partial proof and its probe block in t_part.
        aux_kind (nullable): RSG_TERM_* / RSG_AUX_POLY per ABSOLUTE auxiliary index, as rsg_groth16_prove takes it."""
        full = unpack(recv, self.world, self.per, self.L_R, self.S).contiguous() if recv is not None else self.t_full
        self._full = full   # keep alive until the kernels have run
        ptrs = (C.c_void_p * 6)(*[full[v].data_ptr() for v in range(5)], self.t_aux.data_ptr())
        used = (C.c_size_t * 3)()
        if aux_kind is not None:
            aux_kind = np.ascontiguousarray(aux_kind, dtype=np.uint8)
            assert aux_kind.size == self.aux
        self._aux_kind = aux_kind
        self._ptrs = ptrs
        check(self.ctxP.lib.rsg_groth16_lincombs_shard(self.ctxP.h, self.crs.h, C.byref(self.layout), self.n, self.aux, ptrs,
                                                       aux_kind.ctypes.data_as(C.c_void_p) if aux_kind is not None else None,
                                                       C.c_void_p(self.t_part.data_ptr()), self.pstride, used))
        return [int(u) for u in used]

    def combine(self, all_parts):
        """Phase 4 after the all-gather of the `world` records [partial proof | probe block]: modular sum of the partial
        proofs into t_final (asynchronous), then the global transparent-prefix check on the probe blocks.  Returns the
        verdict: 0 = t_final is the proof; 1 = run the chain (chain_step on ranks 0..G-1 in order, chain_finish on rank 0)."""
        E3 = 3 * self.ctxP.enc_words
        check(self.ctxP.lib.rsg_enc_sum_strided(self.ctxP.h, C.c_void_p(all_parts.data_ptr()), self.world, 3, self.part_words,
                                                C.c_void_p(self.t_final.data_ptr())))
        if self._h_blocks is None:
            self._h_blocks = self.torch.empty(self.world, self.block_words, dtype=self.torch.int64).pin_memory()
        self._h_blocks.copy_(all_parts.view(self.world, self.part_words)[:, E3:], non_blocking=True)   # one strided D2H copy
        self.torch.cuda.current_stream().synchronize()
        blocks = self._h_blocks.numpy().view(np.uint64)
        verdict = C.c_int(0)
        check(self.ctxP.lib.rsg_groth16_shard_check(blocks.ctypes.data_as(C.c_void_p), self.world, self.L_R, self.pstride,
                                                    int(self.ctxP.Q[0]), C.byref(verdict)))
        return verdict.value

    def new_carry(self):
        """The chain's state before rank 0: six empty inner products."""
        return self.torch.zeros(6 * self.ctxP.enc_words, dtype=self.torch.int64, device=self.t_part.device), np.zeros(6, dtype=np.uint8)

    def chain_step(self, carry, present):
        """Exact continuation of the six inner products over this rank's terms (after lincomb_phase of the same proof);
        carry / present are updated in place and travel to the next rank."""
        ak = self._aux_kind
        check(self.ctxP.lib.rsg_groth16_lincombs_chain(self.ctxP.h, self.crs.h, self.carry_first, C.byref(self.layout), self.n,
                                                       self.aux, self._ptrs, ak.ctypes.data_as(C.c_void_p) if ak is not None else None,
                                                       C.c_void_p(carry.data_ptr()), present.ctypes.data_as(C.c_void_p)))

    def chain_finish(self, carry, present):
        """Rank 0 (holds alpha and beta): operator+= chain of groth16.tcc:89-112 over the complete inner products -> t_final."""
        check(self.ctxP.lib.rsg_groth16_chain_finish(self.ctxP.h, self.crs.h, C.byref(self.layout), C.c_void_p(carry.data_ptr()),
                                                     present.ctypes.data_as(C.c_void_p), C.c_void_p(self.t_final.data_ptr())))

    def close(self):
        for ctx, h in self._wraps:
            ctx.lib.rsg_ringvec_destroy(h)
        self._wraps = []
        # device objects first, then their contexts (their __del__ skips the free once the context handle is gone)
        for obj in (self.r1csW, self.crs):
            obj.__del__()
        self.ctxW.close()
        self.ctxP.close()


def run_chain(sp, dist):
    """The exact chain over NCCL ranks (torch.distributed point-to-point): the carry visits rank 0 .. G-1 and returns to
    rank 0, which finishes and broadcasts the proof.  Collective: every rank calls it after combine() returned 1."""
    torch = sp.torch
    carry, present = sp.new_carry()
    flags = torch.zeros(6, dtype=torch.uint8, device=carry.device)
    if sp.rank > 0:
        dist.recv(carry, src=sp.rank - 1)
        dist.recv(flags, src=sp.rank - 1)
        present[:] = flags.cpu().numpy()
    sp.chain_step(carry, present)
    flags.copy_(torch.from_numpy(present))
    if sp.world > 1:
        dst = sp.rank + 1 if sp.rank + 1 < sp.world else 0
        if sp.rank == 0:
            dist.send(carry, dst=dst)
            dist.send(flags, dst=dst)
            dist.recv(carry, src=sp.world - 1)
            dist.recv(flags, src=sp.world - 1)
            present[:] = flags.cpu().numpy()
        else:
            dist.send(carry, dst=dst)
            dist.send(flags, dst=dst)
    if sp.rank == 0:
        sp.chain_finish(carry, present)
    dist.broadcast(sp.t_final, src=0)
