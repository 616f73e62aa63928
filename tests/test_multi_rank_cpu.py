"""world_size-2 (and 3) host-side logic of the N-GPU path on CPU with the gloo backend: the term sharding of the
proving key partitions every CRS vector exactly, and partial proofs (computed here by the C oracle, standing in for
each rank's GPU) all-gathered and summed with modular addition equal the reference's proof."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, golden, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import oracle_lib as O
    from rsgv import Case
    from ringsnark_b200.backend import NONE, groth16_shard_layout
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = Case(golden)
        n, aux = case.n, case.aux
        L = groth16_shard_layout(n, aux, rank, world)
        # (1) layouts of all ranks partition the term ranges
        mine = torch.tensor([L["s_pows_lo"], L["s_pows_hi"], L["delta_ts_lo"], L["delta_ts_hi"], L["delta_mid_lo"],
                             L["delta_mid_hi"], int(L["alpha_idx"] != NONE)], dtype=torch.int64)
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        allv = torch.stack(allv).numpy()
        for lo_c, hi_c, total in ((0, 1, n + 1), (2, 3, n + 1), (4, 5, aux)):
            assert allv[0, lo_c] == 0 and allv[-1, hi_c] == total
            assert all(allv[r, hi_c] == allv[r + 1, lo_c] for r in range(world - 1))
        assert allv[:, 6].sum() == 1 and allv[0, 6] == 1
        # (2) partial proofs from the rank's term ranges; all-gather; modular add == reference proof
        s_pows, _ = case.enc("crs_s_pows")
        delta_ts, _ = case.enc("crs_delta_ts")
        delta_mid, _ = case.enc("crs_delta_mid")
        alpha, _ = case.enc("crs_alpha")
        beta, _ = case.enc("crs_beta")

        def part(crs, name, lo, hi, limit):
            words, tag, scalar = case.ring(name)
            tags = O.term_tags(words, tag, scalar)
            hi = min(hi, limit)
            if hi <= lo:
                return np.zeros(case.enc_words, dtype=np.uint64)
            out, _ = O.inner_product(crs[lo:hi], words[lo:hi], tags[lo:hi], case.N_R, case.L_R, case.q, case.N_E, case.L_E, case.Q)
            return out

        def add(a, b):
            return O.enc_add(a, b, case.L_R, case.N_E, case.L_E, case.Q)

        A = add(part(s_pows, "wit_A_io", L["s_pows_lo"], L["s_pows_hi"], n), part(s_pows, "wit_A_mid", L["s_pows_lo"], L["s_pows_hi"], n))
        B = add(part(s_pows, "wit_B_io", L["s_pows_lo"], L["s_pows_hi"], n), part(s_pows, "wit_B_mid", L["s_pows_lo"], L["s_pows_hi"], n))
        Cc = add(part(delta_ts, "wit_H", L["delta_ts_lo"], L["delta_ts_hi"], n + 1),
                 part(delta_mid, "auxiliary_input", L["delta_mid_lo"], L["delta_mid_hi"], aux))
        if L["alpha_idx"] != NONE:
            A, B = add(A, alpha[0]), add(B, beta[0])
        mine = torch.from_numpy(np.stack([A, B, Cc]).view(np.int64))
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        total = None
        for g in gathered:
            g = g.numpy().view(np.uint64)
            total = g if total is None else np.stack([add(total[e], g[e]) for e in range(3)])
        ok = np.array_equal(total, case.enc("proof")[0])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_prover_gloo(world):
    golden = os.path.join(HERE, "golden", "tiny_quirks.rsgv")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, golden, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(world))
    assert res == [(r, True) for r in range(world)]


def _exchange_worker(rank, world, port, golden, q):
    """The slots<->terms exchange of ringsnark_b200/distributed.py on CPU: every rank starts from its SLOT block of the
    reference's witness (all coefficients) and must end with ALL slots of the coefficients of ITS term range."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    from rsgv import Case
    from ringsnark_b200.distributed import VEC_ROWS, send_rows, slot_shard, unpack
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = Case(golden)
        n, L_R, N_R = case.n, case.L_R, case.N_R
        S = N_R // world
        order = ["A_io", "B_io", "C_io", "A_mid", "B_mid", "C_mid"]
        full = np.concatenate([case.ring("wit_" + k)[0] for k in order] + [case.ring("wit_H")[0]])   # [7n+1][L_R*N_R]
        wit = np.concatenate([slot_shard(full, L_R, N_R, rank, world), np.zeros((1, L_R * S), dtype=np.uint64)])
        idx, per = send_rows(n, world)
        send = torch.from_numpy(wit[idx].view(np.int64))
        # gloo has no all_to_all: gather everything, keep the blocks addressed to this rank
        allsend = [torch.zeros_like(send) for _ in range(world)]
        dist.all_gather(allsend, send)
        blk = 5 * per
        recv = torch.stack([s[rank * blk:(rank + 1) * blk] for s in allsend])
        got = unpack(recv, world, per, L_R, S).numpy().view(np.uint64)                    # [5][per][L_R*N_R]
        ok = True
        for v, name in enumerate(VEC_ROWS):
            ref = case.ring("wit_" + name)[0]
            for i in range(per):
                k = rank * per + i
                want = ref[k] if k < ref.shape[0] else np.zeros(L_R * N_R, dtype=np.uint64)
                ok = ok and np.array_equal(got[v, i], want)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_slot_to_term_exchange_gloo(world):
    golden = os.path.join(HERE, "golden", "tiny_fast.rsgv")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, golden, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(world))
    assert res == [(r, True) for r in range(world)]


def _probe_worker(rank, world, port, golden, q):
    """The global transparent-prefix check of the N-GPU path (include/rsgpu.h, rsg_groth16_shard_check) on CPU: every rank
    builds the probe block its GPU would ship -- the running sums of its inner products at NTT slot (c1, limb 0, x = 0), here from
    the C oracle's per-term products --, the blocks are all-gathered with gloo and every rank must reach the same verdict."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import ctypes as C
    import torch
    import oracle_lib as O
    from rsgv import Case
    import ringsnark_b200 as rs
    from ringsnark_b200.backend import NONE, groth16_shard_layout
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = Case(golden)
        n, aux, L_R, Q0 = case.n, case.aux, case.L_R, int(case.Q[0])
        lib = rs.load_library()
        L = groth16_shard_layout(n, aux, rank, world)
        pstride = max((n + 1 + world - 1) // world, (aux + world - 1) // world, 1)
        bw = int(lib.rsg_groth16_shard_block_words(L_R, pstride))
        block = np.full(bw, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
        block[0:8] = 0
        block[7] = pstride
        ct_words = 2 * case.L_E * case.N_E
        probe = lambda enc, j: int(enc[j * ct_words + case.L_E * case.N_E])          # c1, limb 0, x = 0 of ring limb j
        vecs = [("crs_s_pows", "wit_A_io", L["s_pows_lo"], min(L["s_pows_hi"], n)), ("crs_s_pows", "wit_A_mid", L["s_pows_lo"], min(L["s_pows_hi"], n)),
                ("crs_s_pows", "wit_B_io", L["s_pows_lo"], min(L["s_pows_hi"], n)), ("crs_s_pows", "wit_B_mid", L["s_pows_lo"], min(L["s_pows_hi"], n)),
                ("crs_delta_ts", "wit_H", L["delta_ts_lo"], min(L["delta_ts_hi"], n + 1)),
                ("crs_delta_mid", "auxiliary_input", L["delta_mid_lo"], min(L["delta_mid_hi"], aux))]
        for ip, (cname, vname, lo, hi) in enumerate(vecs):
            crs = case.enc(cname)[0]
            words, tag, scalar = case.ring(vname)
            tags = O.term_tags(words, tag, scalar)
            run = [0] * L_R
            live = 0
            for t in range(lo, hi):
                if tags[t] == 0:
                    continue
                live += 1
                term, _ = O.inner_product(crs[t:t + 1], words[t:t + 1], tags[t:t + 1], case.N_R, L_R, case.q, case.N_E, case.L_E, case.Q)
                for j in range(L_R):
                    run[j] = (run[j] + probe(term, j)) % Q0
                    block[8 + 8 * L_R + (ip * L_R + j) * pstride + (t - lo)] = run[j]
            block[1 + ip] = live
            for j in range(L_R):
                block[8 + ip * L_R + j] = run[j]
        if L["alpha_idx"] != NONE:
            for e, name in enumerate(("crs_alpha", "crs_beta")):
                for j in range(L_R):
                    block[8 + 6 * L_R + e * L_R + j] = probe(case.enc(name)[0][0], j)
        mine = torch.from_numpy(block.view(np.int64))
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        blocks = np.ascontiguousarray(torch.stack(gathered).numpy().view(np.uint64))
        verdict = C.c_int(-1)
        rc = lib.rsg_groth16_shard_check(blocks.ctypes.data_as(C.c_void_p), world, L_R, pstride, Q0, C.byref(verdict))
        q.put((rank, rc, verdict.value))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("golden,want", [("tiny_fast", 0), ("tiny_quirks", 0), ("tiny_transp", 1)])
def test_global_transparent_prefix_check_gloo(golden, want):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000) + want + len(golden)
    procs = [ctx.Process(target=_probe_worker, args=(r, world, port, os.path.join(HERE, "golden", golden + ".rsgv"), q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(world))
    assert res == [(r, 0, want) for r in range(world)]
