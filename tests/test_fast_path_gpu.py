"""GPU tests of the static-plan prover (csrc/prover_fast.cuh): the one-launch-sequence lincomb phase, the TMA-fed
streaming kernel and the two-stream overlap must give the reference's proof bit for bit -- and fall back to the exact
host-driven path whenever a transparent-ciphertext candidate shows up (seal_ring.tcc:493-504)."""
import glob
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle_lib as O
from rsgv import Case

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.rsgv")))
IDS = [os.path.basename(g)[:-5] for g in GOLD]
REF_HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")

MODES = {
    "exact": {"RSG_FAST": "0"},
    "fast": {},
    "fast_tma": {"RSG_LIN": "tma"},
    "fast_overlap": {"RSG_OVERLAP": "1"},
    "fast_tma_overlap": {"RSG_LIN": "tma", "RSG_OVERLAP": "1"},
    "fast_tma_1cta": {"RSG_LIN": "tma", "RSG_LT_CTAS": "1", "RSG_FAST_SPLITS": "3"},
    "fast_overlap_fullreg": {"RSG_OVERLAP": "2", "RSG_OVERLAP_CHUNKS": "1"},
    "fast_ntt_half": {"RSG_NTT_HALF": "1"},
    # round-2 kernels against the ones they replaced
    "fast_lin_narrow": {"RSG_LIN": "narrow"},                              # k_crs_lincomb<2> instead of k_crs_lincomb_wide
    "fast_ntt_single_cta": {"RSG_NTT_CLUSTER": "0"},                       # k_lift_fwd_ntt_f64 / k_lift_fwd_ntt instead of the cluster kernels
    "fast_lift_barrett": {"RSG_NTT_CLUSTER": "0", "RSG_LIFT": "barrett"},  # LiftIoF64 instead of LiftIoSmallQ
    "fast_ntt_int": {"RSG_NTT": "int"},                                    # integer cluster transform also where the FP64 one applies
    "fast_ntt_int_single_cta": {"RSG_NTT": "int", "RSG_NTT_CLUSTER": "0"},
    "fast_enc_full": {"RSG_ENC_ROWS": "0"},                                # encode_body instead of the compact-row batch encode
    "fast_lin_unpaired": {"RSG_LIN_PAIR": "0"},                            # plan order instead of the A / B splits side by side
}


def _set_mode(monkeypatch, mode):
    for k in ("RSG_FAST", "RSG_LIN", "RSG_OVERLAP", "RSG_LT_CTAS", "RSG_FAST_SPLITS", "RSG_OVERLAP_CHUNKS", "RSG_NTT_HALF", "RSG_NTT_CLUSTER",
              "RSG_LIFT", "RSG_NTT", "RSG_ENC_ROWS", "RSG_LIN_PAIR"):
        monkeypatch.delenv(k, raising=False)
    for k, v in MODES[mode].items():
        monkeypatch.setenv(k, v)


def _aux_kind(case):
    _, tag, scalar = case.ring("auxiliary_input")
    kind = np.full(case.aux, 0xFF, dtype=np.uint8)
    for i in range(case.aux):
        if int(tag[i]) == 0:
            s = int(scalar[i])
            kind[i] = 0 if s == 0 else (1 if s == 1 else 2)
    return kind


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_golden_proofs_every_mode(path, mode, monkeypatch):
    """Every golden case of the reference (fast / slow lift, scalar 0 / 1 / k inputs, zero-prefix elements, the constant wire,
    a transparent prefix) through rsg_groth16_prove in every mode of the lincomb phase."""
    import ringsnark_b200 as rs
    _set_mode(monkeypatch, mode)
    case = Case(path)
    ctx = rs.Context(case.N_R, case.q, case.N_E, case.Q)
    try:
        r1cs = rs.R1cs(ctx, case.n, case.io, case.aux, case.d["r1cs_row_ptr"], case.d["r1cs_col"], case.d["r1cs_coeff"])
        pk = rs.Groth16ProvingKey(ctx, r1cs)
        pk.load(case.enc("crs_s_pows")[0], case.enc("crs_delta_ts")[0], case.enc("crs_delta_mid")[0],
                case.enc("crs_alpha")[0], case.enc("crs_beta")[0])
        assignment = np.concatenate([case.ring("primary_input")[0], case.ring("auxiliary_input")[0]])
        for _ in range(2):   # the second call reuses the cached plan
            proof, used = pk.prove(assignment, _aux_kind(case))
            assert np.array_equal(proof, case.enc("proof")[0])
        if mode == "exact":
            assert ctx.stat("fast_proofs") == 0
        else:
            assert ctx.stat("fast_proofs") == 2
            # tiny_transp: a prefix of <s_pows, A_io> is transparent -> the probe must send the call to the exact path
            assert (ctx.stat("fast_fallbacks") > 0) == (int(case.seed) == 11)
        del pk, r1cs
    finally:
        ctx.close()


@pytest.mark.parametrize("cfg_name", ["c4m", "c5s", "c1"])
def test_synthetic_proofs_every_mode(cfg_name, monkeypatch):
    """Reference-sized parameter sets (N_E = 2^14 on the FP64 pipe, 2^15 split transforms on the integer pipe, 2 ring limbs at
    2^13) with a synthetic CRS: every mode gives the words of the exact path, and a 3-way term sharding sums to them."""
    import torch
    import ringsnark_b200 as rs
    from ringsnark_b200.params import CONFIGS, synthetic_r1cs
    cfg = CONFIGS[cfg_name]
    n, io, aux = cfg["n"], cfg["io"], cfg["aux"]
    row_ptr, col, coeff = synthetic_r1cs(n, io, aux, seed=3)
    proofs, useds = {}, {}
    for mode in MODES:
        _set_mode(monkeypatch, mode)
        ctx = rs.Context(cfg["N_R"], cfg["q"], cfg["N_E"], cfg["Q"])
        try:
            r1cs = rs.R1cs(ctx, n, io, aux, row_ptr, col, coeff)
            pk = rs.Groth16ProvingKey(ctx, r1cs)
            pk.fill_synthetic(5)
            pk.assignment.fill_uniform(6)
            proofs[mode], used = pk.prove()
            useds[mode] = used
            if mode == "fast":   # term shards of the same key: partial proofs sum to the proof
                E = ctx.enc_words
                parts = torch.zeros(3 * 3 * E, dtype=torch.int64, device="cuda")
                for r in range(3):
                    pkr = rs.Groth16ProvingKey(ctx, r1cs, r, 3)
                    pkr.fill_synthetic(5)
                    pkr.assignment.fill_uniform(6)
                    p, _ = pkr.prove()
                    parts[r * 3 * E:(r + 1) * 3 * E] = torch.from_numpy(p.reshape(-1).view(np.int64))
                    del pkr
                total = torch.zeros(3 * E, dtype=torch.int64, device="cuda")
                ctx.enc_sum(parts.data_ptr(), 3, 3, total.data_ptr())
                ctx.sync()
                torch.cuda.synchronize()
                assert np.array_equal(total.cpu().numpy().view(np.uint64), proofs[mode].reshape(-1))
            del pk, r1cs
        finally:
            ctx.close()
    for mode in MODES:
        assert np.array_equal(proofs[mode], proofs["exact"]), mode
        assert useds[mode] == useds["exact"], (mode, useds[mode], useds["exact"])


def test_fill_uniform_at_matches_whole_arena():
    import ringsnark_b200 as rs
    from ringsnark_b200.params import CONFIGS
    cfg = CONFIGS["c1"]
    ctx = rs.Context(cfg["N_R"], cfg["q"], cfg["N_E"], cfg["Q"])
    try:
        whole = ctx.crs(7)
        whole.fill_uniform(99)
        piece = ctx.crs(3)
        piece.fill_uniform_at(0, 2, 4, 99)     # elements 4, 5 of the virtual arena
        piece.fill_uniform_at(2, 1, 1, 99)     # element 1
        w = whole.download()
        p = piece.download()
        assert np.array_equal(p[0], w[4]) and np.array_equal(p[1], w[5]) and np.array_equal(p[2], w[1])
        Q = np.asarray(cfg["Q"], dtype=np.uint64)
        assert (w.reshape(7, ctx.L_R, 2, ctx.L_E, ctx.N_E) < Q[None, None, None, :, None]).all()
        del whole, piece
    finally:
        ctx.close()


def test_c2p_ring_element_coefficients():
    """SURVEY.md 8(d) C2' (benchmarks/bench_ntt_SEAL.cpp:29-55 restated): ONE constraint whose linear combination has
    RING-ELEMENT coefficients, no auxiliary input -- the domain is {0}, H = 0 and the proof's C is the EMPTY encoding.  The
    system has no CSR form with scalar coefficients, so the path runs from the reference's evaluation vectors: witness map,
    the prover's inner products and the += chain against the reference's dump (built with -DNDEBUG: the reference asserts
    on copying an empty encoding)."""
    if not os.path.exists(REF_HARNESS):
        pytest.skip("oracle/_ref/ref_harness not built (needs /root/reference at build time)")
    import ringsnark_b200 as rs
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "c2p.rsgv")
        subprocess.check_call([REF_HARNESS, "dump", "c2p", path, "31"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
        case = Case(path)
    assert int(case.d["r1cs_scalar_coeffs"][0]) == 0 and case.n == 1 and case.aux == 0
    ctx = rs.Context(case.N_R, case.q, case.N_E, case.Q)
    try:
        n = case.n
        order = ["A_mid", "B_mid", "C_mid", "A_io", "B_io", "C_io", "A_full", "B_full", "C_full"]
        ev = ctx.ringvec_from(np.concatenate([case.ring("eval_" + k)[0] for k in order]))
        coeffs, H = ctx.witness_map(n, ev)
        got = coeffs.download()
        for idx, k in enumerate(["A_io", "B_io", "C_io", "A_mid", "B_mid", "C_mid"]):
            words, tag, scalar = case.ring("wit_" + k)
            for i in range(n):   # the reference keeps scalar coefficients as scalars: compare slot-wise values
                want = words[i] if int(tag[i]) else np.concatenate([np.full(case.N_R, int(scalar[i]) % int(p), dtype=np.uint64) for p in case.q])
                assert np.array_equal(got[idx * n + i], want), (k, i)
        s_pows = ctx.crs_from(case.enc("crs_s_pows")[0])
        delta_ts = ctx.crs_from(case.enc("crs_delta_ts")[0])
        ip, ip_size = case.enc("ip")
        jobs = [(s_pows, coeffs, 0, "wit_A_io"), (s_pows, coeffs, 3 * n, "wit_A_mid"), (s_pows, coeffs, n, "wit_B_io"),
                (s_pows, coeffs, 4 * n, "wit_B_mid")]
        outs = []
        for k, (crs, vec, first, name) in enumerate(jobs):
            _, tag, scalar = case.ring(name)
            tags = ctx.term_tags(vec, tag, scalar, first=first, count=n)
            out, used = ctx.inner_product(crs, vec, tags, coeff_first=first)
            assert (used == 0) == (int(ip_size[k][0]) == 2 ** 64 - 1), k
            if used:
                assert np.array_equal(out, ip[k]), k
            outs.append((out, used))
        # H = 0: every term of <delta_ts, H> is skipped -> the empty encoding, like the reference's C
        _, tag, scalar = case.ring("wit_H")
        tags = ctx.term_tags(H, tag, scalar, first=0, count=len(tag))
        out, used = ctx.inner_product(delta_ts, H, tags)
        assert used == 0 and int(ip_size[4][0]) == 2 ** 64 - 1
        proof, psize = case.enc("proof")
        assert int(psize[2][0]) == 2 ** 64 - 1          # the reference's C is empty
        # A = ip0 + ip1 + alpha, B = ip2 + ip3 + beta through the C ABI's operator+=
        import torch
        for e, extra in enumerate(("crs_alpha", "crs_beta")):
            acc = None
            for (o, used) in outs[2 * e:2 * e + 2]:
                if not used:
                    continue
                t = torch.from_numpy(o.view(np.int64)).cuda()
                if acc is None:
                    acc = t.clone()
                else:
                    assert ctx.lib.rsg_enc_add(ctx.h, acc.data_ptr(), t.data_ptr()) == 0
            t = torch.from_numpy(np.ascontiguousarray(case.enc(extra)[0][0]).view(np.int64)).cuda()
            if acc is None:
                acc = t.clone()
            else:
                assert ctx.lib.rsg_enc_add(ctx.h, acc.data_ptr(), t.data_ptr()) == 0
            ctx.sync()
            torch.cuda.synchronize()
            assert np.array_equal(acc.cpu().numpy().view(np.uint64), proof[e]), e
        del s_pows, delta_ts, ev, coeffs, H
    finally:
        ctx.close()


def _seeds(case_seed, count, L_R):
    """cases.hpp::make_enc_contexts seeds limb j's factory with {seed, j + 1, 0xB200, 0...}; a seeded factory hands the same
    seed to every PRNG it creates (randomgen.h:440-448)."""
    s = np.zeros((count, L_R, 8), dtype=np.uint64)
    for j in range(L_R):
        s[:, j, 0], s[:, j, 1], s[:, j, 2] = int(case_seed), j + 1, 0xB200
    return s


@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_encode_reproduces_reference_crs(path):
    """EncodingElem::encode on the GPU (rsg_encode): decode the reference's CRS elements with the C oracle, encode the ring
    elements again on the GPU under the same secret keys and seeds -> SEAL's ciphertext words, bit for bit (Blake2xb PRNG,
    uniform and centred-binomial samplers, c0 = -(a s + t e) + m)."""
    import ringsnark_b200 as rs
    case = Case(path)
    ctx = rs.Context(case.N_R, case.q, case.N_E, case.Q)
    try:
        sk = case.d["dec_sk"]
        want = np.concatenate([case.enc("crs_s_pows")[0], case.enc("crs_delta_mid")[0][:3], case.enc("crs_alpha")[0], case.enc("crs_beta")[0]])
        rings = np.stack([O.decode(w, sk, case.N_R, case.L_R, case.q, case.N_E, case.L_E, case.Q)[0] for w in want])
        vec = ctx.ringvec_from(rings)
        crs = ctx.encode(sk, vec, _seeds(case.seed, len(want), case.L_R))
        got = crs.download()
        for i in range(len(want)):
            assert np.array_equal(got[i], want[i]), i
        # and the verifier's front half gives the ring elements back
        ring2, budget = ctx.decode(sk, got)
        assert np.array_equal(ring2, rings)
        del vec, crs
    finally:
        ctx.close()


def test_encode_redraws_against_oracle():
    """sample_poly_uniform's rejection loop (rlwe.cpp:120-127) practically never fires for SEAL's default primes (just below a
    power of two: ~2^-31 per word); with primes near 0.75 * 2^60 one word in 64 is redrawn, some of them twice.  GPU against
    the C oracle (pinned to SEAL on the bulk path by the golden CRS)."""
    import ringsnark_b200 as rs
    N_R, N_E = 128, 256

    def primes(start, count):
        out, k = [], start // 512
        while len(out) < count:
            p = k * 512 + 1
            if pow(2, p - 1, p) == 1 and pow(3, p - 1, p) == 1 and pow(5, p - 1, p) == 1:
                out.append(p)
            k += 1
        return out

    Q = primes(3 << 58, 4)
    q = primes(1 << 24, 2)
    ctx = rs.Context(N_R, q, N_E, Q)
    try:
        rng = np.random.default_rng(4)
        L_R, L_E, count = 2, 4, 5
        sk = np.stack([np.stack([rng.integers(0, Ql, size=N_E, dtype=np.uint64) for Ql in Q]) for _ in range(L_R)])
        rings = np.stack([np.concatenate([rng.integers(0, qj, size=N_R, dtype=np.uint64) for qj in q]) for _ in range(count)])
        seeds = rng.integers(0, 2 ** 63, size=(count, L_R, 8), dtype=np.uint64)
        n_rej = 0
        for i in range(count):
            pub = O.prng_bytes(seeds[i, 0], 0, 64).view(np.uint64)
            bulk = O.prng_bytes(pub, 0, L_E * N_E * 8).view(np.uint64).reshape(L_E, N_E)
            for l, Ql in enumerate(Q):
                n_rej += int((bulk[l] >= np.uint64((2 ** 64 - 1) - ((2 ** 64 - 1) % Ql) - 1)).sum())
        assert n_rej > 20          # the case really exercises the redraws
        vec = ctx.ringvec_from(rings)
        got = ctx.encode(sk, vec, seeds).download()
        for i in range(count):
            want = O.encode(rings[i], sk, seeds[i], N_R, L_R, q, N_E, L_E, Q)
            assert np.array_equal(got[i], want), i
        del vec
    finally:
        ctx.close()


def test_encode_round_trip_full_size():
    """C4 parameters (N_E = 2^14, 8 limbs): decode(encode(r)) = r with fresh noise budget, 40 elements, distinct seeds."""
    import ringsnark_b200 as rs
    from ringsnark_b200.params import CONFIGS
    cfg = CONFIGS["c4"]
    ctx = rs.Context(cfg["N_R"], cfg["q"], cfg["N_E"], cfg["Q"])
    try:
        rng = np.random.default_rng(9)
        count = 40
        sk = np.stack([rng.integers(0, int(Ql), size=ctx.N_E, dtype=np.uint64) for Ql in cfg["Q"]])[None]
        vec = ctx.ringvec(count)
        vec.fill_uniform(11)
        seeds = rng.integers(0, 2 ** 63, size=(count, 1, 8), dtype=np.uint64)
        crs = ctx.encode(sk, vec, seeds)
        ring, budget = ctx.decode(sk, crs.download())
        assert np.array_equal(ring, vec.download())
        assert (budget > 200).all()
        one = O.encode(vec.download(3, 1)[0], sk, seeds[3], ctx.N_R, 1, cfg["q"], ctx.N_E, ctx.L_E, cfg["Q"])
        assert np.array_equal(crs.download(3, 1)[0], one)
        del vec, crs
    finally:
        ctx.close()


def _ip_exact(ctx, crs, first, vec, vfirst, count, tags=None):
    """one EncodingElem::inner_product through the exact per-inner-product entry point"""
    if tags is None:
        tags = ctx.term_tags(vec, first=vfirst, count=count)
    return ctx.inner_product(crs, vec, tags, crs_first=first, coeff_first=vfirst)


@pytest.mark.parametrize("cfg_name,zk", [("c4m", True), ("c1", True), ("c4m", False)])
def test_fused_rinocchio_equals_per_inner_product_sequence(cfg_name, zk):
    """rsg_rinocchio_prove (one transform per coefficient, shared by the s_pows and alpha_s_pows streams, shifts fused) against
    the reference's own sequence rinocchio.tcc:74-190 assembled from the exact per-inner-product entry points: ten / eleven
    rsg_inner_product calls, d_k * z_enc as a one-term inner product over the intermediate encoding, rsg_enc_add."""
    import torch
    import ringsnark_b200 as rs
    from ringsnark_b200.backend import rinocchio_prove, groth16_prove_refs
    from ringsnark_b200.capi import check
    from ringsnark_b200.params import CONFIGS, synthetic_r1cs
    cfg = CONFIGS[cfg_name]
    n, io, aux = cfg["n"], cfg["io"], cfg["aux"]
    row_ptr, col, coeff = synthetic_r1cs(n, io, aux, seed=5)
    ctx = rs.Context(cfg["N_R"], cfg["q"], cfg["N_E"], cfg["Q"])
    try:
        W, E = ctx.ring_words, ctx.enc_words
        r1cs = rs.R1cs(ctx, n, io, aux, row_ptr, col, coeff)
        s_pows, alpha_s = ctx.crs(n + 1), ctx.crs(n + 1)
        beta_prods, beta_ts = ctx.crs(aux), ctx.crs(3)
        for k, a in enumerate((s_pows, alpha_s, beta_prods, beta_ts)):
            a.fill_uniform(100 + k)
        assignment = ctx.ringvec(io + aux)
        assignment.fill_uniform(7)
        dvec = ctx.ringvec(3)
        dvec.fill_uniform(8)
        h_d = dvec.download() if zk else None
        refs = [(s_pows, 0), (alpha_s, 0), (beta_prods, 0), (beta_ts, 0), (beta_ts, 1), (beta_ts, 2)]
        got, used = rinocchio_prove(ctx, r1cs, refs, assignment, h_d)
        assert ctx.stat("fast_fallbacks") == 0
        # ---- the reference's sequence, exact entry points
        evals = r1cs.evaluate(assignment)
        coeffs, H = ctx.ringvec(6 * n), ctx.ringvec(n + 1)
        import ctypes as C
        check(ctx.lib.rsg_witness_map_r1cs(ctx.h, r1cs.h, evals.h, h_d.ctypes.data_as(C.c_void_p) if zk else None, coeffs.h, H.h))
        Zc = ctx.vanishing(n)                                   # [L_R][n+1]
        zw = np.repeat(Zc.T.reshape(n + 1, ctx.L_R, 1), ctx.N_R, axis=2).reshape(n + 1, W)
        Z = ctx.ringvec_from(zw)
        ztags = ctx.term_tags(Z)
        ztags[n] = 1                                            # the leading coefficient is the scalar 1
        want = []
        for vfirst in (3 * n, 4 * n, 5 * n):                   # a_mid, b_mid, c_mid
            for crs in (s_pows, alpha_s):
                want.append(_ip_exact(ctx, crs, 0, coeffs, vfirst, n)[0])
        for crs in (s_pows, alpha_s):
            want.append(_ip_exact(ctx, crs, 0, H, 0, n + 1)[0])
        z_enc = [_ip_exact(ctx, crs, 0, Z, 0, n + 1, ztags)[0] for crs in (s_pows, alpha_s)]
        aux_vec = ctx.ringvec_from(assignment.download(io, aux))
        want.append(_ip_exact(ctx, beta_prods, 0, aux_vec, 0, aux)[0])

        def shifted(acc_words, enc_words_list, ks):
            acc = torch.from_numpy(acc_words.view(np.int64)).cuda()
            for words, k in zip(enc_words_list, ks):
                one = ctx.crs_from(words.reshape(1, E))
                prod, _ = ctx.inner_product(one, dvec, np.array([2], dtype=np.uint8), coeff_first=k)
                t = torch.from_numpy(prod.view(np.int64)).cuda()
                assert ctx.lib.rsg_enc_add(ctx.h, acc.data_ptr(), t.data_ptr()) == 0
                ctx.sync()
                del one
            torch.cuda.synchronize()
            return acc.cpu().numpy().view(np.uint64)

        if zk:
            for e in range(6):
                want[e] = shifted(want[e], [z_enc[e & 1]], [e // 2])
            bt = beta_ts.download()
            want[8] = shifted(want[8], [bt[0], bt[1], bt[2]], [0, 1, 2])
        for e in range(9):
            assert np.array_equal(got[e], want[e]), e
        # ringGroth16 over separate arenas == over one arena
        pk = rs.Groth16ProvingKey(ctx, r1cs)
        pk.fill_synthetic(9)
        pk.assignment.upload(assignment.download())
        one_arena, used1 = pk.prove()
        L = pk.layout
        parts = [ctx.crs_from(pk.crs.download(L.s_pows_off, n + 1)), ctx.crs_from(pk.crs.download(L.delta_ts_off, n + 1)),
                 ctx.crs_from(pk.crs.download(L.delta_mid_off, aux)), ctx.crs_from(pk.crs.download(L.alpha_idx, 2))]
        refs5 = [(parts[0], 0), (parts[1], 0), (parts[2], 0), (parts[3], 0), (parts[3], 1)]
        many, used5 = groth16_prove_refs(ctx, r1cs, refs5, assignment)
        assert np.array_equal(many, one_arena) and used5 == used1
        del pk, parts, r1cs, s_pows, alpha_s, beta_prods, beta_ts, assignment, dvec, evals, coeffs, H, Z, aux_vec
    finally:
        ctx.close()
