#!/usr/bin/env python
"""bench.py -- ringGroth16 prover on B200s (one rank per GPU) and the reference's SEAL CPU prover beside it.

A step = one ringGroth16 proof (groth16::prover, zk_proof_systems/groth16/groth16.tcc:69-115) of a synthetic,
satisfying R1CS of the named shape: linear_combination::evaluate -> QRP witness map -> the CRS linear combinations
A, B, C.  Default workload "c4" = the logistic-regression shape of BASELINE.json (N_R = 2048, one 54-bit ring prime,
N_E = 2^14, 8 RNS limbs of 48/49 bit, n = 1031 constraints, 517 primary + 1538 auxiliary inputs), with a synthetic CRS
(uniform residues, generated on the device).

  value : ms per proof, CRS + assignment resident in HBM, timed with CUDA events on the launching stream
  e2e   : ms per proof through the C-ABI call rsg_groth16_prove with HOST buffers (pinned): H2D of the assignment and
          D2H of the proof inside the timed region
  N > 1 : strong scaling of ONE proof (ringsnark_b200/distributed.py): witness map sharded by slot, one NCCL all-to-all
          (slots <-> terms), every CRS vector sharded by term, one NCCL all-gather of the partial proofs + the modular-add
          kernel
  --impl reference : the reference's own CPU prover (oracle/_ref/ref_harness = unmodified ringSNARK + SEAL 4.1.1),
          bounded sample, extrapolated to the workload as stated in `sample`
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ringGroth16 prove time"
UNIT = "ms"
REF_HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


# ------------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(cfg_name, cfg, budget="default"):
    """Times the UNMODIFIED reference on this box's host cores on a bounded sample and extrapolates to one proof.
    lincomb: EncodingElem::inner_product is linear in the number of non-zero terms (seal_ring.tcc:415-431);
    witness map: 44 n^2 N_R L_R modular multiplications, i.e. quadratic in n."""
    if not os.path.exists(REF_HARNESS):
        return None
    n, io, aux = cfg["n"], cfg["io"], cfg["aux"]
    terms_total = 4 * n + (n - 1) + aux
    t_terms = 512 if budget == "default" else 128
    n_s = min(n, 129 if budget == "default" else 65)
    case = "c4" if cfg_name.startswith("c4") else cfg_name
    ncpu = os.cpu_count() or 1

    def run(args):
        out = subprocess.run([REF_HARNESS, "time", case] + args, capture_output=True, text=True, timeout=1200)
        if out.returncode != 0:
            raise RuntimeError(out.stderr[-400:])
        return json.loads(out.stdout.strip().splitlines()[-1])

    lin = run(["lincomb", f"terms={t_terms}", "reps=1", "threads=1"])
    wit = run(["witness", f"n={n_s}", "reps=1"])
    lincomb_ms = lin["seconds"] * 1e3 / t_terms * terms_total
    witness_ms = wit["seconds"] * 1e3 * (n / n_s) ** 2
    # SURVEY.md 8(d) "best-effort parallel": what the host cores could do if groth16::prover ran its inner products the way
    # rinocchio.tcc:106-163 does (OpenMP sections) -- `ncpu` concurrent inner products of t_par terms each; the witness map
    # stays on one thread (Polytools' pragmas are inert).  Reported next to the faithful number, not instead.
    par = None
    try:
        t_par = max(16, t_terms // 8)
        lp = run(["lincomb", f"terms={t_par}", "reps=1", f"threads={ncpu}"])
        par = {"cores": ncpu, "lincomb_terms_per_s": lp["terms_per_s"], "lincomb_ms": terms_total / lp["terms_per_s"] * 1e3,
               "value": terms_total / lp["terms_per_s"] * 1e3 + witness_ms,
               "sample": f"{ncpu} concurrent inner products x {t_par} terms; witness map single-threaded as in the reference"}
    except Exception as ex:
        par = {"error": str(ex)[:200]}
    return {
        "value": lincomb_ms + witness_ms, "unit": UNIT, "cores": 1, "kind": "reference", "host_cpus": ncpu, "extrapolated": True,
        "lincomb_ms_per_term": lin["seconds"] * 1e3 / t_terms, "lincomb_ms": lincomb_ms, "witness_ms": witness_ms,
        "parallel_best_effort": par,
        "sample": (f"unmodified reference (SEAL 4.1.1, g++ -O3), 1 thread (groth16::prover has no OpenMP): inner_product on "
                   f"{t_terms} of {terms_total} terms x{terms_total / t_terms:.1f} (linear) + witness map at n={n_s} "
                   f"x{(n / n_s) ** 2:.1f} (quadratic in n) -> one {cfg_name} proof"),
    }


REF_FULL_CASES = ("c1", "c3p", "c4s", "c4m", "c4")     # shapes oracle/cases.hpp knows: groth16::prover is run WHOLE on these


def cpu_reference_full(cfg_name, cfg):
    """ONE whole run of the unmodified reference's groth16::prover (zk_proof_systems/groth16/groth16.tcc:69-115) on this
    box's host cores: witness map + six inner products + additions, synthetic CRS of the right shape, CRS generation
    excluded.  Measured, not extrapolated (C4: 4-10 minutes on one core -- the reference prover is single-threaded)."""
    if not os.path.exists(REF_HARNESS) or cfg_name not in REF_FULL_CASES:
        return None
    out = subprocess.run([REF_HARNESS, "time", cfg_name, "prover", "reps=1"], capture_output=True, text=True, timeout=3400)
    if out.returncode != 0:
        raise RuntimeError(out.stderr[-400:])
    r = json.loads(out.stdout.strip().splitlines()[-1])
    return {"value": r["seconds"] * 1e3, "unit": UNIT, "cores": 1, "kind": "reference", "host_cpus": os.cpu_count() or 1,
            "extrapolated": False,
            "sample": (f"measured: one whole groth16::prover call of the unmodified reference (SEAL 4.1.1, g++ -O3) on the "
                       f"{cfg_name} shape, n={r['n']}, io={r['io']}, aux={r['aux']}, 1 thread (groth16::prover has no OpenMP)")}


def run_reference_arm(args, cfg_name, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    if not os.path.exists(REF_HARNESS):
        emit({"impl": "reference", "unavailable": "oracle/_ref/ref_harness not built (needs /root/reference at build time)"})
        return
    last = None
    if os.environ.get("RSG_REF_SAMPLE", "full") == "full":
        try:
            last = cpu_reference_full(cfg_name, cfg)    # one measured run, whatever --steps says: a step is 4-10 minutes
        except Exception as ex:
            sys.stderr.write(f"bench.py: full reference run failed ({ex}); falling back to the bounded sample\n")
    steps_run = 1
    if last is None:
        vals = []
        for _ in range(max(1, min(args.steps, 2))):      # each "step" is a fresh bounded sample
            last = cpu_reference_sample(cfg_name, cfg, budget="small")
            vals.append(last["value"])
        last["value"] = statistics.median(vals)
        last["extrapolated"] = True
        steps_run = len(vals)
    v = last["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps_run,
        "warmup": 0, "ms_per_step": v, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": workload_config(cfg_name, cfg, args.gpus),
        "extrapolated": bool(last.get("extrapolated")),
        "cpu_baseline": last,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
    }
    emit(line)


def workload_config(cfg_name, cfg, gpus):
    return {"workload": f"{cfg_name}: ringGroth16 prover, N_R={cfg['N_R']}, L_R={len(cfg['q'])}, N_E={cfg['N_E']}, "
                        f"L_E={len(cfg['Q'])}, n={cfg['n']} constraints, io={cfg['io']}, aux={cfg['aux']}",
            "crs": "synthetic uniform residues", "sharding": f"witness map by slot/{gpus}, all-to-all, lincombs by term/{gpus}, all-gather + modular add" if gpus > 1 else "none",
            "l2": "inputs larger than L2 (CRS streamed per proof >> 126 MB)"}


# ------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for ln in self.proc.stdout:
                self.rows.append([x.strip() for x in ln.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "sampled_over": "warm-up proofs (>= 0.5 s, same load) + both timed regions, 20 ms period"}


def make_assignment(cfg, row_ptr, col, coeff, seed):
    """Satisfying assignment of the synthetic circuit on the host (setup, untimed): free variables uniform in
    [0, q_j), every constraint's output variable = <A_i, x> * <B_i, x> slot-wise."""
    import numpy as np
    n, io, aux, N_R = cfg["n"], cfg["io"], cfg["aux"], cfg["N_R"]
    q = [int(x) for x in cfg["q"]]
    nv, nfree = io + aux, io + aux - n
    rng = np.random.default_rng(seed)
    x = np.zeros((nv, len(q), N_R), dtype=object)
    for j, p in enumerate(q):
        x[:nfree, j, :] = rng.integers(0, p, size=(nfree, N_R), dtype=np.uint64).astype(object)

    def lc(m, i):
        acc = [np.zeros(N_R, dtype=object) for _ in q]
        r = m * n + i
        for t in range(row_ptr[r], row_ptr[r + 1]):
            for j, p in enumerate(q):
                term = (coeff[t] % p) if col[t] == 0 else (coeff[t] % p) * x[col[t] - 1, j]
                acc[j] = (acc[j] + term) % p
        return acc

    for i in range(n):
        a, b = lc(0, i), lc(1, i)
        for j, p in enumerate(q):
            x[nfree + i, j] = (a[j] * b[j]) % p
    return x.astype(np.uint64).reshape(nv, len(q) * N_R)


def proof_checksum(words):
    """64-bit position-weighted checksum of the proof words: equal for every world size when the CRS is seeded by global
    term index (Groth16ProvingKey.fill_synthetic)."""
    import numpy as np
    w = np.ascontiguousarray(words, dtype=np.uint64).reshape(-1)
    k = (np.arange(w.size, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) | np.uint64(1)
    with np.errstate(over="ignore"):
        return "%016x" % int(np.bitwise_xor.reduce(w * k))


def make_assignment_gpu(cfg, row_ptr, col, coeff, seed, device):
    """The same satisfying assignment built on the device (setup, untimed) for shapes where the host loop of make_assignment
    is hopeless (C5: 8193 variables x 32768 slots): free variables uniform (k_fill_uniform), every constraint's output
    variable = <A_i, x> * <B_i, x> through the library's element-wise ring operators (rsg_ring_binop / rsg_ring_scalar_op).
    Returns the host words [io+aux][L_R*N_R]."""
    import ctypes as C
    import ringsnark_b200 as rs
    from ringsnark_b200.capi import check
    n, io, aux = cfg["n"], cfg["io"], cfg["aux"]
    nv, nfree = io + aux, io + aux - n
    ctx = rs.Context(cfg["N_R"], cfg["q"], cfg["N_E"], cfg["Q"], device=device)
    try:
        x = ctx.ringvec(nv)
        x.fill_uniform(seed)
        tmp = ctx.ringvec(3)     # 0: <A_i, x>, 1: <B_i, x>, 2: one scaled term
        lib = ctx.lib

        def lc(m, i, dst):
            r = m * n + i
            first = True
            const = 0
            for t in range(row_ptr[r], row_ptr[r + 1]):
                v, k = int(col[t]), int(coeff[t])
                if v == 0:
                    const += k
                    continue
                if k == 1:
                    src, sidx = x, v - 1
                else:
                    check(lib.rsg_ring_scalar_op(ctx.h, 2, x.h, v - 1, k, tmp.h, 2, 1))
                    src, sidx = tmp, 2
                if first:
                    check(lib.rsg_ring_scalar_op(ctx.h, 0, src.h, sidx, 0, tmp.h, dst, 1))      # copy (add the scalar 0)
                    first = False
                else:
                    check(lib.rsg_ring_binop(ctx.h, 0, tmp.h, dst, src.h, sidx, tmp.h, dst, 1))
            assert not first
            if const:
                check(lib.rsg_ring_scalar_op(ctx.h, 0, tmp.h, dst, const, tmp.h, dst, 1))

        for i in range(n):
            lc(0, i, 0)
            lc(1, i, 1)
            check(lib.rsg_ring_binop(ctx.h, 2, tmp.h, 0, tmp.h, 1, x.h, nfree + i, 1))
        ctx.sync()
        return x.download()
    finally:
        ctx.close()


def parity_single(ctx, r1cs, pk, cfg, proof, rs, seed=7):
    """Outside every timed region: is the proof the bench just timed the reference's proof?
      witness_identity        A(r) B(r) - C(r) = H(r) Z(r) at a random point, on the device's witness-map output
                              (r1cs_to_qrp.tcc:148-259), for a few slots of every ring limb;
      oracle_subrange_terms   the lincomb kernels over a 32-term sub-range of every CRS vector (the same static launch
                              sequence, through rsg_groth16_lincombs) against the C oracle (seal_ring.tcc:361-433);
      shard_sum_equals_proof  the partial proofs of 4 term shards sum (mod Q_l) to the timed proof, word for word."""
    import ctypes as C
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from ringsnark_b200.backend import Groth16Layout, NONE, groth16_shard_layout
    from ringsnark_b200.capi import check
    n, io, aux = cfg["n"], cfg["io"], cfg["aux"]
    q, Q = [int(x) for x in cfg["q"]], [int(x) for x in cfg["Q"]]
    N_R, L_R, W, E = cfg["N_R"], len(q), ctx.ring_words, ctx.enc_words
    out = {}
    evals = r1cs.evaluate(pk.assignment)
    coeffs, H = ctx.ringvec(6 * n), ctx.ringvec(n + 1)
    check(ctx.lib.rsg_witness_map_r1cs(ctx.h, r1cs.h, evals.h, None, coeffs.h, H.h))
    cw = coeffs.download().reshape(6, n, L_R, N_R)      # A_io, B_io, C_io, A_mid, B_mid, C_mid
    hw = H.download().reshape(n + 1, L_R, N_R)
    rng = np.random.default_rng(seed)
    ok = True
    for j, p in enumerate(q):
        for slot in [int(x) for x in rng.integers(0, N_R, size=4)]:
            r = int(rng.integers(n + 1, p))

            def ev(col):
                acc = 0
                for v in reversed(col):
                    acc = (acc * r + int(v)) % p
                return acc
            A = (ev(cw[0, :, j, slot]) + ev(cw[3, :, j, slot])) % p
            B = (ev(cw[1, :, j, slot]) + ev(cw[4, :, j, slot])) % p
            Cc = (ev(cw[2, :, j, slot]) + ev(cw[5, :, j, slot])) % p
            Z = 1
            for i in range(n):
                Z = Z * (r - i) % p
            ok = ok and (A * B - Cc) % p == ev(hw[:, j, slot]) * Z % p
    out["witness_identity"] = bool(ok)

    base = {0: 0, 1: 3 * n, 2: n, 3: 4 * n}               # A_io, A_mid, B_io, B_mid inside `coeffs`
    d_c, d_h, d_a = coeffs.device_ptr(), H.device_ptr(), pk.assignment.device_ptr() + io * W * 8

    def lincombs(s, t, m, alpha):
        """partial proof over s_pows[s0:s1), delta_ts[t0:t1), delta_mid[m0:m1) of the one arena"""
        L = Groth16Layout()
        L.s_pows_off, L.s_pows_lo, L.s_pows_hi = s[0], s[0], s[1]
        L.delta_ts_off, L.delta_ts_lo, L.delta_ts_hi = n + 1 + t[0], t[0], t[1]
        L.delta_mid_off, L.delta_mid_lo, L.delta_mid_hi = 2 * n + 2 + m[0], m[0], m[1]
        L.alpha_idx, L.beta_idx = (2 * n + 2 + aux, 2 * n + 3 + aux) if alpha else (NONE, NONE)
        ptrs = [d_c + (base[k] + s[0]) * W * 8 for k in range(4)] + [d_h + t[0] * W * 8, d_a + m[0] * W * 8]
        return ctx.groth16_lincombs(pk.crs, L, n, aux, ptrs)[0]

    T = max(1, min(32, n - 2, aux))                      # tiny circuits (C1: n = 2): one term
    a = min(100, max(0, n - 2 - 32))
    part = lincombs((a, a + T), (a, a + T), (a, a + T), False)
    tags = np.full(T, 2, dtype=np.uint8)
    ipk = lambda crs_first, words: O.inner_product(pk.crs.download(crs_first, T), words, tags, N_R, L_R, q, ctx.N_E, ctx.L_E, Q)[0]
    cw2 = cw.reshape(6, n, W)
    aux_w = pk.assignment.download(io + a, T)
    want = [O.enc_add(ipk(a, cw2[0, a:a + T]), ipk(a, cw2[3, a:a + T]), L_R, ctx.N_E, ctx.L_E, Q),
            O.enc_add(ipk(a, cw2[1, a:a + T]), ipk(a, cw2[4, a:a + T]), L_R, ctx.N_E, ctx.L_E, Q),
            O.enc_add(ipk(n + 1 + a, hw.reshape(n + 1, W)[a:a + T]), ipk(2 * n + 2 + a, aux_w), L_R, ctx.N_E, ctx.L_E, Q)]
    out["oracle_subrange_terms"] = T if all(np.array_equal(part[e], want[e]) for e in range(3)) else 0

    import torch
    world = 4
    parts = torch.zeros(world * 3 * E, dtype=torch.int64, device="cuda")
    for r_ in range(world):
        d = groth16_shard_layout(n, aux, r_, world)
        p_ = lincombs((d["s_pows_lo"], d["s_pows_hi"]), (d["delta_ts_lo"], d["delta_ts_hi"]), (d["delta_mid_lo"], d["delta_mid_hi"]), r_ == 0)
        parts[r_ * 3 * E:(r_ + 1) * 3 * E] = torch.from_numpy(p_.reshape(-1).view(np.int64))
    total = torch.zeros(3 * E, dtype=torch.int64, device="cuda")
    ctx.enc_sum(parts.data_ptr(), world, 3, total.data_ptr())
    ctx.sync()
    torch.cuda.synchronize()
    out["shard_sum_equals_proof"] = bool(np.array_equal(total.cpu().numpy().view(np.uint64), np.asarray(proof).reshape(-1)))
    out["ok"] = bool(out["witness_identity"] and out["oracle_subrange_terms"] and out["shard_sum_equals_proof"])
    return out


def run_gpu_arm(args, cfg_name, cfg):
    import numpy as np
    import torch
    import ringsnark_b200 as rs
    from ringsnark_b200.params import synthetic_r1cs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, io, aux = cfg["n"], cfg["io"], cfg["aux"]
    SEED = 0xB200
    stream = torch.cuda.Stream()
    row_ptr, col, coeff = synthetic_r1cs(n, io, aux, seed=1)
    big = (io + aux) * cfg["N_R"] * len(cfg["q"]) > (1 << 25)
    h_assign_np = make_assignment_gpu(cfg, row_ptr, col, coeff, SEED, local) if big else make_assignment(cfg, row_ptr, col, coeff, seed=SEED)
    # NTT-domain plaintexts of one rank's term shard beyond the library's default 8 GiB scratch: raise the budget so that the
    # static plan (one launch sequence per proof) still applies -- the shape must then fit HBM, which c5m on one B200 does
    slots = 2 * n + (n + 1) + aux
    need = (slots + world - 1) // world * len(cfg["q"]) * len(cfg["Q"]) * cfg["N_E"]
    if need > (8 << 27) and "RSG_PNTT_BUDGET_WORDS" not in os.environ:
        os.environ["RSG_PNTT_BUDGET_WORDS"] = str(int(need * 1.05))
    import ctypes as C
    from ringsnark_b200.capi import check
    single = world == 1

    def make_single():
        cx = rs.Context(cfg["N_R"], cfg["q"], cfg["N_E"], cfg["Q"], device=local)
        cx.set_stream(stream.cuda_stream)
        r1 = rs.R1cs(cx, n, io, aux, row_ptr, col, coeff)
        key = rs.Groth16ProvingKey(cx, r1, 0, 1)
        key.fill_synthetic(SEED)            # the SAME CRS words for every world size (seeded by global term index)
        key.assignment.upload(h_assign_np)
        return cx, r1, key

    if single:
        ctx, r1cs, pk = make_single()
        ctxs = [ctx]
        layout = pk.layout
        h_assign = torch.from_numpy(h_assign_np.view(np.int64)).pin_memory()
        h_proof = torch.empty(3 * ctx.enc_words, dtype=torch.int64).pin_memory()
        d_final = torch.zeros(3 * ctx.enc_words, dtype=torch.int64, device="cuda")
        h_assign_ptr = h_assign.numpy().view(np.uint64)
        h_proof_np = h_proof.numpy().view(np.uint64)
        h2d_bytes = int(h_assign.numel() * 8)

        def prove(host_io):
            """one step; leaves the proof in d_final / h_proof"""
            with torch.cuda.stream(stream):
                used = (C.c_size_t * 3)()
                check(ctx.lib.rsg_groth16_prove(
                    ctx.h, r1cs.h, pk.crs.h, C.byref(pk.layout), pk.assignment.h,
                    C.c_void_p(h_assign_ptr.ctypes.data) if host_io else None, None,
                    C.c_void_p(h_proof_np.ctypes.data) if host_io else None,
                    C.c_void_p(d_final.data_ptr()), used))
                return [int(u) for u in used]

        def final_words():
            torch.cuda.synchronize()
            return d_final.cpu().numpy().view(np.uint64)
    else:
        # slot-sharded witness map -> all-to-all -> term-sharded lincombs -> all-gather + modular add (distributed.py)
        from ringsnark_b200.distributed import ShardedGroth16Prover, slot_shard
        # peer memory (torch symmetric memory: CUDA IPC mappings + device-side barriers) turns the two exchange steps into single
        # kernels over NVLink (csrc/p2p.cuh); RSG_P2P=0 or an unavailable symmetric-memory backend falls back to NCCL collectives
        symm = None
        if os.environ.get("RSG_P2P", "1") != "0":
            try:
                import torch.distributed._symmetric_memory as symm
            except Exception:
                symm = None
        alloc = (lambda numel: symm.empty(numel, dtype=torch.int64, device=f"cuda:{local}")) if symm else None
        sp = ShardedGroth16Prover(cfg, (row_ptr, col, coeff), rank, world, device=local, stream=stream.cuda_stream, alloc=alloc)
        p2p = False
        if symm:
            try:
                hd = [symm.rendezvous(t, dist.group.WORLD.group_name) for t in (sp.t_full_raw, sp.t_part, sp.t_final)]
                sp.set_peers(list(hd[0].buffer_ptrs), list(hd[1].buffer_ptrs), list(hd[2].buffer_ptrs), lambda: hd[0].barrier())
                sp.t_part.zero_(); sp.t_final.zero_(); sp.t_full_raw.zero_()
                p2p = True
            except Exception as ex:
                sys.stderr.write(f"bench.py: symmetric memory unavailable ({ex!r}); NCCL collectives\n")
        ctx = sp.ctxP
        ctxs = [sp.ctxP, sp.ctxW]
        sp.fill_synthetic(SEED)
        layout = sp.layout
        h_shard = torch.from_numpy(slot_shard(h_assign_np, sp.L_R, sp.N_R, rank, world).view(np.int64)).pin_memory()
        h_aux = torch.from_numpy(np.ascontiguousarray(h_assign_np[io + sp.m_lo:io + sp.m_hi]).view(np.int64)).pin_memory()
        h_proof = torch.empty(3 * ctx.enc_words, dtype=torch.int64).pin_memory()
        sp.load_assignment_shards(h_shard, h_aux, non_blocking=False)
        d_all = torch.zeros(world * sp.part_words, dtype=torch.int64, device="cuda")
        h2d_bytes = int((h_shard.numel() + h_aux.numel()) * 8)

        def prove(host_io):
            with torch.cuda.stream(stream):
                if host_io:
                    sp.load_assignment_shards(h_shard, h_aux)
                if p2p:
                    sp.witness_phase_p2p()
                    used = sp.lincomb_phase()
                    if sp.combine_p2p():
                        from ringsnark_b200.distributed import run_chain
                        run_chain(sp, dist)
                    if host_io:
                        h_proof.copy_(sp.t_final, non_blocking=True)
                    return used
                send = sp.witness_phase()
                recv = torch.empty_like(send)
                dist.all_to_all_single(recv, send)
                used = sp.lincomb_phase(recv)
                dist.all_gather_into_tensor(d_all, sp.t_part)
                if sp.combine(d_all):    # a global prefix vanished at the probe slot (never with uniform CRS words): exact chain
                    from ringsnark_b200.distributed import run_chain
                    run_chain(sp, dist)
                if host_io:
                    h_proof.copy_(sp.t_final, non_blocking=True)
                return used

        def prove_phases(reps=5):
            """GPU-timeline share of each phase of the sharded proof (CUDA events on the launching stream, separate pass)."""
            names = ("witness_map", "all_to_all", "lincombs", "all_gather", "sum_and_check")
            acc = dict.fromkeys(names, 0.0)
            for _ in range(reps if p2p else 0):
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                with torch.cuda.stream(stream):
                    ev[0].record(stream)
                    sp.witness_phase_p2p()
                    ev[1].record(stream)
                    sp.lincomb_phase()
                    ev[2].record(stream)
                    sp.combine_p2p()
                    ev[3].record(stream)
                torch.cuda.synchronize()
                for k, nm in enumerate(("witness_map", "lincombs", "sum_and_check")):
                    acc[nm] += ev[k].elapsed_time(ev[k + 1]) / reps
            if p2p:
                acc["all_to_all"] = acc["all_gather"] = None     # folded into the neighbouring phases (one kernel + barrier each)
                return acc
            for _ in range(reps):
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
                with torch.cuda.stream(stream):
                    ev[0].record(stream)
                    send = sp.witness_phase()
                    recv = torch.empty_like(send)
                    ev[1].record(stream)
                    dist.all_to_all_single(recv, send)
                    ev[2].record(stream)
                    sp.lincomb_phase(recv)
                    ev[3].record(stream)
                    dist.all_gather_into_tensor(d_all, sp.t_part)
                    ev[4].record(stream)
                    sp.combine(d_all)
                    ev[5].record(stream)
                torch.cuda.synchronize()
                for k, nm in enumerate(names):
                    acc[nm] += ev[k].elapsed_time(ev[k + 1]) / reps
            return acc

        def final_words():
            torch.cuda.synchronize()
            return sp.t_final.cpu().numpy().view(np.uint64)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(v):
        if not dist:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # clocks are sampled from here to the end of the second timed region: nvidia-smi needs ~0.1 s to deliver its first line and
    # a timed region at 8 GPUs lasts ~30 ms, so the warm-up (the same proofs under the same load) is stretched to >= 0.5 s
    sampler = ClockSampler(local)
    sampler.start()
    t_warm = time.perf_counter()
    used = None
    for _ in range(max(3, args.warmup)):
        used = prove(False)
        prove(True)
    while True:
        torch.cuda.synchronize()
        go = torch.tensor([1.0 if time.perf_counter() - t_warm < 0.5 else 0.0], device="cuda")
        if dist:
            dist.all_reduce(go, op=dist.ReduceOp.MAX)      # every rank runs the same number of (collective) warm-up proofs
        if go.item() == 0.0:
            break
        used = prove(False)
    barrier()

    # ---- timed region 1: device-resident inputs ("value"); NO per-launch instrumentation inside it
    l0 = sum(cx.launch_count() for cx in ctxs)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        used = prove(False)
    ev1.record(stream)
    barrier()
    ms_dev = reduce_max(ev0.elapsed_time(ev1) / args.steps)
    launches = sum(cx.launch_count() for cx in ctxs) - l0

    # ---- timed region 2: through the C ABI with host buffers ("e2e")
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        prove(True)
        torch.cuda.synchronize()
    barrier()
    ms_e2e = reduce_max((time.perf_counter() - t0) * 1e3 / args.steps)
    clocks = sampler.stop()

    # ---- separate profiled pass (per-launch CUDA events on the launching streams): kernel breakdown + roofline inputs
    psteps = max(1, min(args.steps, 5))
    stat_names = ("lincomb_terms", "lincomb_plain_terms", "lincomb_shared_terms", "lincomb_launches", "ntt_forward_polys",
                  "ntt_inverse_polys", "merged_lincombs", "exact_fallbacks", "fast_proofs", "fast_fallbacks")
    for cx in ctxs:
        cx.enable_timing(True)
    st0 = {k: sum(cx.stat(k) for cx in ctxs) for k in stat_names}
    barrier()
    for _ in range(psteps):
        prove(False)
    barrier()
    stats = {k: (sum(cx.stat(k) for cx in ctxs) - st0[k]) / psteps for k in stat_names}   # per proof, this rank
    kern = {}
    for name in ("k_crs_lincomb", "k_lift_fwd_ntt", "k_encode_intt", "k_interp_fast", "k_quotient_fast", "k_modmat_interp",
                 "k_modmat_divZ", "k_conv_top", "k_centre_add",
                 "k_r1cs_eval", "k_enc_sum", "k_enc_add", "k_is_zero_prefix", "k_probe", "k_probe_eval", "k_full_from_parts",
                 "k_c1_nonzero", "k_zero_transparent"):
        ms = cnt = 0
        for cx in ctxs:
            m_, c_ = cx.timing(name)
            ms, cnt = ms + m_, cnt + c_
        kern[name] = {"ms_per_step": ms / psteps, "launches_per_step": cnt / psteps}
    for cx in ctxs:
        cx.enable_timing(False)

    phases = None
    if not single:
        barrier()
        phases = prove_phases()
        barrier()

    # ---- the other proof system on the same shape (single GPU): rinocchio::prover (rinocchio.tcc:74-190) in zero-knowledge
    # mode through rsg_rinocchio_prove -- eleven inner products over s_pows / alpha_s_pows / beta_prods with every coefficient
    # encoded and transformed once, nine proof elements.  Reported beside the headline, not instead of it.
    rino = None
    rino_key_bytes = (2 * (n + 1) + aux + 3) * ctx.enc_words * 8
    if single and os.environ.get("RSG_BENCH_RINOCCHIO", "1") != "0" and rino_key_bytes < (24 << 30):   # a second key must fit beside the first
        try:
            from ringsnark_b200.backend import CrsRef
            arenas = [ctx.crs(n + 1), ctx.crs(n + 1), ctx.crs(max(aux, 1)), ctx.crs(3)]
            for k_, a_ in enumerate(arenas):
                a_.fill_uniform(SEED + 100 + k_)
            refs = (CrsRef * 6)()
            for k_, (a_, f_) in enumerate(((arenas[0], 0), (arenas[1], 0), (arenas[2], 0), (arenas[3], 0), (arenas[3], 1), (arenas[3], 2))):
                refs[k_].crs, refs[k_].first = a_.h, f_
            dvec = ctx.ringvec(3)
            dvec.fill_uniform(SEED + 200)
            h_d = dvec.download()
            d_rino = torch.zeros(9 * ctx.enc_words, dtype=torch.int64, device="cuda")
            used9 = (C.c_size_t * 9)()

            def prove_rino():
                with torch.cuda.stream(stream):
                    check(ctx.lib.rsg_rinocchio_prove(ctx.h, r1cs.h, refs, pk.assignment.h, None, None, C.c_void_p(h_d.ctypes.data), None,
                                                      C.c_void_p(d_rino.data_ptr()), used9))
            for _ in range(3):
                prove_rino()
            torch.cuda.synchronize()
            rsteps = max(3, min(args.steps, 10))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(rsteps):
                prove_rino()
            e1.record(stream)
            torch.cuda.synchronize()
            rms = e0.elapsed_time(e1) / rsteps
            rino = {"metric": "Rinocchio prove time (zero-knowledge)", "value": rms, "unit": UNIT, "steps": rsteps, "proof_elements": 9,
                    "ratio_to_ringGroth16": rms / ms_dev if ms_dev else None, "terms": [int(u) for u in used9],
                    "fast_fallbacks": int(ctx.stat("fast_fallbacks")),
                    "proof_checksum": proof_checksum(d_rino.cpu().numpy().view(np.uint64)),
                    "parity": "tests/test_fast_path_gpu.py::test_fused_rinocchio_equals_per_inner_product_sequence (c1, c4m); full C4 against "
                              "the reference's prover: profiles/r2_dropin_c4_full_rinocchio.json"}
            del arenas, dvec, d_rino
        except Exception as ex:
            rino = {"error": repr(ex)[:300]}

    # ---- parity, outside every timed region
    proof_words = final_words()
    checksum = proof_checksum(proof_words)
    parity = None
    if single:
        try:
            parity = parity_single(ctx, r1cs, pk, cfg, proof_words, rs)
        except Exception as ex:
            parity = {"ok": False, "error": repr(ex)[:300]}
    else:
        # rank 0 rebuilds the SAME key unsharded and proves on one GPU: the N-rank proof must equal it word for word
        parity = {"ok": True}
        if rank == 0:
            try:
                cx1, r11, pk1 = make_single()
                one, _ = pk1.prove()
                cx1.sync()
                same = bool(np.array_equal(one.reshape(-1), proof_words))
                parity = {"ok": same, "n_rank_proof_equals_1gpu_proof": same, "checksum_1gpu": proof_checksum(one)}
                del pk1, r11
                cx1.close()
            except Exception as ex:
                parity = {"ok": False, "error": repr(ex)[:300]}
        barrier()

    # ---- roofline of the HBM-bound kernel (k_crs_lincomb): algorithmic bytes per step / its device time per step
    L_R, L_E, N_E = len(cfg["q"]), len(cfg["Q"]), cfg["N_E"]
    logN = N_E.bit_length() - 1
    row = L_R * L_E * N_E * 8
    # every streamed term: 2 CRS words per slot, + 1 NTT-domain plaintext word unless the term is a bare ciphertext (alpha /
    # beta / scalar-1 inputs), + one 2-word output per output encoding; counted by the library
    lin_launches = stats["lincomb_launches"]
    n_outputs = 3 if stats["fast_proofs"] else lin_launches
    # A CRS element that two paired splits both multiply (ringGroth16's A and B inner products run over the same s_pows; their
    # CTAs are launched side by side so that the second read is an L2 hit) has to come from HBM once: it is counted once.
    alg_bytes = row * (2 * (stats["lincomb_terms"] - stats["lincomb_shared_terms"]) + stats["lincomb_plain_terms"] + 2 * n_outputs)
    lin_ms = kern["k_crs_lincomb"]["ms_per_step"]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (lin_ms * 1e-3) / 1e9 if lin_ms > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "lincomb_traffic.json"))).get("dram_bytes_per_step")
    except Exception:
        pass
    # NTT work actually done: forward transforms are dense; the batch-encode inverse transform skips the structurally zero
    # quarters when N_R <= N_E / 2 (ntt.cuh: 12 levels at half width, level 1 as N/4 products, level 0 dense)
    dense = (N_E // 2) * logN
    inv_each = dense if (2 * cfg["N_R"] > N_E or logN > 14) else (logN - 2) * (N_E // 4) + N_E // 4 + N_E // 2
    fwd_bfly = stats["ntt_forward_polys"] * dense
    inv_bfly = stats["ntt_inverse_polys"] * inv_each
    fwd_ms, inv_ms = kern["k_lift_fwd_ntt"]["ms_per_step"], kern["k_encode_intt"]["ms_per_step"]
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    pipe_peak = 148 * 64 * sm_max * 1e6          # lane-instructions / s of the FP64 (= of the IMAD) pipe at the maximum clock
    f64 = all(int(x) < (1 << 49) for x in cfg["Q"]) and os.environ.get("RSG_NTT") != "int"
    per_bfly = 10.0 if f64 else 13.0             # pipe instructions per butterfly incl. re-centring / corrections
    ntt_ach = fwd_bfly * per_bfly / (fwd_ms * 1e-3) if fwd_ms else 0.0

    if rank == 0:
        line = {
            "metric": METRIC, "value": ms_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_dev, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "config": workload_config(cfg_name, cfg, world),
            "e2e": {"value": ms_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": int(h_proof.numel() * 8)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity_checked": bool(parity and parity.get("ok")), "parity": parity, "proof_checksum": checksum,
            # per launch: algorithmic bytes / average launch duration, both from the separate profiled pass (CUDA events on
            # the launching stream); traffic = DRAM bytes per launch from the committed ncu --set full capture
            "roofline": {"kernel": "k_crs_lincomb" if os.environ.get("RSG_LIN") == "narrow" else "k_crs_lincomb_wide", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None,
                         "traffic": traffic / lin_launches if traffic and lin_launches and world == 1 else None,
                         "algorithmic_bytes_per_launch": alg_bytes / lin_launches if lin_launches else None,
                         "launch_ms": lin_ms / lin_launches if lin_launches else None, "launches_per_step": lin_launches,
                         "algorithmic_bytes_per_step": alg_bytes, "kernel_ms_per_step": lin_ms,
                         "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s"},
            "roofline_ntt": {"kernel": ("k_lift_fwd_ntt_f64_cl" if logN == 14 and os.environ.get("RSG_NTT_CLUSTER", "1") != "0" else "k_lift_fwd_ntt_f64") if f64 else "k_lift_fwd_ntt", "bound": "fp64 pipe" if f64 else "int32 pipe",
                             "achieved": ntt_ach / 1e12, "peak": pipe_peak / 1e12, "unit": "T lane-instr/s",
                             "frac": ntt_ach / pipe_peak if pipe_peak else None, "instr_per_butterfly_model": per_bfly,
                             "butterflies_per_step": fwd_bfly, "kernel_ms_per_step": fwd_ms,
                             "peak_source": f"148 SM x 64 lanes x {sm_max:.0f} MHz (maximum SM clock)"},
            "kernels_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in kern.items()},
            "phases_ms": {k: (round(v, 4) if v is not None else None) for k, v in phases.items()} if phases else None,
            "rinocchio": rino,
            "exchange": ("peer memory (NVLink loads/stores, csrc/p2p.cuh)" if p2p else "NCCL collectives") if not single else None,
            "ntt": {"forward_butterflies_per_step": fwd_bfly, "inverse_butterflies_per_step": inv_bfly,
                    "forward_gbutterflies_per_s": fwd_bfly / (fwd_ms * 1e-3) / 1e9 if fwd_ms else None,
                    "inverse_gbutterflies_per_s": inv_bfly / (inv_ms * 1e-3) / 1e9 if inv_ms else None},
            "terms_per_step": used,
            "work_per_step": stats,
            "modes": {k: os.environ.get(k) for k in ("RSG_FAST", "RSG_LIN", "RSG_OVERLAP", "RSG_NTT", "RSG_LT_CTAS") if os.environ.get(k)},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb = cpu_reference_sample(cfg_name, cfg)
                line["cpu_baseline"] = cb
                if cb and cb.get("parallel_best_effort"):
                    line["cpu_baseline_parallel"] = cb["parallel_best_effort"]     # inner products on all host cores
            except Exception as ex:  # the baseline is a report, never a reason to lose the GPU number
                line["cpu_baseline"] = {"error": str(ex)[:200]}
        emit(line)
    if single:
        del pk, r1cs
        ctx.close()
    else:
        sp.close()
    if dist:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of the contract goes to the real stdout; everything else a library prints on fd 1 during the
    run (e.g. NCCL's version banner) has been routed to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c4")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    from ringsnark_b200.params import CONFIGS
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference_arm(args, args.config, cfg)
    else:
        run_gpu_arm(args, args.config, cfg)


if __name__ == "__main__":
    main()
