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
        "value": lincomb_ms + witness_ms, "unit": UNIT, "cores": 1, "kind": "reference", "host_cpus": ncpu,
        "lincomb_ms_per_term": lin["seconds"] * 1e3 / t_terms, "lincomb_ms": lincomb_ms, "witness_ms": witness_ms,
        "parallel_best_effort": par,
        "sample": (f"unmodified reference (SEAL 4.1.1, g++ -O3), 1 thread (groth16::prover has no OpenMP): inner_product on "
                   f"{t_terms} of {terms_total} terms x{terms_total / t_terms:.1f} (linear) + witness map at n={n_s} "
                   f"x{(n / n_s) ** 2:.1f} (quadratic in n) -> one {cfg_name} proof"),
    }


def run_reference_arm(args, cfg_name, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    vals, last = [], None
    for _ in range(max(1, min(args.steps, 2))):      # each "step" is a fresh bounded sample
        last = cpu_reference_sample(cfg_name, cfg, budget="small")
        if last is None:
            emit({"impl": "reference", "unavailable": "oracle/_ref/ref_harness not built (needs /root/reference at build time)"})
            return
        vals.append(last["value"])
    v = statistics.median(vals)
    last["value"] = v
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": v, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": workload_config(cfg_name, cfg, args.gpus),
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
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
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
                "reasons": sorted(reasons), "samples": len(sm)}


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
    stream = torch.cuda.Stream()
    row_ptr, col, coeff = synthetic_r1cs(n, io, aux, seed=1)
    h_assign_np = make_assignment(cfg, row_ptr, col, coeff, seed=0xB200)
    import ctypes as C
    from ringsnark_b200.capi import check
    single = world == 1
    if single:
        ctx = rs.Context(cfg["N_R"], cfg["q"], cfg["N_E"], cfg["Q"], device=local)
        ctx.set_stream(stream.cuda_stream)
        ctxs = [ctx]
        r1cs = rs.R1cs(ctx, n, io, aux, row_ptr, col, coeff)
        pk = rs.Groth16ProvingKey(ctx, r1cs, 0, 1)
        pk.crs.fill_uniform(0xB200)
        layout = pk.layout
        h_assign = torch.from_numpy(h_assign_np.view(np.int64)).pin_memory()
        h_proof = torch.empty(3 * ctx.enc_words, dtype=torch.int64).pin_memory()
        pk.assignment.upload(h_assign_np)
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
    else:
        # slot-sharded witness map -> all-to-all -> term-sharded lincombs -> all-gather + modular add (distributed.py)
        from ringsnark_b200.distributed import ShardedGroth16Prover, slot_shard
        sp = ShardedGroth16Prover(cfg, (row_ptr, col, coeff), rank, world, device=local, stream=stream.cuda_stream)
        ctx = sp.ctxP
        ctxs = [sp.ctxP, sp.ctxW]
        sp.crs.fill_uniform(0xB200 + rank)
        layout = sp.layout
        h_shard = torch.from_numpy(slot_shard(h_assign_np, sp.L_R, sp.N_R, rank, world).view(np.int64)).pin_memory()
        h_aux = torch.from_numpy(np.ascontiguousarray(h_assign_np[io + sp.m_lo:io + sp.m_hi]).view(np.int64)).pin_memory()
        h_proof = torch.empty(3 * ctx.enc_words, dtype=torch.int64).pin_memory()
        sp.load_assignment_shards(h_shard, h_aux, non_blocking=False)
        d_all = torch.zeros(world * 3 * ctx.enc_words, dtype=torch.int64, device="cuda")
        h2d_bytes = int((h_shard.numel() + h_aux.numel()) * 8)

        def prove(host_io):
            with torch.cuda.stream(stream):
                if host_io:
                    sp.load_assignment_shards(h_shard, h_aux)
                send = sp.witness_phase()
                recv = torch.empty_like(send)
                dist.all_to_all_single(recv, send)
                used = sp.lincomb_phase(recv)
                dist.all_gather_into_tensor(d_all, sp.t_part)
                sp.combine(d_all)
                if host_io:
                    h_proof.copy_(sp.t_final, non_blocking=True)
                return used

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

    used = None
    for _ in range(args.warmup):
        used = prove(False)
        prove(True)
    barrier()

    # ---- timed region 1: device-resident inputs ("value"), with per-kernel event timing for the roofline
    sampler = ClockSampler(local)
    sampler.start()
    for cx in ctxs:
        cx.enable_timing(True)
    l0 = sum(cx.launch_count() for cx in ctxs)
    stat_names = ("lincomb_terms", "lincomb_plain_terms", "lincomb_launches", "ntt_forward_polys", "ntt_inverse_polys",
                  "merged_lincombs", "exact_fallbacks")
    st0 = {k: sum(cx.stat(k) for cx in ctxs) for k in stat_names}
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        used = prove(False)
    ev1.record(stream)
    barrier()
    ms_dev = reduce_max(ev0.elapsed_time(ev1) / args.steps)
    launches = sum(cx.launch_count() for cx in ctxs) - l0
    stats = {k: (sum(cx.stat(k) for cx in ctxs) - st0[k]) / args.steps for k in stat_names}   # per proof, this rank
    kern = {}
    for name in ("k_crs_lincomb", "k_lift_fwd_ntt", "k_encode_intt", "k_interp_fast", "k_quotient_fast", "k_modmat_interp",
                 "k_modmat_divZ", "k_conv_top",
                 "k_r1cs_eval", "k_enc_sum", "k_enc_add", "k_is_zero_prefix", "k_probe", "k_probe_eval", "k_full_from_parts",
                 "k_c1_nonzero", "k_zero_transparent"):
        ms = cnt = 0
        for cx in ctxs:
            m_, c_ = cx.timing(name)
            ms, cnt = ms + m_, cnt + c_
        kern[name] = {"ms_per_step": ms / args.steps, "launches_per_step": cnt / args.steps}
    for cx in ctxs:
        cx.enable_timing(False)

    # ---- timed region 2: through the C ABI with host buffers ("e2e")
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        prove(True)
        torch.cuda.synchronize()
    barrier()
    ms_e2e = reduce_max((time.perf_counter() - t0) * 1e3 / args.steps)
    clocks = sampler.stop()

    # ---- roofline of the dominant kernel (k_crs_lincomb): algorithmic bytes per step / its device time per step
    L_R, L_E, N_E = len(cfg["q"]), len(cfg["Q"]), cfg["N_E"]
    row = L_R * L_E * N_E * 8
    ones = [1 if layout.alpha_idx != rs.backend.NONE else 0, 1 if layout.beta_idx != rs.backend.NONE else 0, 0]
    # alpha / beta are added by k_enc_add, every other term is streamed by k_crs_lincomb: 3 words per slot per term
    # (2 CRS + 1 NTT-domain plaintext) + one 2-word output per launch
    # counted by the library (rsg_context_stat): the merged A / B passes stream each s_pows element ONCE for io + mid
    lin_launches = stats["lincomb_launches"]
    alg_bytes = row * (2 * stats["lincomb_terms"] + stats["lincomb_plain_terms"] + 2 * lin_launches)
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
    ntt_butterflies = (stats["ntt_forward_polys"] + stats["ntt_inverse_polys"]) * (N_E // 2) * (N_E.bit_length() - 1)
    ntt_ms = kern["k_lift_fwd_ntt"]["ms_per_step"] + kern["k_encode_intt"]["ms_per_step"]

    if rank == 0:
        line = {
            "metric": METRIC, "value": ms_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "config": workload_config(cfg_name, cfg, world),
            "e2e": {"value": ms_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": int(h_proof.numel() * 8)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            # per launch: algorithmic bytes / average launch duration (the per-step sums divided by the launches per step);
            # traffic = DRAM bytes per launch from the committed ncu --set full capture (profiles/lincomb_traffic.json)
            "roofline": {"kernel": "k_crs_lincomb", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None,
                         # the committed capture is the 1-GPU workload; a sharded run has no capture of its own
                         "traffic": traffic / lin_launches if traffic and lin_launches and world == 1 else None,
                         "algorithmic_bytes_per_launch": alg_bytes / lin_launches if lin_launches else None,
                         "launch_ms": lin_ms / lin_launches if lin_launches else None, "launches_per_step": lin_launches,
                         "algorithmic_bytes_per_step": alg_bytes, "kernel_ms_per_step": lin_ms,
                         "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s"},
            "kernels_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in kern.items()},
            "ntt": {"butterflies_per_step": ntt_butterflies, "gbutterflies_per_s": ntt_butterflies / (ntt_ms * 1e-3) / 1e9 if ntt_ms else None},
            "terms_per_step": used,
            "work_per_step": stats,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_reference_sample(cfg_name, cfg)
            except Exception as ex:  # the baseline is a report, never a reason to lose the GPU number
                line["cpu_baseline"] = {"error": str(ex)[:200]}
        emit(line)
    if single:
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
