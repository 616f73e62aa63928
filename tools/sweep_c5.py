#!/usr/bin/env python3
"""SURVEY.md 8(d) config C5 (N_R = N_E = 2^15, one 54-bit ring prime, 8 x 55-bit limbs): CRS-lincomb and witness-map sweep on
one B200, with the parity checks the survey asks for at sizes the reference cannot reach:
  * lincomb: T terms over a synthetic uniform CRS; the first 8 terms are re-derived by the C oracle (multiply_plain + add);
    linearity over the whole range: ip(crs, a + b) == ip(crs, a) + ip(crs, b) is NOT an identity of the reference (the lift
    is centred per plaintext), so the check used is split invariance: ip over [0,T) == ip[0,T/2) + ip[T/2,T).
  * witness map at n up to 8192 on 2^15 slots and 16384 on 2^11 slots (quasi-linear path): the identity A(r) B(r) - C(r) = H(r) Z(r) at a random point r per
    sampled slot, evaluated on the host with Python integers from the downloaded coefficients, and equality with the
    dense path at the n where the dense path is affordable.
Writes one JSON object to stdout (and to argv[1] if given).  GPU only: the product path has no CPU fallback."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import ctypes as C
    import ringsnark_b200 as rs
    from ringsnark_b200.params import CONFIGS
    cfg = CONFIGS["c5s"]
    N_R, q, N_E, Q = cfg["N_R"], cfg["q"], cfg["N_E"], cfg["Q"]
    L_R, L_E = len(q), len(Q)
    p = int(q[0])
    out = {"config": "c5: N_R=N_E=32768, L_R=1 (54-bit), L_E=8 (55-bit)", "lincomb": [], "witness": []}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))

    # ---- lincomb sweep
    ctx = rs.Context(N_R, q, N_E, Q)
    for T in (256, 1024, 4096):
        crs = ctx.crs(T); crs.fill_uniform(7)
        vec = ctx.ringvec(T); vec.fill_uniform(8)
        tags = np.full(T, 2, dtype=np.uint8)
        full, used = ctx.inner_product(crs, vec, tags)
        ctx.sync(); ctx.enable_timing(True)
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.inner_product(crs, vec, tags, to_host=False)
        ctx.sync()
        wall = (time.perf_counter() - t0) * 1e3 / reps
        lin = ctx.timing("k_crs_lincomb")[0] / reps
        ntt = (ctx.timing("k_lift_fwd_ntt")[0] + ctx.timing("k_encode_intt")[0]) / reps
        ctx.enable_timing(False)
        alg = L_R * L_E * N_E * 8 * (3 * T + 2)
        bfly = T * (1 + L_E) * (N_E // 2) * 15
        # split invariance
        half = T // 2
        a, _ = ctx.inner_product(crs, vec, tags[:half])
        b, _ = ctx.inner_product(crs, vec, tags[half:], crs_first=half, coeff_first=half)
        Qw = np.repeat(np.array(Q, dtype=np.uint64), N_E)
        Qw = np.tile(Qw, 2 * L_R)
        s = a.astype(np.uint64) + b.astype(np.uint64)
        s = np.where(s >= Qw, s - Qw, s)
        rec = {"terms": T, "inner_product_ms": round(wall, 3), "k_crs_lincomb_ms": round(lin, 4),
               "lincomb_GBs": round(alg / lin / 1e6, 1), "frac_of_measured_hbm_peak": round(alg / lin / 1e6 / peak, 4),
               "ntt_pipeline_ms": round(ntt, 3), "gbutterflies_per_s": round(bfly / ntt / 1e6, 1),
               "split_invariant": bool(np.array_equal(s, full))}
        try:
            import oracle_lib as O
            k = 8
            want, _ = O.inner_product(crs.download(0, k), vec.download(0, k), tags[:k], N_R, L_R, q, N_E, L_E, Q)
            got, _ = ctx.inner_product(crs, vec, tags[:k])
            rec["first_8_terms_equal_oracle"] = bool(np.array_equal(got, want))
        except Exception as ex:  # the oracle library is test infrastructure; absent -> reported, not fatal
            rec["first_8_terms_equal_oracle"] = f"oracle unavailable: {ex}"
        out["lincomb"].append(rec)
        del crs, vec
    ctx.close()

    # ---- witness-map sweep (quasi-linear path), identity at a random point
    rng = np.random.default_rng(5)
    for n, N_R in ((256, N_R), (1024, N_R), (2048, N_R), (4096, N_R), (8192, N_R), (16384, 2048)):
        # n = 16384 (both buffers of a slot in global memory) on 2048 slots only: at 2^15 slots its vectors need > 130 GB
        os.environ["RSG_WITNESS"] = "fast"
        ctx = rs.Context(N_R, q, N_E, Q)
        ev = ctx.ringvec(9 * n); ev.fill_uniform(100 + n)
        # make the full assignment satisfy A*B = C slot-wise: C_full <- A_full * B_full (rsg_ring_binop, op 2 = mul)
        assert ctx.lib.rsg_ring_binop(ctx.h, 2, ev.h, 6 * n, ev.h, 7 * n, ev.h, 8 * n, n) == 0
        coeffs, H = ctx.witness_map(n, ev)
        ctx.sync(); ctx.enable_timing(True)
        reps = 3
        for _ in range(reps):
            ctx.witness_map(n, ev, coeffs, H)
        ctx.sync()
        t_int = ctx.timing("k_interp_fast")[0] / reps
        t_quo = ctx.timing("k_quotient_fast")[0] / reps
        ctx.enable_timing(False)
        full = ctx.interpolate(n, ev, batch=3, y_first=6 * n)          # A_full, B_full, C_full coefficients
        Z = [int(z) for z in ctx.vanishing(n)[0]]
        slots = [0, 1, N_R // 2 + 3, N_R - 1]
        Fc = full.download()[:, slots].astype(object)
        Hc = H.download()[:, slots].astype(object)
        ok = True
        for si in range(len(slots)):
            r = int(rng.integers(1 << 40)) % p
            def ev_poly(col, length):
                acc = 0
                for k in range(length - 1, -1, -1):
                    acc = (acc * r + int(col[k])) % p
                return acc
            A_r, B_r, C_r = (ev_poly(Fc[m * n:(m + 1) * n, si], n) for m in range(3))
            H_r = ev_poly(Hc[:, si], n + 1)
            Z_r = 0
            for k in range(n, -1, -1):
                Z_r = (Z_r * r + Z[k]) % p
            ok = ok and (A_r * B_r - C_r) % p == H_r * Z_r % p
        rec = {"n": n, "N_R": N_R, "interp_ms_8_vectors": round(t_int, 3), "quotient_ms": round(t_quo, 3),
               "witness_ms": round(t_int + t_quo, 3), "identity_AB_minus_C_eq_HZ": bool(ok),
               "Wref_modmuls": 44 * n * n * N_R * L_R,
               "Wref_Gmodmul_per_s": round(44 * n * n * N_R * L_R / (t_int + t_quo) / 1e6, 1)}
        ctx.close()
        if n <= 1024:
            os.environ["RSG_WITNESS"] = "dense"
            ctx = rs.Context(N_R, q, N_E, Q)
            ev2 = ctx.ringvec(9 * n); ev2.fill_uniform(100 + n)
            assert ctx.lib.rsg_ring_binop(ctx.h, 2, ev2.h, 6 * n, ev2.h, 7 * n, ev2.h, 8 * n, n) == 0
            c2, H2 = ctx.witness_map(n, ev2)
            ctx.sync(); ctx.enable_timing(True)
            ctx.witness_map(n, ev2, c2, H2)
            ctx.sync()
            rec["dense_witness_ms"] = round(sum(ctx.timing(k)[0] for k in ("k_modmat_interp", "k_modmat_divZ", "k_conv_top")), 3)
            rec["equals_dense_path"] = bool(np.array_equal(c2.download(), coeffs.download()) and np.array_equal(H2.download(), H.download()))
            ctx.close()
        out["witness"].append(rec)
        del ev, coeffs, H
    os.environ.pop("RSG_WITNESS", None)
    text = json.dumps(out, indent=1)
    print(text)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text + "\n")


if __name__ == "__main__":
    main()
