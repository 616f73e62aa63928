#!/usr/bin/env python3
"""SURVEY.md 8(f) rank 2 measurement: EncodingElem::decode of the three elements of a ringGroth16 proof at the C4
parameters (N_E = 2^14, 8 limbs, one 54-bit ring limb): rsg_decode on the GPU (wall clock of the C call with host
buffers, and the sum of its kernels) next to the unmodified reference on one host core.  Prints one JSON object."""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


def main():
    import ringsnark_b200 as rs
    from ringsnark_b200.capi import RsgError
    from ringsnark_b200.params import CONFIGS
    cfg = CONFIGS["c4"]
    ctx = rs.Context(cfg["N_R"], cfg["q"], cfg["N_E"], cfg["Q"])
    rng = np.random.default_rng(3)
    sk = np.concatenate([rng.integers(0, int(p), size=cfg["N_E"], dtype=np.uint64) for p in cfg["Q"]])
    # timing only: zero ciphertexts decode without a noise error; the arithmetic does not depend on the values
    enc = np.zeros((3, ctx.enc_words), dtype=np.uint64)
    ctx.decode(sk, enc)
    ctx.enable_timing(True)
    reps = 20
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.decode(sk, enc)
    wall = (time.perf_counter() - t0) * 1e3 / reps
    kern = {k: round(ctx.timing(k)[0] / reps, 4) for k in ("k_dec_phase", "k_ntt_inv", "k_dec_modt", "k_ntt_fwd", "k_dec_gather")}
    res = {"workload": "decode of 3 encodings (one ringGroth16 proof), C4 parameters",
           "gpu": {"call_wall_ms": round(wall, 3), "kernels_ms": kern, "kernels_sum_ms": round(sum(kern.values()), 4),
                   "h2d_bytes": int(enc.nbytes + sk.nbytes)}}
    if os.path.exists(REF):
        o = subprocess.run([REF, "time", "c4", "decode", "terms=3", "reps=3"], capture_output=True, text=True, timeout=600)
        r = json.loads(o.stdout.strip().splitlines()[-1])
        res["cpu_reference"] = {"cores": 1, "ms_for_3_encodings": round(r["seconds"] * 1e3, 2)}
        res["speedup_vs_1_core"] = round(r["seconds"] * 1e3 / wall, 1)
    ctx.close()
    text = json.dumps(res, indent=1)
    print(text)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text + "\n")


if __name__ == "__main__":
    main()
