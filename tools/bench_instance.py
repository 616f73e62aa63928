#!/usr/bin/env python3
"""SURVEY.md 8(f) rank 3 measurement: r1cs_to_qrp_instance_map_with_evaluation at the C4 shape (n = 1031, io = 517,
aux = 1538, N_R = 2048) on the GPU (rsg_instance_map, CUDA events around its two kernels + wall clock of the C call),
next to the unmodified reference on one host core at n = 129 and 257 (quadratic in n: extrapolated to 1031, stated).
Prints one JSON object; argv[1] (optional) = output file."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


def main():
    import ringsnark_b200 as rs
    from ringsnark_b200.params import CONFIGS, synthetic_r1cs
    cfg = CONFIGS["c4"]
    n, io, aux = cfg["n"], cfg["io"], cfg["aux"]
    ctx = rs.Context(cfg["N_R"], cfg["q"], cfg["N_E"], cfg["Q"])
    r1cs = rs.R1cs(ctx, n, io, aux, *synthetic_r1cs(n, io, aux, seed=1))
    t = ctx.ringvec(1); t.fill_uniform(9)
    r1cs.instance_map(t); ctx.sync()
    ctx.enable_timing(True)
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        out = r1cs.instance_map(t)
    ctx.sync()
    wall = (time.perf_counter() - t0) * 1e3 / reps
    k1, k2 = ctx.timing("k_lagrange_at")[0] / reps, ctx.timing("k_instance_accum")[0] / reps
    res = {"workload": f"c4 instance map: n={n}, variables={io + aux}, N_R={cfg['N_R']}, L_R=1",
           "gpu": {"k_lagrange_at_ms": round(k1, 4), "k_instance_accum_ms": round(k2, 4), "call_wall_ms": round(wall, 4)},
           "reference_algorithm_modmuls": 2 * n * n * cfg["N_R"]}
    if os.path.exists(REF):
        pts = []
        for ns in (129, 257):
            o = subprocess.run([REF, "time", "c4", "instance", f"n={ns}", "reps=1"], capture_output=True, text=True, timeout=900)
            pts.append((ns, json.loads(o.stdout.strip().splitlines()[-1])["seconds"]))
        ns, sec = pts[-1]
        res["cpu_reference"] = {"cores": 1, "samples_s": {str(a): b for a, b in pts},
                                "extrapolated_ms_at_n": round(sec * 1e3 * (n / ns) ** 2, 1),
                                "sample": f"unmodified reference, 1 thread, n={ns} x{(n / ns) ** 2:.1f} (quadratic; the n=129 point checks the exponent)"}
        res["speedup_vs_1_core"] = round(res["cpu_reference"]["extrapolated_ms_at_n"] / wall, 1)
    ctx.close()
    text = json.dumps(res, indent=1)
    print(text)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text + "\n")


if __name__ == "__main__":
    main()
