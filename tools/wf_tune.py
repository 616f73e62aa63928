import os, sys, itertools, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tools/ -> repo root
if len(sys.argv) > 1 and sys.argv[1] == "one":
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    import ringsnark_b200 as rs
    from ringsnark_b200.params import CONFIGS
    cfg = CONFIGS["c4"]
    n = int(os.environ.get("WF_N", cfg["n"]))
    ctx = rs.Context(cfg["N_R"], cfg["q"], cfg["N_E"], cfg["Q"])
    ev = ctx.ringvec(9 * n); ev.fill_uniform(1)
    coeffs, H = ctx.witness_map(n, ev)
    ctx.sync()
    ctx.enable_timing(True)
    for _ in range(10):
        ctx.witness_map(n, ev, coeffs, H)
    ctx.sync()
    out = {k: round(ctx.timing(k)[0] / 10, 4) for k in ("k_interp_fast", "k_quotient_fast")}
    print(json.dumps({"sl": os.environ.get("RSG_WF_SL"), "thr": os.environ.get("RSG_WF_THREADS"), "n": n, **out}))
    sys.exit(0)
for sl, thr in [(4, 512), (2, 256), (2, 128), (1, 128)]:
    env = dict(os.environ, RSG_WF_SL=str(sl), RSG_WF_THREADS=str(thr), RSG_WITNESS="fast")
    r = subprocess.run([sys.executable, __file__, "one"], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-300:], flush=True)
