import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tools/ -> repo root
if len(sys.argv) > 1 and sys.argv[1] == "one":
    sys.path.insert(0, ROOT)
    import numpy as np
    import ringsnark_b200 as rs
    from ringsnark_b200.params import CONFIGS
    cfg = CONFIGS["c4"]
    T = 2063
    ctx = rs.Context(cfg["N_R"], cfg["q"], cfg["N_E"], cfg["Q"])
    crs = ctx.crs(T); crs.fill_uniform(3)
    vec = ctx.ringvec(T); vec.fill_uniform(4)
    tags = np.full(T, 2, dtype=np.uint8)
    ctx.inner_product(crs, vec, tags, to_host=False)
    ctx.sync(); ctx.enable_timing(True)
    for _ in range(6):
        ctx.inner_product(crs, vec, tags, to_host=False)
    ctx.sync()
    ms = ctx.timing("k_crs_lincomb")[0] / 6
    alg = 8 * 16384 * 8 * (3 * T + 2)
    print(json.dumps({"splits": os.environ.get("RSG_LIN_SPLITS"), "unroll": os.environ.get("RSG_LIN_UNROLL"), "th": os.environ.get("RSG_LIN_THREADS"), "ms": round(ms, 4), "GBs": round(alg / ms / 1e6, 1)}))
    sys.exit(0)
for splits, unroll, th in [(5, 2, 256), (5, 1, 256), (5, 3, 256), (4, 2, 256), (6, 2, 256), (8, 2, 256), (9, 2, 256), (10, 2, 256), (12, 2, 256), (5, 2, 128), (10, 2, 128), (5, 2, 512), (3, 2, 512), (12, 1, 256), (9, 1, 256)]:
    env = dict(os.environ, RSG_LIN_SPLITS=str(splits), RSG_LIN_UNROLL=str(unroll), RSG_LIN_THREADS=str(th))
    r = subprocess.run([sys.executable, __file__, "one"], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-300:], flush=True)
