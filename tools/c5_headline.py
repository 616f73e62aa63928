"""C5's headline shape (n = 2^16 constraints, N_R = N_E = 2^15, 55-bit ring prime) through the blocked witness kernels:
  * rsg_interpolate of ONE vector of n ring elements at the full ring degree (16 GiB in, 16 GiB out);
  * rsg_witness_map (6 interpolations + 2 by linearity or 8 without an R1CS, quotient) at N_R = 4096 slots (1/8 of the ring:
    what one rank of an 8-GPU slot-sharded run holds), with the identity A(r)B(r) - C(r) = H(r)Z(r) checked on two slots.
Prints one JSON object.  python tools/c5_headline.py [--slots 4096] [--full 1]"""
import argparse
import json
import random
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import ringsnark_b200 as rs                        # noqa: E402
from ringsnark_b200.params import CONFIGS          # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--slots", type=int, default=4096)
ap.add_argument("--full", type=int, default=1)
ap.add_argument("--n", type=int, default=65536)
ap.add_argument("--check", type=int, default=1)
ap.add_argument("--sweep-ctas", type=int, default=0)
args = ap.parse_args()
cfg = CONFIGS["c5s"]
n, p = args.n, int(cfg["q"][0])
out = {"n": n, "prime_bits": p.bit_length(), "N_E": cfg["N_E"]}


def kernel_ms(ctx):
    res = {}
    for k in ("k_interp_fast", "k_interp_fast_const", "k_quotient_fast", "k_full_from_parts"):
        ms, cnt = ctx.timing(k)
        if cnt:
            res[k] = {"ms": round(ms, 2), "launches": cnt}
    return res


if args.full:
    ctx = rs.Context(cfg["N_R"], cfg["q"], cfg["N_E"], cfg["Q"])
    y = ctx.ringvec(n)
    y.fill_uniform(1)
    c = ctx.ringvec(n)
    t0 = time.time()
    ctx.interpolate(n, y, out=c)                   # first call builds the per-(n, q) tables on the host
    ctx.sync()
    out["first_call_s_with_host_tables"] = round(time.time() - t0, 2)
    ctx.enable_timing(True)
    t0 = time.time()
    ctx.interpolate(n, y, out=c)
    ctx.sync()
    out["interpolate_one_vector_full_ring"] = {"N_R": cfg["N_R"], "wall_s": round(time.time() - t0, 3), "kernels_ms": kernel_ms(ctx)}
    # cheap check on every slot: the constant coefficient is the evaluation at node 0 (full parity: tests/test_gpu_parity.py)
    assert np.array_equal(c.download(0, 1), y.download(0, 1))
    del y, c
    ctx.close()

N_R = args.slots
ctx = rs.Context(N_R, cfg["q"], cfg["N_E"], cfg["Q"])
if args.sweep_ctas:
    import os
    y, c = ctx.ringvec(n), ctx.ringvec(n)
    y.fill_uniform(3)
    ctx.interpolate(n, y, out=c)
    ctx.sync()
    out["ctas_per_sm_sweep"] = {"N_R": N_R, "ms_one_vector": {}}
    for k in (1, 2, 3, 4, 6, 8):
        os.environ["RSG_WF_BIG_CTAS"] = str(k)
        ctx.interpolate(n, y, out=c)
        ctx.sync()
        t0 = time.time()
        ctx.interpolate(n, y, out=c)
        ctx.sync()
        out["ctas_per_sm_sweep"]["ms_one_vector"][k] = round((time.time() - t0) * 1e3, 1)
    del os.environ["RSG_WF_BIG_CTAS"]
    del y, c
ev = ctx.ringvec(9 * n)
ev.fill_uniform(2)
e = ev.download(6 * n, 3 * n)
a, b = e[:n, :2].astype(object), e[n:2 * n, :2].astype(object)
# C = A o B on the domain for slots 0 and 1 (the identity is checked there); the other slots keep random C
e[2 * n:, :2] = ((a * b) % p).astype(np.uint64)
ev.upload(e, first=6 * n)
coeffs, H = ctx.ringvec(6 * n), ctx.ringvec(n + 1)
ctx.witness_map(n, ev, coeffs, H)
ctx.sync()
ctx.enable_timing(True)
t0 = time.time()
ctx.witness_map(n, ev, coeffs, H)
ctx.sync()
wall = time.time() - t0
out["witness_map"] = {"N_R": N_R, "wall_s": round(wall, 3), "kernels_ms": kernel_ms(ctx),
                      "full_ring_estimate_s": round(wall * cfg["N_R"] / N_R, 2)}
if not args.check:
    print(json.dumps(out))
    sys.exit(0)
full = ctx.interpolate(n, ev, batch=3, y_first=6 * n).download()
Hg = H.download()
rnd = random.Random(9)


def horner(col, x):
    acc = 0
    for v in reversed(col):
        acc = (acc * x + int(v)) % p
    return acc


for slot in (0, 1):
    r = rnd.randrange(n, p)
    A, B, Cc = (horner(full[k * n:(k + 1) * n, slot], r) for k in range(3))
    Z = 1
    for i in range(n):
        Z = Z * (r - i) % p
    assert (A * B - Cc) % p == horner(Hg[:n - 1, slot], r) * Z % p, slot
out["witness_map"]["identity_checked_slots"] = 2
print(json.dumps(out))
