import os, sys, subprocess, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from rsgv import Case
import ringsnark_b200 as rs
seed = sys.argv[1] if len(sys.argv) > 1 else "11"
subprocess.check_call([os.path.join(ROOT, "oracle/_ref/ref_harness"), "dump", "tiny_quirks", "/tmp/q.rsgv", seed], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
case = Case("/tmp/q.rsgv")
ctx = rs.Context(case.N_R, case.q, case.N_E, case.Q)
n = case.n
r1cs = rs.R1cs(ctx, n, case.io, case.aux, case.d["r1cs_row_ptr"], case.d["r1cs_col"], case.d["r1cs_coeff"])
assign = ctx.ringvec_from(np.concatenate([case.ring("primary_input")[0], case.ring("auxiliary_input")[0]]))
ev = r1cs.evaluate(assign)
evd = ev.download()
order = ["A_mid", "B_mid", "C_mid", "A_io", "B_io", "C_io", "A_full", "B_full", "C_full"]
for k, name in enumerate(order):
    print("eval", name, np.array_equal(evd[k * n:(k + 1) * n], case.ring("eval_" + name)[0]))
coeffs, H = ctx.witness_map(n, ev)
got = coeffs.download()
for idx, k in enumerate(["A_io", "B_io", "C_io", "A_mid", "B_mid", "C_mid"]):
    print("wit", k, np.array_equal(got[idx * n:(idx + 1) * n], case.ring("wit_" + k)[0]))
print("H", np.array_equal(H.download(), case.ring("wit_H")[0]))
s_pows = ctx.crs_from(case.enc("crs_s_pows")[0])
ip, ip_size = case.enc("ip")
for k, (first, nm) in enumerate([(0, "A_io"), (3 * n, "A_mid"), (n, "B_io"), (4 * n, "B_mid")]):
    tags = ctx.term_tags(coeffs, first=first, count=n)
    w, t, s = case.ring("wit_" + nm)
    import oracle_lib as O
    print(nm, "tags gpu", list(tags), "ref", list(O.term_tags(w, t, s)))
    out, used = ctx.inner_product(s_pows, coeffs, tags, coeff_first=first)
    print("ip", nm, used, np.array_equal(out, ip[k]), int(ip_size[k][0]))
