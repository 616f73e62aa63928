#!/bin/bash
# tools/bench_modes.sh TAG "ENV1=a ENV2=b" "ENV3=c" ...  -- one bench.py run per environment setting (GPU box), one JSON line each
# in gpurun_out/TAG_bench_<setting>.json, plus a one-line summary per run on stdout.
tag=$1; shift
mkdir -p gpurun_out
for mode in "$@"; do
  name=$(echo "$mode" | tr ' =' '__')
  env $mode python bench.py --steps 10 --warmup 3 > "gpurun_out/${tag}_bench_${name}.json" 2> "gpurun_out/${tag}_bench_${name}.err"
  python - "gpurun_out/${tag}_bench_${name}.json" "$mode" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["kernels_ms_per_step"]
    print(f"{sys.argv[2]:44s} value {d['value']:.3f} e2e {d['e2e']['value']:.3f} ntt {k['k_lift_fwd_ntt']:.3f} lin {k['k_crs_lincomb']:.3f} enc {k['k_encode_intt']:.3f} "
          f"interp {k['k_interp_fast']:.3f} parity {d.get('parity', {}).get('ok')} checksum {d.get('proof_checksum')}")
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
done
