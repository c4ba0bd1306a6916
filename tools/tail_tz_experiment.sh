#!/bin/bash
# One-GPU A/B of the two-layer tail kernels on the headline workload: window-GEMM tcgen05 tail (default) in both tile
# orders against the warp-MMA tail; prints step time and the per-layer times of each variant.
run() {
  name=$1; shift
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-io-stage --no-side > gpurun_out/r02_tz_$name.json 2> gpurun_out/r02_tz_$name.err
  python - $name <<'PY'
import json, sys
n = sys.argv[1]
try:
  d = json.load(open(f"gpurun_out/r02_tz_{n}.json"))
  print("%-22s step %.4f ms  " % (n, d["ms_per_step"]), json.dumps(d["config"]["layers_ms"]))
except Exception as e:
  print(n, "FAILED", e, open(f"gpurun_out/r02_tz_{n}.err").read()[-800:])
PY
}
run mma SNTC_TAIL_TZ=0
run tz_reverse SNTC_TAIL_TZ=1 SNTC_TAIL_TZ_REVERSE=1
run tz_forward SNTC_TAIL_TZ=1 SNTC_TAIL_TZ_REVERSE=0
run mma2 SNTC_TAIL_TZ=0
