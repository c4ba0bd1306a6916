#!/bin/bash
# One-GPU A/B runs of environment switches on the headline workload: prints step time and the per-layer times of each variant.
run() {
  name=$1; shift
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-io-stage --no-side > gpurun_out/r02_env_$name.json 2> gpurun_out/r02_env_$name.err
  python - $name <<'PY'
import json, sys
n = sys.argv[1]
try:
  d = json.load(open(f"gpurun_out/r02_env_{n}.json"))
  print("%-22s step %.4f ms  " % (n, d["ms_per_step"]), json.dumps(d["config"]["layers_ms"]))
except Exception as e:
  print(n, "FAILED", e, open(f"gpurun_out/r02_env_{n}.err").read()[-500:])
PY
}
run base A=1
run narrow_waves_100 SNTC_TC_NARROW_WAVES=100
run narrow_waves_200 SNTC_TC_NARROW_WAVES=200
run narrow_waves_400 SNTC_TC_NARROW_WAVES=400
run pdl_on SNTC_TC_PDL=1
run base2 A=1
