#!/bin/bash
# One-GPU bench lines of the side configurations (BASELINE configs 0, 2, 3, 4): full JSON lines into gpurun_out/r02_side_<name>.json,
# a one-line summary each on stdout.
mkdir -p gpurun_out
for c in "jpegl 1 512 768" "jpegl 24 512 768" "two_layer_syn2 8 1200 1200" "two_layer_syn2:24 8 1200 1200" "two_layer_syn2:48 8 1200 1200" "mbt2018 24 512 768" "bls2017 2 2160 3840"; do
  set -- $c
  name=$(echo "$1_b$2" | tr ':' '_')
  steps=20; [ "$2" = "1" ] && steps=400   # batch 1: the 4 rotating input sets are captured into CUDA graphs on their second use -- amortise that
  timeout 300 python bench.py --config $1 --batch $2 --height $3 --width $4 --steps $steps --warmup 3 --no-cpu-baseline --no-io-stage --no-side > gpurun_out/r02_side_${name}.json 2> gpurun_out/r02_side_${name}.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
  d = json.load(open(f"gpurun_out/r02_side_{name}.json"))
  r = d["roofline"] or {}
  print(name, "value %.0f Mpx/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], "| dominant", r.get("kernel"),
        "%.0f TFLOP/s" % r.get("achieved", 0), "frac %.3f (%s)" % (r.get("frac", 0), r.get("peak_regime")), "|", json.dumps(d["config"]["layers_ms"]))
except Exception as e:
  print(name, "FAILED", e, open(f"gpurun_out/r02_side_{name}.err").read()[-800:])
PY
done
