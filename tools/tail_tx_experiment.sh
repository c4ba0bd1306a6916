#!/bin/bash
# Tail-kernel tile width experiment: the same bench with libsntc built at -DSNTC_TAIL_TX=16 / 32 (build/libsntc_tx*.so) and the default 64.
for tx in 64 32 16; do
  lib=""; [ $tx != 64 ] && lib="$PWD/build/libsntc_tx$tx.so"
  SNTC_LIB_PATH=$lib python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-io-stage --no-side > gpurun_out/r02_tail_tx$tx.json 2> gpurun_out/r02_tail_tx$tx.err
  python - $tx <<'PY'
import json, sys
tx = sys.argv[1]
try:
  d = json.load(open(f"gpurun_out/r02_tail_tx{tx}.json"))
  print("TX", tx, "step %.4f ms" % d["ms_per_step"], "value %.0f" % d["value"], json.dumps(d["config"]["layers_ms"]))
except Exception as e:
  print("TX", tx, "FAILED", e, open(f"gpurun_out/r02_tail_tx{tx}.err").read()[-600:])
PY
done
SNTC_LIB_PATH=$PWD/build/libsntc_tx16.so python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "warp_mma_tail or full_size_config2 or decode_matches_oracle" 2>&1 | tail -3
