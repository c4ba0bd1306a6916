#!/bin/bash
# The multi-GPU bench lines of one box:  tools/multi_gpu_runs.sh N   (N = 2, 4, 8; under `gpurun --gpus N`)
#   headline (configs[1]), configs[2] two_layer_syn2 1200x1200 sharded, configs[4] bls2017 4K frames sharded (whole frames per GPU)
#   and split inside the frame (--tile-frames: one latent-row band per GPU).
N=${1:-2}
run() {  # name port args...
  name=$1; port=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-io-stage "$@" \
    > gpurun_out/r02_scale_${N}gpu_${name}_v5.json 2> gpurun_out/r02_scale_${N}gpu_${name}_v5.err
  python - <<PY
import json
try:
  d = json.load(open("gpurun_out/r02_scale_${N}gpu_${name}_v5.json"))
  print("${name}", "N=%d" % d["n_gpus"], "value %.0f Mpx/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], d["scaling"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
  print("${name} FAILED", e); print(open("gpurun_out/r02_scale_${N}gpu_${name}_v5.err").read()[-1500:])
PY
}
run headline 29531
run syn2_1200 29532 --config two_layer_syn2 --batch 8 --height 1200 --width 1200 --no-side
run bls2017_4k 29533 --config bls2017 --batch 2 --height 2160 --width 3840 --no-side
run bls2017_4k_tiled 29534 --config bls2017 --batch 2 --height 2160 --width 3840 --no-side --tile-frames
