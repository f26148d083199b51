#!/bin/bash
# BASELINE configs[1], [2] and points of [4] on N GPUs of one box (N = $1): one bench.py line each into gpurun_out/scale_N.jsonl
# usage (from the repo root, on the GPU box):  bash tools/scale_runs.sh 8      (REDUCED=1: configs[1], [2] and one point of [4] only)
N=${1:-1}
OUT=gpurun_out/scale_$N.jsonl
mkdir -p gpurun_out; : > $OUT
run() {
  if [ "$N" = "1" ]; then python bench.py --gpus 1 "$@" >> $OUT 2>> gpurun_out/scale_$N.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" >> $OUT 2>> gpurun_out/scale_$N.err; fi
}
run --steps 100 --warmup 5 --no-cpu-baseline                                   # configs[1] headline (with the full-ring and lazy_smooth=0 legs)
run --steps 50 --warmup 3 --no-cpu-baseline --no-extras --config radial4k      # configs[2]
if [ -z "$REDUCED" ]; then
for pt in sweep:512:1280x720 sweep:2048:1920x1080 sweep:8192:1920x1080 sweep:16384:3840x2160 sweep:4096:7680x4320; do   # configs[4]
  run --steps 30 --warmup 3 --no-cpu-baseline --no-extras --config $pt
done
run --steps 50 --warmup 3 --no-cpu-baseline --no-extras --config graph720      # configs[3] modules, throughput
run --steps 50 --warmup 3 --no-cpu-baseline --no-extras --config wave720
else
run --steps 30 --warmup 3 --no-cpu-baseline --no-extras --config sweep:8192:1920x1080   # REDUCED=1: configs[1], [2] and one point of [4]
fi
python - <<PY
import json
for l in open("$OUT"):
    try: d = json.loads(l)
    except Exception: continue
    print("%-52s n_gpus %d  value %10.0f f/s  e2e %10.0f f/s  whole-step frac %.3f  raster frac %.3f" % (d["metric"][22:], d["n_gpus"], d["value"], d["e2e"]["value"], d["roofline"]["whole_step_frac"], d["roofline"]["frac"]))
PY
