#!/bin/bash
# The round's measurements on ONE GPU of a fresh box (tools/scale_runs.sh N for N > 1); small outputs only (gpurun_out is
# capped at 64 MiB).  part A: benches, sweeps, cadence.  part B (tools/final_profiles.sh): ncu launch lists + full captures.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/gpu_tests_r2.log
[ -n "$SKIP_REF" ] || python bench.py --impl reference --steps 8 --warmup 1 > gpurun_out/bench_ref_r2.json 2> gpurun_out/bench_ref_r2.err
python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
[ -n "$SKIP_SCALE1" ] || bash tools/scale_runs.sh 1 > gpurun_out/scale_1.log 2>&1
python tools/sweep_configs.py > gpurun_out/sweep_r2.log 2>&1
rm -f gpurun_out/modules_r2.jsonl
for m in circle1080 radial1080 graph1080 wave1080; do python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-extras --config $m >> gpurun_out/modules_r2.jsonl 2>> gpurun_out/modules_r2.err; done
python tools/cadence.py > gpurun_out/cadence.log 2>&1
python tools/config1_latency.py > gpurun_out/config1.log 2>&1
[ -n "$SKIP_PROBES" ] || bash tools/prio_probe.sh > gpurun_out/prio_probe.log 2>&1
ls -la gpurun_out
