#!/bin/bash
# Everything the round's profiles/ records come from, on ONE GPU of a fresh box (tools/scale_runs.sh N for N > 1).
# Writes under gpurun_out/: bench_r2.json, bench_ref_r2.json, scale_1.jsonl, sweep.json, cadence.json, config1_latency.json,
# r2_launches.csv (+ per-config launch lists), r2_full.ncu-rep, r2_modules.ncu-rep
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/gpu_tests_r2.log
python bench.py --impl reference --steps 8 --warmup 1 > gpurun_out/bench_ref_r2.json 2> gpurun_out/bench_ref_r2.err
python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
bash tools/scale_runs.sh 1 > gpurun_out/scale_1.log 2>&1
python tools/sweep_configs.py > gpurun_out/sweep_r2.log 2>&1
for m in circle1080 radial1080 graph1080 wave1080; do python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-extras --config $m >> gpurun_out/modules_r2.jsonl 2>> gpurun_out/modules_r2.err; done
GLAVA_B200_NO_CTILE=1 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-extras --config circle1080 > gpurun_out/circle_noctile.json 2>/dev/null
python tools/cadence.py > gpurun_out/cadence.log 2>&1
python tools/config1_latency.py > gpurun_out/config1.log 2>&1
# launch lists (per-launch device time, cold cache, serialised): the kernels' SHARES of a step
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_radial4k.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --config radial4k > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_sweep8192.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --config sweep:8192:1920x1080 > /dev/null 2>&1
# full captures: the headline's two kernels, then one launch of every other kernel of the path
ncu --set full --clock-control none --import-source on -k regex:"raster_bars|spectrum_kernel" -s 6 -c 4 -f -o gpurun_out/r2_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"raster_radial_geo|spectrum_kernel|epilogue_b|av_transpose|k5_need" -s 10 -c 5 -f -o gpurun_out/r2_radial4k python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --config radial4k > /dev/null 2>&1
for m in circle graph wave; do
  ncu --set full --clock-control none --import-source on -k regex:"raster_${m}|texmm" -s 4 -c 2 -f -o gpurun_out/r2_$m python tools/sweep_configs.py $m:4096:1920x1080 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:"k5_table|fifo_ingest" -c 2 -f -o gpurun_out/r2_misc python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out
