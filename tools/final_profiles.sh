#!/bin/bash
# ncu evidence of the round (one GPU).  (tools/sass_segments.py wants the source page of a SINGLE-kernel capture:
# ncu -k regex:spectrum_kernel … tools/spec_probe.py --one 8192, as for profiles/r2a_before_spectrum8192_stall_segments.txt.)  Reports are exported to CSV on the box and deleted (64 MiB cap on gpurun_out).
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
# launch lists (per-launch device time, cold cache, serialised): the kernels' SHARES of a step
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_radial4k.csv $B --config radial4k > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_sweep8192.csv $B --config sweep:8192:1920x1080 > /dev/null 2>&1
cap() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o /tmp/$name "$@" > /dev/null 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  if [ "$5" != "" ]; then :; fi
}
cap r2_headline "raster_bars|spectrum_kernel|epilogue_b|av_transpose|k5_need" 15 5 $B
cap r2_radial4k "raster_radial_geo|spectrum_kernel|epilogue_b|av_transpose|k5_need" 15 5 $B --config radial4k
for m in circle graph wave; do cap r2_$m "raster_${m}|texmm" 4 2 $B --config ${m}1080; done
cap r2_misc "k5_table|fifo_ingest" 0 2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline
ls -la gpurun_out
