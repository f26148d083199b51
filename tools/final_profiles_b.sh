#!/bin/bash
# ncu evidence for the kernels that changed late in the round (one GPU): launch lists of the headline / graph / wave / circle
# steps, full captures of the new kernels.  Reports are exported to CSV on the box and deleted (64 MiB cap on gpurun_out).
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
for c in headline graph1080 wave1080 circle1080; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_$c.csv $B --config $c > /dev/null 2>&1
done
cap() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o /tmp/$name "$@" > /dev/null 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
}
cap r2b_headline "k5_need_smem|epilogue_b|spectrum_kernel" 9 3 $B
cap r2b_graph "raster_graph|column_table" 4 2 $B --config graph1080
cap r2b_wave "raster_wave|column_table" 4 2 $B --config wave1080
cap r2b_circle "raster_circle" 2 1 $B --config circle1080
ls -la gpurun_out | tail -12
