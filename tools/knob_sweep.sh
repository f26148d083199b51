#!/bin/bash
# Development aid: the tuning-knob sweeps behind DESIGN.md §6.2, one sub-command each (run on a B200 via gpurun).
#   tools/knob_sweep.sh rows | overlap | spec_ctas | resident | circle_tiles | module_rows
run() { python bench.py --steps 100 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.0f  step %.3f ms  raster %.3f ms (frac %.3f, iso %.3f)  spectrum %.3f ms  e2e %.0f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline'].get('frac_isolated', 0), d['roofline']['spectrum_kernel_ms'], d['e2e']['value']))"; }
case "$1" in
  rows)         echo -n "default: "; run; echo -n "NO_TAPTAB: "; GLAVA_B200_NO_TAPTAB=1 run
                for r in 135 540 1080; do echo -n "ROWS=$r: "; GLAVA_B200_ROWS=$r run; done ;;
  overlap)      echo -n "default (prio high): "; run; echo -n "NO_OVERLAP: "; GLAVA_B200_NO_OVERLAP=1 run
                for pr in equal low; do echo -n "prio $pr: "; GLAVA_B200_SPEC_PRIO=$pr run; done ;;
  spec_ctas)    for k in 0 1 2 3 4; do echo -n "SPEC_CTAS_PER_SM=$k: "; GLAVA_B200_SPEC_CTAS_PER_SM=$k run; done ;;
  resident)     for c in 0 1 2 3; do for pr in high low; do echo -n "RESIDENT=$c prio=$pr: "; GLAVA_B200_SPEC_RESIDENT=$c GLAVA_B200_SPEC_PRIO=$pr run; done; done ;;
  circle_tiles) for t in 4 16 32; do echo "CIRCLE_TILES=$t"; GLAVA_B200_CIRCLE_TILES=$t python tools/gpu_probe.py 2>&1 | grep circle; done ;;
  module_rows)  for r in 135 240 360 720; do echo "ROWS=$r"; GLAVA_B200_ROWS=$r python tools/gpu_probe.py 2>&1 | grep -E "graph|wave|radial"; done ;;
  *) sed -n 2,3p "$0" ;;
esac
