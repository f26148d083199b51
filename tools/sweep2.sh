#!/bin/bash
run() { python bench.py --steps 100 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.0f  step %.3f ms  raster %.3f ms (frac %.3f)  spectrum %.3f ms  e2e %.0f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['spectrum_kernel_ms'], d['e2e']['value']))"; }
echo -n "default: "; run
echo -n "NO_TAPTAB: "; GLAVA_B200_NO_TAPTAB=1 run
for r in 135 540 1080; do echo -n "ROWS=$r: "; GLAVA_B200_ROWS=$r run; done
for b in 160 192; do echo -n "BX=$b: "; GLAVA_B200_BX=$b run; done
