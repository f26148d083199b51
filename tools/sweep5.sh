#!/bin/bash
run() { python bench.py --steps 100 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.0f  step %.3f ms  raster %.3f ms (frac %.3f, iso %.3f)  spectrum %.3f ms  e2e %.0f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['frac_isolated'], d['roofline']['spectrum_kernel_ms'], d['e2e']['value']))"; }
for c in 0 1 2 3; do for pr in high low; do echo -n "RESIDENT=$c prio=$pr: "; GLAVA_B200_SPEC_RESIDENT=$c GLAVA_B200_SPEC_PRIO=$pr run; done; done
