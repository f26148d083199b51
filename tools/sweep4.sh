#!/bin/bash
run() { python bench.py --steps 100 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.0f  step %.3f ms  raster %.3f ms (frac %.3f, iso %.3f)  spectrum %.3f ms  e2e %.0f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['frac_isolated'], d['roofline']['spectrum_kernel_ms'], d['e2e']['value']))"; }
echo -n "default (prio high): "; run
echo -n "NO_OVERLAP: "; GLAVA_B200_NO_OVERLAP=1 run
echo -n "prio equal: "; GLAVA_B200_SPEC_PRIO=equal run
echo -n "prio low: "; GLAVA_B200_SPEC_PRIO=low run
