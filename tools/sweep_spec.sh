#!/bin/bash
# development sweep: persistent-grid size of the spectrum kernel vs whole-step throughput
for k in 0 1 2 3 4; do
  echo -n "SPEC_CTAS_PER_SM=$k: "
  GLAVA_B200_SPEC_CTAS_PER_SM=$k python bench.py --steps 100 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.0f  step %.3f ms  raster %.3f ms (frac %.3f)  spectrum %.3f ms  e2e %.0f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['spectrum_kernel_ms'], d['e2e']['value']))"
done
