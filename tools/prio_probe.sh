#!/bin/bash
# development aid: does the three-kernel spectrum path overlap better with the raster kernel at another stream priority?
PTS="bars:4096:1920x1080 bars:8192:1920x1080 bars:16384:1920x1080 bars:8192:1280x720 radial:8192:3840x2160"
for pr in low equal high; do
  echo "== GLAVA_B200_SPEC_PRIO=$pr"
  GLAVA_B200_SPEC_PRIO=$pr python tools/sweep_configs.py $PTS 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print('  %-6s n=%5d %dx%4d step %.3f ms  whole-step %.3f' % (r['module'], r['bufsize'], r['width'], r['height'], r['step_ms'], r['whole_step_frac_of_hbm_peak']))"
done
