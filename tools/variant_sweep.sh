#!/bin/bash
# Whole-step time of the spectrum-kernel variants at the points where the spectrum kernel matters (development aid).
PTS="bars:2048:1280x720 bars:4096:1280x720 bars:4096:1920x1080 bars:8192:1280x720 bars:8192:1920x1080 bars:16384:1280x720 bars:16384:1920x1080 graph:4096:1920x1080"
run() { echo "== $*"; env "$@" python tools/sweep_configs.py $PTS 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print(f\"  {r['module']:6s} n={r['bufsize']:5d} {r['width']}x{r['height']:4d} step {r['step_ms']:.3f} ms  whole-step {r['whole_step_frac_of_hbm_peak']:.3f}\")"; }
run X=0
run GLAVA_B200_K5_SPLIT=1
run GLAVA_B200_SPEC_OOP=1 GLAVA_B200_SPEC_T=256
run GLAVA_B200_SPEC_OOP=1 GLAVA_B200_SPEC_T=128
run GLAVA_B200_SPEC_OOP=1 GLAVA_B200_SPEC_T=256 GLAVA_B200_K5_SPLIT=1
run GLAVA_B200_SPEC_OOP=1 GLAVA_B200_SPEC_T=128 GLAVA_B200_K5_SPLIT=1
run GLAVA_B200_SPEC_OOP=1 GLAVA_B200_SPEC_T=512 GLAVA_B200_K5_SPLIT=1
