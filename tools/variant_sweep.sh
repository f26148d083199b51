#!/bin/bash
# Whole-step time of the FFT-kernel variants within the three-kernel spectrum path (development aid).
run() { pts=$1; shift; echo "== $* [$pts]"; env "$@" python tools/sweep_configs.py $pts 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print('  %-6s n=%5d %dx%4d step %.3f ms  whole-step %.3f' % (r['module'], r['bufsize'], r['width'], r['height'], r['step_ms'], r['whole_step_frac_of_hbm_peak']))"; }
P16="bars:16384:1920x1080 bars:16384:1280x720"
P4="bars:4096:1920x1080 bars:4096:1280x720 bars:2048:1280x720"
run "$P16" X=0
run "$P16" GLAVA_B200_SPEC_OOP=1 GLAVA_B200_SPEC_T=512
run "$P16" GLAVA_B200_SPEC_OOP=1 GLAVA_B200_SPEC_T=1024
run "$P16" GLAVA_B200_SPEC_OOP=1 GLAVA_B200_SPEC_T=256
run "$P16" GLAVA_B200_SPEC_OOP=0 GLAVA_B200_SPEC_T=256
run "$P4" X=0
run "$P4" GLAVA_B200_SPEC_OOP=1 GLAVA_B200_SPEC_T=256
run "$P4" GLAVA_B200_SPEC_OOP=1 GLAVA_B200_SPEC_T=128
run "bars:8192:1920x1080 bars:8192:1280x720" GLAVA_B200_K5N_WARPS=8
run "bars:8192:1920x1080 bars:8192:1280x720" GLAVA_B200_K5N_WARPS=16
