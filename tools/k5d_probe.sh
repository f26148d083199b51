#!/bin/bash
# development aid: need-list K5 out of shared-memory tiles (k5_need_smem_kernel) against the L2 form (GLAVA_B200_K5N_SMEM=0):
# parity, per-kernel times from ncu launch lists at several tile heights (GLAVA_B200_K5N_ROWS: -per cent of the tallest window),
# then the whole step at the points it matters
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_spectrum.py tests/test_llvmpipe_golden.py tests/test_gpu_masked.py -m gpu -q -x 2>&1 | tail -2
ll() { ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv python tools/spec_probe.py --one $1 2>/dev/null | grep -E "k5_need|av_transpose" | sed -E "s/.*glb::([a-z0-9_]+)(<[^>]*>)?.*\"ns\",\"([0-9]+)\"/\1 \3/" | tail -2 | tr '\n' ' '; echo; }
for n in ${SIZES:-4096 8192 16384}; do
  [ -n "$WITH_L2" ] && { echo -n "n=$n L2 form: "; GLAVA_B200_K5_SPLIT=1 GLAVA_B200_K5N_SMEM=0 ll $n; }
  for rows in ${ROWS:--105 -150}; do for wpc in ${WPC:-16}; do
    echo -n "n=$n tiles rows=$rows warps=$wpc: "; GLAVA_B200_K5_SPLIT=1 GLAVA_B200_K5N_ROWS=$rows GLAVA_B200_K5N_WARPS=$wpc ll $n
  done; done
done
PTS="bars:4096:1920x1080 bars:8192:1920x1080 bars:8192:1280x720 bars:16384:1920x1080 bars:4096:1280x720 radial:8192:3840x2160"
[ -n "$WITH_L2" ] && {
echo "== whole step, L2 form"; GLAVA_B200_K5N_SMEM=0 python tools/sweep_configs.py $PTS 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  %-6s n=%5d %4dx%4d step %.4f ms whole %.3f' % (d['module'], d['bufsize'], d['width'], d['height'], d['step_ms'], d.get('whole_step_frac_of_hbm_peak', 0)))
"
}
for rows in ${STEP_ROWS:--105}; do
echo "== whole step, tiles rows=$rows"; GLAVA_B200_K5N_ROWS=$rows python tools/sweep_configs.py $PTS 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  %-6s n=%5d %4dx%4d step %.4f ms whole %.3f' % (d['module'], d['bufsize'], d['width'], d['height'], d['step_ms'], d.get('whole_step_frac_of_hbm_peak', 0)))
"
done
