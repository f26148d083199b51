#!/bin/bash
for t in 4 16 32; do echo "CIRCLE_TILES=$t"; GLAVA_B200_CIRCLE_TILES=$t python tools/gpu_probe.py 2>&1 | grep circle; done
for r in 135 240 360 720; do echo "ROWS=$r"; GLAVA_B200_ROWS=$r python tools/gpu_probe.py 2>&1 | grep -E "graph|wave|radial"; done
