#!/bin/bash
# Development aid: build the host-only sources of libglava_b200 (config reader + colour compiler, --pipe parser, audio
# boundary) with AddressSanitizer + UBSan against CUDA-free stubs and drive them with generated and hostile input.
#   tools/san_host_check.sh        -> "sanitizer run clean" or an ASan / UBSan report
set -e
cd "$(dirname "$0")/.."
OUT=${TMPDIR:-/tmp}/libglava_host_san.so
g++ -std=c++17 -g -O1 -fPIC -shared -fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer \
    -ffp-contract=off -Wall -Wno-unused-function -Wno-unknown-pragmas -include tools/san/cuda_types.h \
    -o "$OUT" glava_b200/csrc/config.cpp glava_b200/csrc/pipe.cpp glava_b200/csrc/audio.cpp tools/san/stubs.cpp -lpthread
LD_PRELOAD=$(g++ -print-file-name=libasan.so):$(g++ -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 \
    python tools/san/drive.py "$OUT"
