#!/bin/bash
# usage: tools/gpu.sh <timeout-seconds> '<command>'   -- rebuild the library here (cross-compile), then run on a B200 box
set -e
cd "$(dirname "$0")/.."
make -s -j8 -C deep-gan-encoders_b200/csrc all 2>&1 | grep -v "spill" || true
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
