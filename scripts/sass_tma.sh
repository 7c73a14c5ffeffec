#!/bin/bash
# development helper: compile quantize_tma.cu (benchmark variant only) and list the SASS of the kernel
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Iinclude -DVBQ_DEV_ONE -cubin -Xptxas -v \
  vbq_b200/csrc/quantize_tma.cu -o /tmp/tma.cubin 2>&1 | grep -E "error|registers|spill" | grep -v " 0 bytes spill" || true
cuobjdump -sass /tmp/tma.cubin | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//' > /tmp/tma.txt
wc -l /tmp/tma.txt
