#!/bin/bash
# Per-phase cycle counts of CTA 0 (clock64 between the barriers), printed by a -DLDPC_PHASE_TIMING build.
#   here:        scripts/phase_timing.sh build
#   on the GPU:  scripts/phase_timing.sh run [codes...]
set -e
cd "$(dirname "$0")/.."
LIB=scratch/libldpc_timing.so
if [ "$1" = build ]; then
  mkdir -p scratch
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DLDPC_PHASE_TIMING \
    -o $LIB fixedpointldpc_b200/csrc/ldpc_decoder.cu fixedpointldpc_b200/csrc/ldpc_encode.cu fixedpointldpc_b200/csrc/ldpc_code.cpp
  exit 0
fi
shift || true
for c in ${@:-wifi a5}; do
  echo "== $c (30 iterations, then the operating point)"
  LDPC_B200_LIB=$PWD/$LIB python bench.py --code $c --steps 1 --warmup 3 --no-cpu --frames 65536 --e2e-frames 1024 2>&1 | grep "phase cycles" | sed -n '4p;6p'
done
