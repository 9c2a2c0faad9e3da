#!/bin/bash
# Build K2 variants (template parameters of the queue kernel) into tune/*.so -- run HERE (CPU, nvcc) with "build",
# then on the GPU box with "run": parity on a small case first (under timeout), then the C3-sized timing.
set -u
cd "$(dirname "$0")/.."
VARIANTS=${VARIANTS:-"16x2 12x3 8x4 10x4"}          # heavy[-light] as QGxNG, e.g. 10x4-8x5
if [ "${1:-}" = "build" ]; then
  mkdir -p tune
  for v in $VARIANTS; do
    h=${v%-*}; l=${v#*-}
    qg=${h%x*}; ng=${h#*x}; qgl=${l%x*}; ngl=${l#*x}
    touch tiktaalik_b200/csrc/assemble.cu
    make -s -C tiktaalik_b200/csrc EXTRA="-DTKB_QQG=$qg -DTKB_QQN=$ng -DTKB_QQG_L=$qgl -DTKB_QQN_L=$ngl ${XFLAGS:-}" OUT=../../tune/lib_${PREFIX:-q}$v.so || exit 1
    grep -A2 "assemble_cst_queue_kernelILi18" tiktaalik_b200/csrc/assemble.ptxas.log | tail -2
  done
  touch tiktaalik_b200/csrc/assemble.cu
  make -s -C tiktaalik_b200/csrc || exit 1           # the default build back in place
  exit 0
fi
mkdir -p gpurun_out
out=gpurun_out/${TUNE_LOG:-r2_tune_k2.log}
: > $out
for lib in ${LIBS:-tune/lib_q*.so}; do
  echo "=== $lib" >> $out
  TKB_LIB=$PWD/$lib timeout 180 python tools/knob_case.py >> $out 2>&1 || { echo "PARITY RUN FAILED rc=$?" >> $out; continue; }
  TKB_LIB=$PWD/$lib timeout 300 python tools/profile_case.py 256 160 2>&1 | grep -E "stage ms|K2 queue|K2 pipe" >> $out
done
echo "=== pipe (round-1 form)" >> $out
TKB_K2_MODE=pipe timeout 300 python tools/profile_case.py 256 160 2>&1 | grep -E "stage ms" >> $out
