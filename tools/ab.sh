#!/bin/bash
# A/B of two builds of the engine on the GPU box: tools/ab.sh TAG "libA libB ..." "workload ..."   (paths relative to the
# package directory; the empty name "-" is the in-tree libflicb200.so).  Per-kernel CUDA-event times, round-trip checked.
TAG=$1; LIBS=${2:-"exp/base.so -"}; WL=${3:-"C2x64"}
mkdir -p gpurun_out
PKG=$PWD/fast-losless-image-compression-format_b200
for rep in 1 2; do
  for w in $WL; do
    for l in $LIBS; do
      if [ "$l" = "-" ]; then unset FLIC_LIB; else export FLIC_LIB=$PKG/$l; fi
      echo -n "[$l] " | tee -a gpurun_out/${TAG}_ab.txt
      timeout 300 python tools/ktime.py $w 20 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_ab.txt
    done
  done
done
unset FLIC_LIB
