#!/bin/bash
# ncu --set full of our GEMM and cuBLAS on the same shape.  usage: gpurun -- bash scripts/profile_gemm.sh <tag> M N K epi
TAG=$1; shift
CMD="python scripts/prof_gemm_vs_cublas.py $*"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|nvjet|cutlass|sm100|xmma|gemm' -s 4 -c 2 -f \
    -o gpurun_out/${TAG} $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/${TAG}_ncu.log
