#!/bin/bash
# ncu --set full of the attention kernel alone.  usage: gpurun -- bash scripts/profile_attention.sh <tag> [B T H hd]
TAG=$1; shift
CMD="python scripts/prof_attention.py ${*:-32 1499 16 80} 2"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'attention_v' -s 2 -c 1 -f \
    -o gpurun_out/${TAG} $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/${TAG}_ncu.log
