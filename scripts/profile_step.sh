#!/bin/bash
# ncu evidence for one bench step (run under gpurun, 1 GPU):
#   1. launch list of the whole bench command (gpu__time_duration per launch, cold-cache and serialised: SHARES matter)
#   2. one `--set full` capture of the first kernels of a step (FE layer 0, the 6 conv GEMMs, LN, projection, pos-conv,
#      one full encoder layer: LN, QKV, attention, out-proj, LN, FFN1, FFN2)
# usage: gpurun --timeout 900 -- bash scripts/profile_step.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 1 --skip-e2e --no-cpu-baseline"
# kernels of ours matched per step: fe_layer0 1 + conv 6 + LN 1 + proj 1 + posconv 1 + 48 x 7 + final LN 1 + CTC 1 = 348
$CMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${TAG}_launches.csv $CMD \
    > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > $OUT/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on \
    -k regex:'gemm_kernel|attention_v|layernorm|fe_layer0' -s 348 -c 17 -f -o $OUT/${TAG}_full $CMD \
    > $OUT/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la $OUT
