#!/bin/bash
# Multi-GPU evidence of one box (run under `gpurun --gpus N`):  bash scripts/multi_gpu_evidence.sh N [tests]
#   1. (with "tests") tensor-parallel parity against the oracle on 2 .. N GPUs, and the one-process engine pool
#   2. bench.py --gpus N: data-parallel line + the omniASR_CTC_7B tensor-parallel leg (config4_tp)
#   3. BASELINE config 5 (1B, 9.5 h) and config 3 (3B, 1 h) through the drop-in pipeline, ONE process driving N GPUs
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
if [ "$2" = "tests" ]; then
  timeout 900 python -m pytest tests/test_tp_multi_gpu.py "tests/test_gpu_engine.py::test_one_process_drives_every_gpu" -m gpu -q -s -rA \
      > $OUT/r2_multi_tests_$N.log 2>&1
  echo "tests rc=$?"; tail -4 $OUT/r2_multi_tests_$N.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 5 --warmup 3 > $OUT/r2_bench_dp${N}.json 2> $OUT/r2_bench_dp${N}.err
echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$OUT/r2_bench_dp${N}.json"))
print("dp$N:", round(d["value"]), "audio-s/s,", round(d["ms_per_step"], 2), "ms/step, e2e", round(d["e2e"]["value"]))
c = d.get("config4_tp")
if c:
    print("7B tp$N:", {k: c.get(k) for k in ("ms_per_step", "value", "encoder_gemm_frac_of_sustained_peak", "error")})
    print("   stages:", {k: round(v, 1) for k, v in (c.get("stages_ms_per_step") or {}).items()})
PY
timeout 600 python scripts/long_audio_stress.py --hours 9.5 --model omniASR_CTC_1B --devices $N --repeat 2 \
    > $OUT/r2_config5_pool_${N}gpu.json 2> $OUT/r2_config5_pool_${N}gpu.err
echo "config5 rc=$?"; cat $OUT/r2_config5_pool_${N}gpu.json
timeout 600 python scripts/long_audio_stress.py --hours 1 --model omniASR_CTC_3B --devices $N --repeat 3 \
    > $OUT/r2_config3_pool_${N}gpu.json 2> $OUT/r2_config3_pool_${N}gpu.err
echo "config3 rc=$?"; cat $OUT/r2_config3_pool_${N}gpu.json
