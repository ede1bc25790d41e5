"""BASELINE configs 3 / 5: one long synthetic recording through the drop-in pipeline (transcribe_chunked).
    python scripts/long_audio_stress.py --hours 9.5 --model omniASR_CTC_1B --devices 8     ONE process, an engine and a
                                                       worker thread per GPU (engine_pool.py) - how the reference's web
                                                       app holds the pipeline (workflows/wav2elan_web/app.py:38-54)
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/long_audio_stress.py --hours 1 \
           --model omniASR_CTC_3B                      one process per GPU, windows sharded over the ranks, token ids
                                                       gathered on the host
Prints one JSON line (rank 0): audio seconds per wall second end to end (host PCM16 in, segments out), with the SM
clocks sampled during the timed call."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "omnilingual-asr_b200"))
from omnilingual_asr.models.inference.ctc_pipeline import CTCASRPipeline  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--hours", type=float, default=9.5)
ap.add_argument("--model", default="omniASR_CTC_1B")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--devices", type=int, default=0, help="single process driving this many GPUs (0: torchrun / one GPU)")
ap.add_argument("--repeat", type=int, default=1, help="timed repetitions (best and all are reported)")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

n = int(args.hours * 3600 * 16000)
rng = np.random.default_rng(1234)            # every rank holds the same recording (a shared file in production)
pcm = np.empty(n, dtype=np.int16)
t0 = time.perf_counter()
step = 16000 * 600
tone = (2000 * np.sin(2 * np.pi * 220.0 * np.arange(step) / 16000.0)).astype(np.float32)
for s in range(0, n, step):
    m = min(step, n - s)
    pcm[s:s + m] = (rng.standard_normal(m, dtype=np.float32) * 3000 + tone[:m]).astype(np.int16)
gen_s = time.perf_counter() - t0

import bench  # noqa: E402  (the nvidia-smi clock sampler)

if args.devices > 0:
    devs = [torch.device("cuda", i) for i in range(args.devices)]
    pipe = CTCASRPipeline(args.model, weights="random", batch_windows=args.batch, devices=devs, distributed=False)
    n_gpus = args.devices
else:
    pipe = CTCASRPipeline(args.model, weights="random", batch_windows=args.batch, device=torch.device("cuda", local))
    n_gpus = world
# warm-up: workspace, tensor maps, pinned staging buffers - two full batches per GPU
pipe.transcribe_chunked(pcm[: 16000 * 30 * args.batch * 2 * max(args.devices, 1)], sample_rate=16000)
walls = []
clk = {}
for rep in range(args.repeat):
    if world > 1:
        dist.barrier()
    for d in range(max(args.devices, 1)):
        torch.cuda.synchronize(d if args.devices > 0 else local)
    sampler = bench.ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    t0 = time.perf_counter()
    res = pipe.transcribe_chunked(pcm, sample_rate=16000)
    walls.append(time.perf_counter() - t0)
    if sampler:
        clk = sampler.stop()
wall = min(walls)
if rank == 0:
    print(json.dumps({"workload": f"{args.model}, {args.hours} h synthetic PCM16, 30 s windows, "
                                  + (f"one process x {args.devices} GPUs (engine pool)" if args.devices > 0 else f"dp{world} (torchrun)"),
                      "windows": int(np.ceil(n / (16000 * 30))), "segments": len(res.segments),
                      "audio_s": n / 16000.0, "wall_s": wall, "walls_s": walls, "audio_s_per_s": n / 16000.0 / wall,
                      "host_generate_s": gen_s, "n_gpus": n_gpus, "pool": pipe.pool.stats, "clocks": clk}), flush=True)
pipe.close()
if world > 1:
    dist.destroy_process_group()
