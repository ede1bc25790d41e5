"""BASELINE configs 3 / 5: one long synthetic recording through the drop-in pipeline (transcribe_chunked), windows
sharded over the ranks of a torchrun job, token ids gathered on the host.
    python scripts/long_audio_stress.py --hours 9.5 --model omniASR_CTC_1B
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/long_audio_stress.py --hours 1 --model omniASR_CTC_3B
Prints one JSON line (rank 0): audio seconds per wall second end to end (host PCM16 in, segments out)."""
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

pipe = CTCASRPipeline(args.model, weights="random", batch_windows=args.batch, device=torch.device("cuda", local))
pipe.transcribe_chunked(pcm[: 16000 * 30 * args.batch], sample_rate=16000)     # warm-up: workspace, tensor maps
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
res = pipe.transcribe_chunked(pcm, sample_rate=16000)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"workload": f"{args.model}, {args.hours} h synthetic PCM16, 30 s windows, dp{world}",
                      "windows": int(np.ceil(n / (16000 * 30))), "segments": len(res.segments),
                      "audio_s": n / 16000.0, "wall_s": wall, "audio_s_per_s": n / 16000.0 / wall,
                      "host_generate_s": gen_s, "n_gpus": world}), flush=True)
if world > 1:
    dist.destroy_process_group()
