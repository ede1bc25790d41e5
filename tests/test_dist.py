"""N>1 path on the CPU: two gloo ranks shard the windows, gather token ids on the host, and every rank ends
with the result of the single-process run."""
import os
import pickle
import sys
from pathlib import Path

import numpy as np
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, out_dir, overlap=0.0):
    for p in (str(ROOT), str(ROOT / "omnilingual-asr_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    from omnilingual_asr import CTCASRPipeline
    from tests._fake_engine import OracleEngine
    torch.set_num_threads(2)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = OracleEngine("tiny")
    pipe = CTCASRPipeline(eng.cfg, engine=eng, window_seconds=1.0, batch_windows=2, overlap_seconds=overlap)
    x = np.random.default_rng(9).standard_normal(int(5.3 * 16000)).astype(np.float32)
    res = pipe.transcribe_chunked(x)
    windows_here = sum(c[0][0] for c in eng.calls)
    with open(os.path.join(out_dir, f"r{rank}.pkl"), "wb") as f:
        pickle.dump(([(s.start, s.end, s.text) for s in res.segments], windows_here), f)
    dist.destroy_process_group()


def test_two_rank_shard_and_host_gather(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    outs = [pickle.load(open(tmp_path / f"r{r}.pkl", "rb")) for r in range(world)]
    assert outs[0][0] == outs[1][0]                       # every rank holds the merged transcript
    assert outs[0][1] + outs[1][1] == 6 and abs(outs[0][1] - outs[1][1]) <= 1   # 6 windows, block partition
    # equals the single-process result
    for p in (str(ROOT), str(ROOT / "omnilingual-asr_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from omnilingual_asr import CTCASRPipeline
    from tests._fake_engine import OracleEngine
    eng = OracleEngine("tiny")
    pipe = CTCASRPipeline(eng.cfg, engine=eng, window_seconds=1.0, batch_windows=2, distributed=False)
    x = np.random.default_rng(9).standard_normal(int(5.3 * 16000)).astype(np.float32)
    single = [(s.start, s.end, s.text) for s in pipe.transcribe_chunked(x).segments]
    assert single == outs[0][0]
    starts = [s[0] for s in single]
    assert starts == sorted(starts)


def test_two_rank_overlapping_windows_stitch_after_the_gather(tmp_path):
    """Overlap-and-stitch needs both neighbours of a cut: it runs after the host gather, on every rank alike."""
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path), 0.25), nprocs=world, join=True)
    outs = [pickle.load(open(tmp_path / f"r{r}.pkl", "rb")) for r in range(world)]
    assert outs[0][0] == outs[1][0]
    assert outs[0][1] + outs[1][1] == 7                   # 5.3 s in 1 s windows 0.75 s apart
    for p in (str(ROOT), str(ROOT / "omnilingual-asr_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from omnilingual_asr import CTCASRPipeline
    from tests._fake_engine import OracleEngine
    eng = OracleEngine("tiny")
    pipe = CTCASRPipeline(eng.cfg, engine=eng, window_seconds=1.0, batch_windows=2, overlap_seconds=0.25, distributed=False)
    x = np.random.default_rng(9).standard_normal(int(5.3 * 16000)).astype(np.float32)
    single = [(s.start, s.end, s.text) for s in pipe.transcribe_chunked(x).segments]
    assert single == outs[0][0]


def _tp_worker(rank, world, port, out_dir, fail_rank):
    for p in (str(ROOT), str(ROOT / "omnilingual-asr_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    from omnilingual_asr import CTCASRPipeline
    from tests._fake_engine import FlakyEngine, OracleEngine
    torch.set_num_threads(2)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = FlakyEngine(10 ** 6, name="tiny") if rank == fail_rank else OracleEngine("tiny")
    if fail_rank < 0:
        eng.tp_world = world          # a tensor-parallel engine: both ranks are ONE data-parallel replica
    pipe = CTCASRPipeline(eng.cfg, engine=eng, window_seconds=1.0, batch_windows=2)
    x = np.random.default_rng(9).standard_normal(int(3.3 * 16000)).astype(np.float32)
    try:
        res = pipe.transcribe_chunked(x)
        out = ("ok", [(s.start, s.end, s.text) for s in res.segments], sum(c[0][0] for c in eng.calls))
    except RuntimeError as e:
        out = ("error", str(e), 0)
    with open(os.path.join(out_dir, f"r{rank}.pkl"), "wb") as f:
        pickle.dump(out, f)
    dist.destroy_process_group()


def test_tensor_parallel_ranks_run_identical_batches(tmp_path):
    """the ranks of a tensor-parallel group feed the same cross-GPU reductions, so the pipeline must give
    every one of them the SAME windows (data-parallel rank = rank // tp_world) instead of a shard each."""
    world = 2
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_tp_worker, args=(world, port, str(tmp_path), -1), nprocs=world, join=True)
    outs = [pickle.load(open(tmp_path / f"r{r}.pkl", "rb")) for r in range(world)]
    assert outs[0][0] == outs[1][0] == "ok"
    assert outs[0][2] == outs[1][2] == 4                  # both ranks ran all four windows
    assert outs[0][1] == outs[1][1]
    starts = [s[0] for s in outs[0][1]]
    assert starts == sorted(starts) and len(set(starts)) == len(starts)      # the group's windows appear once


def test_failing_rank_does_not_strand_the_others(tmp_path):
    """A rank whose device step raises still takes part in the host gather (it sends an error marker): every rank
    raises instead of one of them waiting in all_gather_object for ever."""
    world = 2
    port = 35500 + (os.getpid() % 2000)
    mp.spawn(_tp_worker, args=(world, port, str(tmp_path), 1), nprocs=world, join=True)
    outs = [pickle.load(open(tmp_path / f"r{r}.pkl", "rb")) for r in range(world)]
    assert outs[0][0] == "error" and "rank 1" in outs[0][1]
    assert outs[1][0] == "error" and "injected device failure" in outs[1][1]
