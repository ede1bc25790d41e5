#!/usr/bin/env python
"""Headline benchmark: audio-seconds transcribed per second (inverse RTF) for omniASR CTC on B200.

    python bench.py --gpus 1 --steps 10 --warmup 3            # our arm, one JSON line on stdout
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                       # the CPU oracle timed on the host cores

A step is one pass of the hot path (a8..a16: normalise -> conv FE -> pos-conv -> encoder -> CTC arg-max ->
collapse) over one batch of `--batch` synthetic 30 s windows per GPU (BASELINE.json configs[1]:
omniASR_CTC_1B, 32 x 30 s, bf16 operands / fp32 accumulate, random-init weights).
  value  windows resident in HBM when the timed region starts (device time, CUDA events, max over ranks)
  e2e    the same through oasr_transcribe_host from pinned HOST buffers (H2D + forward + D2H every step)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (str(ROOT), str(ROOT / "omnilingual-asr_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WINDOW_SEC = 30.0
SR = 16000
_JSON_OUT = sys.stdout
METRIC = "audio-sec/sec (inverse RTF)"
UNIT = "audio-s/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="omniASR_CTC_1B")
    ap.add_argument("--batch", type=int, default=32, help="30 s windows per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling aid: only the device-resident leg")
    ap.add_argument("--cpu-windows", type=int, default=2, help="windows in the bounded CPU sample")
    ap.add_argument("--tp-mode", default="fused", choices=["fused", "nccl"],
                    help="fused: all-reduce + residual + LayerNorm as one kernel over NVLink peer memory; nccl: ncclAllReduce")
    ap.add_argument("--tp", type=int, default=1, help="tensor-parallel degree (BASELINE config 4: 7B encoder over "
                                                      "2/4/8 GPUs); must equal --gpus, every rank sees the same batch")
    ap.add_argument("--no-tp-leg", action="store_true",
                    help="with --gpus N >= 2: skip the short omniASR_CTC_7B tensor-parallel leg (config4_tp key)")
    ap.add_argument("--tp-leg-steps", type=int, default=3)
    return ap.parse_args()


def synthetic_windows(batch: int, seed: int) -> torch.Tensor:
    """[batch, 480000] fp32: unit Gaussian noise plus a few tones (non-degenerate LayerNorm statistics)."""
    g = torch.Generator().manual_seed(seed)
    L = int(WINDOW_SEC * SR)
    x = torch.randn(batch, L, generator=g)
    t = torch.arange(L, dtype=torch.float32) / SR
    for f, a in ((220.0, 0.5), (1333.0, 0.25), (3100.0, 0.1)):
        x += a * torch.sin(2 * np.pi * f * t)[None]
    return x


def flops_per_window(cfg, T: int) -> dict:
    """Algorithmic FLOPs (2 * MAC) of each stage for one 30 s window."""
    d, F, Lyr, V = cfg.d_model, cfg.d_ffn, cfg.n_layers, cfg.vocab
    n = int(WINDOW_SEC * SR)
    lens = []
    for _, k, s in cfg.fe_layers:
        n = (n - k) // s + 1
        lens.append(n)
    fe = [2.0 * lens[0] * 512 * 10] + [2.0 * lens[i] * 512 * 512 * cfg.fe_layers[i][1] for i in range(1, len(lens))]
    cg = d // cfg.pos_groups
    return {
        "fe_layer0": fe[0], "fe_conv_1_6": sum(fe[1:]), "feature_proj": 2.0 * T * 512 * d,
        "posconv": 2.0 * T * d * cg * cfg.pos_kernel,
        "qkv_gemm": Lyr * 2.0 * T * d * 3 * d, "outproj_gemm": Lyr * 2.0 * T * d * d,
        "ffn1_gemm": Lyr * 2.0 * T * d * F, "ffn2_gemm": Lyr * 2.0 * T * d * F,
        "attention": Lyr * 4.0 * T * T * d, "ctc_head_argmax": 2.0 * T * d * V,
    }


def measured_peaks() -> tuple[dict, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text()), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def kernels_sha() -> str:
    """Content hash of the CUDA sources liboasr.so is built from: profiles/ncu_traffic.json records the hash of the
    binary it was captured on (scripts/profile_step.sh), and a capture of other kernels is not reported as `traffic`."""
    import hashlib
    h = hashlib.sha256()
    src = ROOT / "omnilingual-asr_b200" / "csrc"
    for f in sorted(list(src.glob("*.cu")) + list(src.glob("*.cuh")) + list(src.glob("*.h"))):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_throughput(model: str, windows: int, reps: int = 1, warmup: int = 0, bf16: bool = False):
    """Times the CPU oracle (PyTorch eager, all host threads) on `windows` 30 s windows: fp32 (the headline baseline) or,
    with bf16=True, the same graph under torch.autocast(cpu, bfloat16) - the AMX path BASELINE.md section 3 asks for
    beside it (contractions in bf16, LayerNorm / softmax in fp32)."""
    from oracle import ctc_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.PRESETS[model]
    w = O.init_weights(cfg, seed=0)
    wave = synthetic_windows(windows, 1234)
    ns = [wave.shape[1]] * windows
    times = []
    import contextlib
    ctx = torch.autocast("cpu", dtype=torch.bfloat16) if bf16 else contextlib.nullcontext()
    with torch.no_grad(), ctx:
        for i in range(warmup + reps):
            t0 = time.perf_counter()
            wn = O.wave_layer_norm(wave, ns)
            out = O.forward(w, wn, ns, cfg)
            for b in range(windows):
                O.greedy_collapse(out.frame_ids[b], out.n_frames[b])
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return windows * WINDOW_SEC / (sum(times) / len(times)), cores, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    windows = 1
    t_all0 = time.perf_counter()
    val, cores, times = cpu_oracle_throughput(args.model, windows, reps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model}, 32 x 30 s synthetic 16 kHz windows (BASELINE configs[1])",
                   "sample": f"{windows} x 30 s window per step"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{windows} x 30 s window per step, {args.steps} steps, PyTorch eager fp32 oracle"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_all0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def tp_leg(args, dist, dev, rank, world, local):
    """BASELINE configs[3]: omniASR_CTC_7B with the encoder split tensor-parallel over all `world` GPUs, every rank
    feeding the same 32 x 30 s batch (peer-memory reductions, tp_fused.cu).  A few steps, reported under `config4_tp`
    of the data-parallel line so that the driver's scaling record carries it."""
    from omnilingual_asr.models.config import get_model_config
    from omnilingual_asr.models.inference.ctc_engine import CtcEngine
    from omnilingual_asr.models.weights import random_weights
    model = "omniASR_CTC_7B"
    cfg = get_model_config(model)
    B, L = args.batch, int(WINDOW_SEC * SR)
    T = cfg.feature_length(L)
    steps = max(1, args.tp_leg_steps)
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(CtcEngine.tp_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    eng = CtcEngine(cfg, device=dev, tp_rank=rank, tp_world=world, tp_id=idt.cpu().numpy().tobytes())
    eng.tp_enable_peer_memory(B, L)
    eng.load_state_dict(random_weights(cfg, 0, dev))
    torch.cuda.empty_cache()
    wave_dev = synthetic_windows(B, 1234).to(dev)
    ns = [L] * B
    for _ in range(2):
        eng.forward(wave_dev, ns, return_frame_ids=False)
    eng.profile_read()
    eng.profile(True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(steps):
        res = eng.forward(wave_dev, ns, return_frame_ids=False)
    ev1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    stage_ms = eng.profile_read()
    eng.profile(False)
    clk = clocks.stop() if rank == 0 else {}
    eng.close()
    fl = flops_per_window(cfg, T)
    gemm_names = ["qkv_gemm", "outproj_gemm", "ffn1_gemm", "ffn2_gemm"]
    g_ms = sum(stage_ms[n][0] for n in gemm_names)
    g_flops = sum(fl[n] for n in gemm_names) * B * steps / world          # this rank's share of the contractions
    peaks, _ = measured_peaks()
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    ach = g_flops / (g_ms / 1e3) / 1e12 if g_ms > 0 else 0.0
    return {
        "workload": f"{model}, {B} x 30 s synthetic windows, encoder tensor-parallel over {world} GPUs (same batch on every rank)",
        "tp": world, "steps": steps, "warmup": 2, "ms_per_step": ms, "value": B * WINDOW_SEC / (ms / 1e3), "unit": UNIT,
        "reduction": "bf16 partial sums; reduce-scatter + residual + LayerNorm + all-gather per half-batch on the copy "
                     "engines beside the other half-batch's GEMMs (tp_fused.cu: tp_dma_reduce_layernorm)",
        "stages_ms_per_step": {k: v[0] / steps for k, v in stage_ms.items() if v[1] > 0},
        "encoder_gemm_tflops_per_gpu": ach, "encoder_gemm_frac_of_sustained_peak": ach / peak,
        "tokens_window0": int(len(res.token_ids[0])), "clocks": clk,
    }


def run_ours(args):
    import torch.distributed as dist

    from omnilingual_asr.models.config import get_model_config
    from omnilingual_asr.models.inference.ctc_engine import CtcEngine
    from omnilingual_asr.models.weights import random_weights

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cfg = get_model_config(args.model)
    B, L = args.batch, int(WINDOW_SEC * SR)
    T = cfg.feature_length(L)
    tp = args.tp
    if tp > 1:
        if tp != world:
            raise ValueError("--tp must equal --gpus (one tensor-parallel group)")
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(CtcEngine.tp_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        eng = CtcEngine(cfg, device=dev, tp_rank=rank, tp_world=world, tp_id=idt.cpu().numpy().tobytes())
        if args.tp_mode == "fused":
            eng.tp_enable_peer_memory(B, L)
    else:
        eng = CtcEngine(cfg, device=dev)
    eng.load_state_dict(random_weights(cfg, 0, dev))
    host = synthetic_windows(B, 1234 + (0 if tp > 1 else rank)).pin_memory()
    ns = [L] * B
    wave_dev = host.to(dev)
    audio_per_step = (1 if tp > 1 else world) * B * WINDOW_SEC

    # ---------------------------------------------------------------- device-resident leg ("value")
    for _ in range(args.warmup):
        eng.forward(wave_dev, ns, return_frame_ids=False)
    eng.profile_read()
    eng.profile(True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        eng.forward(wave_dev, ns, return_frame_ids=False)
    ev1.record()
    barrier()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = eng.launch_count - n0
    stage_ms = eng.profile_read()
    eng.profile(False)
    clk = clocks.stop() if rank == 0 else {}
    value = audio_per_step * args.steps / (dev_ms / 1e3)

    # ---------------------------------------------------------------- end-to-end leg (host buffers)
    e2e_steps = 0 if args.skip_e2e else args.steps
    for _ in range(0 if args.skip_e2e else min(args.warmup, 2)):
        eng.transcribe_host(host, ns)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = eng.transcribe_host(host, ns)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0) if e2e_steps else float("nan")
    barrier()
    e2e_val = audio_per_step * args.steps / e2e_s
    h2d = B * L * 4
    d2h = 2 * B * T * 4 + B * 4

    # ---------------------------------------------------------------- config 4: the 7B tensor-parallel leg
    config4 = None
    if world > 1 and tp == 1 and not args.no_tp_leg:
        eng.close()
        del wave_dev
        torch.cuda.empty_cache()
        try:
            config4 = tp_leg(args, dist, dev, rank, world, local)
        except Exception as e:  # noqa: BLE001 - the data-parallel line is still valid
            config4 = {"error": f"{type(e).__name__}: {e}"[:400]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- roofline of the dominant kernel
    peaks, peak_src = measured_peaks()
    fl = flops_per_window(cfg, T)
    stages = {}
    total_stage_ms = sum(ms for ms, _ in stage_ms.values()) or 1.0
    for name, (ms, cnt) in stage_ms.items():
        if cnt == 0:
            continue
        entry = {"ms_per_step": ms / args.steps, "share": ms / total_stage_ms, "launch_groups": cnt}
        if name in fl and ms > 0:
            entry["tflops"] = fl[name] * B * args.steps / (ms / 1e3) / 1e12 / (tp if name not in ("fe_layer0", "fe_conv_1_6", "feature_proj", "posconv", "ctc_head_argmax") else 1)
        stages[name] = entry
    gemm_names = ["qkv_gemm", "outproj_gemm", "ffn1_gemm", "ffn2_gemm"]
    g_ms = sum(stage_ms[n][0] for n in gemm_names)
    g_cnt = sum(stage_ms[n][1] for n in gemm_names)
    g_flops = sum(fl[n] for n in gemm_names) * B * args.steps / max(tp, 1)   # a tensor-parallel rank computes 1/tp of them
    achieved = g_flops / (g_ms / 1e3) / 1e12 if g_ms > 0 else 0.0
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    # DRAM bytes per launch of the same kernels from the committed `ncu --set full` capture (profiles/ncu_traffic.json,
    # 1B shapes only): mean over the four encoder GEMMs, like flops_per_launch
    traffic, ncu_note = None, None
    try:
        tj = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
        if tj.get("kernels_sha") != kernels_sha():
            ncu_note = (f"profiles/ncu_traffic.json was captured on other kernel sources ({tj.get('kernels_sha')} vs "
                        f"{kernels_sha()}): regenerate with scripts/profile_step.sh")
        elif args.model == "omniASR_CTC_1B" and B == 32 and tp == 1:
            per = [tj["kernels"][n]["traffic_bytes"] for n in gemm_names]
            traffic = sum(per) / len(per)
            ncu_note = {n: {"traffic_bytes": tj["kernels"][n]["traffic_bytes"],
                            "tensor_pipe_active_pct": tj["kernels"][n]["tensor_pipe_active_pct"]} for n in gemm_names}
    except Exception:
        pass
    roofline = {
        "kernel": "gemm_kernel<256,*> (tcgen05 GEMM: encoder QKV / out-proj / FFN1 / FFN2)",
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)",
        "traffic": traffic, "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu)",
        "ncu": ncu_note,
        "flops_per_launch": g_flops / max(g_cnt, 1), "avg_launch_ms": g_ms / max(g_cnt, 1),
        "share_of_step": g_ms / total_stage_ms,
        "whole_path_tflops": sum(fl.values()) * B * (1 if tp > 1 else world) * args.steps / (dev_ms / 1e3) / 1e12,
        "stages": stages,
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.model}, {B} x 30 s synthetic 16 kHz windows per GPU (BASELINE configs[1]), "
                               "random-init weights, bf16 operands / fp32 accumulate",
                   "windows_per_gpu": B, "frames_per_window": T, "parallelism": (f"tp{world} (encoder split over heads / FFN columns; per row-parallel GEMM "
                                    + ("one fused reduce + residual + LayerNorm kernel over NVLink peer memory)"
                                       if args.tp_mode == "fused" else "one NCCL all-reduce)") if tp > 1
                                   else f"dp{world} (window shards, no collective)"),
                   "l2": "inputs re-read from HBM every step: per-step working set (FE activations ~6 GB, encoder "
                         "activations ~1.5 GB, weights 1.9 GB) is far larger than the 126 MB L2"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_s / args.steps, "api": "oasr_transcribe_host (C-ABI, pinned host buffers)"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roofline,
    }
    if config4 is not None:
        line["config4_tp"] = config4
    if not args.no_cpu_baseline and world == 1:
        val, cores, times = cpu_oracle_throughput(args.model, args.cpu_windows, reps=1)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{args.cpu_windows} x 30 s windows, one pass, PyTorch eager fp32 oracle "
                                          f"({times[0]:.1f} s)"}
        try:   # the same oracle with bf16 contractions (AMX where the host has it): reported beside the fp32 baseline
            v16, _, t16 = cpu_oracle_throughput(args.model, args.cpu_windows, reps=1, warmup=1, bf16=True)
            line["cpu_baseline"]["bf16_autocast"] = {"value": v16, "unit": UNIT,
                                                     "sample": f"{args.cpu_windows} x 30 s windows, second of two passes ({t16[0]:.1f} s)"}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"]["bf16_autocast"] = {"error": str(e)[:200]}
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    # stdout carries exactly ONE JSON line: libraries that write to fd 1 from native code (NCCL prints its version
    # banner there) are pointed at stderr, the JSON line goes to the saved descriptor
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    global _JSON_OUT
    _JSON_OUT = os.fdopen(saved, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
