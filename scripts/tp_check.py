"""Tensor-parallel parity on N GPUs: every rank runs its slice of the encoder; rank 0 compares the result with the CPU
ORACLE computed with the tensor-parallel rounding points (bf16-operand emulation, partial sums rounded to bf16 and
added in rank order: ctc_oracle.forward(tp_world=N)) and with an unsplit engine.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/tp_check.py [preset] [fused|nccl] [B]
B > 3 repeats the golden windows (with different lengths) so that the half-batch pipeline of the peer-memory path has
work for both halves."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "omnilingual-asr_b200"))
from omnilingual_asr.models.config import CtcModelConfig  # noqa: E402
from omnilingual_asr.models.inference.ctc_engine import CtcEngine  # noqa: E402
from oracle import ctc_oracle as O  # noqa: E402
from tests.golden.make_golden import golden_inputs  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
name = sys.argv[1] if len(sys.argv) > 1 else "wide2l"
ocfg = O.PRESETS[name]
cfg = CtcModelConfig(ocfg.name, ocfg.d_model, ocfg.n_layers, ocfg.n_heads, ocfg.d_ffn, vocab=ocfg.vocab, pos_groups=ocfg.pos_groups)
w = O.init_weights(ocfg, seed=0)
wave, ns = golden_inputs()
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
if reps > 1:                       # a longer, still ragged batch: copies of the golden windows cut to other lengths
    waves, nss = [], []
    for k in range(reps):
        wk = wave.clone()
        nk = [max(400, n - 517 * k) for n in ns]
        for b, n in enumerate(nk):
            wk[b, n:] = 0
        waves.append(wk)
        nss += nk
    wave, ns = torch.cat(waves), nss

idt = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(CtcEngine.tp_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
eng = CtcEngine(cfg, device=dev, tp_rank=rank, tp_world=world, tp_id=bytes(idt.cpu().numpy().tobytes()))
eng.load_state_dict(w)
mode = sys.argv[2] if len(sys.argv) > 2 else "fused"
if mode == "fused":
    eng.tp_enable_peer_memory(wave.shape[0], wave.shape[1])
res = eng.forward(wave.to(dev), ns, normalised=True, return_hidden=True)
res2 = eng.forward(wave.to(dev), ns, normalised=True, return_hidden=False)   # second call: epochs keep counting
assert all((a == b).all() for a, b in zip(res.token_ids, res2.token_ids)), "forward is not repeatable"
torch.cuda.synchronize()
# every rank holds the same replicated result
ids = torch.from_numpy(res.frame_ids.astype(np.int64)).to(dev)
ids0 = ids.clone()
dist.broadcast(ids0, 0)
assert bool((ids == ids0).all()), "ranks disagree on the frame ids"
if rank == 0:
    ref = CtcEngine(cfg, device=dev)
    ref.load_state_dict(w)
    r0 = ref.forward(wave.to(dev), ns, normalised=True, return_hidden=True)
    err = max(float((res.hidden[b, :nf] - r0.hidden[b, :nf]).norm() / r0.hidden[b, :nf].norm()) for b, nf in enumerate(r0.n_frames))
    same = float(np.mean([np.mean(res.frame_ids[b, :nf] == r0.frame_ids[b, :nf]) for b, nf in enumerate(r0.n_frames)]))
    print(f"tp{world} {name} [{mode}]: hidden rel err vs unsplit {err:.2e}, frame-id agreement {same:.4f}")
    assert err < 8e-3 and same >= 0.95      # the unsplit engine does not round partial sums: one bf16 rounding apart
    # the oracle with this run's rounding points (the NCCL mode sums fp32 partials: plain bf16-operand emulation)
    emu = O.forward(w, wave, ns, ocfg, emulate_bf16=True, return_logits=True, tp_world=world if mode == "fused" else 1)
    margin = O.top2_margin(emu.logits)
    tot = ok = tot_all = ok_all = 0
    e_or = 0.0
    for b, nf in enumerate(emu.n_frames):
        eq = res.frame_ids[b, :nf] == emu.frame_ids[b, :nf].numpy()
        keep = (margin[b, :nf] > 0.05).numpy()
        tot += int(keep.sum()); ok += int(eq[keep].sum()); tot_all += nf; ok_all += int(eq.sum())
        e_or = max(e_or, float((res.hidden[b, :nf].cpu() - emu.hidden[b, :nf]).norm() / emu.hidden[b, :nf].norm()))
    print(f"tp{world} {name} [{mode}]: vs ORACLE (tp rounding points): ids {ok_all}/{tot_all}, on margin>0.05 {ok}/{tot}, "
          f"hidden rel err {e_or:.2e}")
    assert ok == tot and e_or < 8e-3
    print("TP OK")
dist.barrier()
dist.destroy_process_group()
