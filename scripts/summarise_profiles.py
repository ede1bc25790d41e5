#!/usr/bin/env python
"""Turns the ncu output of scripts/profile_step.sh into the committed summaries under profiles/.

    python scripts/summarise_profiles.py r1          (reads gpurun_out/r1_launches.csv, gpurun_out/r1_full.ncu-rep)

Writes profiles/<tag>_launches_last_step.csv (every launch of the LAST bench step with its device time),
profiles/<tag>_launch_shares.md (per-kernel share of that step) and profiles/<tag>_full_summary.csv
(selected `ncu --set full` counters per captured launch).
"""
from __future__ import annotations

import collections
import csv
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"
SRC = ROOT / "gpurun_out"
DST = ROOT / "profiles"
DST.mkdir(exist_ok=True)


def short(name: str) -> str:
    m = re.search(r"(gemm_kernel<[^>]*>|attention_v\d_kernel<[^>]*>|attention_v\d_kernel|[A-Za-z_0-9]+_kernel)", name)
    s = m.group(1) if m else name.split("(")[0][-60:]
    return s.replace("(oasr::Epilogue)", "")


def launches():
    f = SRC / f"{TAG}_launches.csv"
    if not f.exists():
        return
    lines = [ln for ln in f.read_text().splitlines() if ln.startswith('"')]
    rows = list(csv.DictReader(lines))
    # the last step = everything from the last wave_stats launch on
    start = max(i for i, r in enumerate(rows) if "wave_stats" in r["Kernel Name"])
    step = rows[start:]
    with open(DST / f"{TAG}_launches_last_step.csv", "w", newline="") as out:
        w = csv.writer(out)
        w.writerow(["id", "kernel", "grid", "block", "time_us"])
        for r in step:
            unit = r["Metric Unit"]
            v = float(r["Metric Value"].replace(",", ""))
            us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
            w.writerow([r["ID"], short(r["Kernel Name"]), r["Grid Size"], r["Block Size"], f"{us:.2f}"])
            r["_us"] = us
    agg = collections.OrderedDict()
    for r in step:
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += r["_us"]
    total = sum(a[1] for a in agg.values())
    with open(DST / f"{TAG}_launch_shares.md", "w") as out:
        out.write(f"# {TAG}: launches of one bench step (omniASR_CTC_1B, 32 x 30 s), `ncu --metrics gpu__time_duration.sum "
                  f"--clock-control none`\n\nSerialised, cold-cache times: compare shares, not absolutes.  "
                  f"{len(step)} launches, {total / 1e3:.2f} ms in total.\n\n| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|\n")
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.write(f"| `{k}` | {n} | {us / 1e3:.3f} | {us / n:.1f} | {100 * us / total:.1f} % |\n")
    print((DST / f"{TAG}_launch_shares.md").read_text())


METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_active.avg",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg",
]


def full():
    rep = SRC / f"{TAG}_full.ncu-rep"
    if not rep.exists():
        return
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = [ln for ln in raw.splitlines() if ln.startswith('"')]
    rd = csv.reader(lines)
    header = next(rd)
    units = next(rd)
    cols = {h: i for i, h in enumerate(header)}
    tensor_cols = [h for h in header if "pipe_tensor" in h and "pct" in h]
    keep = [m for m in METRICS if m in cols]
    for h in tensor_cols:
        if h not in keep:
            keep.append(h)
    with open(DST / f"{TAG}_full_summary.csv", "w", newline="") as out:
        w = csv.writer(out)
        w.writerow(["id", "kernel"] + [f"{m} [{units[cols[m]]}]" for m in keep])
        for r in rd:
            w.writerow([r[cols["ID"]], short(r[cols["Kernel Name"]])] + [r[cols[m]] for m in keep])
    print((DST / f"{TAG}_full_summary.csv").read_text())


def traffic():
    """profiles/ncu_traffic.json: DRAM bytes and tensor-pipe activity per captured launch, keyed by the stage the
    launch belongs to (scripts/profile_step.sh captures the first 17 kernels of a step, in this order), stamped with
    the content hash of the kernel sources so that bench.py only reports a `traffic` measured on the current kernels."""
    import json
    sys.path.insert(0, str(ROOT))
    import bench
    f = DST / f"{TAG}_full_summary.csv"
    if not f.exists():
        return
    rows = list(csv.reader(f.read_text().splitlines()))
    head, rows = rows[0], rows[1:]
    col = {h.split(" [")[0]: i for i, h in enumerate(head)}
    order = ["fe_layer0", "fe_conv_layer1", "fe_conv_layer2", "fe_conv_layer3", "fe_conv_layer4", "fe_conv_layer5",
             "fe_conv_layer6", "layernorm_proj", "feature_proj", "posconv", "layernorm_attn", "qkv_gemm", "attention",
             "outproj_gemm", "layernorm_ffn", "ffn1_gemm", "ffn2_gemm"]
    if len(rows) < len(order):
        print(f"traffic: only {len(rows)} captured launches, expected {len(order)}; ncu_traffic.json not written")
        return

    def num(r, k):
        return float(r[col[k]].replace(",", ""))

    unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    units = {h.split(" [")[0]: h.split(" [")[1].rstrip("]") if " [" in h else "" for h in head}
    tens = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    kernels = {}
    for name, r in zip(order, rows):
        rd = num(r, "dram__bytes_read.sum") * unit_scale.get(units["dram__bytes_read.sum"], 1.0)
        wr = num(r, "dram__bytes_write.sum") * unit_scale.get(units["dram__bytes_write.sum"], 1.0)
        dur = num(r, "gpu__time_duration.sum")
        du = units["gpu__time_duration.sum"]
        dur_us = dur / 1e3 if du.startswith("n") else (dur if du.startswith("u") else dur * 1e3)
        kernels[name] = {"kernel": r[1], "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr,
                         "duration_us": dur_us, "tensor_pipe_active_pct": num(r, tens) if tens in col else None,
                         "xu_pipe_pct": num(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active")
                         if "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active" in col else None}
    out = {"source": f"profiles/{TAG}_full_summary.csv (ncu --set full --clock-control none, one launch each, "
                     "omniASR_CTC_1B, 32 x 30 s; scripts/profile_step.sh)",
           "kernels_sha": bench.kernels_sha(), "kernels": kernels}
    (DST / "ncu_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
    print(f"ncu_traffic.json written for kernel sources {out['kernels_sha']}")


launches()
full()
traffic()
