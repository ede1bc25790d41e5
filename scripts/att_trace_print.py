"""Prints the OASR_ATT_TRACE timeline of the attention kernels (CTA 0).  Rows of the trace file: 0 = MMA-issuing warp of
tile A (even slots: S issued, odd: P.V issued), 1.. = first softmax warp of tile A, B (, C); per key block j the softmax
slots are j*6 + {0 block start, 1 S in registers / handed back, 2 first chunk done, 4 all chunks done, 5 P handed over}.
The stamps are compiled in only with `make -C omnilingual-asr_b200/csrc clean all EXTRA=-DOASR_ATT_TRACING` (they cost ~8 %).
    OASR_ATT_TRACE=trace.txt python scripts/prof_attention.py 32 1499 16 80 1 ; python scripts/att_trace_print.py trace.txt"""
import sys

rows = [list(map(int, l.split())) for l in open(sys.argv[1])]
t0 = min(x for r in rows for x in r if x > 0)
tiles = [i for i in range(1, min(len(rows), 4)) if any(v > 0 for v in rows[i])]
first = int(sys.argv[2]) if len(sys.argv) > 2 else 6
for j in range(first, first + 6):
    parts = []
    for ti in tiles:
        r = rows[ti]
        if (j + 1) * 6 > len(r) or r[j * 6] <= 0:
            continue
        ev = [r[j * 6 + k] - t0 if r[j * 6 + k] > 0 else None for k in (0, 1, 2, 4, 5)]
        parts.append("ABC"[ti - 1] + " " + str(ev))
    m = rows[0]
    mma = (m[2 * j] - t0, m[2 * j + 1] - t0) if 2 * j + 1 < len(m) and m[2 * j] > 0 else None
    print(j, "  ".join(parts), " issuer A (S, P.V):", mma)
