"""Prints the OASR_ATT_TRACE timeline of attention_v4 (CTA 0): per key block and query tile the SM-clock stamps
start / first two chunks loaded / chunk 0 done / S consumed / all chunks done / P handed over."""
import sys
rows = [list(map(int, l.split())) for l in open(sys.argv[1])]
t0 = min(x for r in rows for x in r if x)
for role in (1, 2):
    print("tile", "AB"[role - 1], "(start, ld01, c0, sfree, c12, pdone) relative; deltas; period")
    r = rows[role]
    for j in range(0, 20):
        ev = r[j * 6:(j + 1) * 6]
        if not ev[0]:
            break
        nxt = r[(j + 1) * 6] if (j + 1) * 6 < len(r) else 0
        print(j, [e - t0 for e in ev], [ev[i + 1] - ev[i] for i in range(5)], (nxt - ev[0]) if nxt else None)
print("mma (before p_full_A wait, after P.V_A issue):", [(rows[0][2 * j] - t0, rows[0][2 * j + 1] - t0) for j in range(16)])
