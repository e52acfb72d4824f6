#!/usr/bin/env python
"""Aggregate ncu's SASS source page of one kernel: shared-memory wavefronts (actual / ideal / excess) per opcode and
the warp-stall sample distribution.  usage: python tools/ncu_source_summary.py <report.ncu-rep> <kernel regex>"""
import collections, csv, subprocess, sys

rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(raw))
name = rows[0][1]
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def num(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
per = collections.OrderedDict()
stalls = collections.Counter()
tot_samples = 0.0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    op = r[ix["Source"]].split()
    op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
    w, wi, we = num(r, "L1 Wavefronts Shared"), num(r, "L1 Wavefronts Shared Ideal"), num(r, "L1 Wavefronts Shared Excessive")
    if w:
        d = per.setdefault(op, [0, 0.0, 0.0, 0.0, 0.0])
        d[0] += 1; d[1] += w; d[2] += wi; d[3] += we; d[4] += num(r, "Instructions Executed")
    tot_samples += num(r, "# Samples")
    for k in hdr:
        if k.startswith("stall_") and not k.endswith("(Not Issued)"):
            stalls[k] += num(r, k)
print(f"kernel: {name}")
print("opcode | static instr | executed | smem wavefronts | ideal | excess | excess share")
tw = sum(d[1] for d in per.values())
for op, d in sorted(per.items(), key=lambda kv: -kv[1][1]):
    print(f"{op} | {d[0]} | {d[4]:.0f} | {d[1]:.0f} | {d[2]:.0f} | {d[3]:.0f} | {d[3] / max(1.0, d[1]):.2f}")
print(f"total shared wavefronts {tw:.0f}")
print("warp stall samples:", ", ".join(f"{k[6:]} {v / max(1.0, sum(stalls.values())):.1%}" for k, v in stalls.most_common(8)))
