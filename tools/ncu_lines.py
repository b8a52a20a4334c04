#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel launch in a `ncu --set full --import-source on` report.
  python tools/ncu_lines.py gpurun_out/prof_X.ncu-rep [kernel-regex] [launch-skip] [top-n]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
rx = sys.argv[2] if len(sys.argv) > 2 else "."
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[h]
iL, iI, iS = 0, hdr.index("Instructions Executed"), hdr.index("# Samples")
agg = collections.defaultdict(lambda: [0, 0, ""])
tot = tots = 0
for r in rows[h + 1:]:
    if len(r) <= iI:
        continue
    try:
        ln, ins, sm = int(r[iL]), int(float(r[iI] or 0)), int(float(r[iS] or 0))
    except ValueError:
        continue
    a = agg[ln]
    a[0] += ins; a[1] += sm; a[2] = r[1]
    tot += ins; tots += sm
print("# %s  kernel ~ %s (launch %s): %d warp instructions, %d stall samples" % (rep.split("/")[-1], rx, skip, tot, tots))
for ln, (ins, sm, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5d %6.2f%% inst %6.2f%% samp | %s" % (ln, 100.0 * ins / max(tot, 1), 100.0 * sm / max(tots, 1), src.strip()[:120]))
