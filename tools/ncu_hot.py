#!/usr/bin/env python
"""What limits one kernel launch of an `ncu --set full --import-source on` report: headline counters, stall mix,
per-source-line shares of instructions / stall samples, and the lines with excess shared-memory wavefronts (bank
conflicts) or many L1 tag requests per global load.
  python tools/ncu_hot.py REPORT.ncu-rep [kernel-regex] [launch-skip] [top-n]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
rx = sys.argv[2] if len(sys.argv) > 2 else "."
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 14
sel = ["--kernel-name", "regex:" + rx, "--launch-skip", skip, "--launch-count", "1"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, vals = rows[0], rows[2]
m = dict(zip(hdr, vals))
def g(k):
    try: return float(m[k].replace(",", ""))
    except (KeyError, ValueError): return float("nan")
print("# %s  %s grid %s block %s" % (rep.split("/")[-1], m.get("Kernel Name"), m.get("Grid Size"), m.get("Block Size")))
print("time %.1f us | regs %s smem/blk %s KB | CTAs/SM limit smem %s regs %s | warps active %.1f%% | issue active %.1f%% | inst %.2fM" % (
    g("gpu__time_duration.sum"), m.get("launch__registers_per_thread"), m.get("launch__shared_mem_per_block_dynamic"),
    m.get("launch__occupancy_limit_shared_mem"), m.get("launch__occupancy_limit_registers"),
    g("sm__warps_active.avg.pct_of_peak_sustained_active"), g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    g("smsp__inst_executed.sum") / 1e6))
print("L1 hit %.1f%% L2 hit %.1f%% | dram rd %s %s wr %s %s | smem pipe %.1f%% | tensor pipe %s%%" % (
    g("l1tex__t_sector_hit_rate.pct"), g("lts__t_sector_hit_rate.pct"), m.get("dram__bytes_read.sum"), "", m.get("dram__bytes_write.sum"), "",
    g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    m.get("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", m.get("sm__inst_executed_pipe_tensor.sum", "-"))))
st = {k.split("issue_stalled_")[1].split("_per_issue")[0]: g(k) for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio")}
print("stalls per issue:", ", ".join("%s %.2f" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
H = rows[h]
ix = {n: H.index(n) for n in H if n in ("# Samples", "Instructions Executed", "L1 Tag Requests Global", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal",
                               "stall_barrier", "stall_long_sb", "stall_short_sb", "stall_mio", "stall_wait", "stall_math")}
for n in ("L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "L1 Tag Requests Global"):
    ix.setdefault(n, None)
agg = collections.defaultdict(collections.Counter)
text = {}
cur = None
for r in rows[h + 1:]:
    if len(r) <= max(v for v in ix.values() if v is not None):
        continue
    if r[0].strip():
        try: cur = int(r[0]); text[cur] = r[1].strip()
        except ValueError: pass
        continue
    a = agg[cur]
    for n, i in ix.items():
        if i is None: continue
        try: a[n] += float(r[i] or 0)
        except ValueError: pass
    if "LDG" in r[3]:
        try: a["ldg"] += float(r[ix["Instructions Executed"]] or 0)
        except ValueError: pass
tot = collections.Counter()
for a in agg.values():
    tot.update(a)
print("-- lines by stall samples (%d samples, %.2fM warp instructions)" % (tot["# Samples"], tot["Instructions Executed"] / 1e6))
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    print("%5d samp %5.1f%% inst %5.1f%% | bar %3d long %3d short %3d mio %3d wait %3d math %3d | %s" % (
        ln, 100 * a["# Samples"] / max(tot["# Samples"], 1), 100 * a["Instructions Executed"] / max(tot["Instructions Executed"], 1),
        a["stall_barrier"], a["stall_long_sb"], a["stall_short_sb"], a["stall_mio"], a["stall_wait"], a["stall_math"], text.get(ln, "")[:95]))
print("-- shared-memory wavefronts: %.2fM, ideal %.2fM" % (tot["L1 Wavefronts Shared"] / 1e6, tot["L1 Wavefronts Shared Ideal"] / 1e6))
for ln, a in sorted(agg.items(), key=lambda kv: -(kv[1]["L1 Wavefronts Shared"] - kv[1]["L1 Wavefronts Shared Ideal"]))[:6]:
    ex = a["L1 Wavefronts Shared"] - a["L1 Wavefronts Shared Ideal"]
    if ex <= 0: break
    print("%5d wavefronts %.0fk ideal %.0fk | %s" % (ln, a["L1 Wavefronts Shared"] / 1e3, a["L1 Wavefronts Shared Ideal"] / 1e3, text.get(ln, "")[:95]))
if tot["ldg"]:
    print("-- global loads: %.0fk warp instructions, %.2f L1 tag requests each" % (tot["ldg"] / 1e3, tot["L1 Tag Requests Global"] / tot["ldg"]))
