#!/usr/bin/env python
"""Summaries of ncu output for profiles/ (text, small enough to commit).

  python tools/ncu_summary.py list gpurun_out/launches.csv            # per-kernel launch list summary
  python tools/ncu_summary.py full gpurun_out/prof_X.ncu-rep          # key metrics of a --set full capture
"""
import collections
import csv
import subprocess
import sys

KEY = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
       "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
       "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
       "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
       "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
       "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "gpc__cycles_elapsed.avg",
       "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
       "sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
       "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
       "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
       "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
       "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
       "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum"]


def short(name):
    return name.split("(")[0].replace("<unnamed>::", "").replace("void ", "")


def do_list(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            hdr, start = r, i + 1
            break
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        a = agg[short(r[ki])]
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none: %d launches, %.1f us total (cold-cache, serialised)" %
          (len(rows) - start, tot / 1e3))
    print("%-46s %8s %12s %7s" % ("kernel", "launches", "total_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-46s %8d %12.1f %6.1f%%" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))


def do_full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full --clock-control none --import-source on :", path.split("/")[-1])
    for r in rows[2:]:
        print("kernel:", short(r[hdr.index("Kernel Name")]), " grid", r[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
        for k in KEY:
            if k in hdr:
                i = hdr.index(k)
                print("  %-84s %s %s" % (k, r[i], units[i]))
        # every tensor-pipe counter the capture holds (names differ between ncu versions: sm__pipe_tensor*, sm__inst_executed_pipe_tensor*)
        for i, k in enumerate(hdr):
            if "pipe_tensor" in k and k not in KEY and r[i] not in ("", "n/a"):
                print("  %-84s %s %s" % (k, r[i], units[i]))


if __name__ == "__main__":
    (do_list if sys.argv[1] == "list" else do_full)(sys.argv[2])
