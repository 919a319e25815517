#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/rNN_launches.md  [title]
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/rNN_kernel.md [title]
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg.per_second", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fma.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "TPC.TriageCompute.sm__cycles_active.avg",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("hbsm_b200::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    return name.strip()


def launches(src, dst, title):
    rows = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            ns *= 1e6
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += ns
        total += ns
    with open(dst, "w") as f:
        f.write("# %s\n\nSource: `ncu --metrics gpu__time_duration.sum --clock-control none` (per-launch times are cold-cache and "
                "serialised: read the SHARES).  %d launches, %.3f ms total.\n\n" % (title, sum(a[0] for a in agg.values()), total / 1e6))
        f.write("| kernel | launches | total ms | share | last grid | block |\n|---|---:|---:|---:|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.4f | %.1f%% | %s | %s |\n" % (k, a[0], a[1] / 1e6, 100 * a[1] / total, a[2], a[3]))
    print("wrote", dst)


def full(src, dst, title):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# %s\n\nSource: `ncu --set full --clock-control none --import-source on` (%s); values per launch.\n\n" % (title, src))
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write("## `%s` grid %s block %s\n\n| metric | value | unit |\n|---|---:|---|\n" %
                    (short(d.get("Kernel Name", "?")), d.get("Grid Size"), d.get("Block Size")))
            for i, h in enumerate(hdr):
                want = (h in KEEP or "dmma" in h or "pipe_tensor_cycles_active" in h or "inst_executed_pipe_tc" in h
                        or ("utc" in h and ".sum.pct" in h) or ("stalled" in h and h.endswith("per_issue_active.ratio")))
                if want and r[i] not in ("0", "", "n/a"):
                    f.write("| %s | %s | %s |\n" % (h, r[i], units[i]))
            rd = d.get("dram__bytes_read.sum"); wr = d.get("dram__bytes_write.sum")
            f.write("\n")
    print("wrote", dst)


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    (launches if mode == "launches" else full)(src, dst, title)
