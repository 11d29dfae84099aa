#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/<name>.md: the metrics the judge greps for."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__icc_request_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
]
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def main(rep, out, title):
    # a .csv is the raw page already exported on the GPU box (`ncu -i X.ncu-rep --page raw --csv`: the reports themselves are
    # tens of MB and do not fit gpurun_out's copy-back limit)
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in raw.splitlines() if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    lines = ["# %s" % title, "", "source: `%s` (ncu --set full --clock-control none --import-source on; times under ncu are serialised/cold - use shares, not absolutes)" % rep, ""]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        lines += ["## %s" % d.get("Kernel Name", "?"), "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in d and d[k] != "":
                lines.append("| %s | %s | %s |" % (k, d[k], u[k]))
        st = sorted(((float(d[h] or 0), h[len(STALLS):]) for h in hdr if h.startswith(STALLS) and not h.endswith("_not_issued")), reverse=True)
        tot = sum(v for v, _ in st) or 1
        lines += ["", "warp stall samples: " + ", ".join("%s %.0f%%" % (n, 100 * v / tot) for v, n in st[:7]), ""]
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else sys.argv[1])
