#!/usr/bin/env python
"""Summarise an .ncu-rep into markdown for profiles/ (run in the build container):

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.md "title" "command"
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
    "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep, out, title, cmd = sys.argv[1:5]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["# " + title, "", "Command: `%s`" % cmd, ""]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if d.get("gpu__time_duration.sum", "").startswith(("-", "n")):
            continue
        lines += ["## %s  grid %s block %s" % (d["Kernel Name"], d["Grid Size"], d["Block Size"]),
                  "", "| metric | value | unit |", "|---|---|---|"]
        for k in METRICS:
            if k in d:
                lines.append("| %s | %s | %s |" % (k, d[k], units[hdr.index(k)]))
        st = []
        for k in hdr:
            if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
                try:
                    st.append((float(d[k].replace(",", "")), k.split("stalled_")[1]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1.0
        lines += ["", "Warp-state samples: " + ", ".join(
            "%s %.1f %%" % (k, 100 * v / tot) for v, k in sorted(st, reverse=True)[:8]), ""]
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
