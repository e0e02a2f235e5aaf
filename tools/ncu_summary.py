"""Summarise .ncu-rep captures (read here, no GPU needed) into profiles/*.txt: key metrics + stall reasons."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# %s  (ncu --set full --clock-control none; per launch)\n" % rep)
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            f.write("\n== %s  grid %s block %s\n" % (d.get("Kernel Name", "?")[:110], d.get("Grid Size", ""), d.get("Block Size", "")))
            for k in KEYS:
                if k in d:
                    f.write("%-78s %18s %s\n" % (k, d[k], units[hdr.index(k)]))
            stalls = []
            for h, v in zip(hdr, vals):
                if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                    try:
                        stalls.append((float(v), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                    except ValueError:
                        pass
            f.write("stall reasons (warps per issue-active cycle): " + ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)[:8]) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
