"""Summarise an ncu report (read here, no GPU needed) into profiles/<name>.md:
   python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_mppi_tc_v1 "note"."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sm__cycles_active.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum.per_cycle_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
    "sm__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "launch__cluster_dim_x", "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
]


def main():
    rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ["# ncu --set full summary: %s" % rep, "", note, ""]
    for r in rows[2:]:
        lines.append("## launch id %s: %s  grid %s block %s" % (r[idx["ID"]], r[idx["Kernel Name"]], r[idx.get("Grid Size", 0)],
                                                               r[idx.get("Block Size", 0)]))
        lines.append("")
        lines.append("| metric | unit | value |")
        lines.append("|---|---|---|")
        for k in KEYS:
            if k in idx:
                lines.append("| %s | %s | %s |" % (k, units[idx[k]], r[idx[k]]))
        lines.append("")
    open(out + ".md", "w").write("\n".join(lines))
    print("wrote", out + ".md")


if __name__ == "__main__":
    main()
