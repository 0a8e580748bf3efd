"""ncu report -> profiles/: writes the raw page of a .ncu-rep as CSV and prints the figures the round's summary quotes.
    python tools/ncu_extract.py gpurun_out/r2c_k2_lean2_dense.ncu-rep profiles/r2c_k2_lean2_dense_full_raw.csv"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "smsp__warps_active.avg.per_cycle_active", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    open(out, "w").write(txt)
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:90s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    main()
