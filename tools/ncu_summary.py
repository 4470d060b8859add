"""Summarise an ncu --set full report (.ncu-rep) into a small csv: one row per profiled launch with the
metrics the roofline discussion needs.  usage: ncu_summary.py report.ncu-rep out.csv"""
import csv
import subprocess
import sys

WANT = [
    ("Kernel Name", "kernel"),
    ("gpu__time_duration.sum", "duration_us"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_limit_regs_blocks"),
    ("launch__occupancy_limit_shared_mem", "occ_limit_smem_blocks"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__inst_executed.sum", "warp_instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active_lanes_per_inst"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("lts__t_sectors_op_red.sum", "l2_red_sectors"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1_data_pipe_lsu_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1_data_pipe_shared_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared_bank_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared_wavefronts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_pipe"),
]

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
# ncu picks a unit per column and report (us / ms, Mbyte / Gbyte ...): normalise to microseconds and megabytes
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6}
NORM = {"duration_us": "us", "dram_read": "Mbyte", "dram_write": "Mbyte"}
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([n for _, n in WANT] + ["units: duration_us=us; dram_read=Mbyte; dram_write=Mbyte; percentages in %; stall_* = warps per issue-active cycle"])
    for r in rows[2:]:
        vals = []
        for m, n in WANT:
            v = r[idx[m]] if m in idx else ""
            if n == "kernel":
                v = v.split("(")[0].replace("dmr::", "")
            elif n in NORM and v:
                v = "%.6g" % (float(v.replace(",", "")) * SCALE.get(units[idx[m]], 1.0))
            vals.append(v)
        w.writerow(vals)
print("wrote", out, len(rows) - 2, "launches")
