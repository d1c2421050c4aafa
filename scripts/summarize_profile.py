"""Write a text summary of an ncu report (key metrics of the captured kernel + hottest source lines)
for committing under profiles/.

    python scripts/summarize_profile.py gpurun_out/step_exact.ncu-rep profiles/r1_step_exact.txt
"""
import csv
import subprocess
import sys

rep, out_path = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]
lines = ["ncu summary of %s" % rep.split("/")[-1],
         "(captured with: ncu --set full --clock-control none --import-source on -k regex:smc_step -s 4 -c 1 ...;",
         " absolute times under ncu are cold-cache and serialised -- bench.py's CUDA-event numbers are the timings)", ""]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        lines.append("%-86s %s %s" % (w, vals[i], units[i]))
rd = float(vals[hdr.index("dram__bytes_read.sum")]) if "dram__bytes_read.sum" in hdr else 0
wr = float(vals[hdr.index("dram__bytes_write.sum")]) if "dram__bytes_write.sum" in hdr else 0
lines.append("%-86s %.1f Mbyte" % ("DRAM traffic per launch (read + write)", rd + wr))
lines.append("")
src = subprocess.run([sys.executable, __file__.replace("summarize_profile.py", "ncu_lines.py"), rep, "25"],
                     capture_output=True, text=True).stdout
lines.append("hottest source lines (warp-stall samples / executed warp instructions):")
lines.append(src)
open(out_path, "w").write("\n".join(lines))
print("wrote", out_path)
