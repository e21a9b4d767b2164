"""ncu report (.ncu-rep) -> markdown table of the metrics the profiles/ summaries quote, one column per captured launch.
    python scripts/ncu_summary.py gpurun_out/prof_band4.ncu-rep profiles/r1f_ncu_band4_summary.md "title" """
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_dim_x", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warp_latency_per_inst_issued.ratio"]
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
kn = hdr.index("Kernel Name")
names = [r[kn].replace("<unnamed>::", "").split("(")[0] for r in data]
md = [f"# {title}", "", f"Source: `{rep}` (scratch), read with `ncu -i ... --page raw --csv`.  One column per captured launch.", "",
      "| metric | unit | " + " | ".join(f"`{n}`" for n in names) + " |", "|---|---|" + "---:|" * len(names)]
for k in KEYS:
    if k in hdr:
        i = hdr.index(k)
        md.append(f"| `{k}` | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
open(out, "w").write("\n".join(md) + "\n")
print(out, len(data), "launches")
