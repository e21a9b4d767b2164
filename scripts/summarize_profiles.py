"""Turn the scratch captures under gpurun_out/ into the tracked summaries under profiles/ (round tag as argv[1])."""
import csv, collections, json, os, re, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r1e"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)

def launches(path, out, title, cmd):
    rows = list(csv.reader(open(path, errors="ignore")))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg, tot, n = collections.OrderedDict(), 0.0, 0
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        name = re.split(r"[<(]", re.sub(r"^void ", "", r[ki].replace("<unnamed>::", "").replace("(anonymous namespace)::", "")))[0].strip()
        try:
            v = float(r[vi].replace(",", "")) / 1000.0
        except ValueError:
            continue
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v; n += 1
    md = [f"# {title}", "", f"Command (under gpurun, 1x B200): `{cmd}`", "",
          "Times are cold-cache and serialised under the profiler: compare shares, not absolutes.", "",
          f"Total {tot/1000:.2f} ms over {n} launches.", "", "| kernel | launches | total us | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
        md.append(f"| `{k[-70:]}` | {c} | {t:.1f} | {100*t/tot:.1f}% | {t/c:.2f} |")
    open(out, "w").write("\n".join(md) + "\n")
    return agg, tot

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warp_latency_per_inst_issued.ratio"]

def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return {k: (units[hdr.index(k)], data[0][hdr.index(k)]) for k in KEYS if k in hdr}

agg, tot = launches("gpurun_out/%s_launches.csv" % tag, "profiles/%s_launches.md" % tag,
                    "Round 1, capture %s: launch list of the tracked frame (banded Cholesky v3, canonical J^T J)" % tag[-1].upper(),
                    "ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/%s_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline" % tag)
md = ["# Round 1, capture %s: `ncu --set full --clock-control none` summaries (one launch each)" % tag[-1].upper(), "",
      "Commands: `ncu --set full --clock-control none --import-source on -k regex:data_jtj_kernel -s 30 -c 1 ... python bench.py --steps 4 --warmup 3`",
      "and `scripts/ncu_band3.sh` (band_chol3_kernel alone, n=1862, bw=370, 148 CTAs).  The .ncu-rep files stay in gpurun_out/ (scratch).", ""]
traffic = None
for title, path in (("data_jtj_kernel", "gpurun_out/prof_jtj.ncu-rep"), ("band_chol3_kernel", "gpurun_out/prof_band3.ncu-rep")):
    m = full(path)
    md += [f"## {title}", "", "| metric | unit | value |", "|---|---|---:|"] + [f"| `{k}` | {u} | {v} |" for k, (u, v) in m.items()] + [""]
    if title == "data_jtj_kernel":
        f = lambda x: float(x[1]) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}[x[0]]
        traffic = {"read": f(m["dram__bytes_read.sum"]), "write": f(m["dram__bytes_write.sum"])}
open("profiles/%s_ncu_full_summary.md" % tag, "w").write("\n".join(md) + "\n")
print(open("profiles/%s_launches.md" % tag).read()[:2200])
print(traffic)
