"""Text summary of one `ncu --set full` report for profiles/: selected raw metrics of the first profiled launch,
the instruction-cache footprint (tools/ncu_footprint.py) and executed instructions per source file / hottest lines.
usage: ncu_lines.py file.ncu-rep [warp_substeps]   (warp_substeps: divides the per-line counts, e.g. warps x substeps)"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
W = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h, r = raw[0], raw[2]
print("--- launch 0", r[h.index("Kernel Name")])
for w in ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
          "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
          "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "l1tex__t_sector_hit_rate.pct",
          "dram__bytes_read.sum", "dram__bytes_write.sum",
          "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
          "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]:
    if w in h:
        print("   %s %s %s" % (w, raw[1][h.index(w)], r[h.index(w)]))
print(subprocess.run([sys.executable, __file__.replace("ncu_lines", "ncu_footprint"), rep], capture_output=True, text=True).stdout.strip())
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout.splitlines()))
byline, samp, src = collections.Counter(), collections.Counter(), {}
fil, hd = None, None
for x in rows:
    if not x:
        continue
    if x[0] == "File Path":
        fil = x[1].split("/")[-1]
        continue
    if x[0] == "Line No":
        hd = x
        ie, isamp = hd.index("Instructions Executed"), hd.index("# Samples")
        continue
    if hd is None or len(x) <= ie:
        continue
    try:
        ln, n, s = int(x[0]), int(x[ie]), int(x[isamp] or 0)
    except ValueError:
        continue
    byline[(fil, ln)] += n; samp[(fil, ln)] += s; src[(fil, ln)] = x[1].strip()[:100]
ts = max(1, sum(samp.values()))
byf = collections.Counter()
for (f, l), v in byline.items():
    byf[f] += v
print("executed warp-instructions per source file%s:" % (" / %g" % W if W != 1 else ""), {k: round(v / W, 1) for k, v in byf.items()})
print("hottest source lines (executed%s, share of stall samples):" % (" / %g" % W if W != 1 else ""))
for k, n in byline.most_common(16):
    print("   %-28s %8.0f  %5.1f%%  %s" % ("%s:%d" % k, n / W, 100 * samp[k] / ts, src[k]))
