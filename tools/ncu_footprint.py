"""Instruction-cache footprint of a profiled kernel: 128-byte lines of SASS that were executed at all, and those
executed on (nearly) every loop trip, from an `ncu --set full` report.  usage: ncu_footprint.py file.ncu-rep [loop_trips]
(tools/icache_probe.cu: a loop body beyond ~32 KB costs ~2.8 cycles per instruction in instruction fetch alone)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ii = hdr.index("Instructions Executed")
ex = [int(r[ii] or 0) for r in data]
n = len(ex)
mx = max(ex)
lines = {}
for i, e in enumerate(ex):
    lines.setdefault(i // 8, []).append(e)
touched = sum(1 for v in lines.values() if max(v) > 0)
for frac in (0.9, 0.5, 0.1, 0.01):
    hot = sum(1 for v in lines.values() if max(v) >= frac * mx)
    print("lines with an instruction executed >= %4.0f%% of the hottest count: %4d = %5.1f KB" % (100 * frac, hot, hot * 128 / 1024))
print("static instructions %d (%.1f KB); lines touched at all %d = %.1f KB; executed warp-instructions %d" % (n, n * 16 / 1024, touched, touched * 128 / 1024, sum(ex)))
