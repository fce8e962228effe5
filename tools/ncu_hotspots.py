"""Join ncu per-instruction stall samples (sass source page) with nvdisasm line info of the same build,
and aggregate by source function.  usage: ncu_hotspots.py file.ncu-rep [kernel-substring]"""
import collections, csv, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else "SawyerTraitsENS_11ConstParamsELb0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
si, ii, ti = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "mjmpc_b200", "libmjmpc_b200.so")], cwd=tmp, stdout=subprocess.DEVNULL)
txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, "rollout_reacher.sm_100a.cubin")], capture_output=True, text=True).stdout.split("\n")
src = open(os.path.join(ROOT, "mjmpc_b200", "csrc", "chain_dynamics.cuh")).read().split("\n")
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r"(?:template <[^>]*>\s*)?MJB_(?:HD|NOINLINE) \S+.*?\b(\w+)\(", l)
    if m and not l.startswith(" "): marks.append((i, m.group(1)))
def fn_of(line):
    name = "?"
    for ln, n in marks:
        if ln <= line: name = n
    return name
sec = None; cur = None; seq = []
for l in txt:
    if l.startswith(".text."): sec = l.strip(); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        f = os.path.basename(m.group(1)); cur = (fn_of(int(m.group(2))) if f == "chain_dynamics.cuh" else f, int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?(\S+)", l)
    if m and sec and want in sec: seq.append((cur, m.group(2)))
print("ncu instrs", len(data), "nvdisasm instrs", len(seq))
n = min(len(data), len(seq))
S = collections.Counter(); I = collections.Counter(); T = collections.Counter(); L = collections.Counter()
for (key, op), r in zip(seq[:n], data[:n]):
    s, ie, te = int(r[si] or 0), int(r[ii] or 0), int(r[ti] or 0)
    S[key[0]] += s; I[key[0]] += ie; T[key[0]] += te; L[key] += s
tot = sum(S.values()); toti = sum(I.values())
print("%-24s %8s %7s %12s %7s %9s %6s" % ("function", "samples", "%", "warp-instrs", "%", "smp/kinst", "lanes"))
for k, v in S.most_common(24):
    print("%-24s %8d %6.1f%% %12d %6.1f%% %9.2f %6.1f" % (k, v, 100 * v / tot, I[k], 100 * I[k] / toti, 1000 * v / max(1, I[k]), T[k] / max(1, I[k])))
print("hottest source lines:")
for (k, ln), v in L.most_common(14): print("   %-22s line %4d  %5.1f%%" % (k, ln, 100 * v / tot))
