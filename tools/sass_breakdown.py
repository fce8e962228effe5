"""Attribute the SASS of the rollout kernel to source functions (development aid).
usage: python tools/sass_breakdown.py   (after `python -m mjmpc_b200.build`)"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "mjmpc_b200", "libmjmpc_b200.so")], cwd=tmp, stdout=subprocess.DEVNULL)
txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, "rollout_reacher.sm_100a.cubin")], capture_output=True, text=True).stdout.split("\n")
src = open(os.path.join(ROOT, "mjmpc_b200", "csrc", "chain_dynamics.cuh")).read().split("\n")
# function start lines in chain_dynamics.cuh
marks = []
for i, l in enumerate(src, 1):
    m = re.match(r"(?:template <[^>]*>\s*)?MJB_(?:HD|NOINLINE) \S+.*?\b(\w+)\(", l)
    if m and not l.startswith(" "):
        marks.append((i, m.group(1)))
def fn_of(line):
    name = "?"
    for ln, n in marks:
        if ln <= line: name = n
    return name
want = sys.argv[1] if len(sys.argv) > 1 else "SawyerTraitsENS_11ConstParamsELb0"
sec = None; cur = None
per = collections.defaultdict(collections.Counter); tot = collections.Counter()
for l in txt:
    if l.startswith(".text."):
        sec = l.strip(); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        f = os.path.basename(m.group(1)); cur = fn_of(int(m.group(2))) if f == "chain_dynamics.cuh" else f; continue
    if re.match(r"\s+/\*[0-9a-f]+\*/\s+\S+", l) and sec:
        per[sec][cur] += 1; tot[sec] += 1
for s in tot:
    if want in s or "soft_row" in s or "line_search" in s or "contact_row" in s:
        print(s[:110], tot[s])
        if want in s:
            for k, v in per[s].most_common(): print("    %-28s %d" % (k, v))

# opcode histogram of the hot part (cold out-of-line helpers excluded)
cold = {"soft_row", "line_search", "contact_row"}
sec = None; cur = None
ops = collections.Counter(); byfn = collections.defaultdict(collections.Counter)
for l in txt:
    if l.startswith(".text."):
        sec = l.strip(); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        f = os.path.basename(m.group(1)); cur = fn_of(int(m.group(2))) if f == "chain_dynamics.cuh" else f; continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?(\S+)", l)
    if m and sec and want in sec and cur not in cold:
        op = m.group(2).split(".")[0].rstrip(";")
        ops[op] += 1; byfn[cur][op] += 1
print("hot opcode histogram:", sum(ops.values()))
print("  ", ops.most_common(30))
fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU"))
print("   fp64-pipe-ish:", fp64)
for fn in ("chain_substep", "make_rows", "rollout_reacher.cu", "rot", "sincos_joint", "?"):
    print("  ", fn, byfn[fn].most_common(12))
