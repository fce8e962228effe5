"""Per-source-line opcode breakdown of a line range of chain_dynamics.cuh / rollout_reacher.cu in the production kernel.
usage: sass_lines.py <file> <first> <last>"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fname, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "mjmpc_b200", "libmjmpc_b200.so")], cwd=tmp, stdout=subprocess.DEVNULL)
txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, "rollout_reacher.sm_100a.cubin")], capture_output=True, text=True).stdout.split("\n")
sec = None; cur = None
cnt = collections.Counter()
for l in txt:
    if l.startswith(".text."): sec = l; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", l)
    if m and sec and "SawyerTraitsENS_11ConstParamsELb0" in sec and cur and cur[0] == fname and lo <= cur[1] <= hi:
        ins = m.group(1)
        op = ins.split()[1] if ins.startswith('@') else ins.split()[0]
        cnt[(cur[1], op.split('.')[0])] += 1
src = open(os.path.join(ROOT, "mjmpc_b200", "csrc", fname)).read().split('\n')
byline = collections.Counter()
for (ln, op), c in cnt.items(): byline[ln] += c
for ln, c in sorted(byline.items()):
    ops = ", ".join("%s:%d" % (op, n) for (l2, op), n in sorted(cnt.items(), key=lambda x: -x[1]) if l2 == ln)
    print("%4d %4d | %-72s | %s" % (ln, c, src[ln - 1].strip()[:72], ops[:100]))
print("total", sum(byline.values()))
