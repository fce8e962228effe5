"""Development aid: per-basic-block dependency analysis of the rollout kernel's SASS.

For every large basic block of the production instantiation: instruction count, issue-bound cycles (FP64
instructions occupy the FP64 pipe for 2 cycles per warp on sm_100, others 1) and the length of the longest
register-dependency chain under the latencies below (measured by tools/lat_probe.cu).  A block whose chain is
much longer than its issue time is latency-bound for a lone warp; the chain listing shows which source lines
sit on it.
usage: python tools/sass_critpath.py [lib.so] [kernel-substring] [--chain]
"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = [a for a in sys.argv[1:] if not a.startswith("--")]
lib = args[0] if args else os.path.join(ROOT, "mjmpc_b200", "libmjmpc_b200.so")
want = args[1] if len(args) > 1 else "SawyerTraitsENS_11ConstParamsELb0ELb0"
show_chain = "--chain" in sys.argv
LAT = dict(DFMA=float(os.environ.get("LAT_FP64", 8)), DMUL=float(os.environ.get("LAT_FP64", 8)),
           DADD=float(os.environ.get("LAT_FP64", 8)), DSETP=float(os.environ.get("LAT_DSETP", 10)),
           MUFU=float(os.environ.get("LAT_MUFU", 20)), LDS=29, LDC=8, LDCU=8, LDG=400, LDL=30)
DEFAULT_LAT = 4.5
FP64 = {"DFMA", "DMUL", "DADD", "DSETP"}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.startswith("rollout_reacher") and f.endswith(".cubin")]
cub = cub[0] if cub else [f for f in os.listdir(tmp) if "rollout_reacher" in f][0]
txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
on = False; cur_line = None; blocks = [[]]
for l in txt:
    if l.startswith(".text."):
        on = want in l; continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\.L_x_\d+:", l.strip()):
        blocks.append([]); continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if not m: continue
    text = m.group(2).strip()
    pm = re.match(r"(@!?U?P\d+)\s+(.*)", text)
    pred = pm.group(1) if pm else None
    body = pm.group(2) if pm else text
    op = body.split()[0]
    base = op.split(".")[0]
    blocks[-1].append(dict(addr=int(m.group(1), 16), op=op, base=base, pred=pred, body=body, line=cur_line))
    if base in ("BRA", "EXIT", "RET", "CALL", "BSYNC", "BSSY"):
        if base in ("BRA", "EXIT", "RET", "CALL"): blocks.append([])

def regs_of(tok, wide):
    out = []
    for m in re.finditer(r"\b(U?R)(\d+)\b", tok):
        n = int(m.group(2)); out.append((m.group(1), n))
        if wide: out.append((m.group(1), n + 1))
    for m in re.finditer(r"\b(U?P)(\d)\b", tok):
        out.append((m.group(1), int(m.group(2))))
    return out

def analyse(b):
    ready = collections.defaultdict(float); src_of = {}
    best = (0.0, None); finish = []; parent = []
    # in-order issue of a lone warp: an instruction issues when the previous one has issued (FP64: the pipe
    # is busy 2 cycles) and its operands are ready
    t_issue = 0.0; ready_io = collections.defaultdict(float); stall_by_line = collections.Counter()
    for i, ins in enumerate(b):
        ops = ins["body"][len(ins["op"]):].split(",")
        wide = ins["base"] in ("DFMA", "DMUL", "DADD", "DSETP") or ".64" in ins["op"] or ins["base"] == "MUFU" and "64" in ins["op"]
        dst_tok, src_toks = (ops[0], ops[1:]) if ops else ("", [])
        if ins["base"] in ("STS", "STG", "STL", "BRA", "BSSY", "BSYNC", "EXIT", "LDGSTS"): src_toks = ops; dst_tok = ""
        n_dst = 2 if ins["base"] in ("DSETP", "ISETP", "FSETP", "PLOP3") else 1
        if n_dst == 2: dst_tok = ",".join(ops[:2]); src_toks = ops[2:]
        srcs = []
        for t in src_toks: srcs += regs_of(t, wide and not t.strip().startswith("P"))
        if ins["pred"]: srcs += regs_of(ins["pred"].lstrip("@!"), False)
        start, par = 0.0, None
        for r in srcs:
            if ready[r] > start: start, par = ready[r], src_of.get(r)
        lat = LAT.get(ins["base"], DEFAULT_LAT)
        op_ready = max([ready_io[r] for r in srcs] + [0.0])
        if op_ready > t_issue:
            stall_by_line["%s:%d" % ins["line"] if ins["line"] else "?"] += op_ready - t_issue
            t_issue = op_ready
        fin_io = t_issue + lat
        t_issue += 2.0 if (ins["base"] in FP64 or ins["base"] == "MUFU") else 1.0
        fin = start + lat
        finish.append(fin); parent.append(par)
        for r in regs_of(dst_tok, wide and ins["base"] not in ("DSETP",)):
            ready[r] = fin; src_of[r] = i; ready_io[r] = fin_io
        if fin > best[0]: best = (fin, i)
    chain = []
    i = best[1]
    while i is not None:
        chain.append(i); i = parent[i]
    return best[0], chain[::-1], t_issue, stall_by_line

tot_issue = tot_cp = 0
print("%-10s %6s %6s %8s %8s %8s  %s" % ("addr", "instr", "fp64", "issue", "chain", "in-order", "largest in-order stalls by source line"))
for b in blocks:
    if len(b) < 40: continue
    nf = sum(1 for i in b if i["base"] in FP64 or i["base"] == "MUFU")
    issue = len(b) + nf
    cp, chain, t_io, stalls = analyse(b)
    src = collections.Counter()
    for i in chain:
        if b[i]["line"]: src["%s:%d" % b[i]["line"]] += LAT.get(b[i]["base"], DEFAULT_LAT)
    print("0x%06x   %6d %6d %8d %8.0f %8.0f  %s" % (b[0]["addr"], len(b), nf, issue, cp, t_io,
          ", ".join("%s(%d)" % (k.replace("chain_dynamics.cuh", "cd"), v) for k, v in stalls.most_common(8))))
    if show_chain:
        for i in chain: print("      %-60s %s" % (b[i]["body"][:60], b[i]["line"]))
