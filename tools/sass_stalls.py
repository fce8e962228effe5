"""Development aid: the stall counts ptxas encoded in the control bits of the rollout kernel (sm_100a: bits
105-108 of each 128-bit instruction), summed per basic block = the issue time of a lone warp through that
block as the compiler scheduled it (variable-latency waits on scoreboards come on top).
usage: python tools/sass_stalls.py [lib.so] [kernel-substring]"""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = [a for a in sys.argv[1:] if not a.startswith("--")]
lib = args[0] if args else os.path.join(ROOT, "mjmpc_b200", "libmjmpc_b200.so")
want = args[1] if len(args) > 1 else "SawyerTraitsENS_11ConstParamsELb0ELb0"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
on = False; ins = []
for i, l in enumerate(txt):
    if "Function :" in l: on = want in l; continue
    if not on: continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", l)
    if m:
        hi = int(re.search(r"0x([0-9a-f]{16})", txt[i + 1]).group(1), 16)
        body = m.group(2).strip()
        body = re.sub(r"^@!?U?P\d+\s+", "", body)
        ins.append(dict(addr=int(m.group(1), 16), op=body.split()[0].split(".")[0], body=body, stall=(hi >> 41) & 0xF,
                        yld=(hi >> 45) & 1, wait=(hi >> 52) & 0x3F))
# basic blocks: split after control flow and at branch targets
targets = set()
for x in ins:
    m = re.search(r"\b(?:BRA|BSSY|CALL)\S*\s+.*?0x([0-9a-f]+)", x["body"])
    if m: targets.add(int(m.group(1), 16))
blocks = [[]]
for x in ins:
    if x["addr"] in targets and blocks[-1]: blocks.append([])
    blocks[-1].append(x)
    if x["op"] in ("BRA", "EXIT", "RET", "CALL"): blocks.append([])
print("%-10s %6s %6s %8s %8s %6s" % ("addr", "instr", "fp64", "stalls", "cyc/ins", "waits"))
show_all = "--all" in sys.argv      # every block with its last instruction (to follow the hot path by hand)
for b in blocks:
    if not b or (len(b) < 40 and not show_all): continue
    nf = sum(1 for x in b if x["op"] in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU"))
    st = sum(max(x["stall"], 1) for x in b)
    print("0x%06x   %6d %6d %8d %8.2f %6d  %s" % (b[0]["addr"], len(b), nf, st, st / len(b), sum(1 for x in b if x["wait"]),
                                                 b[-1]["body"][:70] if show_all else ""))
