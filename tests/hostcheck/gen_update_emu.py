"""TEST HARNESS (CPU): the textual transform that turns a product .cu file into a host translation unit for the
block emulator (block_emu.h) -- kernel<<<grid, block, smem, stream>>>(args) becomes emu::launch(grid, block, smem,
stream, [&]{ kernel(args); }), `extern __shared__ T x[];` becomes a pointer to the emulated dynamic shared
memory.  The source text is otherwise untouched: the kernels and their launch code are the product's.
gen_lib_emu.py applies it to every source of the library and builds libmjmpc_b200_emu.so."""
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def transform(src: str) -> str:
    src = re.sub(r"extern\s+__shared__\s+(\w+)\s+(\w+)\[\];", r"\1* \2 = (\1*)emu::g_dyn_smem;", src)
    out, pos = "", 0
    pat = re.compile(r"([A-Za-z_][\w:]*(?:<[^<>;()]*>)?)\s*<<<")
    while True:
        m = pat.search(src, pos)
        if not m:
            out += src[pos:]
            break
        end_cfg = src.index(">>>", m.end())
        cfg = _split_top(src[m.end():end_cfg])
        while len(cfg) < 4:
            cfg.append("0")
        i = src.index("(", end_cfg)
        depth, j = 0, i
        while True:
            depth += src[j] == "("
            depth -= src[j] == ")"
            if depth == 0:
                break
            j += 1
        args = src[i + 1:j].replace("\\\n", " ")
        out += src[pos:m.start()]
        out += "emu::launch(dim3(%s), dim3(%s), (size_t)(%s), (cudaStream_t)(%s), [&] { %s(%s); })" % (
            cfg[0], cfg[1], cfg[2], cfg[3], m.group(1), args)
        pos = j + 1
    return src if not out else out
