"""TEST INFRASTRUCTURE: rewrites the inline PTX of the device headers into calls of a PTX interpreter, so that the DEVICE
code paths (everything under `#if defined(__CUDA_ARCH__)`: the carry-chain field arithmetic of fe.cuh, the long-number
helpers of hgcd.cuh, the funnel-shift SHA-512, 128-bit loads, cp.async staging) can be compiled for the host and checked
on the CPU by the same unit tests as the portable host paths (tests/test_host_sim.py).

    python ptx_rewrite.py <src dir> <dst dir>      # every *.cuh / *.h / *.cu of src, asm statements (and, in the .cu files,
                                                   # kernel launches: see rewrite_launches) replaced, written to dst

Every statement

    asm [volatile] ( "template" [: outputs [: inputs [: clobbers]]] );

becomes

    { static const edg_ptx::Prog p_ = edg_ptx::compile("template");
      edg_ptx::Arg a_[] = { edg_ptx::out(lvalue, "+r"), ..., edg_ptx::in(value, "r"), ... };
      edg_ptx::run(p_, a_, count); }

with the operands numbered as in the original (%0 .. : outputs first, then inputs).  ptx_emul.h holds the interpreter and
the shims for the CUDA keywords and intrinsics the headers use.  The product never sees any of this.
"""
import os
import re
import sys


def split_top(s, sep):
    """Split s at top-level occurrences of the character sep (outside parentheses, brackets and string literals)."""
    parts, depth, cur, i, in_str = [], 0, [], 0, False
    while i < len(s):
        ch = s[i]
        if in_str:
            cur.append(ch)
            if ch == "\\":
                cur.append(s[i + 1])
                i += 1
            elif ch == '"':
                in_str = False
        elif ch == '"':
            in_str = True
            cur.append(ch)
        elif ch in "([{":
            depth += 1
            cur.append(ch)
        elif ch in ")]}":
            depth -= 1
            cur.append(ch)
        elif ch == sep and depth == 0:                  # (no C++ "::" occurs in the operand expressions of these headers)
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
        i += 1
    parts.append("".join(cur))
    return parts


def find_matching_paren(s, start):
    depth, i, in_str = 0, start, False
    while i < len(s):
        ch = s[i]
        if in_str:
            if ch == "\\":
                i += 1
            elif ch == '"':
                in_str = False
        elif ch == '"':
            in_str = True
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced parentheses in an asm statement")


def operands(section):
    """'"=r"(x), "+r"(y[0])' -> [('=r', 'x'), ('+r', 'y[0]')]"""
    out = []
    for item in split_top(section, ","):
        item = item.strip()
        if not item:
            continue
        m = re.match(r'"([^"]*)"\s*\((.*)\)\s*$', item, re.S)
        if not m:
            raise ValueError("cannot parse asm operand: " + item)
        out.append((m.group(1), m.group(2).strip()))
    return out


ASM_RE = re.compile(r"\basm\s*(?:volatile\s*)?\(")


def rewrite(text, name):
    out, pos, count = [], 0, 0
    while True:
        m = ASM_RE.search(text, pos)
        if not m:
            break
        line_start = text.rfind("\n", 0, m.start()) + 1
        if "//" in text[line_start:m.start()]:          # the word in a comment
            out.append(text[pos:m.end()])
            pos = m.end()
            continue
        open_paren = m.end() - 1
        close = find_matching_paren(text, open_paren)
        semi = text.index(";", close)
        body = text[open_paren + 1:close]
        sections = split_top(body, ":")
        template = "".join(re.findall(r'"((?:[^"\\]|\\.)*)"', sections[0]))
        outs = operands(sections[1]) if len(sections) > 1 else []
        ins = operands(sections[2]) if len(sections) > 2 else []
        args = [f'edg_ptx::out({e}, "{c}")' for c, e in outs] + [f'edg_ptx::in({e}, "{c}")' for c, e in ins]
        count += 1
        if template.strip() == "":                      # an optimisation barrier only
            repl = "{ }"
        else:
            arr = "edg_ptx::Arg a_[] = { " + ", ".join(args) + " }; " if args else "edg_ptx::Arg *a_ = nullptr; "
            repl = ('{ static const edg_ptx::Prog p_ = edg_ptx::compile("' + template + '"); ' + arr +
                    f"edg_ptx::run(p_, a_, {len(args)}); }}")
        out.append(text[pos:m.start()])
        out.append(repl)
        pos = semi + 1
    out.append(text[pos:])
    return "".join(out), count


LAUNCH_RE = re.compile(r"([A-Za-z_][A-Za-z_0-9]*(?:<[^<>;(){}]*>)?)\s*<<<")
DYN_SMEM_RE = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\[\];")


def rewrite_launches(text):
    """kernel<<<grid, block[, smem[, stream]]>>>(args);  ->  edg_simt::launch(grid, block, smem, stream, [=] { kernel(args); });
    and  extern __shared__ T name[];  ->  T *name = (T *)the block's dynamic shared memory  (simt_emul.h)."""
    out, pos, count = [], 0, 0
    while True:
        m = LAUNCH_RE.search(text, pos)
        if not m:
            break
        close = text.index(">>>", m.end())
        cfg = [c.strip() for c in split_top(text[m.end():close], ",")]
        while len(cfg) < 4:
            cfg.append("0")
        args_open = text.index("(", close)
        args_close = find_matching_paren(text, args_open)
        semi = text.index(";", args_close)
        out.append(text[pos:m.start()])
        out.append(f"edg_simt::launch({cfg[0]}, {cfg[1]}, {cfg[2]}, (void *)({cfg[3]}), [=] {{ {m.group(1)}{text[args_open:args_close + 1]}; }});")
        pos = semi + 1
        count += 1
    out.append(text[pos:])
    text = "".join(out)
    text = DYN_SMEM_RE.sub(lambda d: f"{d.group(1)} *{d.group(2)} = ({d.group(1)} *)edg_simt::cur->blk->dyn;", text)
    return text, count


COUNT_OLD = """#if defined(EDG_COUNT_OPS) && !defined(__CUDA_ARCH__)
static unsigned long edg_cnt_mul = 0, edg_cnt_sq = 0;"""
COUNT_NEW = """#if defined(EDG_EMUL_COUNT)       /* (inserted by ptx_rewrite.py) the SIMT emulator counts field operations of the real kernels */
#define EDG_COUNT_MUL() ((void)edg_ptx::cnt_mul.fetch_add(1, std::memory_order_relaxed))
#define EDG_COUNT_SQ() ((void)edg_ptx::cnt_sq.fetch_add(1, std::memory_order_relaxed))
#elif defined(EDG_COUNT_OPS) && !defined(__CUDA_ARCH__)
static unsigned long edg_cnt_mul = 0, edg_cnt_sq = 0;"""


def main(src, dst):
    os.makedirs(dst, exist_ok=True)
    total = 0
    launches = 0
    for f in sorted(os.listdir(src)):
        if f.endswith((".cuh", ".h", ".cu")):
            new, n = rewrite(open(os.path.join(src, f)).read(), f)
            total += n
            if f == "fe.cuh":
                assert new.count(COUNT_OLD) == 1
                new = new.replace(COUNT_OLD, COUNT_NEW)
            if f.endswith(".cu"):
                new, k = rewrite_launches(new)
                launches += k
            with open(os.path.join(dst, f + ".cpp" if f.endswith(".cu") else f), "w") as fh:
                fh.write("// GENERATED by tests/host_sim/ptx_rewrite.py from libeddsa_b200/csrc/" + f + " — test infrastructure, do not edit\n" + new)
    print(f"rewrote {total} asm statements and {launches} kernel launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
