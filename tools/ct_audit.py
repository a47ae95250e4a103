#!/usr/bin/env python3
"""Constant-time SASS audit of the secret-key kernels (sm_100a).

Static taint analysis over the disassembly (cuobjdump -sass) of each kernel that touches a secret
(k_x25519, k_comb<0/1>, k_expand_key, k_sign_nonce<>, k_sign_finish<>, k_sk_convert):

  * sources : every value loaded from global memory through a pointer derived from a SECRET kernel
              parameter (the secret key / scalar arrays); everything computed from such values;
              local-memory (stack) loads once any secret has been stored to the stack.
  * sinks   : (1) any branch / jump / call / exit whose guard predicate or target is tainted,
              (2) any memory instruction (LDG/STG/LDS/STS/LDL/STL/LDC/ATOM/RED) whose ADDRESS
                  registers or guard predicate are tainted,
              (3) variable-latency arithmetic (MUFU, integer/FP division helpers, I2F/F2I ...) on
                  tainted operands,
              (4) shared memory: a tainted value may be STORED to shared memory at an untainted address (the tensor-core
                  table lookup hands each lane its entry through a lane-indexed exchange area), but from then on EVERY
                  shared-memory load is treated as secret, exactly like the stack.
  The analysis is flow-sensitive (per-instruction register / predicate / uniform-register state,
  iterated to a fixed point over the control-flow graph including CALL/RET edges) and conservative:
  a guard-predicated write merges the old and new taint, RET goes to every call-return site, BRX is
  treated as a sink.

Exit status 0 and "PASS" for every kernel means: no secret-dependent branch, no secret-dependent
address, no variable-latency instruction on secrets — the property BASELINE.json's north_star asks
for ("masked table selects and cswap with no secret-dependent branches or addressing, checked by a
SASS audit").  The verify kernel is public-data-only and is reported for information.

usage: python tools/ct_audit.py [--lib libeddsa_b200/libeddsa_b200.so] [--json out.json] [-v]
"""
import argparse
import json
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# kernel-name substring -> byte offsets (within the parameter block) of the SECRET pointer parameters
# signatures: see libeddsa_b200/csrc/kernels_*.cu   (every parameter is 8 bytes wide)
SECRET_PARAMS = {
    "k_x25519E": {"params": ["n", "out", "scalar", "point"], "secret": ["scalar"]},
    # the comb kernel shared by genpub / sign (MODE 0: scalars reduced mod L by the hash kernels) and x25519_base (MODE 1)
    "k_combILi0": {"params": ["n", "out", "out_stride", "scalars", "wipe", "comb_g"], "secret": ["scalars"]},
    "k_combILi1": {"params": ["n", "out", "out_stride", "scalars", "wipe", "comb_g"], "secret": ["scalars"]},
    "k_expand_key": {"params": ["n", "a_out", "sec"], "secret": ["sec"]},
    "k_sign_nonceILb0": {"params": ["n", "a_out", "r_out", "sec", "msgs", "off", "fixed_len"], "secret": ["sec"]},          # fixed-length batches
    "k_sign_nonceILb1": {"params": ["n", "a_out", "r_out", "sec", "msgs", "off", "fixed_len"], "secret": ["sec"]},          # ragged batches (length-sorted tiles)
    "k_sign_finishILb0": {"params": ["n", "sig", "a_in", "r_in", "pub", "msgs", "off", "fixed_len"], "secret": ["a_in", "r_in"]},
    "k_sign_finishILb1": {"params": ["n", "sig", "a_in", "r_in", "pub", "msgs", "off", "fixed_len"], "secret": ["a_in", "r_in"]},
    "k_sk_convert": {"params": ["n", "out", "in"], "secret": ["in"]},
}
PUBLIC_KERNELS = ["k_verify", "k_pk_convert"]

PTR = "ptr"     # derived from a secret POINTER parameter (an address, not itself secret)
SEC = 0xFFFFFFFF  # secret data: taint is a 32-bit mask of the bits that may depend on a secret, so that
                  # predicates ptxas packs into spare bits of a general register (@P LOP3 R, R, 0x10 ... /
                  # LOP3 P, RZ, R, 0x2000) are tracked bit by bit instead of poisoning each other


def is_sec(v):
    return isinstance(v, int) and v != 0


def tjoin(a, b):
    """least upper bound of two taint values (None < PTR < bit masks ordered by inclusion)"""
    if is_sec(a) or is_sec(b):
        return (a if is_sec(a) else 0) | (b if is_sec(b) else 0)
    return a or b

BRANCH_OPS = ("BRA", "BRX", "JMP", "JMX", "CALL", "RET", "EXIT", "BREAK", "BSYNC", "WARPSYNC", "YIELD", "NANOSLEEP", "KILL", "BPT", "RTT")
MEM_OPS = ("LDG", "STG", "LDS", "STS", "LDL", "STL", "LD", "ST", "ATOM", "ATOMS", "ATOMG", "RED", "LDSM", "LDGSTS", "LDGDEPBAR", "CCTL")
VARLAT_OPS = ("MUFU", "IDIV", "I2F", "F2I", "F2F", "I2FP", "FCHK", "DFMA", "DMUL", "DADD", "RRO")
PRED_ONLY_DEST = ("ISETP", "FSETP", "DSETP", "HSETP2", "PLOP3", "UISETP", "UPLOP3", "VOTE", "VOTEU", "R2P", "PSETP", "UP2UR")
CARRY_OUT = ("IADD3", "LEA", "IMAD.WIDE", "IMAD.X", "UIADD3", "ULEA", "VIADD", "IMAD.HI", "UIMAD.WIDE")

REG_RE = re.compile(r"^(-|~|!|\|)?(U?R\d+|U?RZ|U?P\d+|U?PT)(\.64|\.reuse|\.H0_H0|\.H1_H1|\.X\d|\.F32|\.B\d|\.U32|\.S32|\|)*$")


class Insn:
    __slots__ = ("addr", "guard", "neg", "op", "operands", "text")

    def __init__(self, addr, guard, neg, op, operands, text):
        self.addr, self.guard, self.neg, self.op, self.operands, self.text = addr, guard, neg, op, operands, text


def disassemble(path):
    """{function name: [Insn]} for every kernel in the object / shared library."""
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    line_re = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*(/\*.*)?$")
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            continue
        m = line_re.match(line)
        if not m or cur is None:
            continue
        addr, body = int(m.group(1), 16), m.group(2).strip()
        guard, neg = None, False
        g = re.match(r"^@(!?U?P\d+|!?U?PT)\s+(.*)$", body)
        if g:
            guard, neg, body = g.group(1).lstrip("!"), g.group(1).startswith("!"), g.group(2)
        parts = body.split(None, 1)
        op = parts[0]
        operands = split_operands(parts[1]) if len(parts) > 1 else []
        cur.append(Insn(addr, guard, neg, op, operands, m.group(2).strip()))
    return funcs


def split_operands(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "[(":
            depth += 1
        elif ch in "])":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def regs_in(token):
    """register names (R#, UR#, P#, UP#) mentioned in an operand token, expanding .64 pairs."""
    names = []
    for m in re.finditer(r"(U?R)(\d+)(\.64)?|(U?P)(\d+)", token):
        if m.group(1):
            base, n = m.group(1), int(m.group(2))
            names.append(f"{base}{n}")
            if m.group(3):
                names.append(f"{base}{n + 1}")
        else:
            names.append(f"{m.group(4)}{m.group(5)}")
    return names


def is_plain_reg(token):
    return REG_RE.match(token) is not None


def width_regs(op):
    if ".128" in op:
        return 4
    if ".64" in op and not op.startswith(("SHF", "IMAD", "ISETP", "LEA", "USHF")):
        return 2
    return 1


def expand(name, n):
    m = re.match(r"(U?R)(\d+)$", name)
    if not m or n == 1:
        return [name]
    return [f"{m.group(1)}{int(m.group(2)) + i}" for i in range(n)]


def classify(ins):
    """-> (dest names, source names, address names, store-data names)"""
    op, ops = ins.op, ins.operands
    base = op.split(".")[0]
    dests, srcs, addr, data = [], [], [], []
    if base in ("STG", "STS", "STL", "ST", "RED", "ATOMS", "ATOMG", "ATOM") or base in ("LDG", "LDS", "LDL", "LD", "LDSM"):
        mem = [t for t in ops if "[" in t]
        for t in mem:
            addr += regs_in(t)
        rest = [t for t in ops if "[" not in t]
        if base.startswith("ST") or base == "RED":
            for t in rest:
                for r in regs_in(t):
                    data += expand(r, width_regs(op))
        else:
            seen_dest = False
            for t in rest:
                rs = regs_in(t)
                if not seen_dest and rs and is_plain_reg(t) and not rs[0].lstrip("U").startswith("P"):
                    dests += expand(rs[0], width_regs(op))
                    seen_dest = True
                elif not seen_dest and rs and rs[0].lstrip("U").startswith("P"):
                    dests += rs
                else:
                    data += rs if base.startswith("ATOM") else []
                    srcs += rs
        return dests, srcs, addr, data
    if base in ("LDC", "ULDC", "LDCU"):
        if ops:
            n = 2 if ".64" in op else (4 if ".128" in op else 1)
            dests += expand(regs_in(ops[0])[0], n) if regs_in(ops[0]) else []
            for t in ops[1:]:
                inner = re.search(r"\[(?:0x[0-9a-f]+|\w+)\]\[(.*)\]", t)
                if inner:
                    addr += regs_in(inner.group(1))
                addr += [r for r in regs_in(t) if r not in addr]
        return dests, srcs, addr, data
    if base in PRED_ONLY_DEST or op.startswith(PRED_ONLY_DEST):
        dests = [r for t in ops[:2] for r in regs_in(t)]
        srcs = [r for t in ops[2:] for r in regs_in(t)]
        return dests, srcs, addr, data
    if base in BRANCH_OPS or base in ("BSSY", "BAR", "NOP", "DEPBAR", "MEMBAR", "ERRBAR", "BMOV", "CCTL", "FENCE", "ACQBULK"):
        srcs = [r for t in ops for r in regs_in(t)]
        return dests, srcs, addr, data
    if base in ("IMMA", "HMMA"):
        # D (4 registers), A (2 for .16816 u8 / 4 for .16832), B (1 / 2), C (4 or RZ): everything is data
        na, nb = (4, 2) if ".16832" in op else (2, 1)
        regs = [regs_in(t) for t in ops]
        if regs and regs[0]:
            dests += expand(regs[0][0], 4)
        for t, n_ in zip(regs[1:], (na, nb, 4)):
            if t:
                srcs += expand(t[0], n_)
        return dests, srcs, addr, data
    # generic ALU form: leading predicate dests, one register dest, optional carry-out predicates
    i = 0
    while i < len(ops) and is_plain_reg(ops[i]) and regs_in(ops[i]) and regs_in(ops[i])[0].lstrip("U").startswith("P"):
        dests += regs_in(ops[i])
        i += 1
    if i < len(ops) and is_plain_reg(ops[i]):
        rs = regs_in(ops[i])
        if rs:
            wide = 2 if (op.startswith(("IMAD.WIDE", "UIMAD.WIDE", "CS2R", "DADD", "DFMA", "DMUL")) or (base in ("MOV", "UMOV") and ".64" in op)) else 1
            dests += expand(rs[0], wide)
        i += 1
        if op.startswith(CARRY_OUT):
            while i < len(ops) and is_plain_reg(ops[i]) and regs_in(ops[i]) and regs_in(ops[i])[0].lstrip("U").startswith("P") and len(ops) - i > 2:
                dests += regs_in(ops[i])
                i += 1
    for k, t in enumerate(ops[i:]):
        rs = regs_in(t)
        srcs += rs
        # 64-bit addend of IMAD.WIDE (third source) occupies a register pair
        if op.startswith(("IMAD.WIDE", "UIMAD.WIDE")) and k == 2 and rs and is_plain_reg(t):
            srcs += expand(rs[0], 2)[1:]
    return dests, srcs, addr, data


def param_base(insns):
    """lowest constant-bank-0 offset that looks like the parameter block (0x380 on sm_100, 0x210 older)."""
    offs = []
    for ins in insns:
        for t in ins.operands:
            m = re.search(r"c\[0x0\]\[0x([0-9a-f]+)\]", t)
            if m:
                offs.append(int(m.group(1), 16))
    cand = [o for o in offs if o >= 0x160]
    for base in (0x380, 0x210, 0x160):
        if any(base <= o < base + 0x80 for o in cand):
            return base
    return 0x380


def audit(name, insns, secret_offsets, verbose=False):
    index = {ins.addr: k for k, ins in enumerate(insns)}
    n = len(insns)
    succ = [[] for _ in range(n)]
    call_returns = []
    for k, ins in enumerate(insns):
        base = ins.op.split(".")[0]
        tgt = None
        for t in ins.operands:
            m = re.match(r"^\(?\*?\"?(?:BRANCH_TARGETS)?.*$", t)
            mm = re.search(r"(?<![\[\w])0x([0-9a-f]+)$", t)
            if mm and base in ("BRA", "CALL", "BSSY", "JMP", "RET", "BREAK"):
                tgt = int(mm.group(1), 16)
        if base == "EXIT" or base == "KILL":
            if ins.guard and k + 1 < n:
                succ[k].append(k + 1)
            continue
        if base == "RET":
            continue    # filled below
        if base in ("BRA", "JMP"):
            if tgt is not None and tgt in index:
                succ[k].append(index[tgt])
            if ins.guard or tgt is None or any(t in ("DIV",) for t in ins.operands) or len([t for t in ins.operands if regs_in(t) and regs_in(t)[0].startswith(("P", "UP"))]) > 0:
                if k + 1 < n:
                    succ[k].append(k + 1)
            continue
        if base == "CALL":
            if tgt is not None and tgt in index:
                succ[k].append(index[tgt])
            if k + 1 < n:
                call_returns.append(k + 1)
                succ[k].append(k + 1)   # conservative: also fall through
            continue
        # BSSY only records a reconvergence point; control does not transfer there
        if k + 1 < n:
            succ[k].append(k + 1)
    # match every RET with the return sites of the calls that can reach it (one level of context:
    # a callee's body is everything reachable from its entry without following nested call edges)
    calls_to = defaultdict(list)          # callee entry index -> [return-site index]
    call_target = {}
    for k, ins in enumerate(insns):
        if ins.op.split(".")[0] == "CALL":
            tg = [s_ for s_ in succ[k] if s_ != k + 1]
            if tg:
                calls_to[tg[0]].append(k + 1)
                call_target[k] = tg[0]
    rets_of = defaultdict(set)
    for entry in calls_to:
        seen, todo = set(), [entry]
        while todo:
            j = todo.pop()
            if j in seen or j >= n:
                continue
            seen.add(j)
            b = insns[j].op.split(".")[0]
            if b == "RET":
                rets_of[j].add(entry)
                if insns[j].guard and j + 1 < n:
                    todo.append(j + 1)
                continue
            if b == "CALL":
                if j + 1 < n:
                    todo.append(j + 1)      # skip over the nested callee
                continue
            todo.extend(succ[j])
    for k, ins in enumerate(insns):
        if ins.op.split(".")[0] == "RET":
            sites = []
            for entry in rets_of.get(k, ()):
                sites += calls_to[entry]
            succ[k] = sorted(set(sites)) if sites else list(call_returns)
            if ins.guard and k + 1 < n:
                succ[k].append(k + 1)
    for k in call_target:                   # a call does not fall through: control comes back via RET
        succ[k] = [call_target[k]]

    state = [None] * n          # dict name -> SEC / PTR
    state[0] = {}
    stack_secret = [False]
    smem_secret = [False]
    violations = {}
    work = [0]
    inwork = {0}

    def get(st, r):
        return st.get(r)

    def join(a, b):
        out = dict(a)
        changed = False
        for r in [r for r in out if isinstance(r, tuple)]:      # pending half-writes must agree on both paths
            if b.get(r) != out[r]:
                del out[r]
                changed = True
        for r, v in b.items():
            if isinstance(r, tuple):
                continue
            if out.get(r) != v:
                nv = tjoin(out.get(r), v)
                if out.get(r) != nv:
                    out[r] = nv
                    changed = True
        return out, changed

    def flag(ins, kind, regs):
        violations[(ins.addr, kind)] = (ins.text, sorted(set(regs)))

    iters = 0
    while work:
        k = work.pop()
        inwork.discard(k)
        iters += 1
        ins = insns[k]
        st = dict(state[k])
        base = ins.op.split(".")[0]
        dests, srcs, addr, data = classify(ins)
        guard_t = st.get(ins.guard) if ins.guard else None

        def rd(r):
            # value seen on THIS instruction's guard path: a half-write under the same predicate and
            # polarity is what the instruction reads
            pend = st.get(("pend", r))
            if ins.guard and pend and pend[0] == ins.guard and pend[1] == ins.neg:
                return pend[2]
            return st.get(r)

        def taint_of(names):
            kinds = [rd(r) for r in names]
            if any(is_sec(k_) for k_ in kinds):
                return SEC
            if PTR in kinds:
                return PTR
            return None

        def lop3_imm():
            """(source register, immediate, lut) of `LOP3.LUT [P,] Rd, Ra, imm, RZ, lut, !PT`, else None"""
            if not ins.op.startswith("LOP3"):
                return None
            ops_ = [t for t in ins.operands]
            while ops_ and re.match(r"^U?P(\d+|T)$", ops_[0]):
                ops_ = ops_[1:]
            if len(ops_) >= 5 and re.match(r"^R\d+$|^RZ$", ops_[0]) and re.match(r"^R\d+(\.reuse)?$", ops_[1]) \
                    and re.match(r"^0x[0-9a-f]+$", ops_[2]) and ops_[3] == "RZ" and re.match(r"^0x[0-9a-f]+$", ops_[4]):
                return ops_[1].split(".")[0], int(ops_[2], 16), int(ops_[4], 16)
            return None

        # constant-bank reads of a secret pointer parameter
        reads_secret_param = False
        for t in ins.operands:
            m = re.search(r"c\[0x0\]\[0x([0-9a-f]+)\]", t)
            if m:
                off = int(m.group(1), 16)
                width = 8 if (".64" in ins.op or base in ("LDC", "ULDC", "LDCU") and ".64" in ins.op) else 4
                for so in secret_offsets:
                    if off < so + 8 and so < off + max(width, 4):
                        reads_secret_param = True

        # ---- sinks
        if base in BRANCH_OPS or base == "BSSY":
            bad = [r for r in ([ins.guard] if ins.guard else []) + srcs if is_sec(rd(r))]
            if bad:
                flag(ins, "secret-dependent control flow", bad)
            if base == "BRX" or base == "JMX":
                flag(ins, "indirect branch (not analysable)", srcs)
        if base in MEM_OPS or base in ("LDC", "ULDC", "LDCU"):
            bad = [r for r in addr + ([ins.guard] if ins.guard else []) if is_sec(rd(r))]
            if bad:
                flag(ins, "secret-dependent memory address / predicate", bad)
        if base == "SHFL" and len(ins.operands) >= 4:
            bad = [r for r in regs_in(ins.operands[3]) if is_sec(rd(r))]         # the source-lane operand
            if bad:
                flag(ins, "secret-dependent shuffle lane", bad)
        if base in VARLAT_OPS and is_sec(taint_of(srcs)):
            flag(ins, "variable-latency instruction on secret data", srcs)

        # ---- transfer
        new = None
        if base in ("LDG", "LD"):
            new = SEC if taint_of(addr) in (PTR, SEC) else None
        elif base == "LDL":
            new = SEC if stack_secret[0] else None
        elif base in ("LDS", "LDSM"):
            new = SEC if smem_secret[0] else None
        elif base in ("LDC", "ULDC", "LDCU"):
            new = PTR if reads_secret_param else None
        elif base in ("STS",):
            if is_sec(taint_of(data)) and not smem_secret[0]:
                smem_secret[0] = True
                # shared memory became secret: re-run everything that loads from it
                for j, other in enumerate(insns):
                    if other.op.startswith("LDS") and state[j] is not None and j not in inwork:
                        work.append(j)
                        inwork.add(j)
        elif base in ("STL",):
            if is_sec(taint_of(data)) and not stack_secret[0]:
                stack_secret[0] = True
                # stack became secret: re-run everything that loads from the stack
                for j, other in enumerate(insns):
                    if other.op.startswith("LDL") and state[j] is not None and j not in inwork:
                        work.append(j)
                        inwork.add(j)
        elif base in ("S2R", "CS2R", "S2UR"):
            new = None
        else:
            new = taint_of(srcs)
            li = lop3_imm()
            pred_new = new
            if li is not None:
                ra, imm, lut = li
                tv = rd(ra)
                bits = tv if is_sec(tv) else 0
                if lut == 0xC0:                      # Ra & imm
                    bits &= imm
                elif lut in (0xFC, 0x3C):            # Ra | imm, Ra ^ imm: the constant bits carry no secret
                    pass
                else:
                    bits = SEC if bits else 0
                new = bits if bits else (PTR if tv == PTR else None)
                pred_new = new                        # P = (result != 0)
            if reads_secret_param and not is_sec(new):
                new = PTR
        if dests:
            for d in dests:
                if d in ("RZ", "URZ", "PT", "UPT"):
                    continue
                val = new
                if d.lstrip("U").startswith("P"):
                    val = SEC if is_sec(val) else val        # a predicate is one bit
                if is_sec(guard_t):
                    li2 = lop3_imm()
                    if li2 is not None and li2[2] == 0xFC and not d.lstrip("U").startswith("P"):
                        val = (val if is_sec(val) else 0) | li2[1]   # @secret R |= imm: only the imm bits become secret
                    else:
                        val = SEC
                if ins.guard and ins.guard not in ("PT", "UPT"):
                    # a predicated write keeps the old value on the other path ... unless the
                    # complementary write (@P / @!P of the same, unmodified predicate) came just before
                    pend = st.get(("pend", d))
                    if pend and pend[0] == ins.guard and pend[1] != ins.neg:
                        val = tjoin(pend[2], val)
                        st.pop(("pend", d), None)
                    else:
                        old = st.get(d)
                        st[("pend", d)] = (ins.guard, ins.neg, val)
                        val = tjoin(old, val)
                else:
                    st.pop(("pend", d), None)
                if val is None:
                    st.pop(d, None)
                else:
                    st[d] = val
                # redefining a predicate invalidates half-writes that were guarded by it
                if d.lstrip("U").startswith("P"):
                    for key in [key for key in st if isinstance(key, tuple) and st[key][0] == d]:
                        del st[key]
        for s in succ[k]:
            if state[s] is None:
                state[s] = dict(st)
                changed = True
            else:
                state[s], changed = join(state[s], st)
            if changed and s not in inwork:
                work.append(s)
                inwork.add(s)
        if iters > 40 * n + 100000:
            raise RuntimeError("taint analysis did not converge")

    secret_loads = sum(1 for k, ins in enumerate(insns) if ins.op.startswith(("LDG", "LD.")) and state[k] is not None and
                       any(state[k].get(r) is not None for r in classify(ins)[2]))
    tainted_instrs = 0
    for k, ins in enumerate(insns):
        if state[k] is None:
            continue
        d, s, a, dd = classify(ins)
        if any(is_sec(state[k].get(r)) for r in s + dd):
            tainted_instrs += 1
    if os.environ.get("CT_EXPLAIN"):
        want = int(os.environ["CT_EXPLAIN"], 16)
        k0 = index.get(want)
        seen = set()

        def explain(k, reg, depth):
            for j in range(k - 1, -1, -1):
                d, s_, a_, dd = classify(insns[j])
                if reg in d and state[j] is not None:
                    srcs = s_ + a_ + ([insns[j].guard] if insns[j].guard else [])
                    tag = {r: state[j].get(r) for r in srcs if state[j].get(r)}
                    print("  " * depth + f"{hex(insns[j].addr)}: {insns[j].text}    tainted-in: {tag}")
                    if depth < 6:
                        for r, v in tag.items():
                            if is_sec(v) and (j, r) not in seen:
                                seen.add((j, r))
                                explain(j, r, depth + 1)
                    return
        if k0 is not None and name in os.environ.get("CT_KERNEL", name):
            d, s_, a_, dd = classify(insns[k0])
            print("EXPLAIN", name, hex(want), insns[k0].text)
            for r in s_ + a_ + ([insns[k0].guard] if insns[k0].guard else []):
                if is_sec(state[k0].get(r)):
                    explain(k0, r, 1)
    return {"kernel": name, "instructions": n, "secret_loads": secret_loads, "instructions_on_secret_data": tainted_instrs,
            "stack_holds_secrets": stack_secret[0], "shared_memory_holds_secrets": smem_secret[0], "reached": sum(1 for s in state if s is not None),
            "violations": [{"addr": hex(a), "kind": kind, "sass": text, "regs": regs} for (a, kind), (text, regs) in sorted(violations.items())]}


def run(lib, verbose=False, specs=None):
    funcs = disassemble(lib)
    results = []
    for key, spec in (specs or SECRET_PARAMS).items():
        match = [f for f in funcs if key in f]
        if not match:
            results.append({"kernel": key, "error": "kernel not found in " + lib, "violations": [{"kind": "missing"}]})
            continue
        insns = funcs[match[0]]
        base = param_base(insns)
        offs = [base + 8 * spec["params"].index(p) for p in spec["secret"]]
        r = audit(key.rstrip("E"), insns, offs, verbose)
        r["param_base"] = hex(base)
        r["secret_param_offsets"] = [hex(o) for o in offs]
        results.append(r)
    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "libeddsa_b200", "libeddsa_b200.so"))
    ap.add_argument("--json")
    ap.add_argument("-v", action="store_true")
    args = ap.parse_args()
    results = run(args.lib, args.v)
    ok = True
    for r in results:
        bad = r["violations"]
        status = "PASS" if not bad else "FAIL"
        ok &= not bad
        print(f"{status}  {r['kernel']:14s} instrs={r.get('instructions')} reached={r.get('reached')} secret-loads={r.get('secret_loads')} "
              f"instrs-on-secret-data={r.get('instructions_on_secret_data')} stack-secret={r.get('stack_holds_secrets')} violations={len(bad)}")
        for v in bad[: (1000 if args.v else 8)]:
            print("      ", v)
    if args.json:
        json.dump(results, open(args.json, "w"), indent=1)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
