"""Per-kernel SASS statistics of libhimgcu.so (static instruction counts by opcode and pipe, registers,
spills, the mnemonics that prove TMA bulk copies / mbarriers / clusters / dp4a / DPX).

    python tools/sass_stats.py                 # summary table of every kernel
    python tools/sass_stats.py k_forward3      # opcode histogram of the kernels whose name matches
    python tools/sass_stats.py k_forward3 --dump > fwd.sass
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "himg_b200", "_lib", "libhimgcu.so")

FMA_PIPE = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "HADD2", "HMUL2", "IDP", "IMUL")
ALU_PIPE = ("IADD3", "IADD", "LOP3", "SHF", "PRMT", "VIADD", "VIADDMNMX", "VIMNMX", "LEA", "ISETP", "SEL", "MOV", "IABS",
            "FMNMX", "SGXT", "BMSK", "FLO", "POPC", "VABSDIFF", "LOP", "FSETP", "PLOP3", "P2R", "R2P", "CS2R", "I2I",
            "FSEL", "BREV", "VHMNMX", "IMNMX")
LSU = ("LDS", "STS", "LDG", "STG", "LDL", "STL", "LDSM", "ATOMS", "ATOMG", "RED", "ATOM", "LDGSTS", "LD", "ST", "LDC",
       "LDGDEPBAR", "DEPBAR", "SHFL", "MATCH", "VOTE", "REDUX", "UBLKCP", "SYNCS", "MEMBAR", "ERRBAR", "CCTL")
MARKS = ("UBLKCP", "SYNCS", "UCGABAR", "LDGSTS", "IDP.4A", "VIADDMNMX", "UTMALDG", "STL", "LDL")


def functions():
    global LIB
    for a in sys.argv[1:]:
        if a.startswith("--file="):
            LIB = a[len("--file="):]
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
        if m and cur:
            usage[cur] = tuple(int(x) for x in m.groups())
    out = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            out[cur].append(m.group(2).strip())
    return out, usage


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
    except OSError:
        return name


def opcode(ins):
    toks = ins.split()
    if toks and toks[0].startswith("@"):
        toks = toks[1:]
    return toks[0] if toks else ""


def pipe(op):
    base = op.split(".")[0]
    if base.startswith("U") and base not in ("UBLKCP",):
        return "uniform"
    if base in FMA_PIPE:
        return "fma"
    if base in ALU_PIPE:
        return "alu"
    if base in LSU:
        return "lsu"
    return "other"


def main():
    pat = None
    dump = "--dump" in sys.argv
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if args:
        pat = args[0]
    funcs, usage = functions()
    if pat is None:
        print("%-58s %6s %5s %6s %6s | %s" % ("kernel", "instr", "regs", "smem", "local", " ".join(MARKS)))
        for name, body in funcs.items():
            reg, sh, loc = usage.get(name, (0, 0, 0))
            cnt = [sum(1 for i in body if m in i) for m in MARKS]
            print("%-58s %6d %5d %6d %6d | %s" % (demangle(name)[-58:], len(body), reg, sh, loc,
                                                  " ".join("%*d" % (len(m), c) for m, c in zip(MARKS, cnt))))
        return
    for name, body in funcs.items():
        dn = demangle(name)
        if pat not in dn:
            continue
        if dump:
            print("// " + dn)
            print("\n".join(body))
            continue
        ops = collections.Counter(opcode(i).split(".")[0] for i in body)
        pipes = collections.Counter(pipe(opcode(i)) for i in body)
        print("== %s: %d instructions, usage %s" % (dn, len(body), usage.get(name)))
        print("   pipes: " + ", ".join("%s %d" % kv for kv in pipes.most_common()))
        print("   " + ", ".join("%s %d" % kv for kv in ops.most_common(40)))


if __name__ == "__main__":
    main()
