"""Per-loop instruction counts of one kernel, by issue pipe (development tool, no GPU needed).

usage: sass_loops.py OBJECT KERNEL_SUBSTRING [--min N] [--dump LABEL]

Loops are the ranges between a label and a later branch back to it in the nvdisasm listing.  The pipe of
an opcode follows tools/micro/pipes.cu as measured on a B200: LOP3/SHF/PRMT/IADD3/ISETP/SEL/VIMNMX/LEA share
the ALU pipe (2 warp instructions per clock and SM), IMAD/IDP/VIADD.16x2 the FMA pipe (2; IMAD.HI and
IMAD.WIDE count twice), FLO/POPC the XU pipe (0.5), shared and global memory instructions the LSU.
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ALU = {"LOP3", "SHF", "PRMT", "IADD3", "ISETP", "SEL", "VIMNMX", "VIMNMX3", "LEA", "IABS", "MOV", "BMSK", "SGXT", "PLOP3", "P2R", "R2P",
       "IADD", "ICMP", "CS2R", "VABSDIFF", "VABSDIFF4", "IMNMX", "FSEL", "FMNMX"}
FMA = {"IMAD", "IDP", "FFMA", "FMUL", "FADD", "HFMA2", "HADD2", "HMUL2", "VIADD"}
XU = {"FLO", "POPC", "MUFU", "BREV", "I2F", "F2I"}
LSU = {"LDS", "STS", "ATOMS", "LDG", "STG", "LDL", "STL", "RED", "ATOMG", "LDSM", "LD", "ST", "ATOM", "REDS"}


def classify(op):
    base = op.split(".")[0]
    if base == "IMAD" and (".HI" in op or ".WIDE" in op):
        return "FMA2"
    if base == "VIADD" and "16x2" not in op:
        return "ALU?"  # plain 32-bit VIADD: pipe not measured
    if base in ALU:
        return "ALU"
    if base in FMA:
        return "FMA"
    if base in XU:
        return "XU"
    if base in LSU:
        return "LSU"
    return "other"


def listing(obj, want):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout.split("\n")
        starts = [i for i, l in enumerate(dis) if l.lstrip().startswith(".type") and "@function" in l]
        for k, i in enumerate(starts):
            if want in dis[i]:
                return dis[i:starts[k + 1] if k + 1 < len(starts) else len(dis)]
    raise SystemExit("kernel not found")


def main():
    obj, want = sys.argv[1], sys.argv[2]
    min_n = int(sys.argv[sys.argv.index("--min") + 1]) if "--min" in sys.argv else 20
    dump = sys.argv[sys.argv.index("--dump") + 1] if "--dump" in sys.argv else None
    body = listing(obj, want)
    inst = []      # (op string, full text, source line)
    labels = {}    # label -> index of the next instruction
    line = ""
    for l in body:
        m = re.search(r'//## File ".*/([^/"]+)", line (\d+)', l)
        if m:
            line = "%s:%s" % (m.group(1), m.group(2))
            continue
        m = re.match(r"\s*(\.L_x_\d+):", l)
        if m:
            labels[m.group(1)] = len(inst)
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            txt = m.group(2)
            parts = txt.split()
            op = parts[1] if parts[0].startswith("@") else parts[0]
            inst.append((op, txt, line))
    total = collections.Counter(classify(op) for op, _, _ in inst)
    print("kernel: %d instructions  %s" % (len(inst), dict(total)))
    loops = []
    for i, (op, txt, _) in enumerate(inst):
        if op.startswith("BRA"):
            m = re.search(r"(\.L_x_\d+)", txt)
            if m and m.group(1) in labels and labels[m.group(1)] <= i:
                loops.append((labels[m.group(1)], i, m.group(1)))
    for a, b, lab in sorted(loops):
        n = b - a + 1
        if n < min_n:
            continue
        c = collections.Counter(classify(op) for op, _, _ in inst[a:b + 1])
        ops = collections.Counter(op.split(".")[0] for op, _, _ in inst[a:b + 1])
        lines = [l for _, _, l in inst[a:b + 1] if l]
        print("loop %-10s %5d inst  ALU %4d  ALU? %3d  FMA %4d  FMA2 %3d  XU %3d  LSU %3d  other %3d   %s .. %s" % (
            lab, n, c["ALU"], c["ALU?"], c["FMA"], c["FMA2"], c["XU"], c["LSU"], c["other"], lines[0] if lines else "", lines[-1] if lines else ""))
        print("      " + " ".join("%s:%d" % kv for kv in ops.most_common(14)))
        if dump == lab:
            for op, txt, l in inst[a:b + 1]:
                print("   %-26s %s" % (l, txt))


if __name__ == "__main__":
    main()
