"""Opcode histogram of the compiled kernels (cuobjdump -sass on phylowgs_b200/libpwgs_b200.so), whole kernel and per innermost
loop, so that the FP64-instruction constants bench.py's roofline uses can be read off the cubin (VERDICT r1 'weak' 2).

    python scripts/sass_histogram.py [kernel-name-substring] > profiles/<round>_mh_kernel_sass_histogram.txt

A loop is a backward branch [target, branch]; counts are exclusive of the loops nested inside a loop.  For every loop:
the number of FP64-pipe instructions (DFMA / DADD / DMUL / DSETP / DMNMX; they issue on the FP64 pipe), of the conversions and
reciprocal seeds that go to the XU pipe (I2F.F64, F2F, MUFU.*), of shared / global loads, and of everything else."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("PWGS_LIB") or os.path.join(ROOT, "phylowgs_b200", "libpwgs_b200.so")
FP64 = ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX")


def functions(text):
    out, name, body = [], None, []
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                out.append((name, body))
            name, body = m.group(1), []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            body.append((int(m.group(1), 16), m.group(2).strip()))
    if name:
        out.append((name, body))
    return out


def opcode(ins):
    ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
    return ins.split()[0]


def klass(op):
    base = op.split(".")[0]
    if base in FP64:
        return "fp64"
    if base == "MUFU" or op.startswith("I2F") or op.startswith("F2F") or op.startswith("F2I") or base == "I2F":
        return "xu"
    if base in ("LDL", "STL"):
        return "spill"
    if base in ("LDS", "LDSM"):
        return "lds"
    if base in ("LDG", "LD", "LDC", "LDCU", "ULDC"):
        return "ldg/ldc"
    if base in ("STS", "STG", "ST", "ATOMG", "RED", "ATOMS"):
        return "store/atomic"
    if base in ("BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "WARPSYNC", "BAR", "NANOSLEEP", "YIELD"):
        return "control"
    return "int/other"


def loops(body):
    addr_index = {a: i for i, (a, _) in enumerate(body)}
    found = []
    for i, (a, ins) in enumerate(body):
        if opcode(ins).startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", ins)
            if m:
                t = int(m.group(1), 16)
                if t <= a and t in addr_index:
                    found.append((addr_index[t], i))
    return sorted(set(found))


def main():
    pat = sys.argv[1] if len(sys.argv) > 1 else "mh_kernelILb1"
    text = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    for name, body in functions(text):
        if pat not in name:
            continue
        print("== %s: %d instructions" % (name, len(body)))
        hist = collections.Counter(opcode(i).split(".")[0] + ("." + opcode(i).split(".")[1] if opcode(i).startswith(("MUFU", "I2F", "F2F")) and "." in opcode(i) else "")
                                   for _, i in body)
        kl = collections.Counter(klass(opcode(i)) for _, i in body)
        print("   classes:", dict(kl))
        print("   top opcodes:", ", ".join("%s %d" % kv for kv in hist.most_common(24)))
        print("   loops (address range, instructions exclusive of nested loops; fp64 / xu / lds / ldg+ldc / int+other / control) with their FP64 mix:")
        all_loops = loops(body)
        for lo, hi in all_loops:
            nested = [o for o in all_loops if o != (lo, hi) and lo <= o[0] and o[1] <= hi]
            skip = set()
            for o in nested:
                skip.update(range(o[0], o[1] + 1))
            seg = [body[i] for i in range(lo, hi + 1) if i not in skip]      # exclusive of the loops nested inside
            if len(seg) < 12:
                continue
            c = collections.Counter(klass(opcode(i)) for _, i in seg)
            f = collections.Counter(opcode(i).split(".")[0] for _, i in seg if klass(opcode(i)) == "fp64")
            x = collections.Counter(opcode(i) for _, i in seg if klass(opcode(i)) == "xu")
            print("   [%05x..%05x]%s %4d instr: fp64 %4d  xu %3d  lds %3d  ldg/ldc %3d  int/other %4d  control %2d  spill %2d | %s | %s"
                  % (body[lo][0], body[hi][0], " +%d nested" % len(nested) if nested else "", len(seg), c["fp64"], c["xu"], c["lds"], c["ldg/ldc"], c["int/other"], c["control"] + c["store/atomic"], c["spill"],
                     " ".join("%s %d" % kv for kv in sorted(f.items())), " ".join("%s %d" % kv for kv in sorted(x.items()))))


if __name__ == "__main__":
    main()
