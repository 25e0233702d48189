"""Warp-stall samples and executed instructions of an `ncu --set full --import-source on` report, attributed to the innermost
SASS loop each instruction belongs to (loops = backward branches of the source page).  For every hot loop: share of the samples,
executed warp-instructions, FP64-pipe instructions among them, dispatch slots (2 per FP64 instruction + 1 per other, see
scripts/micro/plain_round.cu) and the leading stall reasons.

    python scripts/ncu_loops.py gpurun_out/<report>.ncu-rep [kernel-substring] > profiles/<tag>_mh_kernel_loops.txt"""
import collections
import csv
import re
import subprocess
import sys

FP64 = ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
want = sys.argv[2] if len(sys.argv) > 2 else "mh_kernel"
rows = list(csv.reader(out.splitlines()))
# the page is a sequence of blocks: ["Kernel Name", name], header, instruction rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
for blk in blocks:
    if want not in blk["name"]:
        continue
    hdr, data = blk["hdr"], blk["rows"]
    ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = int(data[0][ia], 16)
    offs = [int(r[ia], 16) - base for r in data]
    loops = []
    for k, r in enumerate(data):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:\w+,\s*)?0x([0-9a-f]+)", r[isrc])
        if m and int(m.group(1), 16) - base < offs[k]:
            loops.append((int(m.group(1), 16) - base, offs[k]))

    def smallest(off):
        best = None
        for a, b in loops:
            if a <= off <= b and (best is None or b - a < best[1] - best[0]):
                best = (a, b)
        return best
    agg, tot = collections.defaultdict(collections.Counter), 0
    for k, r in enumerate(data):
        c = agg[smallest(offs[k])]
        s = int(r[isamp] or 0)
        tot += s
        c["samples"] += s
        c["inst"] += int(r[iex] or 0)
        c["n"] += 1
        for i in stalls:
            c[hdr[i]] += int(r[i] or 0)
        op = [o for o in r[isrc].split() if not o.startswith("@")]
        if op and op[0].startswith(FP64):
            c["fp64"] += int(r[iex] or 0)
        if op and op[0].startswith(("LDS", "STS")):
            c["lds"] += int(r[iex] or 0)
    ti, tf = sum(c["inst"] for c in agg.values()), sum(c["fp64"] for c in agg.values())
    print("%s\n  %d samples, %d warp-instructions executed, %d of them FP64-pipe (%.1f %%), dispatch slots %d" % (blk["name"], tot, ti, tf, 100.0 * tf / ti, ti + tf))
    for l, c in sorted(agg.items(), key=lambda x: -x[1]["samples"])[:16]:
        print("  %-16s static %4d  samples %5.1f %%  executed %10d  fp64 %10d  lds %9d  slots %5.1f %% | %s" % (
            "[%05x..%05x]" % l if l else "(straight-line)", c["n"], 100.0 * c["samples"] / tot, c["inst"], c["fp64"], c["lds"],
            100.0 * (c["inst"] + c["fp64"]) / (ti + tf), " ".join("%s %d" % (k[6:], v) for k, v in c.most_common(12) if k.startswith("stall_") and v)[:150]))
