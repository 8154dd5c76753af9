"""Executed-instruction mix of one kernel from an .ncu-rep captured with --import-source on:
  python scripts/ncu_opmix.py rep.ncu-rep [top]
Sums "Instructions Executed" (warp level) by opcode over the SASS source page, and the stall samples by opcode."""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    i_src, i_ex, i_smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ex, smp = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) <= i_ex:
            continue
        m = re.match(r"\s*(@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", r[i_src])
        if not m:
            continue
        op = m.group(2)
        if op in ("IMAD", "MOV") and ("MOV" in (m.group(3) or "") or op == "MOV"):
            op = "MOVE(IMAD.MOV/MOV)"
        elif op == "IMAD" and m.group(3) in (".U32",) and "RZ, RZ" in r[i_src]:
            op = "MOVE(IMAD.MOV/MOV)"
        ex[op] += int(r[i_ex] or 0)
        smp[op] += int(r[i_smp] or 0)
    tot, tots = sum(ex.values()), sum(smp.values())
    fp64 = sum(ex[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
    print(f"{rows[0][1][:90]}\nwarp instructions executed {tot}, FP64-pipe {fp64} ({100.0 * fp64 / tot:.1f} %), stall samples {tots}")
    for op, n in ex.most_common(top):
        print(f"  {op:22s} {n:12d} {100.0 * n / tot:6.2f} %   samples {100.0 * smp[op] / max(tots, 1):6.2f} %")


if __name__ == "__main__":
    main()
