"""Instruction-class histogram of the loops of one kernel's SASS (cuobjdump -sass -fun NAME obj > file).
Usage: python scripts/sass_loops.py file.sass [min_len]
Finds every backward branch, reports the instruction mix between its target and the branch."""
import collections
import re
import sys

pat = re.compile(r"^\s+/\*([0-9a-f]+)\*/\s+(@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)([.A-Z0-9_]*)\s*(.*?);")


def main():
    path = sys.argv[1]
    min_len = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    ins = []
    for ln in open(path):
        m = pat.match(ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(3), m.group(4), m.group(5)))
    addr_idx = {a: i for i, (a, *_r) in enumerate(ins)}
    loops = []
    for i, (a, op, mod, args) in enumerate(ins):
        if op == "BRA":
            t = re.search(r"0x([0-9a-f]+)", args)
            if t:
                ta = int(t.group(1), 16)
                if ta < a and ta in addr_idx and i - addr_idx[ta] >= min_len:
                    loops.append((addr_idx[ta], i))
    print(f"{len(ins)} instructions, {len(loops)} backward loops >= {min_len}")
    for lo, hi in loops:
        h = collections.Counter(op for _, op, _, _ in ins[lo:hi + 1])
        n = hi - lo + 1
        fp64 = sum(h[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
        lsu = sum(h[k] for k in ("LDS", "STS", "LDL", "STL", "LDG", "STG", "LDSM"))
        print(f"loop {ins[lo][0]:#x}..{ins[hi][0]:#x}: {n} instr, FP64-pipe {fp64}, LSU {lsu}, MUFU {h['MUFU']}, SHFL {h['SHFL']}, BAR {h['BAR']}")
        print("   " + ", ".join(f"{k} {v}" for k, v in h.most_common(24)))


if __name__ == "__main__":
    main()
