#!/usr/bin/env python
"""Per-kernel SASS census of libb200knn.so (cuobjdump -sass): how many tcgen05 MMA / TMA / tensor-memory instructions
every kernel holds -- the evidence that the hot path is hand-written tcgen05 / TMEM / TMA code.

    python tools/sass_census.py > profiles/r02_sass_census.txt
"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "image-retrieval---thesis-2026_b200", "libb200knn.so")
MNEMONICS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "UTMALDG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "UTCATOMSWS", "SYNCS",
             "FFMA2", "FFMA", "DADD", "POPC", "ATOMS", "SHFL", "REDUX", "BAR.SYNC", "LDS", "STS", "LDG", "STG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = cur.replace("(anonymous namespace)::", "")
            cur = re.sub(r"_GLOBAL__N__\w+::", "", cur).replace("knn::", "").replace("void ", "")
            depth, cut = 0, len(cur)
            for pos in range(len(cur) - 1, -1, -1):      # drop the parameter list (the last top-level parenthesis)
                if cur[pos] == ")":
                    depth += 1
                elif cur[pos] == "(":
                    depth -= 1
                    if depth == 0:
                        cut = pos
                        break
            cur = cur[:cut]
            while cur in kernels:
                cur += "'"
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["_total"] += 1
        for mn in MNEMONICS:
            if op == mn or op.startswith(mn + "."):
                kernels[cur][mn] += 1
                break
    print(f"# SASS census of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a); counts are static instructions per kernel")
    cols = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "UTMALDG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "FFMA2", "FFMA", "POPC", "ATOMS", "DADD"]
    print(f"{'kernel':78s} {'instr':>7s} " + " ".join(f"{c:>12s}" for c in cols))
    tot = collections.Counter()
    for name, c in kernels.items():
        print(f"{name[:78]:78s} {c['_total']:7d} " + " ".join(f"{c[x]:12d}" for x in cols))
        tot.update(c)
    print(f"{'TOTAL':78s} {tot['_total']:7d} " + " ".join(f"{tot[x]:12d}" for x in cols))


if __name__ == "__main__":
    main()
