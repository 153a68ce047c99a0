#!/usr/bin/env python
"""SASS-level view of an ncu report: groups consecutive instructions with similar execution counts into
regions and prints each region's share of issued instructions, average active threads and stall samples.
    python tests/tools/ncu_regions.py report.ncu-rep [kernel-index] [top]"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else -1
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    kern, cur = [], None
    for r in csv.reader(raw.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kern.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    seen = set()
    uniq = []
    for k in kern:  # ncu prints each kernel twice (SASS and SASS+source views)
        key = (k["name"], len(k["rows"]))
        if key not in seen:
            seen.add(key)
            uniq.append(k)
    for idx, k in enumerate(uniq):
        if which >= 0 and idx != which:
            continue
        h = k["hdr"]
        ie, te, ss, src = (h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples"),
                           h.index("Source"))
        R = [(int(r[ie]), int(r[te]), int(r[ss]), r[src].strip()) for r in k["rows"]]
        tot, tt, ts = sum(x[0] for x in R), sum(x[1] for x in R), max(1, sum(x[2] for x in R))
        print(f"== {k['name'][:70]}  SASS={len(R)} warp-instr={tot} avg-threads={tt / tot:.1f}")
        byop = collections.Counter()
        for e, t, s, txt in R:
            w = txt.split()
            op = (w[1] if w[0].startswith("@") else w[0]).split(".")[0]
            byop[op] += e
        print("   opcode mix %:", [(o, round(100 * c / tot, 1)) for o, c in byop.most_common(12)])
        regions = []
        for i, (e, t, s, txt) in enumerate(R):
            if regions and abs(e - regions[-1]["e0"]) <= 0.15 * max(e, regions[-1]["e0"], 1):
                g = regions[-1]
                g["n"] += 1; g["e"] += e; g["t"] += t; g["s"] += s; g["last"] = i
            else:
                regions.append(dict(first=i, last=i, n=1, e=e, t=t, s=s, e0=e, txt=txt))
        for g in sorted(regions, key=lambda g: -g["e"])[:top]:
            print(f"   SASS[{g['first']:4d}-{g['last']:4d}] n={g['n']:3d} exec/instr={g['e'] / g['n'] / 1e6:8.2f}M "
                  f"share={100 * g['e'] / tot:5.1f}% threads={g['t'] / max(g['e'], 1):5.1f} stall={100 * g['s'] / ts:5.1f}%  {g['txt'][:44]}")


if __name__ == "__main__":
    main()
