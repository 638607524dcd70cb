#!/usr/bin/env python
"""Development helper: attribute the warp-stall samples of an ncu report (SASS source page) to CUDA source lines.

    ncu -i rep.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all lib.so ; nvdisasm -g x.cubin > all.sass
    python scripts/ncu_lines.py src.csv all.sass <mangled kernel name> [top]

The SASS page lists the kernel's instructions in address order; nvdisasm -g lists the same instructions with
`//## File ..., line N` markers.  Both are matched by instruction index."""
import csv
import re
import sys
from collections import defaultdict


def main():
    src_csv, sass, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(src_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ci = hdr.index("# Samples")
    ii = hdr.index("Instructions Executed")
    samples = [(int(r[ci] or 0), int(r[ii] or 0), r[1].strip()) for r in rows[hi + 1:] if len(r) > ci]
    lines = open(sass).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith(".text." + kern + ":"))
    cur = ("?", 0)
    inl = ""
    tags = []
    for l in lines[start + 1:]:
        if l.startswith("//-----") or l.startswith(".text."):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            inl = m.group(3)
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
            tags.append((cur, inl))
    n = min(len(tags), len(samples))
    print(f"instructions: sass {len(tags)}, ncu {len(samples)}")
    by_line = defaultdict(lambda: [0, 0])
    tot = 0
    for k in range(n):
        by_line[tags[k][0]][0] += samples[k][0]
        by_line[tags[k][0]][1] += samples[k][1]
        tot += samples[k][0]
    print(f"total samples {tot}")
    for (f, ln), (s, ex) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100.0 * s / max(tot, 1):6.2f} %  {s:8d} samples  {ex:12d} inst   {f}:{ln}")


if __name__ == "__main__":
    main()
