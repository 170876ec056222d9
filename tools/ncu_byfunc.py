#!/usr/bin/env python
"""Per-FUNCTION instruction / stall-sample shares from `ncu -i X --page source --csv --print-source
cuda,sass`: every per-line aggregate row is attributed to the function whose definition encloses that
line in the .cu/.cuh file (inlined code is attributed to the inlined function's own lines).
usage: ncu_byfunc.py <src.csv> <csrc dir>"""
import collections
import csv
import os
import re
import sys

PAT = re.compile(r'^(?:template\s*<[^>]*>\s*)?(?:B2_HD|B2_LEAF|B2_DEV|static|__global__|__device__|inline|'
                 r'cudaError_t|int|void|size_t|bool|float|double)\b.*?\b([A-Za-z_]\w*)\s*\(')


def func_ranges(path):
    out = []
    try:
        lines = open(path, errors="replace").read().splitlines()
    except OSError:
        return out
    pend = False
    for i, l in enumerate(lines, 1):
        if l.startswith("template"):
            pend = True
        m = PAT.match(l)
        if m and not l.startswith(" "):
            out.append((i, m.group(1)))
        elif pend and not l.startswith(" ") and "(" in l:
            m2 = re.search(r"\b([A-Za-z_]\w*)\s*\(", l)
            if m2:
                out.append((i, m2.group(1)))
        if not l.startswith("template"):
            pend = False
    return out


def main():
    path, srcdir = sys.argv[1], sys.argv[2]
    ranges = {}
    cur, hdr = None, None
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for r in csv.reader(open(path, errors="replace")):
        if not r:
            continue
        if r[0] in ("File Path", "File Name"):
            cur = os.path.basename(r[1])
            if cur not in ranges:
                ranges[cur] = func_ranges(os.path.join(srcdir, cur))
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) > 8 and r[0].isdigit():
            ie, s, te = (hdr.index("Instructions Executed"), hdr.index("# Samples"),
                         hdr.index("Thread Instructions Executed"))
            try:
                n, smp, thr = int(r[ie]), int(r[s]), int(r[te])
            except ValueError:
                continue
            ln = int(r[0])
            fn = cur
            for start, name in ranges.get(cur, []):
                if start <= ln:
                    fn = name
            a = agg[fn]
            a[0] += n
            a[1] += smp
            a[2] += thr
    tot = sum(a[0] for a in agg.values()) or 1
    ts = sum(a[1] for a in agg.values()) or 1
    print(f"total warp-inst {tot}  samples {ts}")
    print(f'{"function":30s} {"inst%":>7s} {"samples%":>8s} {"lanes":>6s}')
    for fn, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if a[0] * 1000 < tot:
            continue
        print(f"{fn:30s} {100*a[0]/tot:7.2f} {100*a[1]/ts:8.2f} {a[2]/max(a[0],1):6.1f}")


if __name__ == "__main__":
    main()
