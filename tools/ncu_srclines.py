#!/usr/bin/env python
"""Per-CUDA-source-line profile from an ncu report that was captured with --import-source on:
  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; ncu_srclines.py src.csv [topN]
Rows with a line number are ncu's own per-line aggregates (inlined code is attributed to the line
of the inlined function); SASS rows (no line number) are skipped."""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    cur, hdr, out = None, None, []
    for r in csv.reader(open(path, errors="replace")):
        if not r:
            continue
        if r[0] in ("File Path", "File Name"):
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) > 8 and r[0].isdigit():
            ie, s, te = (hdr.index("Instructions Executed"), hdr.index("# Samples"),
                         hdr.index("Thread Instructions Executed"))
            try:
                out.append((int(r[ie]), int(r[s]), int(r[te]), cur, int(r[0]), r[1].strip()[:80]))
            except ValueError:
                pass
    tot = sum(o[0] for o in out) or 1
    ts = sum(o[1] for o in out) or 1
    print("total warp-inst", tot, "samples", ts)
    for o in sorted(out, reverse=True)[:top]:
        print(f"{100*o[0]/tot:5.2f}% inst {100*o[1]/ts:5.2f}% smp lanes {o[2]/max(o[0],1):5.1f} {o[3]}:{o[4]}  {o[5]}")


if __name__ == "__main__":
    main()
