#!/usr/bin/env python
"""Aggregate an ncu SASS source page per enclosing source FUNCTION (by line ranges parsed from the
.cuh/.cu files).  usage: ncu_funcs.py <src.csv> <nvdisasm -g -c out> <kernel substring> <src dir>"""
import collections
import csv
import os
import re
import sys


def func_ranges(path):
    out = []
    pat = re.compile(r'^(?:template\s*<[^>]*>\s*)?(?:B2_HD|B2_LEAF|static|__global__|__device__|inline|cudaError_t|int|void|size_t|bool|float|double)\b.*?\b([A-Za-z_]\w*)\s*\(')
    lines = open(path, errors='replace').read().splitlines()
    pending_template = False
    for i, l in enumerate(lines, 1):
        if l.startswith('template'):
            pending_template = True
            m = pat.match(l)
            if not m:
                continue
        m = pat.match(l)
        if m and not l.startswith(' '):
            out.append((i, m.group(1)))
        elif pending_template and not l.startswith(' ') and '(' in l:
            m2 = re.search(r'\b([A-Za-z_]\w*)\s*\(', l)
            if m2:
                out.append((i, m2.group(1)))
        if not l.startswith('template'):
            pending_template = False
    return out


def main():
    src_csv, dis, kname, srcdir = sys.argv[1:5]
    ranges = {}
    for f in os.listdir(srcdir):
        if f.endswith(('.cuh', '.cu', '.h')):
            ranges[f] = func_ranges(os.path.join(srcdir, f))

    def func_of(f, line):
        r = ranges.get(f)
        if not r:
            return f
        name = f + ':<top>'
        for start, n in r:
            if start <= line:
                name = n
            else:
                break
        return name

    lines = open(dis, errors='replace').read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('.text.') and kname in l][0]
    cur = ('?', 0)
    seq = []
    for l in lines[start + 1:]:
        if l.lstrip().startswith('//-----') and '.text.' in l:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,6}\*/', l):
            seq.append(cur)
    rows = list(csv.reader(open(src_csv)))
    h = rows[1]
    si, ei, ti = h.index('# Samples'), h.index('Instructions Executed'), h.index('Thread Instructions Executed')
    data = rows[2:]
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    for k in range(min(len(seq), len(data))):
        a = agg[func_of(*seq[k])]
        a[0] += int(data[k][si] or 0)
        a[1] += int(data[k][ei] or 0)
        a[2] += int(data[k][ti] or 0)
        a[3] += 1
    ts = sum(a[0] for a in agg.values()) or 1
    te = sum(a[1] for a in agg.values()) or 1
    print(f'{"function":28s} {"samples%":>8s} {"inst%":>7s} {"thr/inst":>8s} {"#sass":>6s}')
    for b, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f'{b:28s} {100*a[0]/ts:8.2f} {100*a[1]/te:7.2f} {a[2]/max(a[1],1):8.1f} {a[3]:6d}')


if __name__ == '__main__':
    main()
