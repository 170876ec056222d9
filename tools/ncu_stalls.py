#!/usr/bin/env python
"""Per source line (and per inlined function): which stall reasons the samples of an ncu SASS source
page fall under.  usage: ncu_stalls.py <src.csv> <nvdisasm -g -c output> <mangled kernel substring>
[topN] [stall column, default all]"""
import collections
import csv
import re
import sys


def main():
    src_csv, dis, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    only = sys.argv[5] if len(sys.argv) > 5 else None
    lines = open(dis, errors='replace').read().splitlines()
    start = None
    for i, l in enumerate(lines):
        if '.text.' in l and kname in l:
            start = i
            break
    assert start is not None, 'kernel not found'
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
    cols = [i for i, c in enumerate(h) if c.startswith('stall_')]
    # the sampled (all) columns come first, the not-issued ones repeat the names: keep the first set
    seen, keep = set(), []
    for i in cols:
        if h[i] not in seen:
            seen.add(h[i])
            keep.append(i)
    data = rows[2:]
    n = min(len(data), len(seq))
    tot = collections.Counter()
    per = collections.defaultdict(collections.Counter)
    for k in range(n):
        for i in keep:
            v = int(data[k][i] or 0)
            if v:
                tot[h[i]] += v
                per[seq[k]][h[i]] += v
    all_s = sum(tot.values()) or 1
    print('stall totals:', ', '.join(f'{k[6:]} {100 * v / all_s:.1f}%' for k, v in tot.most_common(10)))
    key = (lambda c: c[only]) if only else (lambda c: sum(c.values()))
    print(f'{"file:line":30s} {"samples%":>8s}  top reasons')
    for loc, c in sorted(per.items(), key=lambda kv: -key(kv[1]))[:top]:
        s = key(c)
        print(f'{loc[0] + ":" + str(loc[1]):30s} {100 * s / all_s:8.2f}  ' +
              ', '.join(f'{k[6:]} {100 * v / all_s:.2f}' for k, v in c.most_common(4)))


if __name__ == '__main__':
    main()
