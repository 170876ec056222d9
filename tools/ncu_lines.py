#!/usr/bin/env python
"""Join an ncu SASS source page (ncu -i X --page source --csv) with nvdisasm -g line info:
aggregates stall samples and executed instructions per CUDA source line and per inlined function.
usage: ncu_lines.py <src.csv> <nvdisasm -g -c output> <mangled kernel substring> [topN]"""
import collections
import csv
import re
import sys


def main():
    src_csv, dis, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    # --- parse disassembly: sequence of (file,line,inlined_chain) per instruction of the kernel
    lines = open(dis, errors='replace').read().splitlines()
    start = None
    for i, l in enumerate(lines):
        if l.startswith('.text.') and kname in l:
            start = i
            break
    if start is None:
        # section header form: //--------------------- .text.<name>
        for i, l in enumerate(lines):
            if '.text.' in l and kname in l and l.lstrip().startswith('//-----'):
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
    # --- parse ncu csv
    rows = list(csv.reader(open(src_csv)))
    h = rows[1]
    si = h.index('# Samples')
    ei = h.index('Instructions Executed')
    ti = h.index('Thread Instructions Executed')
    data = rows[2:]
    n = min(len(data), len(seq))
    print(f'instructions: ncu {len(data)} nvdisasm {len(seq)}')
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    for k in range(n):
        key = seq[k]
        a = agg[key]
        a[0] += int(data[k][si] or 0)
        a[1] += int(data[k][ei] or 0)
        a[2] += int(data[k][ti] or 0)
        a[3] += 1
    tot_s = sum(a[0] for a in agg.values()) or 1
    tot_e = sum(a[1] for a in agg.values()) or 1
    print(f'total samples {tot_s} warp-inst {tot_e}')
    print(f'{"file:line":32s} {"samples%":>8s} {"inst%":>7s} {"thr/inst":>8s} {"#sass":>6s}')
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f'{key[0]+":"+str(key[1]):32s} {100*a[0]/tot_s:8.2f} {100*a[1]/tot_e:7.2f} {a[2]/max(a[1],1):8.1f} {a[3]:6d}')


if __name__ == '__main__':
    main()
