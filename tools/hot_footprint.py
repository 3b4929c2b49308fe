#!/usr/bin/env python3
"""Hot code footprint of a kernel from an `ncu --page source --csv` export: how many bytes of SASS (in 128-byte i-cache lines)
carry which share of the executed warp instructions.   usage: python tools/hot_footprint.py <source.csv>"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, ie = hdr.index('Address'), hdr.index('Instructions Executed')
base = int(rows[2][ia], 16)
lines = {}
tot = 0
for r in rows[2:]:
    off, e = int(r[ia], 16) - base, int(r[ie] or 0)
    lines[off >> 7] = lines.get(off >> 7, 0) + e
    tot += e
n_all = (int(rows[-1][ia], 16) - base + 16 + 127) >> 7
srt = sorted(lines.values(), reverse=True)
print(f'kernel text {n_all * 128 / 1024:.1f} KB in {n_all} lines of 128 B; executed warp instructions {tot:.4g}; lines ever executed {sum(1 for v in srt if v)} = {sum(1 for v in srt if v) * 128 / 1024:.1f} KB')
acc, k = 0, 0
marks = [0.5, 0.75, 0.9, 0.95, 0.98, 0.99, 0.995, 0.999]
for i, v in enumerate(srt):
    acc += v
    while k < len(marks) and acc >= marks[k] * tot:
        print(f'  {100 * marks[k]:5.1f} % of the executed instructions lie in the hottest {i + 1:4d} lines = {(i + 1) * 128 / 1024:6.1f} KB')
        k += 1
# an LRU model is out of reach without a trace; the stack distance floor: a cache of S lines holding the hottest S lines misses at least the rest
for kb in (6, 16, 32, 48, 64, 96, 128):
    s = kb * 1024 // 128
    print(f'  a {kb:3d} KB cache pinned to the hottest lines would still miss {100 * (1 - sum(srt[:s]) / tot):5.2f} % of the instruction fetches')
