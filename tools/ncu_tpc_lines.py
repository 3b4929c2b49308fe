#!/usr/bin/env python3
"""Per-source-line view of an `ncu --page source --csv` export of the thread-per-cell kernel (line table from `nvdisasm -g` of the
profiled library): warp instructions, active lanes per instruction.
usage: python tools/ncu_tpc_lines.py <source.csv> <kernel substring> <lib.so> [top N]"""
import csv, os, re, subprocess, sys, tempfile
src, KERNEL, lib = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
d = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=d, capture_output=True)
txt = '\n'.join(subprocess.run(['nvdisasm', '-g', '-c', os.path.join(d, c)], capture_output=True, text=True).stdout for c in sorted(os.listdir(d)) if c.endswith('.cubin'))
insec, cur, line_of = False, None, {}
for line in txt.splitlines():
    if line.startswith('//') and '.text.' in line:
        insec = KERNEL in line
        continue
    if not insec:
        continue
    m = re.match(r'^\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'^\s+/\*([0-9a-f]{4,})\*/', line)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src)))
hdr = rows[1]
ia, ie, it = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
base = int(rows[2][ia], 16)
agg = {}
for r in rows[2:]:
    k = line_of.get(int(r[ia], 16) - base)
    a = agg.setdefault(k, [0, 0])
    a[0] += int(r[ie] or 0); a[1] += int(r[it] or 0)
te = sum(a[0] for a in agg.values())
cache = {}
def text(k):
    if not k:
        return ''
    f = None
    for root in ('nb_iot_environment_b200/csrc', 'include'):
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', root, k[0])
        if os.path.exists(p):
            f = p
    if not f:
        return ''
    if f not in cache:
        cache[f] = open(f).read().splitlines()
    return cache[f][k[1] - 1].strip()[:110] if k[1] - 1 < len(cache[f]) else ''
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{100 * a[0] / te:5.1f}% lanes {a[1] / max(a[0], 1):5.2f}  {k[0] if k else "?"}:{k[1] if k else 0:<5d} {text(k)}')
