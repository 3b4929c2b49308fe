#!/usr/bin/env python3
"""Which source lines carry a stall reason: `ncu --page source --csv` export + the line table of the profiled library (nvdisasm -g).
usage: python tools/ncu_stall_lines.py <source.csv> <kernel substring> <lib.so> [stall column, default stall_long_sb] [top N]"""
import collections, csv, os, re, subprocess, sys, tempfile
src, KERNEL, lib = sys.argv[1], sys.argv[2], sys.argv[3]
col = sys.argv[4] if len(sys.argv) > 4 else 'stall_long_sb'
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=d, capture_output=True)
cubs = sorted(f for f in os.listdir(d) if f.endswith('.cubin'))
txt = '\n'.join(subprocess.run(['nvdisasm', '-c', '-g', os.path.join(d, c)], capture_output=True, text=True).stdout for c in cubs)
insec, lines, cur = False, {}, None
for line in txt.splitlines():
    if line.startswith('//') and '.text.' in line:
        insec = KERNEL in line
    m = re.match(r'^\s*//## File "(.*)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'^\s+/\*([0-9a-f]{4,})\*/', line)
    if m and insec:
        lines[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src)))
hdr = rows[1]
ia, il = hdr.index('Address'), hdr.index(col)
base = int(rows[2][ia], 16)
tot = sum(int(r[il] or 0) for r in rows[2:])
by = collections.Counter()
for r in rows[2:]:
    by[lines.get(int(r[ia], 16) - base)] += int(r[il] or 0)
acc = 0
print(f'# {col}: {tot} samples; share per source line of the instruction that waits (the consumer of the load)')
for k, v in by.most_common(top):
    acc += v
    if not k:
        print('%5.2f%% (cum %4.1f)  (no line)' % (100 * v / tot, 100 * acc / tot))
        continue
    f, l = k
    path = [p for p in (os.path.join(ROOT, 'nb_iot_environment_b200', 'csrc', f), os.path.join(ROOT, 'include', f)) if os.path.exists(p)]
    t = ''
    if path:
        L = open(path[0]).read().splitlines()
        t = L[l - 1].strip()[:110] if l - 1 < len(L) else ''
    print('%5.2f%% (cum %4.1f)  %s:%d  %s' % (100 * v / tot, 100 * acc / tot, f, l, t))
