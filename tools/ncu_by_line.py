#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` export of nbiot_sim_kernel per CUDA source line
(line table from `nvdisasm -g` of the built library; the library must be the build that was profiled).
usage: python tools/ncu_by_line.py <source.csv> [lib.so] [top N]"""
import csv, os, re, subprocess, sys, tempfile
src = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'nb_iot_environment_b200', 'libnbiot_b200.so')
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
KERNEL = os.environ.get('NBIOT_PROFILE_KERNEL', 'nbk_c1s16nbiot_sim_kernelILi7')   # the 72-register build (bench workload)
d = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=d, capture_output=True)
cubs = sorted(f for f in os.listdir(d) if f.endswith('.cubin'))      # one per translation unit of the library
txt = '\n'.join(subprocess.run(['nvdisasm', '-g', '-c', os.path.join(d, cub)], capture_output=True, text=True).stdout for cub in cubs)
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
ia, ie, isamp, inoinst = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('stall_no_inst')
base = int(rows[2][ia], 16)
agg = {}
for r in rows[2:]:
    off = int(r[ia], 16) - base
    k = line_of.get(off, ('?', 0))
    a = agg.setdefault(k, [0, 0, 0, 0])
    a[0] += 1; a[1] += int(r[ie] or 0); a[2] += int(r[isamp] or 0); a[3] += int(r[inoinst] or 0)
tot_e = sum(a[1] for a in agg.values()); tot_s = sum(a[2] for a in agg.values())
srcs = {}
def text(k):
    f, l = k
    for root in ('nb_iot_environment_b200/csrc', 'include'):
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', root, f)
        if os.path.exists(p):
            if p not in srcs: srcs[p] = open(p).read().splitlines()
            return srcs[p][l - 1].strip()[:110] if 0 < l <= len(srcs[p]) else ''
    return ''
print(f'{"file:line":28s} {"sass":>5s} {"inst_exec":>11s} {"%exec":>6s} {"samples":>8s} {"%smp":>6s} {"no_inst":>8s}  source')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f'{k[0] + ":" + str(k[1]):28s} {a[0]:5d} {a[1]:11d} {100*a[1]/tot_e:6.2f} {a[2]:8d} {100*a[2]/max(tot_s,1):6.2f} {a[3]:8d}  {text(k)}')
