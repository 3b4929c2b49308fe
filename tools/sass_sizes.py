#!/usr/bin/env python3
"""Per-function SASS size of the built CUDA library (instruction-cache footprint check).
usage: python tools/sass_sizes.py [path/to/lib.so] [kernel substring]"""
import os, re, subprocess, sys, tempfile
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'nb_iot_environment_b200', 'libnbiot_b200.so')
want = sys.argv[2] if len(sys.argv) > 2 else 'nbk_c1s16nbiot_sim_kernelILi7'
d = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=d, capture_output=True)
cubs = sorted(f for f in os.listdir(d) if f.endswith('.cubin'))      # one per translation unit of the library
txt = '\n'.join(subprocess.run(['nvdisasm', '-c', os.path.join(d, cub)], capture_output=True, text=True).stdout for cub in cubs)
cur, sizes, order = None, {}, []
for line in txt.splitlines():
    m = re.match(r'^(\$?[\w\$\.]+):\s*$', line.strip())
    if m and not m.group(1).startswith('.L'):
        cur = m.group(1)
        if cur not in sizes:
            sizes[cur] = 0; order.append(cur)
        continue
    if cur and re.match(r'^\s+/\*[0-9a-f]{4,}\*/', line):
        sizes[cur] += 16
def pretty(n):
    n = n.split('$')[-1] if n.startswith('$_Z') else n        # $<kernel>$<function>
    try:
        return subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()[:90]
    except Exception:
        return n
tot = 0
rows = []
for n in order:
    if want in n or (n.startswith('$__internal') ):
        rows.append((sizes[n], n)); 
for s, n in sorted(rows, reverse=True):
    if want in n: tot += s
    print(f'{s:8d}  {pretty(n)}')
print('total bytes in functions of', want, ':', tot)
