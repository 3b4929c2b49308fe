#!/usr/bin/env python3
"""Per-function view of an `ncu --page source --csv` export of the thread-per-cell kernel: warp instructions executed, thread
instructions executed (their ratio = active lanes per instruction), stall samples.
usage: python tools/ncu_tpc_functions.py <source.csv> <kernel substring, e.g. nbk_tpc_c117nbiot_simt_kernel> [lib.so]"""
import bisect, csv, os, re, subprocess, sys, tempfile
src, KERNEL = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'nb_iot_environment_b200', 'libnbiot_b200.so')
d = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=d, capture_output=True)
txt = '\n'.join(subprocess.run(['nvdisasm', '-c', os.path.join(d, c)], capture_output=True, text=True).stdout for c in sorted(os.listdir(d)) if c.endswith('.cubin'))
starts, names, insec, cur = [], [], False, None
for line in txt.splitlines():
    if line.startswith('//') and '.text.' in line:
        insec = KERNEL in line
    m = re.match(r'^(\$?[\w\$\.]+):\s*$', line.strip())
    if m and not m.group(1).startswith('.L'):
        cur = m.group(1) if insec else None
        continue
    m = re.match(r'^\s+/\*([0-9a-f]{4,})\*/', line)
    if m and cur and insec:
        starts.append(int(m.group(1), 16)); names.append(cur); cur = None


def pretty(n):
    n2 = n.split('$')[-1] if n.startswith('$_Z') else n
    out = subprocess.run(['c++filt', n2], capture_output=True, text=True).stdout.strip()
    return re.sub(r'\(.*', '', re.sub(r'_INTERNAL_\w+::', '', out))[:44]


rows = list(csv.reader(open(src)))
hdr = rows[1]
ia, ie, it, isamp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
base = int(rows[2][ia], 16)
agg = {}
for r in rows[2:]:
    k = bisect.bisect_right(starts, int(r[ia], 16) - base) - 1
    a = agg.setdefault(names[k] if k >= 0 else '?', [0, 0, 0, 0])
    a[0] += 16; a[1] += int(r[ie] or 0); a[2] += int(r[it] or 0); a[3] += int(r[isamp] or 0)
te, tt, ts = sum(a[1] for a in agg.values()), sum(a[2] for a in agg.values()), sum(a[3] for a in agg.values())
print(f'{"function":44s} {"bytes":>7s} {"warp_inst":>13s} {"%":>5s} {"lanes":>6s} {"thread_inst%":>12s} {"samples%":>8s}')
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{pretty(name):44s} {a[0]:7d} {a[1]:13d} {100 * a[1] / te:5.1f} {a[2] / max(a[1], 1):6.2f} {100 * a[2] / tt:12.1f} {100 * a[3] / max(ts, 1):8.1f}')
print(f'total: {te} warp instructions, {tt / te:.2f} active lanes per instruction, {sum(a[0] for a in agg.values())} bytes of SASS')
