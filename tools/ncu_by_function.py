#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` export of nbiot_sim_kernel per device function.
usage: python tools/ncu_by_function.py <source.csv> [lib.so]
Shows, for every function of the kernel's text section: SASS bytes, warp instructions executed, share of
executed instructions, stall samples and the share of samples that are instruction-fetch stalls."""
import csv, os, re, subprocess, sys, tempfile, bisect
src = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'nb_iot_environment_b200', 'libnbiot_b200.so')
KERNEL = os.environ.get('NBIOT_PROFILE_KERNEL', 'nbk_c1s16nbiot_sim_kernelILi7')   # the 72-register build (bench workload)
d = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=d, capture_output=True)
cubs = sorted(f for f in os.listdir(d) if f.endswith('.cubin'))      # one per translation unit of the library
txt = '\n'.join(subprocess.run(['nvdisasm', '-c', os.path.join(d, cub)], capture_output=True, text=True).stdout for cub in cubs)
# function start offsets inside the sim kernel's section
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
    n2 = n.split('$')[-1] if n.startswith('$_Z') else n        # $<kernel>$<function>
    out = subprocess.run(['c++filt', n2], capture_output=True, text=True).stdout.strip()
    out = re.sub(r'_INTERNAL_\w+::', '', out)
    return re.sub(r'\(.*', '', out)[:40]
rows = list(csv.reader(open(src)))
hdr = rows[1]
ia, ie, isamp, inoinst = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('stall_no_inst')
ilong, ishort, iwait, ibr = hdr.index('stall_long_sb'), hdr.index('stall_short_sb'), hdr.index('stall_wait'), hdr.index('stall_branch_resolving')
base = int(rows[2][ia], 16)
agg = {}
for r in rows[2:]:
    off = int(r[ia], 16) - base
    k = bisect.bisect_right(starts, off) - 1
    name = names[k] if k >= 0 else '?'
    a = agg.setdefault(name, [0, 0, 0, 0, 0, 0, 0, 0])
    a[0] += 16; a[1] += int(r[ie] or 0); a[2] += int(r[isamp] or 0); a[3] += int(r[inoinst] or 0)
    a[4] += int(r[ilong] or 0); a[5] += int(r[ishort] or 0); a[6] += int(r[iwait] or 0); a[7] += int(r[ibr] or 0)
tot_e = sum(a[1] for a in agg.values()); tot_s = sum(a[2] for a in agg.values())
print(f'{"function":40s} {"bytes":>7s} {"inst_exec":>12s} {"%exec":>6s} {"samples":>8s} {"%smp":>6s} {"no_inst":>8s} {"long_sb":>8s} {"short":>7s} {"wait":>6s} {"branch":>6s}')
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print(f'{pretty(name):40s} {a[0]:7d} {a[1]:12d} {100*a[1]/tot_e:6.1f} {a[2]:8d} {100*a[2]/max(tot_s,1):6.1f} {a[3]:8d} {a[4]:8d} {a[5]:7d} {a[6]:6d} {a[7]:6d}')
print('total executed', tot_e, 'samples', tot_s)
