#!/usr/bin/env python3
"""Turn an ncu full capture of nbiot_sim_kernel into the tracked summaries under profiles/.
usage: python tools/summarize_profile.py gpurun_out/prof.ncu-rep profiles/r01_sim_kernel
writes <out>_metrics.txt (key raw metrics + per-function table) and profiles/traffic.json"""
import csv, json, os, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
lines = [f'# ncu --set full --clock-control none, one launch of nbiot_sim_kernel (bench workload: 4096 cells x 3000 subframes)', f'# source: {os.path.basename(rep)}', '']
for k in keys:
    if k in m:
        lines.append(f'{k:80s} {m[k][0]:>20s} {m[k][1]}')
lines.append('')
lines.append('# warp stall reasons, average warps stalled per issue-active cycle (smsp__average_warps_issue_stalled_*_per_issue_active.ratio)')
for h in hdr:
    if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
        try:
            if float(m[h][0]) >= 0.01:
                lines.append(f'{h:80s} {float(m[h][0]):20.3f}')
        except ValueError:
            pass
srcp = rep.replace('.ncu-rep', '_src.csv')
open(srcp, 'w').write(subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout)
tab = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_by_function.py'), srcp], capture_output=True, text=True).stdout
lines += ['', '# per device function (SASS bytes, warp instructions executed, stall samples by reason)', tab]
open(out + '_metrics.txt', 'w').write('\n'.join(lines) + '\n')
def to_bytes(v, u):
    f = float(v)
    return f * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
traffic = to_bytes(*m['dram__bytes_read.sum']) + to_bytes(*m['dram__bytes_write.sum'])
json.dump({'dram_bytes_per_launch': traffic, 'kernel': 'nbiot_sim_kernel', 'source': os.path.basename(out) + '_metrics.txt',
           'note': 'dram__bytes_read.sum + dram__bytes_write.sum of one launch (4096 cells x 3000 subframes), ncu --set full'},
          open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w'), indent=1)
print('\n'.join(lines[:40]))
