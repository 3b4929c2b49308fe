#!/usr/bin/env python3
"""Scheduling study for the thread-per-cell kernel: how many lanes of a warp are active per issued instruction when the 32 cells of
a warp run (a) every task class present in turn (what SIMT divergence does) or (b) one voted class per round.
Task traces come from the CPU build of the core cut into tasks (tests/_hostemu, libemu1t.so); costs per class are arguments."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import hostemu_py

NAMES = ['BEGIN', 'STEP_DATA', 'STEP_RARW', 'STEP_RAREND', 'STEP_UPDATE', 'NPRACH_START', 'MSG3', 'NPUSCH_END', 'FINISH', 'GLUE']


def traces(conf, n_cells, T, **caps):
    b = hostemu_py.EmuBatch(conf, list(range(n_cells)), lanes='1t', ctrl=hostemu_py.ControllerTables().build(), **caps)
    L = b.L
    cap = 1 << 24
    buf = np.zeros((cap, 2), dtype=np.int32)
    b.reset()
    L.emu_set_tasklog.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.emu_set_tasklog(buf.ctypes.data, cap)
    b.advance(None, until_time=T)
    n = L.emu_tasklog_len()
    L.emu_set_tasklog(None, 0)
    log = buf[:n]
    return [log[log[:, 0] == i, 1].tolist() for i in range(n_cells)]


def simulate(seqs, cost, policy):
    """32 lanes; returns (lane-instructions, warp-instructions)"""
    pos = [0] * len(seqs)
    lane_i = warp_i = 0
    wait = [0] * len(seqs)
    while True:
        cur = [s[p] if p < len(s) else -1 for s, p in zip(seqs, pos)]
        present = sorted({c for c in cur if c >= 0})
        if not present:
            break
        if policy == 'simt':
            run = present
        else:
            score = {}
            for l, c in enumerate(cur):
                if c >= 0:
                    score[c] = score.get(c, 0) + (1 + (wait[l] / policy if policy > 0 else 0))
            run = [max(score, key=lambda c: (score[c], -c))]
        for c in run:
            lanes = [l for l, x in enumerate(cur) if x == c]
            warp_i += cost[c]
            lane_i += cost[c] * len(lanes)
            for l in lanes:
                pos[l] += 1
                wait[l] = 0
        for l, c in enumerate(cur):
            if c >= 0 and c not in run:
                wait[l] += 1
    return lane_i, warp_i


if __name__ == '__main__':
    CFG1 = {'ratio': 1.0, 'M': 1000, 'buffer_range': [100, 600], 'reward_criteria': 'throughput'}
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
    seqs = traces(CFG1, 64, T, ue_cap=256, grid_w=1024, hist_cap=64)
    flat = np.concatenate([np.asarray(s) for s in seqs])
    freq = np.bincount(flat, minlength=10)
    print('task mix:', {NAMES[k]: int(freq[k]) for k in range(10) if freq[k]})
    for label, cost in (('equal costs', [1] * 10), ('guess', [0, 3, 3, 1, 2, 3, 2, 2, 2, 1])):
        for policy in ('simt', 0, 4, 16):
            tot_l = tot_w = 0
            for w in range(0, 64, 32):
                l, wi = simulate(seqs[w:w + 32], cost, policy)
                tot_l += l; tot_w += wi
            print(f'{label:12s} policy {policy!s:5s}: active lanes per instruction {tot_l / tot_w:5.2f}  (warp instructions {tot_w})')
