"""Experiment (round 2): does stepping the 4096 cells of configs[1] as G groups on G streams, each group's steps
stream-ordered and no barrier between groups, hide the one-wave tail? (cells are independent: results cannot change)
usage: python tools/grouped_exp.py [G ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from nb_iot_environment_b200 import BatchedNBIoTEnv, _cabi

N, K, W = 4096, 20, 3
dev = torch.device('cuda', 0)
acts = torch.from_numpy(bench.ext_actions(W + K, N, seed=1234)).to(dev)
for G in [int(x) for x in sys.argv[1:]] or [1, 2, 4, 8, 16]:
    n = N // G
    envs = [BatchedNBIoTEnv(bench.CONF, agents_conf=bench.AGENTS, ext_agent=bench.EXT_AGENT, num_envs=n, seeds=list(range(g * n, (g + 1) * n)), device=0, **bench.CAPS) for g in range(G)]
    streams = [torch.cuda.Stream(dev) for _ in range(G)]
    a = [[acts[s, g * n:(g + 1) * n].contiguous() for g in range(G)] for s in range(W + K)]
    for e in envs:
        e.reset()
    torch.cuda.synchronize()
    def run(s0, s1):
        for s in range(s0, s1):
            for g in range(G):
                with torch.cuda.stream(streams[g]):
                    e = envs[g]
                    _cabi.check(e.L.nbiot_advance(e._h, a[s][g].data_ptr(), -1, 1 << 30, streams[g].cuda_stream))
                    _cabi.check(e.L.nbiot_gather_obs(e._h, e._items, e._n_items, e._obs.data_ptr(), e._reward.data_ptr(), streams[g].cuda_stream))
    run(0, W)
    torch.cuda.synchronize()
    sub0 = sum(int(e.stats_tensor()[13].item()) for e in envs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(W, W + K)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    sub = sum(int(e.stats_tensor()[13].item()) for e in envs) - sub0
    faults = sum(int(e.stats_tensor()[16].item()) for e in envs)
    print(f'G={G:3d} cells/group {n:5d}: {1e3 * dt / K:7.3f} ms per step of the whole batch, {sub / dt:.4g} sf-inst/s, faults {faults}, geometry {envs[0].launch_geometry()}', flush=True)
    for e in envs:
        e.close()
    del envs
    torch.cuda.empty_cache()
