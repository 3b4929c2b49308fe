#!/usr/bin/env python3
"""Small fixed command for ncu: the bench workload (4096 cells, NPRACH agent external), 3 warm-up steps + 2 steps.
Launch order of nbiot_sim_kernel: #0 construct+reset, #1 reset, #2-4 warm-up steps, #5-6 measured steps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from nb_iot_environment_b200 import BatchedNBIoTEnv  # noqa: E402

N = int(os.environ.get('NBIOT_PROFILE_N', bench.N_PER_GPU))
env = BatchedNBIoTEnv(bench.CONF, agents_conf=bench.AGENTS, ext_agent=bench.EXT_AGENT, num_envs=N, **bench.CAPS)
acts = torch.from_numpy(bench.ext_actions(5, N, seed=1234)).cuda()
env.reset()
for s in range(5):
    obs, rew, _, _, _ = env.step(acts[s])
torch.cuda.synchronize()
st = env.stats()
print('ok', st['subframes'], st['events'], st['faulted'], float(obs.sum()), float(rew.sum()))
