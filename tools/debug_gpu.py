import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nb_iot_environment_b200 import BatchedNBIoTEnv, parameters as par
conf = {'ratio': 1.0, 'M': 1000, 'buffer_range': [100, 600], 'reward_criteria': 'throughput'}
for n in (4, 64):
    env = BatchedNBIoTEnv(conf, num_envs=n, seeds=list(range(n)), all_states_external=True, ue_cap=256, grid_w=1024)
    r = env.records(); print(n, 'after create: state', r['state'][:6], 'conf', r['nprach_conf'][0][:6], 'fault', r['fault'][:6], 'draws', r['draws_lo'][:6])
    env.reset()
    r = env.records(); print(n, 'after reset : state', r['state'][:6], 'conf', r['nprach_conf'][0][:6], 'fault', r['fault'][:6])
    a = np.tile(np.asarray(par.default_action_vector(), dtype=np.uint8), (n, 1))
    for k in range(3):
        env.node_step(a)
        r = env.records(); print(n, 'step', k, 'state', r['state'][:6], 'time', r['time'][:6], 'fault', r['fault'][:6])
    print(env.stats())
    env.close()
