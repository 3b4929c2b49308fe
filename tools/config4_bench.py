#!/usr/bin/env python3
"""BASELINE.json configs[3]: two-carrier cell, NPRACH and NPUSCH parameters both chosen from outside at every decision
(M=3000, ratio=0.5, random traffic: the 'cfg4_two_carriers_bursty' golden's scenario), N instances PER GPU, one launch per
decision of every instance (NodeB.step level, the full 23-item action vector drawn uniformly on the device).
Shards over GPUs under torchrun (262144 instances = 2 x 131072 or 4 x 65536), statistics all-reduced over NCCL.
usage: python tools/config4_bench.py [N_per_gpu] [decisions]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from nb_iot_environment_b200 import BatchedNBIoTEnv
from nb_iot_environment_b200.sharding import shard_seeds, allreduce_stats
CFG4 = {'M': 3000, 'ratio': 0.5, 'buffer_range': [100, 600], 'reward_criteria': 'throughput', 'random': True}
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
D = int(sys.argv[2]) if len(sys.argv) > 2 else 300
W = 50
rank, world, local = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))
dist = None
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
env = BatchedNBIoTEnv(CFG4, num_envs=N, seeds=shard_seeds(N * world, rank, world), n_carriers=2, device=local, all_states_external=True,
                      ue_cap=int(os.environ.get('NBIOT_CFG4_UE_CAP', '3000')), grid_w=2048, hist_cap=64)
# uniform over control_max_values for 2 carriers minus what the reference raises on (tools/gen_golden.py policy_action)
mx = torch.tensor([1, 3, 6, 7, 3, 3, 4, 11, 7, 7, 10, 15, 5, 7, 7, 5, 7, 7, 5, 7, 7, 14, 14], device=dev) + 1
g = torch.Generator(device=dev); g.manual_seed(1 + rank)


def actions():
    a = (torch.rand((N, 23), device=dev, generator=g) * mx).to(torch.uint8)
    a[:, 5] = torch.where(a[:, 5] == 2, torch.full_like(a[:, 5], 3), a[:, 5])                  # sc = 2 is invalid
    rar = env.record_field('state') == 2                                                      # RAR_window: ce_level / rar_Imcs share items 1 / 2
    a[:, 1] = torch.where(rar, a[:, 1] % 3, a[:, 1]); a[:, 2] = torch.where(rar, a[:, 2] % 3, a[:, 2])
    return a


env.reset()
for _ in range(W):
    env.node_step(actions())
s0 = env.stats_tensor().clone()
if dist is not None:
    dist.barrier()
torch.cuda.synchronize(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(D):
    env.node_step(actions())
e1.record(); torch.cuda.synchronize(dev)
d = env.stats_tensor().clone(); d[:20] -= s0[:20]
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
if dist is not None:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    d = allreduce_stats(d)
if rank == 0:
    ms = float(ms.item())
    print(json.dumps({'n_gpus': world, 'N_per_gpu': N, 'decisions': D, 'ms_per_decision': ms / D, 'instance_decisions_per_s': N * world * D / ms * 1e3,
                      'subframe_instances_per_s': int(d[13].item()) / ms * 1e3, 'events_per_subframe': int(d[14].item()) / max(int(d[13].item()), 1),
                      'live_ues_per_cell': int(d[17].item()) / (N * world), 'faulted': int(d[16].item()), 'max_live_ues': int(d[21].item()),
                      'max_lookahead': int(d[20].item()), 'state_GB_per_gpu': env.state.numel() / 1e9}), flush=True)
env.close()
if dist is not None:
    dist.destroy_process_group()
