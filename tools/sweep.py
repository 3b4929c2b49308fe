#!/usr/bin/env python3
"""BASELINE.json configs[4] (the throughput sweep): config-1 traffic (M=1000, ratio=1, default agents on the device, no host
round trip), N instances PER GPU advanced to a horizon in ONE launch, episode statistics reduced on the device and
all-reduced over NCCL when run under torchrun (weak scaling; instance g of the job has Philox key g on any GPU count).
usage: python tools/sweep.py [N ...]                                   (one GPU)
       python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/sweep.py [N ...]
prints one JSON line per N (rank 0)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nb_iot_environment_b200 import BatchedNBIoTEnv
from nb_iot_environment_b200.sharding import shard_seeds, allreduce_stats
CFG1 = {'ratio': 1.0, 'M': 1000, 'buffer_range': [100, 600], 'reward_criteria': 'throughput'}
T = int(os.environ.get('NBIOT_SWEEP_T', '30000'))
rank, world, local = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))
dist = None
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
for N in [int(x) for x in (sys.argv[1:] or ['1000', '10000', '100000'])]:
    env = BatchedNBIoTEnv(CFG1, num_envs=N, seeds=shard_seeds(N * world, rank, world), device=local, ue_cap=128, grid_w=1024, hist_cap=64)
    env.reset()
    env.run_until(3000)                      # warm-up launch
    s0 = env.stats_tensor().clone()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.run_until(T); e1.record(); torch.cuda.synchronize(dev)
    d = env.stats_tensor().clone()
    d[:20] -= s0[:20]                         # sums since the warm-up; the last entries are maxima
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)      # device time, max over ranks
        d = allreduce_stats(d)                         # the one collective of the path
    if rank == 0:
        sf = int(d[13].item())
        print(json.dumps({'n_gpus': world, 'N_per_gpu': N, 'horizon': T, 'ms': float(ms.item()), 'subframe_instances_per_s': sf / float(ms.item()) * 1e3,
                          'events_per_subframe': int(d[14].item()) / max(sf, 1), 'faulted': int(d[16].item()), 'max_lookahead': int(d[20].item()),
                          'max_live_ues': int(d[21].item()), 'state_GB_per_gpu': env.state.numel() / 1e9, 'departures': int(d[3:6].sum().item())}), flush=True)
    env.close(); del env
    torch.cuda.empty_cache()
if dist is not None:
    dist.destroy_process_group()
