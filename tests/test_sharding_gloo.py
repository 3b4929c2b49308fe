"""world_size-2 run of the multi-GPU host logic on CPU (gloo): sharding of instances / seeds and the all-reduce of
the episode statistics. Each rank simulates its shard with the CPU oracle; the reduced vector must equal the
statistics of the unsharded job."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG1 = {'ratio': 1.0, 'M': 1000, 'buffer_range': [100, 600], 'reward_criteria': 'throughput'}
TOTAL, T = 6, 6000


def _stats_of(seeds):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from nb_iot_environment_b200._cabi import N_STATS
    from oracle_controller import OracleController
    v = np.zeros(N_STATS, dtype=np.int64)
    for s in seeds:
        c = OracleController(CFG1, int(s))
        c.reset()
        c.run_until(T)
        k = c.cell.counters()
        v[0:3] += k['ue_arrivals']; v[3:6] += k['ue_departures']; v[6:9] += k['backoffs']
        v[9] += k['ue_connections']; v[10] += k['error_count']; v[11] += k['attempts_count']; v[12] += k['total_bits']
        v[13] += int(c.rec['time']); v[15] += k['draws']
        v[20] = max(v[20], int(c.rec['time']) % 97)        # stand-ins for the two MAX entries
        v[21] = max(v[21], k['ue_count'])
    return v


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from nb_iot_environment_b200.sharding import shard_seeds, allreduce_stats
    seeds = shard_seeds(TOTAL, rank, world, seed_base=40)
    st = torch.from_numpy(_stats_of(seeds))
    allreduce_stats(st)
    if rank == 0:
        np.save(out, st.numpy())
    dist.destroy_process_group()


def test_shard_ranges_cover_the_job():
    sys.path.insert(0, ROOT)
    from nb_iot_environment_b200.sharding import shard_range, shard_seeds
    for total in (1, 7, 4096, 262144):
        for world in (1, 2, 4, 8):
            blocks = [shard_range(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == total
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1
    assert shard_seeds(10, 1, 2, seed_base=5).tolist() == [10, 11, 12, 13, 14]


def test_two_rank_allreduce_of_episode_statistics(tmp_path, oracle_lib):
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / 'stats.npy')
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    want = _stats_of(np.arange(40, 40 + TOTAL))
    assert np.array_equal(got, want)
