"""world_size-2 run of the multi-GPU host logic on CPU (gloo): chemistry is cell-local, so the
N > 1 path is a static shard of globally numbered cells per rank, no data-path collective, and a
max / sum reduction of timings and counters on rank 0 (bench.py).  Checked here: the shard a rank
generates is independent of the partition, per-rank results concatenate to the single-rank result,
and the reductions bench.py uses give the whole-job aggregate."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pflotran_b200 import abi, synth
from oracle.pyoracle import Oracle

N_PER_RANK = 600


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    w = synth.Workload(name)
    cells = synth.make_cells(w, rank * N_PER_RANK, N_PER_RANK)          # bench.py: start = rank * n
    st = synth.host_state(w, cells)
    xx = cells['tran_xx'].copy()
    it, fl = Oracle(w.tables).react(st, xx, 3600.0, abi.RXN_DT_CONSISTENT)
    # the reductions of bench.py (max of times, sum of counters)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    s = torch.tensor([float(it.sum())], dtype=torch.float64)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), xx=xx, it=it, fl=fl, tmax=t.numpy(), itsum=s.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('name', ['calcite', 'hanford300a_eq'])
def test_two_rank_shards_equal_single_rank(tmp_path, name):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    w = synth.Workload(name)
    cells = synth.make_cells(w, 0, world * N_PER_RANK)
    st = synth.host_state(w, cells)
    xx = cells['tran_xx'].copy()
    it, fl = Oracle(w.tables).react(st, xx, 3600.0, abi.RXN_DT_CONSISTENT)
    parts = [np.load(os.path.join(str(tmp_path), 'rank%d.npz' % r)) for r in range(world)]
    np.testing.assert_array_equal(np.concatenate([p['xx'] for p in parts]), xx)
    np.testing.assert_array_equal(np.concatenate([p['it'] for p in parts]), it)
    np.testing.assert_array_equal(np.concatenate([p['fl'] for p in parts]), fl)
    for p in parts:
        assert p['tmax'][0] == float(world)                 # max over ranks
        assert p['itsum'][0] == float(it.sum())             # whole-job counter


def test_shard_generation_is_partition_independent():
    w = synth.Workload('hanford300a_eq')
    whole = synth.make_cells(w, 0, 10000)
    for start, n in ((0, 4096), (4096, 4096), (5000, 123), (8191, 1809)):
        part = synth.make_cells(w, start, n)
        for k in whole:
            a = whole[k][start:start + n] if whole[k].ndim == 1 or k == 'tran_xx' else whole[k][:, start:start + n]
            np.testing.assert_array_equal(part[k], a)


def test_strong_scaling_split_covers_the_single_rank_batch():
    """bench.py --scaling strong: every rank takes ceil(total / world) contiguous, globally numbered cells - the union is the
    single-rank batch (plus at most world - 1 cells beyond it), shard by shard identical to the corresponding slice."""
    w = synth.Workload('calcite')
    total = 10007
    whole = synth.make_cells(w, 0, total + 8)
    for world in (2, 4, 8):
        n = (total + world - 1) // world
        assert n * world >= total and n * world - total < world
        for rank in range(world):
            part = synth.make_cells(w, rank * n, n)
            np.testing.assert_array_equal(part['tran_xx'], whole['tran_xx'][rank * n:(rank + 1) * n])
