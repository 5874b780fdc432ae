"""world_size-2 gloo tests of the N > 1 host logic: clip sharding + the statistics all-reduce (the path's only
collective) give exactly the single-process mean / std."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, feats, out_path):
    sys.path.insert(0, REPO)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from seld_b200 import pipeline
    from seld_b200.sharding import shard_indices
    mine = shard_indices(feats.shape[0], rank, world)
    x = feats[mine].double()
    n = x.shape[-2] * x.shape[-1]
    acc = torch.zeros(2 * n + 1, dtype=torch.float64)
    flat = x.reshape(-1, n)
    acc[:n] = flat.sum(0)
    acc[n:2 * n] = (flat * flat).sum(0)
    acc[2 * n] = flat.shape[0]
    pipeline.allreduce_statistics(acc)
    if rank == 0:
        np.save(out_path, acc.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_statistics_allreduce(tmp_path):
    from oracle import extractor as O
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(7, 50, 8, 7, generator=g) * 5 - 20          # 7 clips: ragged split over 2 ranks
    out = str(tmp_path / 'acc.npy')
    port = _free_port()
    mp.spawn(_worker, args=(2, port, feats, out), nprocs=2, join=True)
    acc = np.load(out)
    n = 8 * 7
    rows = acc[2 * n]
    assert rows == 7 * 50
    mean = acc[:n] / rows
    std = np.sqrt(np.maximum(acc[n:2 * n] / rows - mean ** 2, 0))
    m64, s64 = O.statistics_f64(feats.numpy())
    assert np.abs(mean - m64.reshape(-1)).max() <= 1e-10 and np.abs(std - s64.reshape(-1)).max() <= 1e-10


def test_shard_indices_cover_every_clip_once():
    from seld_b200.sharding import shard_indices
    for n, world in ((600, 8), (600, 1), (7, 2), (3, 8), (0, 4)):
        seen = sorted(i for r in range(world) for i in shard_indices(n, r, world))
        assert seen == list(range(n))
        sizes = [len(shard_indices(n, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    assert shard_indices(600, 3, 8)[:3] == [3, 11, 19]               # clip i -> rank i mod G
