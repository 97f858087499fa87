"""The N>1 path on CPU: world_size-2 `gloo`.  The partition + pack + allreduce logic of
pollen_b200.sharding is the product's; the per-rank compute is substituted by the oracle
here (no GPU in this container), which is the checker role the oracle is allowed."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from pollen_b200 import sharding, synth


def test_lpt_partition_balances_and_covers():
    rng = np.random.default_rng(3)
    lens = rng.integers(1, 10_000, 90)
    for n in (1, 2, 4, 8):
        parts = sharding.lpt_partition(lens, n)
        assert sorted(i for p in parts for i in p) == list(range(90))
        loads = [int(lens[p].sum()) for p in parts]
        assert max(loads) - min(loads) <= int(lens.max())
        assert all(p == sorted(p) for p in parts)
    assert sharding.lpt_partition([5, 5, 5], 5)[3:] == [[], []]


def test_pack_shard_is_bit_identical_to_subset_generation():
    cfg = synth.CONFIGS["tiny"]
    steps, s, e = synth.make_graph(cfg)
    part = [1, 4, 6]
    a = sharding.pack_shard(steps, s, e, part)
    b = synth.make_graph(cfg, path_subset=part)
    assert all((x == y).all() for x, y in zip(a, b))


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = synth.CONFIGS["tiny"]
        start, end = synth.make_spans(cfg.n_paths, cfg.n_steps, cfg.jitter_pct)
        part = sharding.lpt_partition(end - start, world)[rank]
        steps, ls, le = synth.make_graph(cfg, path_subset=part)
        rc, d, u = O.depth_with_uniq(steps, ls, le, cfg.n_segs)
        assert rc == 0
        buf = torch.from_numpy(np.concatenate([d, u]).astype(np.uint32).view(np.int32).copy())
        sharding.allreduce_counts(buf)
        ret[rank] = buf.numpy().view(np.uint32).copy()
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_depth_equals_single():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    cfg = synth.CONFIGS["tiny"]
    steps, s, e = synth.make_graph(cfg)
    rc, d, u = O.depth_with_uniq(steps, s, e, cfg.n_segs)
    want = np.concatenate([d, u]).astype(np.uint32)
    assert (ret[0] == want).all() and (ret[1] == want).all()
