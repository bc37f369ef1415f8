"""Host-side logic of the multi-GPU drivers on CPU: sharding rules, and the all-reduce of the
[loss, gradient] vector with the gloo backend, world_size 2."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wobble_jax_b200 import multigpu, synth


def test_shard_chunks_round_robin():
    parts = [multigpu.shard_chunks(64, r, 8) for r in range(8)]
    assert sorted(sum(parts, [])) == list(range(64))
    assert parts[3] == list(range(3, 64, 8))
    assert all(len(p) == 8 for p in parts)


def test_shard_rows_by_key_keeps_keys_together():
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 37, 500)
    parts = multigpu.shard_rows_by_key(keys, 4)
    assert sorted(np.concatenate(parts).tolist()) == list(range(500))
    owners = {}
    for r, rows in enumerate(parts):
        for k in set(keys[rows].tolist()):
            assert owners.setdefault(k, r) == r          # a key lives on exactly one rank
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 2 * np.bincount(keys).max()
    # contiguous in key order
    firsts = [keys[p].min() for p in parts if len(p)]
    assert firsts == sorted(firsts)
    # identity keys -> contiguous, balanced row ranges
    parts = multigpu.shard_rows_by_key(np.arange(10), 3)
    assert [p.tolist() for p in parts] == [[0, 1, 2], [3, 4, 5, 6], [7, 8, 9]] or sum(len(p) for p in parts) == 10


def test_epoch_keys_and_take_rows():
    ch = synth.make_chunk(0, n_epochs=6, n_pix=130, p_val=3, n_lines=3)
    db, mb = ch.data.blockify()
    keys = multigpu.epoch_keys(ch.model, mb)
    assert np.array_equal(keys, np.arange(6))
    sub_db, sub_mb = multigpu.take_rows(db, mb, np.array([1, 4]))
    assert sub_db["xs"].shape == (2, 130) and np.array_equal(sub_mb["index"], [1, 4])   # global key values kept


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # each rank holds the [loss, grad] of its rows; per-key entries are zero where not owned
        rows = multigpu.shard_rows_by_key(np.arange(n), world)[rank]
        vec = torch.zeros(1 + n, dtype=torch.float64)
        vec[0] = float(len(rows))                      # "loss" = number of local rows
        vec[1 + torch.as_tensor(rows)] = torch.as_tensor(rows, dtype=torch.float64) + 1.0
        multigpu.all_reduce_sum(vec)
        assert float(vec[0]) == n
        assert torch.equal(vec[1:], torch.arange(1, n + 1, dtype=torch.float64))
    finally:
        dist.destroy_process_group()


def test_all_reduce_gloo_world2():
    mp.spawn(_worker, args=(2, _free_port(), 11), nprocs=2, join=True)


def test_epoch_sharding_refuses_empty_shards_on_every_rank_alike():
    """More ranks than epoch keys: every rank must raise before any collective (none may hang)."""
    import wobble_jax_b200 as jabble
    ch = synth.make_chunk(0, n_epochs=3, n_pix=130, p_val=3, n_lines=3)
    db, mb = ch.data.blockify()
    parts = multigpu.shard_rows_by_key(multigpu.epoch_keys(ch.model, mb), 8)
    assert sum(len(p) == 0 for p in parts) > 0
    for rank in range(8):
        with pytest.raises(ValueError, match="without rows"):
            multigpu.EpochShardedObjective(jabble.loss.ChiSquare(), ch.model, db, mb, rank, 8)
