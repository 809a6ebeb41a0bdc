"""CPU (gloo, world_size 2): the N>1 host logic - round-robin sequence sharding (the reference's worker assignment,
lib/test/evaluation/running.py:134-141), the per-step all-gather of boxes and the un-sharding back to sequence order."""
import os
import socket
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)      # spawned workers re-import this module without conftest.py

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_sequences_partition():
    import mmt_b200  # noqa: F401
    from mmt_b200 import runner
    for n, world in ((128, 8), (7, 2), (5, 8), (64, 1)):
        seen = []
        for r in range(world):
            seen += runner.shard_sequences(n, world, r)
        assert sorted(seen) == list(range(n))
    with pytest.raises(ValueError):
        runner.shard_sequences(4, 2, 2)


def _worker(rank, world, port, n_seq, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mmt_b200  # noqa: F401
    from mmt_b200 import runner
    owned = runner.shard_sequences(n_seq, world, rank)
    # stand-in for the forward: the box of sequence s is a function of s only
    boxes = torch.tensor([[s, 2.0 * s, s + 0.5, 1.0] for s in owned], dtype=torch.float32)
    gathered = runner.gather_boxes(boxes)
    full = runner.unshard_boxes(gathered, n_seq)
    ret[rank] = full.clone()
    dist.barrier()
    dist.destroy_process_group()


def test_gather_boxes_world2_gloo():
    world, n_seq = 2, 6
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_seq, ret), nprocs=world, join=True)
    want = torch.tensor([[s, 2.0 * s, s + 0.5, 1.0] for s in range(n_seq)], dtype=torch.float32)
    for r in range(world):
        assert torch.equal(ret[r], want), (r, ret[r])
