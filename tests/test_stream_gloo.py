"""world_size-2 `gloo` test of the multi-GPU path's host logic (runs on CPU): one checkpoint broadcast from rank 0,
round-robin tile shards, max-over-ranks timing."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from samrs_b200.config import geometry
from samrs_b200.stream import broadcast_state_dict, max_over_ranks, shard_indices
from samrs_b200.weights import check_state_dict, synthetic_state_dict


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = geometry("vit_t64")
    loads = []

    def load():            # only the source rank may touch the checkpoint
        loads.append(rank)
        return synthetic_state_dict("vit_t64", 0)

    sd = broadcast_state_dict(g, load, torch.device("cpu"), src=0)
    check_state_dict(g, sd)
    ref = synthetic_state_dict("vit_t64", 0)
    same = all(torch.equal(sd[k], ref[k]) for k in ref)
    slow = max_over_ranks(10.0 + rank, torch.device("cpu"))
    q.put((rank, same, loads, shard_indices(11, rank, world), slow))
    dist.destroy_process_group()


def test_broadcast_and_shard_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, same0, loads0, shard0, slow0), (r1, same1, loads1, shard1, slow1) = res
    assert same0 and same1
    assert loads0 == [0] and loads1 == []
    assert sorted(shard0 + shard1) == list(range(11)) and not set(shard0) & set(shard1)
    assert slow0 == slow1 == 11.0
