"""CPU test of the N > 1 host logic (gloo, world_size 2): round-robin block sharding, per-rank
processing and the gather of per-block results.  The compute function is injected: here it is
the oracle (the product has no CPU path); on the GPU box bench.py passes the CUDA context."""
import hashlib
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from libsais_b200 import gen, sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_blocks, q):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    import _libs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = _libs.oracle()

    def bwt(T):
        rc, U = o.bwt(T)
        return U, rc

    local = sharding.run_batch(n_blocks, rank, world, lambda b: gen.dna(1000 + b, 20000 + 17 * b), bwt)
    merged = sharding.gather_results(local, dist)
    q.put((rank, sorted(local.keys()), merged))
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_assignment_is_a_partition():
    for world in (1, 2, 3, 8):
        for n_blocks in (0, 1, 7, 64):
            seen = []
            for r in range(world):
                seen += sharding.blocks_for_rank(n_blocks, r, world)
            assert sorted(seen) == list(range(n_blocks))
    assert sharding.blocks_for_rank(64, 3, 8) == list(range(3, 64, 8))


def test_two_rank_batch_matches_single_process():
    import _libs
    n_blocks, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_blocks, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = _libs.oracle()
    expect = {}
    for b in range(n_blocks):
        rc, U = o.bwt(gen.dna(1000 + b, 20000 + 17 * b))
        expect[b] = (rc, hashlib.sha256(U.tobytes()).hexdigest())
    for rank, mine, merged in got:
        assert mine == list(range(rank, n_blocks, world))
        assert merged == expect
