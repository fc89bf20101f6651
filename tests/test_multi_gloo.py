"""world_size-2 test of the multi-GPU host logic on CPU processes (gloo backend): block
round-robin, contiguous query chunks and in-order gathering.  No kernel is called here."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from text_compression_b200 import multi
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ws, rk = multi.world()
        assert (ws, rk) == (world, rank)
        lo, hi = multi.query_slice(q, ws, rk)
        local = np.arange(lo, hi, dtype=np.int64) * 3 - 1          # stand-in for per-query counts
        full = multi.gather_in_order(local, q)
        assert full.tolist() == (np.arange(q) * 3 - 1).tolist()
        mine = multi.blocks_of_rank(9, ws, rk)
        out.put((rank, mine, int(full.sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("q", [11, 2, 0])
def test_gather_in_order_world2(q):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == [0, 2, 4, 6, 8] and res[1][1] == [1, 3, 5, 7]
    assert res[0][2] == res[1][2]
