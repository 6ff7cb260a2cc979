"""N>1 host logic on CPU: round-robin page sharding and the single gather of packed records, world_size 2, gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_pages, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from marie_icr_b200.dist import gather_records, shard_indices
    mine = shard_indices(n_pages, rank, world)
    rows = []
    for p in mine:                       # page p contributes (p % 3) + 1 words
        for k in range(p % 3 + 1):
            rows.append([p, k, 10 * p + k, 0, 0, -1, 2, 0, 5, 2])
    rec = torch.tensor(rows, dtype=torch.int32).reshape(-1, 10)
    out = gather_records(rec)
    q.put((rank, mine, out.tolist()))
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    world, n_pages = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pages, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = {r: mine for r, mine, _ in got}
    assert shards[0] == [0, 2, 4, 6] and shards[1] == [1, 3, 5]
    expect = [[p, k, 10 * p + k, 0, 0, -1, 2, 0, 5, 2] for p in range(n_pages) for k in range(p % 3 + 1)]
    for _, _, out in got:
        assert out == expect            # every rank holds all records, ordered by page then detector order


def test_gather_handles_empty_rank():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from marie_icr_b200.dist import gather_records, shard_indices
    assert shard_indices(3, 5, 8) == [] and shard_indices(10, 1, 4) == [1, 5, 9]
    rec = torch.zeros((0, 10), dtype=torch.int32)
    assert gather_records(rec).shape == (0, 10)      # not initialised: identity
