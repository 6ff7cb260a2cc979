"""Data-parallel sharding of a page stream over the GPUs of one box and the single gather of results.

The reference scales by running one executor replica per GPU behind a load balancer
(marie/orchestrate/deployments/__init__.py:1316-1345) and has no collective on this path.
Here: one process per GPU (torchrun), page i -> rank i mod world, no traffic during compute, then one exchange of the
packed per-word records (pipeline.RECORD_HEAD + tokens, int32) over NCCL (NVLink) — or gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_indices(n_pages, rank, world):
    """Round-robin page assignment (keeps density skew balanced)."""
    return list(range(rank, n_pages, world))


def gather_records(rec, group=None, dst=None):
    """rec: [n_local, width] int32 (device tensor for nccl, cpu for gloo) with GLOBAL page ids in column 0.
    Returns the records of all ranks sorted by (page, local order) on every rank (dst=None) — two collectives:
    all_gather of the counts, all_gather of the padded record blocks."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return rec
    world = dist.get_world_size(group)
    count = torch.tensor([rec.shape[0]], dtype=torch.int64, device=rec.device)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c.item()) for c in counts]
    width = rec.shape[1]
    cap = max(max(counts), 1)
    padded = torch.zeros((cap, width), dtype=rec.dtype, device=rec.device)
    padded[:rec.shape[0]] = rec
    blocks = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(blocks, padded, group=group)
    allrec = torch.cat([b[:c] for b, c in zip(blocks, counts)])
    if allrec.shape[0] == 0:
        return allrec
    order = torch.sort(allrec[:, 0].to(torch.int64), stable=True).indices
    return allrec[order]
