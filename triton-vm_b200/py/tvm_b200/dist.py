"""Multi-GPU plumbing (torch.distributed; NCCL on the GPU box, gloo in the CPU tests).

The prove() path shards by COSET of the evaluation domain (it mirrors the
reference's own coset decomposition of the just-in-time LDE, stark.rs:824-885): with r = 8 cosets
of the trace domain, rank g owns evaluation-domain rows i = c + r*k for its cosets c.  Row hashing,
AIR evaluation ("next row" i + r stays inside a coset), DEEP and the quotient are then row-local;
only leaf digests and the quotient codeword are gathered.  This module holds the rank-local index
logic and the collectives' host side; `bench.py --gpus N` currently runs independent replicas.
"""
import torch
import torch.distributed as dist


def cosets_of_rank(num_cosets, world_size, rank):
    """Cosets owned by `rank`: strided assignment c = rank, rank + world, ... (balanced for world | r)."""
    if world_size > num_cosets:
        raise ValueError(f"at most {num_cosets} ranks can share one proof (one coset each)")
    return list(range(rank, num_cosets, world_size))


def row_owner(row_index, num_cosets, world_size):
    """Rank holding evaluation-domain row i (coset c = i mod r)."""
    return (row_index % num_cosets) % world_size


def local_row(row_index, num_cosets, world_size, trace_len):
    """(rank, local memory index) of evaluation-domain row i in the rank's coset-major LDE shard."""
    c, k = row_index % num_cosets, row_index // num_cosets
    rank = c % world_size
    return rank, (c // world_size) * trace_len + k


def max_over_ranks(value_ms, device="cpu"):
    """Timing rule of the bench contract: a step's duration is the max over ranks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value_ms)
    t = torch.tensor([float(value_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_gather_rows(local, world_size):
    """Gathers equally sized per-rank shards [m, w] and interleaves them back to natural row order
    i = c + r*k when every rank owns the cosets c = rank (mod world).  Used for leaf digests (w = 5)
    and the quotient codeword (w = 3)."""
    if not dist.is_initialized() or world_size == 1:
        return local
    parts = [torch.empty_like(local) for _ in range(world_size)]
    dist.all_gather(parts, local.contiguous())
    return torch.stack(parts, dim=0)   # [rank, m, w]; callers map (rank, local index) -> row with local_row()
