"""Multi-GPU plumbing (torch.distributed; NCCL on the GPU box, gloo in the CPU tests).

The prove() path shards by COSET of the evaluation domain (it mirrors the
reference's own coset decomposition of the just-in-time LDE, stark.rs:824-885): with r = 8 cosets
of the trace domain, rank g owns evaluation-domain rows i = c + r*k for its cosets c.  Row hashing,
AIR evaluation ("next row" i + r stays inside a coset), DEEP and the quotient are then row-local;
only interpolant coefficients, leaf digests, the quotient and the DEEP codeword are gathered.  This
module holds the rank-local index logic and `TorchDistComm`, the implementation of the C ABI's
`tvm_comm` collectives on top of torch.distributed (NCCL over NVLink on the GPU box; gloo, staged
through host memory, in tests where several ranks share one GPU).
"""
import ctypes

import torch
import torch.distributed as dist


class _DevMem:
    """Zero-copy view of raw device memory for torch.as_tensor (__cuda_array_interface__)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class TorchDistComm:
    """tvm_comm over a torch.distributed process group.  NCCL: the collectives are enqueued with the library's
    CUDA stream as the current stream, i.e. ordered after the kernels that produced the data, without a host
    synchronisation.  gloo: the stream is synchronised and the buffer staged through host memory."""

    def __init__(self, device, group=None):
        self.group = group
        self.device = torch.device(device)
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        self.calls = {"all_gather": 0, "all_reduce": 0, "bytes": 0}
        self.errors = []
        from . import ALL_GATHER_CB, ALL_REDUCE_CB
        self._ag = ALL_GATHER_CB(self._all_gather)
        self._ar = ALL_REDUCE_CB(self._all_reduce)

    def as_struct(self):
        from . import CommStruct
        return CommStruct(self.rank, self.world, None, self._ag, self._ar)

    def _view(self, ptr, nbytes):
        return torch.as_tensor(_DevMem(ptr, nbytes), device=self.device)

    def _all_gather(self, _user, ptr, bytes_per_rank, stream):
        try:
            total = bytes_per_rank * self.world
            buf = self._view(ptr, total)
            ext = torch.cuda.ExternalStream(int(stream or 0), device=self.device)
            self.calls["all_gather"] += 1
            self.calls["bytes"] += total
            if self.backend == "nccl":
                with torch.cuda.stream(ext):
                    mine = buf[self.rank * bytes_per_rank:(self.rank + 1) * bytes_per_rank]
                    dist.all_gather_into_tensor(buf, mine, group=self.group)
            else:
                ext.synchronize()
                mine = buf[self.rank * bytes_per_rank:(self.rank + 1) * bytes_per_rank].cpu()
                parts = [torch.empty_like(mine) for _ in range(self.world)]
                dist.all_gather(parts, mine, group=self.group)
                with torch.cuda.stream(ext):
                    buf.copy_(torch.cat(parts).to(self.device))
                ext.synchronize()
            return 0
        except Exception as e:  # noqa: BLE001 - must not propagate through the C frame
            self.errors.append(e)
            return 1

    def _all_reduce(self, _user, ptr, count, stream):
        try:
            addr = ctypes.cast(ptr, ctypes.c_void_p).value
            buf = self._view(addr, count * 8).view(torch.int64)   # wrapping 64-bit sum
            ext = torch.cuda.ExternalStream(int(stream or 0), device=self.device)
            self.calls["all_reduce"] += 1
            if self.backend == "nccl":
                with torch.cuda.stream(ext):
                    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            else:
                ext.synchronize()
                host = buf.cpu()
                dist.all_reduce(host, op=dist.ReduceOp.SUM, group=self.group)
                with torch.cuda.stream(ext):
                    buf.copy_(host.to(self.device))
                ext.synchronize()
            return 0
        except Exception as e:  # noqa: BLE001
            self.errors.append(e)
            return 1


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
