"""Coset-sharded prove(): `world` ranks, each with its own context, produce the same
proof as a single GPU.  On a one-GPU box the ranks share cuda:0 and the collectives of `tvm_comm` run over
gloo (staged through host memory); on a multi-GPU box the same test uses NCCL with one GPU per rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from test_gpu_prove import synthetic_instance, synthetic_stir_instance

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, params, q):
    import torch.distributed as dist
    import tvm_b200
    from tvm_b200.dist import TorchDistComm
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    one_gpu_each = torch.cuda.device_count() >= world
    dev = rank if one_gpu_each else 0
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if one_gpu_each else "gloo", rank=rank, world_size=world)
    try:
        security, log2_exp, padded_height, seed, ldt = params
        if ldt == "stir":
            st, d, claim, main, mrand, aux_provider, qrand = synthetic_stir_instance(security, padded_height, seed)
        else:
            st, d, claim, main, mrand, aux_provider, qrand = synthetic_instance(security, log2_exp, padded_height, seed)
        args = ((claim.program_digest, claim.input, claim.output), main, mrand, aux_provider, qrand)
        kw = dict(security_level=security, log2_expansion=log2_exp, padded_height=padded_height,
                  ldt_choice=tvm_b200.LDT_STIR if ldt == "stir" else tvm_b200.LDT_FRI)
        b = tvm_b200.Backend(dev)
        single = b.prove(*args, **kw)
        comm = TorchDistComm(f"cuda:{dev}")
        b.set_comm(comm)
        sharded = b.prove(*args, **kw)
        assert not comm.errors, comm.errors
        assert comm.calls["all_gather"] >= 12 and comm.calls["all_reduce"] == 6, comm.calls
        assert np.array_equal(single, sharded), "sharded proof differs from the single-GPU proof"
        b.set_low_memory(1)                      # sharded AND just-in-time LDE
        assert np.array_equal(single, b.prove(*args, **kw)), "sharded low-memory proof differs"
        b.set_low_memory(0)
        b.set_comm(None)
        assert np.array_equal(single, b.prove(*args, **kw))
        q.put((rank, "ok", len(single)))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e), 0))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,params", [(2, (8, 2, 64, 2, "fri")), (4, (32, 2, 256, 3, "fri")), (8, (8, 2, 64, 7, "fri")),
                                          (2, (6, 2, 256, 12, "stir")),
                                          (2, (6, 2, 4096, 5, "stir"))])     # large enough for the sharded STIR leaf hashing
def test_sharded_proof_equals_single_gpu_proof(world, params):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, params, q)) for r in range(world)]
    for p in ps: p.start()
    for p in ps: p.join(600)
    res = sorted(q.get(timeout=10) for _ in range(world))
    assert all(r[1] == "ok" for r in res), res
    assert all(p.exitcode == 0 for p in ps)
