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


def _worker_tables(rank, world, port, params, q):
    """device-table mode on several ranks: tvm_prove_tables and tvm_prove_aet (every rank holds the whole main trace, runs
    MasterMainTable::extend itself and interpolates its own block of columns) against the single-GPU proof"""
    import torch.distributed as dist
    import tvm_b200
    from tvm_b200.dist import TorchDistComm
    from oracle import tracegen as tg
    import test_vm_programs as tvp
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    one_gpu_each = torch.cuda.device_count() >= world
    dev = rank if one_gpu_each else 0
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl" if one_gpu_each else "gloo", rank=rank, world_size=world)
    try:
        name, ldt = params
        src, inp, ram = tvp._workload(name)
        inst = tvp.program_instance(src, inp, 8, ldt, ram=ram)
        words = tg.assemble(src)
        aet = tg.aet_arrays(words, tg.execute(words, inp, (), ram))
        claim = inst["claim"]
        c = (claim.program_digest, claim.input, claim.output)
        args = (inst["main_rand"], inst["aux_rand"], inst["randomizer_column"], inst["quot_rand"])
        kw = dict(security_level=8, log2_expansion=2, padded_height=inst["padded_height"],
                  ldt_choice=tvm_b200.LDT_FRI if ldt == "fri" else tvm_b200.LDT_STIR)
        b = tvm_b200.Backend(dev)
        single = b.prove_tables(c, inst["main"], *args, **kw)
        comm = TorchDistComm(f"cuda:{dev}")
        b.set_comm(comm)
        assert np.array_equal(single, b.prove_tables(c, inst["main"], *args, **kw)), "sharded tvm_prove_tables differs"
        assert np.array_equal(single, b.prove_aet(c, aet, *args, **kw)), "sharded tvm_prove_aet differs"
        assert not comm.errors, comm.errors
        b.set_low_memory(1)
        assert np.array_equal(single, b.prove_aet(c, aet, *args, **kw)), "sharded low-memory tvm_prove_aet differs"
        b.set_low_memory(0)
        b.set_comm(None)
        q.put((rank, "ok", len(single)))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e), 0))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,params", [(2, ("fib_100", "fri")), (4, ("verifier_3", "stir"))])
def test_sharded_device_table_modes_equal_the_single_gpu_proof(world, params):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker_tables, args=(r, world, port, params, q)) for r in range(world)]
    for p in ps: p.start()
    for p in ps: p.join(900)
    res = sorted(q.get(timeout=10) for _ in range(world))
    assert all(r[1] == "ok" for r in res), res
    assert all(p.exitcode == 0 for p in ps)
