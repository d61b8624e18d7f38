"""world_size-2 gloo test of the multi-GPU logic: N ranks x 1 task each + one allreduce of
the flat outer-gradient buffer + identical clip/Adam on every rank  ==  1 rank accumulating the same
N tasks.  Runs on CPU through RefOps (the collective and host logic are what is under test)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from meta_tts_b200.maml import MamlEngine, batch_from_tuple
from oracle import fs2_oracle as O
from oracle.ops_reference import RefOps

CFG = O.small_model_config(1, 1)


def _engine():
    P = O.init_params(seed=0, model_config=CFG)
    m = MamlEngine(RefOps(split=3), CFG, 16, O.ADAPT_MODULES, 0.001, 1)
    m.load_state_dict({k: v.detach().clone() for k, v in P.items()})
    return m


def _task(m, t, scale):
    sup, qry = O.synth_task(task=t, shots=2, queries=2, L=5, T=12, ragged=True)
    bs = batch_from_tuple(sup, "cpu")
    bq = batch_from_tuple(qry, "cpu", spk_ids=sup[2], average_spk=True)
    return m.task_step(bs, bq, 1, True, accumulate_scale=scale)[0]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    m = _engine()
    _task(m, rank, 1.0 / world)                     # rank r processes task r (1 task per device, base_adaptor.py:128)
    dist.all_reduce(m.g_outer)                      # the single collective of the path
    m.outer_update(1.0, 1.0)
    if rank == 0:
        q.put((m.g_outer.numpy().copy(), m.theta.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_ranks_equal_one_rank_accumulating():
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    g2, th2 = (torch.from_numpy(a) for a in q.get(timeout=500))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    torch.set_num_threads(4)
    m = _engine()
    for t in range(world):
        _task(m, t, 1.0 / world)
    g1 = m.g_outer.clone()
    m.outer_update(1.0, 1.0)
    assert ((g1 - g2).norm() / g1.norm()).item() < 1e-6
    assert ((m.theta - th2).abs().max()).item() < 1e-6
    assert (m.theta - _engine().theta).abs().max().item() > 1e-7      # the update moved the weights (lr(step 0) = 2.5e-7)
