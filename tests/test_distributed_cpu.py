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
    loss = _task(m, rank, 1.0 / world).clone()      # rank r processes task r (1 task per device, base_adaptor.py:128)
    dist.all_reduce(m.g_outer_full)                 # the single collective of the path: gradient + the 6 losses on its tail
    synced = m.g_outer_full[m.layout.total:m.layout.total + 6].clone()
    gathered = [torch.zeros(6) for _ in range(world)]
    dist.all_gather(gathered, loss)                 # (test only) every rank's own losses
    assert torch.allclose(synced, torch.stack(gathered).mean(0), rtol=1e-6, atol=1e-7), "tail != mean of the ranks' losses"
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


# ---- the allreduce issued INSIDE the task step (adapted region under the last Hessian-vector pass' encoder walk) ----------
def _overlap_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    sup, qry = O.synth_task(task=rank, shots=2, queries=2, L=5, T=12, ragged=True)
    bs = batch_from_tuple(sup, "cpu")
    bq = batch_from_tuple(qry, "cpu", spk_ids=sup[2], average_spk=True)
    a = _engine()                                   # reference: task step, THEN one allreduce of the whole buffer
    a.task_step(bs, bq, 1, False, accumulate_scale=1.0 / world)
    dist.all_reduce(a.g_outer_full)
    b = _engine()                                   # the step reduces its own gradient, in two pieces
    calls = []

    def reduce(t):
        calls.append(t.numel())
        dist.all_reduce(t)

    b.task_step(bs, bq, 1, False, accumulate_scale=1.0 / world, reduce=reduce)
    n, a0 = b.layout.total, b.layout.adapt_begin
    assert calls == [n + 8 - a0, a0], calls          # adapted region (+ the loss tail) first, the encoder third last
    assert torch.equal(a.g_outer_full, b.g_outer_full), "in-step reduction differs from reducing after the step"
    c = _engine()                                   # first order: nothing to overlap, one reduction at the end of the step
    calls.clear()
    c.task_step(bs, bq, 1, True, accumulate_scale=1.0 / world, reduce=reduce)
    assert calls == [n + 8], calls
    if rank == 0:
        q.put(True)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_in_step_allreduce_equals_allreduce_after_the_step():
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_overlap_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    assert q.get(timeout=500) is True
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0


# ---- iMAML: clip on each rank, THEN mean-reduce, then a plain Adam step (lightning/systems/imaml.py:123-147) ------------------------
def _imaml_system():
    import copy

    from meta_tts_b200.imaml import IMAMLSystem
    from meta_tts_b200.systems import DEFAULT_ALGORITHM_CONFIG, DEFAULT_TRAIN_CONFIG
    algo = copy.deepcopy(DEFAULT_ALGORITHM_CONFIG)
    algo["adapt"]["train"]["steps"] = 2
    algo["adapt"]["test"]["steps"] = 2
    algo["adapt"]["imaml"] = {"batch_size": 2, "reg_param": 1.0, "K": 2, "stochastic": False}
    s = IMAMLSystem(None, CFG, DEFAULT_TRAIN_CONFIG, algo, n_speaker=16, device="cpu", backend=RefOps(split=3), dropout=False)
    s.load_state_dict({k: v.detach().clone() for k, v in O.init_params(seed=0, model_config=CFG).items()})
    return s


def _imaml_batch(t):
    sup, qry = O.synth_task(task=t + 3, shots=2, queries=2, L=5, T=12, ragged=True)
    return [([sup], [qry])]


def _imaml_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    s = _imaml_system()
    torch.manual_seed(7)
    s.training_step(_imaml_batch(rank), 0)
    if rank == 0:
        q.put(s.maml.theta.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_imaml_two_ranks_clip_then_mean():
    from meta_tts_b200.imaml import Task, imaml_adapt, imaml_hypergradient
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_imaml_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    th2 = torch.from_numpy(q.get(timeout=500))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    torch.set_num_threads(4)
    # one process: each task's hypergradient from the meta parameters, clipped by ITS OWN norm, averaged, plain Adam
    acc = None
    for t in range(world):
        s1 = _imaml_system()
        torch.manual_seed(7)
        b = _imaml_batch(t)
        task = Task(b[0][0][0], b[0][1][0], batch_size=2)
        imaml_adapt(s1.maml, task, 2, 1.0)
        imaml_hypergradient(s1.maml, task, b[0][0][0], b[0][1][0], 2, 1.0, 2, False)
        g = s1.maml.g_task.clone()
        coef = min(1.0, 1.0 / (float(g.norm()) + 1e-6))
        acc = coef * g / world if acc is None else acc + coef * g / world
    ref = _imaml_system()
    ref.maml.g_outer.copy_(acc)
    ref.maml.outer_update(1.0, 0.0)
    assert (ref.maml.theta - th2).abs().max().item() < 1e-6
    assert (th2 - _imaml_system().maml.theta).abs().max().item() > 1e-7
