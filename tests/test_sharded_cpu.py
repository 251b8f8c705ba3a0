"""CPU, world_size 2, gloo: the host-side time-slice sharding protocol (qoc_b200/core/sharded.py:
`slice_bounds`, `sharded_evaluate`, `TorchDistComm`) driven with a NumPy shard engine, against the unsharded
NumPy adjoint model and the torch oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qoc_b200.core.sharded import TorchDistComm, sharded_evaluate, slice_bounds
from tests.problems import Problem


def test_slice_bounds():
    assert slice_bounds(2000, 8) == [0, 250, 500, 750, 1000, 1250, 1500, 1750, 2000]
    b = slice_bounds(10, 4)
    assert b[0] == 0 and b[-1] == 10 and all(b[i] < b[i + 1] for i in range(4))
    assert max(b[i + 1] - b[i] for i in range(4)) - min(b[i + 1] - b[i] for i in range(4)) <= 1
    with pytest.raises(ValueError):
        slice_bounds(3, 4)


def _problem():
    from oracle import adjoint_model as am
    p = Problem(5, 13, 2, 2, 4, complex_controls=True, F=2, seed=21, cost_eval_step=2, step_target=True)
    x = np.concatenate([p.controls.real, p.controls.imag], axis=1)
    dd = p.drives.conj().transpose(0, 2, 1)
    a_ops = np.concatenate([p.drives + dd, 1j * (p.drives - dd)])
    cnt = (p.N - 1) // 2
    terms = [am.CostTerm(0, [p.target_states[s, :, 0][None] for s in range(2)], 1.0, cnt, True),
             am.CostTerm(2, [p.forbidden_states[s, :, :, 0] for s in range(2)], 0.7, cnt * 2, True)]
    return p, x, a_ops, terms


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.numpy_shard_engine import NumpyShardEngine
        p, x, a_ops, terms = _problem()
        eng = NumpyShardEngine(rank, world, x.shape, p.h0, a_ops, p.initial_states[:, :, 0], terms, p.T, p.N, 4,
                               cost_eval_step=2)
        eng.upload(x)
        all_p = torch.zeros(world * eng.GM, dtype=torch.float64)
        all_b = torch.zeros(world * eng.VS, dtype=torch.float64)
        res = sharded_evaluate(eng, TorchDistComm(), all_p, all_b, True)
        cost, grad, finals = eng.unpack(res.numpy())
        res0 = sharded_evaluate(eng, TorchDistComm(), all_p, all_b, False)
        cost0, _, finals0 = eng.unpack(res0.numpy())
        q.put((rank, cost, grad, finals, cost0, finals0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_protocol_gloo(world):
    from oracle import adjoint_model as am
    from oracle import qoc_oracle as orc
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p, x, a_ops, terms = _problem()
    c1, g1, f1 = am.cost_and_grad(x, p.h0, a_ops, p.initial_states[:, :, 0], terms, p.T, p.N, 4, cost_eval_step=2)
    o_err, o_grad, o_fin = orc.schroedinger_cost_and_grad(p.controls, orc.make_hamiltonian(p.h0, p.drives, True),
                                                          p.initial_states, p.costs(orc), p.T, p.N, order=4, cost_eval_step=2)
    for rank, cost, grad, finals, cost0, finals0 in outs:              # every rank holds the full result
        assert abs(cost - c1) < 1e-13 and abs(cost0 - c1) < 1e-13
        assert np.abs(grad - g1).max() < 1e-13
        assert np.abs(finals[:, :, 0] - f1).max() < 1e-13 and np.abs(finals0[:, :, 0] - f1).max() < 1e-13
        assert abs(cost - o_err) < 1e-12
        assert np.linalg.norm(grad[:, :2] + 1j * grad[:, 2:] - o_grad) / np.linalg.norm(o_grad) < 1e-11


def _td_problem():
    from oracle import adjoint_model as am
    from qoc_b200.core.plan import extract_hamiltonian_structure
    p = Problem(4, 9, 2, 2, 4, complex_controls=True, seed=5)
    x = np.concatenate([p.controls.real, p.controls.imag], axis=1)
    g0, channels, offset, gain = extract_hamiltonian_structure(p.hamiltonian_td_numpy(), 2, True, p.T, system_eval_count=p.N,
                                                               magnus_order=4)
    terms = [am.CostTerm(0, [p.target_states[s, :, 0][None] for s in range(2)], 1.0, 1, False)]
    return p, x, g0, channels, (offset, gain), terms


def _td_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.numpy_shard_engine import NumpyShardEngine
        p, x, g0, channels, node_map, terms = _td_problem()
        eng = NumpyShardEngine(rank, world, x.shape, g0, channels, p.initial_states[:, :, 0], terms, p.T, p.N, 4, node_map=node_map)
        eng.upload(x)
        all_p = torch.zeros(world * eng.GM, dtype=torch.float64)
        all_b = torch.zeros(world * eng.VS, dtype=torch.float64)
        res = sharded_evaluate(eng, TorchDistComm(), all_p, all_b, True)
        q.put((rank,) + eng.unpack(res.numpy()))
    finally:
        dist.destroy_process_group()


def test_sharded_time_dependent_gloo():
    """world-size 2: the operator-channel expansion of a time-dependent hamiltonian (host logic of qocb_set_node_map) through the
    sharding protocol, against the oracle calling the torch version of the same callable at every Magnus node"""
    from oracle import qoc_oracle as orc
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_td_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = _td_problem()[0]
    o_err, o_grad, o_fin = orc.schroedinger_cost_and_grad(p.controls, p.hamiltonian_td_torch(), p.initial_states,
                                                          [orc.TargetStateInfidelity(p.target_states)], p.T, p.N, order=4)
    for rank, cost, grad, finals in outs:
        assert abs(cost - o_err) < 1e-12
        assert np.linalg.norm(grad[:, :2] + 1j * grad[:, 2:] - o_grad) / np.linalg.norm(o_grad) < 1e-10
        assert np.abs(finals - o_fin).max() < 1e-12
