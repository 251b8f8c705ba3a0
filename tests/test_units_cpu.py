"""CPU, world_size 2 / 3, gloo: sharding of independent units across ranks (`qoc_b200.core.sharded.units_evaluate`, SURVEY.md
section 8e) - ensemble members (cfg5) and initial states - driven with NumPy engines, against the unsharded NumPy adjoint
model and the torch oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qoc_b200.core.sharded import TorchDistComm, units_evaluate
from tests.problems import Problem


def _problem(coherent):
    from oracle import adjoint_model as am
    p = Problem(4, 9, 2, 5, 4, complex_controls=False, F=2, seed=31, cost_eval_step=2, step_target=True, neglect_phase=not coherent)
    cnt = (p.N - 1) // 2
    terms = [am.CostTerm(0 if coherent else 1, [p.target_states[s, :, 0][None] for s in range(5)], 1.0, cnt, True),
             am.CostTerm(2, [p.forbidden_states[s, :, :, 0] for s in range(5)], 0.7, cnt * 5, True),
             am.CostTerm(0, [p.target_states[s, :, 0][None] for s in range(5)], 0.3, 1, False)]
    return p, terms


def _oracle_costs(p, orc, coherent):
    return p.costs(orc) + [orc.TargetStateInfidelity(p.target_states, cost_multiplier=0.3)]


def _worker(rank, world, port, q, mode, coherent):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.numpy_unit_engine import NumpyMemberEngine, NumpyStateEngine
        p, terms = _problem(coherent)
        if mode == "members":
            z = np.diag(np.arange(4) - 1.5).astype(complex)
            drifts = [p.h0 + d * z for d in np.random.default_rng(2).normal(0, 0.1, 5)]
            eng = NumpyMemberEngine(rank, world, p.controls, drifts, p.drives, p.initial_states[:, :, 0], terms, p.T, p.N, 4, 2)
        else:
            eng = NumpyStateEngine(rank, world, p.controls, p.h0, p.drives, p.initial_states[:, :, 0], terms, p.T, p.N, 4, 2)
        res = units_evaluate(eng, TorchDistComm(), True).numpy().copy()
        res0 = units_evaluate(eng, TorchDistComm(), False).numpy().copy()
        q.put((rank, res, res0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode,world,coherent", [("members", 2, True), ("members", 3, False), ("states", 2, True),
                                                ("states", 3, True), ("states", 2, False)])
def test_units_protocol_gloo(mode, world, coherent):
    from oracle import adjoint_model as am
    from oracle import qoc_oracle as orc
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, mode, coherent)) for r in range(world)]
    for pr in procs:
        pr.start()
    outs = [q.get(timeout=180) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p, terms = _problem(coherent)
    if mode == "members":
        z = np.diag(np.arange(4) - 1.5).astype(complex)
        drifts = [p.h0 + d * z for d in np.random.default_rng(2).normal(0, 0.1, 5)]
    else:
        drifts = [p.h0]
    o_err, o_grad = 0.0, 0.0
    for h0 in drifts:
        v, g, _ = orc.schroedinger_cost_and_grad(p.controls, orc.make_hamiltonian(h0, p.drives, False), p.initial_states,
                                                 _oracle_costs(p, orc, coherent), p.T, p.N, order=4, cost_eval_step=2)
        o_err, o_grad = o_err + v / len(drifts), o_grad + g / len(drifts)
    for rank, res, res0 in outs:                                   # every rank holds the full result
        assert abs(res[-1] - o_err) < 1e-12 and abs(res0[-1] - o_err) < 1e-12, (res[-1], res0[-1], o_err)
        assert np.linalg.norm(res[:-1].reshape(o_grad.shape) - o_grad) / np.linalg.norm(o_grad) < 1e-11
