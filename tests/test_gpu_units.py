"""GPU: sharding of independent units (qoc_b200/core/sharded.py: `EnsembleShardedPlan`, `units_evaluate`).  On a one-GPU box the
NCCL group has a single rank - the device-side pack / weight / all-reduce path is still the one that runs; with >= 2 GPUs
bench.py's ensemble workload runs over two ranks and carries its parity block."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from tests.problems import Problem

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    return np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel()) / max(np.linalg.norm(np.asarray(b).ravel()), 1e-300)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def test_ensemble_sharded_plan_single_rank_group():
    import torch
    import torch.distributed as dist
    import qoc_b200.standard as std
    from oracle import qoc_oracle as orc
    from qoc_b200.core.plan import SchroedingerPlan
    from qoc_b200.core.sharded import EnsembleShardedPlan
    from qoc_b200.models import MagnusPolicy
    own_group = not dist.is_initialized()
    if own_group:
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % _free_port(), rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
    try:
        p = Problem(16, 12, 2, 3, 4, complex_controls=True, F=2, seed=9, cost_eval_step=2, step_target=True)
        E = 5
        z = np.diag(np.linspace(-1, 1, 16)).astype(complex) * 0.3
        drifts = np.stack([p.h0 + d * z for d in np.random.default_rng(1).normal(0, 1.0, E)])
        costs = p.costs(std) + [std.ControlNorm(2, p.M, cost_multiplier=0.05)]
        kw = dict(control_eval_count=p.M, control_count=2, complex_controls=True, magnus_policy=MagnusPolicy.M4, cost_eval_step=2)
        plan = EnsembleShardedPlan(p.hamiltonian_numpy(), p.initial_states, costs, p.T, p.N, drifts, **kw)
        err, grads, finals = plan.cost_and_grad(p.controls)
        err0, finals0 = plan.cost(p.controls)
        total, stages = plan.time_resident(True, warmup=1, iters=2, flush_l2=True)
        plan.close()
        ref = SchroedingerPlan(p.hamiltonian_numpy(), p.initial_states, costs, p.T, p.N, ensemble_drifts=drifts, **kw)
        r_err, r_grads, r_fin = ref.cost_and_grad(p.controls)
        ref.close()
        assert abs(err - r_err) < 1e-13 and abs(err0 - r_err) < 1e-13 and rel(grads, r_grads) < 1e-12
        assert finals.shape == (E, 3, 16, 1) and np.array_equal(finals, r_fin) and np.array_equal(finals0, r_fin)
        assert total > 0 and stages[0] > 0
        o_err, o_grad = 0.0, 0.0
        ocosts = p.costs(orc) + [orc.ControlNorm(2, p.M, cost_multiplier=0.05)]
        for e in range(E):
            v, g, _ = orc.schroedinger_cost_and_grad(p.controls, orc.make_hamiltonian(drifts[e], p.drives, True), p.initial_states,
                                                     ocosts, p.T, p.N, order=4, cost_eval_step=2)
            o_err, o_grad = o_err + v / E, o_grad + g / E
        assert abs(err - o_err) <= 1e-10 * abs(o_err) and rel(grads, o_grad) < 1e-10
    finally:
        if own_group:
            dist.destroy_process_group()


def test_nccl_two_ranks_ensemble():
    """real NCCL run of bench.py's member-sharded ensemble path when the box has >= 2 GPUs (skipped on 1-GPU boxes)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"),
                          "--gpus", "2", "--steps", "2", "--warmup", "3", "--workload", "cfg5_n32_100_S64_E16_M2",
                          "--no-cpu-baseline"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["parity"]["grad_rel_err"] < 1e-10 and line["parity"]["cost_rel_err"] < 1e-10


def _emulate_state_shards(engines, controls, with_grad):
    """ranks emulated in one process: each engine's phases in turn, the all-reduces are plain sums"""
    import torch
    from qoc_b200.core.sharded import units_evaluate

    class SumComm(object):
        """stand-in for the collective: records the buffers of one phase; `finish` replaces them by their sum"""
        def __init__(self):
            self.pending = []

        def all_reduce_sum(self, t):
            self.pending.append(t)

    # phase-by-phase over the engines (units_evaluate interleaves per rank, so drive the phases by hand)
    for e in engines:
        e.plan.upload(controls)
    cohs = []
    for e in engines:
        with torch.cuda.stream(e.stream):
            cohs.append(e.forward(with_grad))
        e.stream.synchronize()
    if cohs[0] is not None and cohs[0].numel():
        tot = sum(c.clone() for c in cohs)
        for c in cohs:
            c.copy_(tot)
    torch.cuda.synchronize()
    outs = []
    for e in engines:
        with torch.cuda.stream(e.stream):
            if with_grad:
                e.backward(e.coh)
            outs.append(e.pack(with_grad).clone())
        e.stream.synchronize()
    torch.cuda.synchronize()
    return sum(outs).cpu().numpy()


@pytest.mark.parametrize("case", [
    # n, slices, K, S, order, complex, F, ces, step_target, neglect_phase, world
    (8, 14, 2, 5, 4, True, 2, 2, True, False, 2),          # coherent step cost + forbid: overlap sums at every cost step
    (16, 9, 2, 6, 2, False, 0, 1, False, False, 3),        # coherent final cost only
    (32, 7, 2, 64, 2, False, 0, 1, False, True, 8),        # cfg5's state count, incoherent cost (no coupling at all)
    (64, 6, 4, 7, 4, False, 3, 1, False, False, 4),        # n = 64: rank-S reverse pass on 1 - 2 states per rank
    (72, 5, 2, 5, 4, True, 2, 2, True, False, 2),          # large-dimension path
], ids=lambda c: "n%d_S%d_w%d" % (c[0], c[3], c[10]))
def test_state_sharding_phases_vs_unsharded_and_oracle(case):
    import qoc_b200.standard as std
    from oracle import qoc_oracle as orc
    from qoc_b200.core.plan import SchroedingerPlan
    from qoc_b200.core.sharded import CudaStateEngine, slice_bounds
    from qoc_b200.models import MagnusPolicy
    pol = {2: MagnusPolicy.M2, 4: MagnusPolicy.M4, 6: MagnusPolicy.M6}
    n, slices, K, S, order, cc, F, ces, step_target, neglect, world = case
    p = Problem(n, slices, K, S, order, complex_controls=cc, F=F, seed=13, cost_eval_step=ces, step_target=step_target,
                neglect_phase=neglect)
    costs = p.costs(std) + ([std.TargetStateInfidelity(p.target_states, cost_multiplier=0.3)] if step_target else [])
    ocosts = p.costs(orc) + ([orc.TargetStateInfidelity(p.target_states, cost_multiplier=0.3)] if step_target else [])
    kw = dict(control_eval_count=p.M, control_count=K, complex_controls=cc, magnus_policy=pol[order], cost_eval_step=ces)
    b = slice_bounds(S, world)
    plans = [SchroedingerPlan(p.hamiltonian_numpy(), p.initial_states, costs, p.T, p.N, state_slice=(b[g], b[g + 1]), **kw)
             for g in range(world)]
    engines = [CudaStateEngine(pl, 0) for pl in plans]
    res = _emulate_state_shards(engines, p.controls, True)
    res0 = _emulate_state_shards(engines, p.controls, False)
    finals = np.concatenate([pl.final_states() for pl in plans])
    for e, pl in zip(engines, plans):
        e.close()
        pl.close()
    g = res[:-1].reshape(p.M, -1)
    grads = g[:, :K] + 1j * g[:, K:] if cc else g
    ref = SchroedingerPlan(p.hamiltonian_numpy(), p.initial_states, costs, p.T, p.N, **kw)
    u_err, u_grad, u_fin = ref.cost_and_grad(p.controls)
    ref.close()
    o_err, o_grad, o_fin = orc.schroedinger_cost_and_grad(p.controls, orc.make_hamiltonian(p.h0, p.drives, cc), p.initial_states,
                                                          ocosts, p.T, p.N, order=order, cost_eval_step=ces)
    assert abs(res[-1] - o_err) <= 1e-10 * abs(o_err) and abs(res0[-1] - o_err) <= 1e-10 * abs(o_err), (res[-1], res0[-1], o_err)
    assert rel(grads, o_grad) < 1e-10 and rel(finals, o_fin) < 1e-10
    assert abs(res[-1] - u_err) <= 1e-12 * abs(u_err) and rel(grads, u_grad) < 1e-11


def test_state_sharded_plan_single_rank_group():
    import torch
    import torch.distributed as dist
    import qoc_b200.standard as std
    from qoc_b200.core.plan import SchroedingerPlan
    from qoc_b200.core.sharded import StateShardedPlan
    from qoc_b200.models import MagnusPolicy
    own_group = not dist.is_initialized()
    if own_group:
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % _free_port(), rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
    try:
        p = Problem(16, 12, 2, 5, 4, complex_controls=True, F=2, seed=9, cost_eval_step=2, step_target=True)
        costs = p.costs(std) + [std.ControlNorm(2, p.M, cost_multiplier=0.05)]
        kw = dict(control_eval_count=p.M, control_count=2, complex_controls=True, magnus_policy=MagnusPolicy.M4, cost_eval_step=2)
        plan = StateShardedPlan(p.hamiltonian_numpy(), p.initial_states, costs, p.T, p.N, **kw)
        err, grads, finals = plan.cost_and_grad(p.controls)
        err0, finals0 = plan.cost(p.controls)
        total, _ = plan.time_resident(True, warmup=1, iters=2)
        plan.close()
        ref = SchroedingerPlan(p.hamiltonian_numpy(), p.initial_states, costs, p.T, p.N, **kw)
        r_err, r_grads, r_fin = ref.cost_and_grad(p.controls)
        ref.close()
        assert abs(err - r_err) < 1e-13 and abs(err0 - r_err) < 1e-13 and rel(grads, r_grads) < 1e-12
        assert finals.shape == r_fin.shape and np.array_equal(finals, r_fin) and np.array_equal(finals0, r_fin) and total > 0
    finally:
        if own_group:
            dist.destroy_process_group()
