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
