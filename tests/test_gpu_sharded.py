"""GPU: the CUDA time-slice sharding phases (qocb_shard_*) against the unsharded CUDA path and the oracle.
On one GPU the ranks are emulated in ONE process: G shard engines run each phase in turn and the collectives are
plain concatenations / sums (no kernels that wait on one another).  With >= 2 visible GPUs the same protocol also
runs over NCCL in test_nccl_two_ranks."""
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


def emulate(engines, controls, with_grad):
    import torch
    for e in engines:
        e.upload(controls)
    ps = []
    for e in engines:
        ps.append(e.forward_local(with_grad).clone()); e.stream.synchronize()
    torch.cuda.synchronize()
    all_p = torch.cat([e.P for e in engines])
    torch.cuda.synchronize()
    for e in engines:
        e.forward_finish(all_p); e.stream.synchronize()
    if with_grad:
        for e in engines:
            e.backward_particular(); e.stream.synchronize()
        all_b = torch.cat([e.b for e in engines])
        torch.cuda.synchronize()
        for e in engines:
            e.backward_finish(all_p, all_b); e.stream.synchronize()
    tot = None
    for e in engines:
        r = e.pack_result(with_grad); e.stream.synchronize()
        tot = r.clone() if tot is None else tot + r
    torch.cuda.synchronize()
    return engines[0].unpack(tot.cpu().numpy())


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("case", [
    (8, 41, 2, 3, 4, True, 2, 1.0, 3, True),
    (16, 64, 2, 2, 6, False, 0, 8.0, 1, False),
    (64, 37, 4, 4, 4, False, 3, 1.0, 1, False),
    (5, 24, 1, 1, 2, True, 1, 1.0, 2, True),
])
def test_sharded_phases_vs_unsharded_and_oracle(world, case):
    import qoc_b200.standard as std
    from oracle import qoc_oracle as orc
    from qoc_b200.core.plan import SchroedingerPlan
    from qoc_b200.core.sharded import CudaShardEngine
    from qoc_b200.models import MagnusPolicy
    pol = {2: MagnusPolicy.M2, 4: MagnusPolicy.M4, 6: MagnusPolicy.M6}
    n, slices, K, S, order, cc, F, stiff, ces, step_target = case
    p = Problem(n, slices, K, S, order, complex_controls=cc, F=F, seed=7, stiff=stiff, cost_eval_step=ces,
                step_target=step_target)
    kw = dict(control_eval_count=p.M, control_count=K, complex_controls=cc, magnus_policy=pol[order], cost_eval_step=ces)
    engines = [CudaShardEngine(r, world, p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, **kw)
               for r in range(world)]
    cost, g, finals = emulate(engines, p.controls, True)
    cost0, _, finals0 = emulate(engines, p.controls, False)
    grads = g[:, :K] + 1j * g[:, K:] if cc else g
    plan = SchroedingerPlan(p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, **kw)
    u_err, u_grad, u_fin = plan.cost_and_grad(p.controls)
    o_err, o_grad, o_fin = orc.schroedinger_cost_and_grad(p.controls, orc.make_hamiltonian(p.h0, p.drives, cc),
                                                          p.initial_states, p.costs(orc), p.T, p.N, order=order,
                                                          cost_eval_step=ces)
    assert abs(cost - o_err) <= 1e-10 * abs(o_err) and abs(cost0 - o_err) <= 1e-10 * abs(o_err)
    assert rel(grads, o_grad) < 1e-10
    assert rel(finals, o_fin) < 1e-10 and rel(finals0, o_fin) < 1e-10
    assert abs(cost - u_err) <= 1e-12 * abs(u_err) and rel(grads, u_grad) < 1e-11
    for e in engines:
        e.close()
    plan.close()


def test_nccl_two_ranks():
    """real NCCL run of bench.py's sharded path when the box has >= 2 GPUs (skipped on 1-GPU boxes)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
                          "--gpus", "2", "--steps", "3", "--warmup", "3", "--workload", "n16_2000_M4", "--no-cpu-baseline"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["parity"]["ok"] is True, line["parity"]            # every N > 1 bench line carries its parity block
    assert line["parity"]["grad_rel_err"] < 1e-10 and line["parity"]["sharded_vs_unsharded"]["grad_rel"] < 1e-10
    assert set(line["stage_ms"]["phases_max_over_ranks"]) >= {"forward_local", "gather_P", "backward_finish"}
