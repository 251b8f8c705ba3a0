"""GPU parity of the large-dimension path (hilbert_size > 64: batched level-3 pipeline, qoc_b200/csrc/large.cuh)
against the oracle, unsharded and time-sharded (ranks emulated in one process, as in test_gpu_sharded.py)."""
import numpy as np
import pytest

from tests.problems import Problem
from tests.test_gpu_sharded import emulate

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def rel(a, b):
    return np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel()) / max(np.linalg.norm(np.asarray(b).ravel()), 1e-300)


CASES = [
    # n, slices, K, S, order, complex, F, stiff, ces, step_target
    (65, 9, 2, 2, 2, False, 0, 1.0, 1, False),
    (72, 11, 2, 3, 4, True, 2, 1.0, 2, True),
    (96, 7, 1, 4, 4, False, 3, 8.0, 1, False),
    (130, 5, 2, 2, 4, True, 0, 30.0, 1, False),
    (68, 401, 2, 2, 2, False, 2, 1.0, 7, True),          # > 160 slices: sweeps run on a coarser level of the propagator tree
    (66, 9, 2, 2, 6, False, 2, 1.0, 1, False),           # Magnus M6 through the batched commutator chain
    (100, 6, 2, 3, 6, True, 0, 8.0, 2, True),
    (72, 5, 1, 5, 6, False, 0, 30.0, 1, False),          # S > 4: dense reverse pass, squarings
    (66, 5, 4, 2, 4, True, 1, 1.0, 1, False),            # KR = 8 > 6: Magnus M4 with explicit commutator products
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "n%d_N%d_M%d_x%g" % (c[0], c[1], c[4], c[7]))
def test_large_dim_vs_oracle(case):
    import qoc_b200.standard as std
    from oracle import qoc_oracle as orc
    from qoc_b200.core.plan import SchroedingerPlan
    from qoc_b200.models import MagnusPolicy
    pol = {2: MagnusPolicy.M2, 4: MagnusPolicy.M4, 6: MagnusPolicy.M6}
    n, slices, K, S, order, cc, F, stiff, ces, step_target = case
    p = Problem(n, slices, K, S, order, complex_controls=cc, F=F, seed=n, stiff=stiff, cost_eval_step=ces, step_target=step_target)
    plan = SchroedingerPlan(p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, control_eval_count=p.M,
                            control_count=K, complex_controls=cc, magnus_policy=pol[order], cost_eval_step=ces)
    err, grads, finals = plan.cost_and_grad(p.controls)
    err_f, finals_f = plan.cost(p.controls)
    o_err, o_grad, o_fin = orc.schroedinger_cost_and_grad(p.controls, orc.make_hamiltonian(p.h0, p.drives, cc), p.initial_states,
                                                          p.costs(orc), p.T, p.N, order=order, cost_eval_step=ces)
    assert abs(err - o_err) <= RTOL * abs(o_err) and abs(err_f - o_err) <= RTOL * abs(o_err)
    assert rel(finals, o_fin) < RTOL and rel(finals_f, o_fin) < RTOL
    assert rel(grads, o_grad) < RTOL, rel(grads, o_grad)
    U = plan.propagators()
    assert U.shape == (slices, n, n)
    assert np.abs(U @ U.conj().transpose(0, 2, 1) - np.eye(n)).max() < 1e-11
    plan.close()


@pytest.mark.parametrize("world,slices", [(2, 13), (3, 13), (2, 700)])
def test_large_dim_sharded(world, slices):
    import qoc_b200.standard as std
    from oracle import qoc_oracle as orc
    from qoc_b200.core.sharded import CudaShardEngine
    from qoc_b200.models import MagnusPolicy
    n, K, S, cc = 80 if slices < 100 else 66, 2, 3, True
    p = Problem(n, slices, K, S, 4, complex_controls=cc, F=2, seed=3, stiff=8.0, cost_eval_step=2, step_target=True)
    kw = dict(control_eval_count=p.M, control_count=K, complex_controls=cc, magnus_policy=MagnusPolicy.M4, cost_eval_step=2)
    engines = [CudaShardEngine(r, world, p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, **kw) for r in range(world)]
    cost, g, finals = emulate(engines, p.controls, True)
    cost0, _, finals0 = emulate(engines, p.controls, False)
    grads = g[:, :K] + 1j * g[:, K:]
    o_err, o_grad, o_fin = orc.schroedinger_cost_and_grad(p.controls, orc.make_hamiltonian(p.h0, p.drives, cc), p.initial_states,
                                                          p.costs(orc), p.T, p.N, order=4, cost_eval_step=2)
    assert abs(cost - o_err) <= RTOL * abs(o_err) and abs(cost0 - o_err) <= RTOL * abs(o_err)
    assert rel(grads, o_grad) < RTOL and rel(finals, o_fin) < RTOL and rel(finals0, o_fin) < RTOL
    for e in engines:
        e.close()


def test_large_dim_limits():
    import qoc_b200.standard as std
    from qoc_b200.core.plan import SchroedingerPlan
    from qoc_b200.models import MagnusPolicy
    p = Problem(520, 3, 1, 1, 2)
    with pytest.raises(RuntimeError):
        SchroedingerPlan(p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, control_eval_count=p.M,
                         control_count=1, magnus_policy=MagnusPolicy.M2)
