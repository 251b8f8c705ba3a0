"""GPU parity at the sizes BASELINE.json names, against the CPU oracle at 1e-10:

  * the FULL headline evaluation, n = 64 x 2000 slices, M4, K = 4, S = 4 (the oracle needs a few seconds);
  * the FULL cfg3 evaluation, n = 60 (3 x 20) x 2000 slices, complex K = 2, TargetStateInfidelity + ForbidStates;
  * cfg4's dimension, n = 256 (large-dimension path), unsharded and time-sharded (ranks emulated in one process);
  * cfg5's shape, n = 32 with S = 64 states (dense reverse pass) and E = 8 ensemble members;
  * LU pivot rule: generators whose Pade denominators have exactly tied / nearly tied pivot candidates."""
import numpy as np
import pytest

from tests.problems import Problem
from tests.test_gpu_sharded import emulate

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def rel(a, b):
    return np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel()) / max(np.linalg.norm(np.asarray(b).ravel()), 1e-300)


def product():
    import qoc_b200.standard as std
    from qoc_b200.core.plan import SchroedingerPlan
    from qoc_b200.models import MagnusPolicy
    return std, SchroedingerPlan, {2: MagnusPolicy.M2, 4: MagnusPolicy.M4, 6: MagnusPolicy.M6}


def oracle(p, drift=None):
    from oracle import qoc_oracle as orc
    h0 = p.h0 if drift is None else drift
    return orc.schroedinger_cost_and_grad(p.controls, orc.make_hamiltonian(h0, p.drives, p.complex_controls), p.initial_states,
                                          p.costs(orc), p.T, p.N, order=p.order, cost_eval_step=p.cost_eval_step)


@pytest.mark.parametrize("shape", [
    (64, 2000, 4, 4, 4, False, 0),            # bench.py default workload n64_2000_M4
    (60, 2000, 2, 4, 4, True, 6),             # cfg3_n60_2000_M4
], ids=["n64_2000_M4", "cfg3_n60_2000_M4"])
def test_full_length_evaluation_vs_oracle(shape):
    n, slices, K, S, order, cc, F = shape
    std, Plan, pol = product()
    p = Problem(n, slices, K, S, order, complex_controls=cc, F=F, seed=0)       # = bench.make_problem(name)
    plan = Plan(p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, control_eval_count=p.M, control_count=K,
                complex_controls=cc, magnus_policy=pol[order])
    err, grads, finals = plan.cost_and_grad(p.controls)
    err_f, finals_f = plan.cost(p.controls)
    plan.close()
    o_err, o_grad, o_fin = oracle(p)
    assert abs(err - o_err) <= RTOL * abs(o_err) and abs(err_f - o_err) <= RTOL * abs(o_err), (err, err_f, o_err)
    assert rel(finals, o_fin) < RTOL and rel(finals_f, o_fin) < RTOL
    assert rel(grads, o_grad) < RTOL, rel(grads, o_grad)


def test_n256_vs_oracle():
    """cfg4's dimension on the large-dimension path: 8 slices, M4, K = 4, S = 4."""
    std, Plan, pol = product()
    p = Problem(256, 8, 4, 4, 4, complex_controls=False, F=0, seed=0)
    plan = Plan(p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, control_eval_count=p.M, control_count=4,
                magnus_policy=pol[4])
    err, grads, finals = plan.cost_and_grad(p.controls)
    plan.close()
    o_err, o_grad, o_fin = oracle(p)
    assert abs(err - o_err) <= RTOL * abs(o_err)
    assert rel(finals, o_fin) < RTOL
    assert rel(grads, o_grad) < RTOL, rel(grads, o_grad)


def test_n256_stiff_forbid_vs_oracle():
    """n = 256 with squarings (s > 0), complex controls and a step cost: the dense reverse pass of the large path."""
    std, Plan, pol = product()
    p = Problem(256, 5, 2, 3, 4, complex_controls=True, F=2, seed=1, stiff=8.0, cost_eval_step=2, step_target=True)
    plan = Plan(p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, control_eval_count=p.M, control_count=2,
                complex_controls=True, magnus_policy=pol[4], cost_eval_step=2)
    err, grads, finals = plan.cost_and_grad(p.controls)
    plan.close()
    o_err, o_grad, o_fin = oracle(p)
    assert abs(err - o_err) <= RTOL * abs(o_err)
    assert rel(finals, o_fin) < RTOL
    assert rel(grads, o_grad) < RTOL, rel(grads, o_grad)


def test_n256_time_sharded_vs_oracle():
    """cfg4 is the shape the time-slice sharding exists for: 2 emulated ranks x 5 slices at n = 256."""
    import qoc_b200.standard as std
    from qoc_b200.core.sharded import CudaShardEngine
    from qoc_b200.models import MagnusPolicy
    p = Problem(256, 10, 4, 4, 4, complex_controls=False, F=0, seed=2)
    kw = dict(control_eval_count=p.M, control_count=4, magnus_policy=MagnusPolicy.M4)
    engines = [CudaShardEngine(r, 2, p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, **kw) for r in range(2)]
    cost, g, finals = emulate(engines, p.controls, True)
    for e in engines:
        e.close()
    o_err, o_grad, o_fin = oracle(p)
    assert abs(cost - o_err) <= RTOL * abs(o_err)
    assert rel(g, o_grad) < RTOL and rel(finals, o_fin) < RTOL


def test_cfg5_shape_vs_oracle():
    """cfg5: n = 32, S = 64 initial states (forces the dense reverse pass: the cotangent has full rank), E = 8 ensemble
    members that differ in the drift, M2.  Oracle = loop over the members, mean."""
    std, Plan, pol = product()
    E = 8
    p = Problem(32, 20, 2, 64, 2, complex_controls=False, F=0, seed=5)     # 64 states = two Haar frames (tests/problems.py)
    assert p.initial_states.shape == (64, 32, 1) and p.target_states.shape == (64, 32, 1)
    rng = np.random.default_rng(12345)
    z = np.diag(np.linspace(-1, 1, 32)).astype(np.complex128) * np.abs(p.h0).max()
    drifts = np.stack([p.h0 + d * z for d in rng.normal(0, 0.1, E)])
    plan = Plan(p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, control_eval_count=p.M, control_count=2,
                magnus_policy=pol[2], ensemble_drifts=drifts)
    err, grads, finals = plan.cost_and_grad(p.controls)
    plan.close()
    o_err, o_grad = 0.0, 0.0
    o_fin = []
    for e in range(E):
        v, g, f = oracle(p, drifts[e])
        o_err += v / E
        o_grad = o_grad + g / E
        o_fin.append(f)
    assert abs(err - o_err) <= RTOL * abs(o_err)
    assert rel(grads, o_grad) < RTOL, rel(grads, o_grad)
    assert finals.shape == (E, 64, 32, 1) and rel(finals, np.stack(o_fin)) < RTOL


@pytest.mark.parametrize("step_target", [False, True], ids=["final_cost", "step_costs"])
def test_single_member_many_states_three_levels(step_target):
    """one member with S = 64 states at n = 32 and enough slices for the three-level boundary scheme: the boundary, fill-in and
    sweep kernels run with eight state groups and more than 48 KB of dynamic shared memory (the parity line of the
    member-sharded cfg5 bench has exactly this shape per rank)."""
    std, Plan, pol = product()
    p = Problem(32, 180, 2, 64, 2, complex_controls=False, F=0, seed=9, step_target=step_target, cost_eval_step=4)
    plan = Plan(p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, control_eval_count=p.M, control_count=2,
                magnus_policy=pol[2], cost_eval_step=4)
    err, grads, finals = plan.cost_and_grad(p.controls)
    plan.close()
    o_err, o_grad, o_fin = oracle(p)
    assert abs(err - o_err) <= RTOL * abs(o_err)
    assert rel(grads, o_grad) < RTOL, rel(grads, o_grad)
    assert rel(finals, o_fin) < RTOL


def _hadamard(n):
    h = np.array([[1.0]])
    while h.shape[0] < n:
        h = np.block([[h, h], [h, -h]])
    return h


@pytest.mark.parametrize("n", [4, 8, 16, 32, 64])
@pytest.mark.parametrize("kind", ["hadamard", "degenerate_drift", "pm1_skew", "permutation"])
def test_expm_pivot_ties(n, kind):
    """The blocked LU picks its pivot from a key that treats magnitudes within 2^-14 as ties (tile.cuh), LAPACK's izamax
    takes the first maximum of |re| + |im|: generators with many equal-magnitude entries (Hadamard / +-1 matrices, exactly
    degenerate drifts, permutations) make the Pade denominator's pivot columns tie exactly or nearly.  Either choice is an
    admissible pivot; the results must agree with the oracle (numpy.linalg.solve = zgesv) to 1e-12."""
    import torch
    from oracle import qoc_oracle as orc
    from qoc_b200.standard.functions import expm, expm_vjp
    rng = np.random.default_rng(n)
    if kind == "hadamard":
        h = _hadamard(n)
        mats = [-1j * h * s for s in (0.05, 0.7, 3.0)] + [h * (0.4 / n)]
    elif kind == "degenerate_drift":
        d = np.diag(np.repeat([1.0, -1.0], n // 2)).astype(complex)
        x = np.ones((n, n)) - np.eye(n)
        mats = [-1j * (d + 0.25 * x), -1j * (3.0 * d + x), -1j * d * 7.0]
    elif kind == "pm1_skew":
        s = np.sign(rng.standard_normal((n, n)))
        a = np.triu(s, 1)
        mats = [(a - a.T) * c for c in (0.02, 0.3, 1.5)] + [1j * (a + a.T) * 0.2]
    else:
        perm = np.roll(np.eye(n), 1, axis=0)
        mats = [perm * c for c in (0.5, 2.0, 9.0)] + [-1j * (perm + perm.T) * 1.3]
    a = np.stack(mats).astype(np.complex128)
    got = expm(a)
    ubar = rng.standard_normal(a.shape) + 1j * rng.standard_normal(a.shape)
    out, abar = expm_vjp(a, ubar)
    for b in range(a.shape[0]):
        at = torch.tensor(a[b], requires_grad=True)
        u = orc.expm_pade(at)
        torch.sum(torch.as_tensor(ubar[b]) * u).real.backward()
        assert rel(got[b], u.detach().numpy()) < 1e-12, (kind, n, b)
        assert rel(out[b], u.detach().numpy()) < 1e-12, (kind, n, b)
        assert rel(abar[b], np.conj(at.grad.numpy())) < 1e-11, (kind, n, b)
