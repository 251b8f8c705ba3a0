"""GPU: `hamiltonian(controls, time)` callables that are NOT affine in the controls through
`NonlinearSchroedingerPlan`: the host evaluates the callable at every Magnus node, the CUDA kernels propagate and
differentiate with respect to the operator-channel coefficients, the chain rule through the callable is a 4-point numeric
Jacobian.  Oracle: the torch version of the same callable under torch.autograd (what the reference does with HIPS autograd).
Cost and final states: 1e-10; gradient: 1e-8 (the numeric Jacobian's accuracy, stated in the class docstring)."""
import numpy as np
import pytest

from tests.problems import Problem, rand_herm

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel()) / max(np.linalg.norm(np.asarray(b).ravel()), 1e-300)


def _callables(p, complex_controls, time_dependent):
    import torch
    rng = np.random.default_rng(77)
    x = rand_herm(rng, p.n) * 0.15
    y = rand_herm(rng, p.n) * 0.2
    h0 = p.h0
    w = 0.9 if time_dependent else 0.0

    def h_np(c, t):
        if complex_controls:
            return h0 + (abs(c[0]) ** 2) * x + np.real(c[1] * np.exp(1j * w * t)) * y + np.sin(np.real(c[0]) + np.imag(c[1])) * (x @ x)
        return h0 + c[0] ** 2 * x + np.sin(c[1] + w * t) * y + c[0] * c[1] * (x @ x)

    cdt = torch.complex128
    th0, tx, ty = (torch.as_tensor(m, dtype=cdt) for m in (h0, x, y))
    txx = tx @ tx

    def h_t(c, t):
        t = float(t)
        if complex_controls:
            ph = complex(np.cos(w * t), np.sin(w * t))
            return th0 + (torch.abs(c[0]) ** 2) * tx + torch.real(c[1] * ph) * ty + torch.sin(torch.real(c[0]) + torch.imag(c[1])) * txx
        return th0 + c[0] ** 2 * tx + torch.sin(c[1] + w * t) * ty + c[0] * c[1] * txx
    return h_np, h_t


@pytest.mark.parametrize("case", [
    # n, slices, order, complex, time_dependent, M
    (4, 12, 2, False, False, 13),
    (6, 10, 4, False, True, 7),
    (5, 9, 4, True, True, 10),
    (8, 8, 6, True, False, 9),
], ids=lambda c: "n%d_M%d_%s_%s" % (c[0], c[2], "c" if c[3] else "r", "td" if c[4] else "ti"))
def test_nonlinear_hamiltonian_vs_oracle(case):
    import qoc_b200.standard as std
    from oracle import qoc_oracle as orc
    from qoc_b200.core.plan import NonlinearHamiltonian, NonlinearSchroedingerPlan, SchroedingerPlan, make_schroedinger_plan
    from qoc_b200.models import MagnusPolicy
    pol = {2: MagnusPolicy.M2, 4: MagnusPolicy.M4, 6: MagnusPolicy.M6}
    n, slices, order, cc, td, M = case
    p = Problem(n, slices, 2, 2, order, complex_controls=cc, F=2, seed=n, M=M, cost_eval_step=2, step_target=True)
    p.T = 3.0
    h_np, h_t = _callables(p, cc, td)
    costs = p.costs(std) + [std.ControlNorm(2, M, cost_multiplier=0.05)]
    ocosts = p.costs(orc) + [orc.ControlNorm(2, M, cost_multiplier=0.05)]
    kw = dict(control_eval_count=M, control_count=2, complex_controls=cc, magnus_policy=pol[order], cost_eval_step=2)
    with pytest.raises(NonlinearHamiltonian):                        # the affine plan refuses loudly (and is a NotImplementedError)
        SchroedingerPlan(h_np, p.initial_states, costs, p.T, p.N, **kw)
    plan = make_schroedinger_plan(h_np, p.initial_states, costs, p.T, p.N, **kw)
    assert isinstance(plan, NonlinearSchroedingerPlan) and plan.KC <= 16
    err, grads, finals = plan.cost_and_grad(p.controls)
    err_f, finals_f = plan.cost(p.controls)
    plan.close()
    o_err, o_grad, o_fin = orc.schroedinger_cost_and_grad(p.controls, h_t, p.initial_states, ocosts, p.T, p.N, order=order,
                                                          cost_eval_step=2)
    assert abs(err - o_err) <= 1e-10 * abs(o_err) and abs(err_f - o_err) <= 1e-10 * abs(o_err)
    assert rel(finals, o_fin) < 1e-10 and rel(finals_f, o_fin) < 1e-10
    assert grads.shape == p.controls.shape and grads.dtype == p.controls.dtype
    assert rel(grads, o_grad) < 1e-8, rel(grads, o_grad)


def test_grape_with_nonlinear_hamiltonian_descends():
    """the public program end to end: grape_schroedinger_discrete builds the non-linear plan by itself"""
    import qoc_b200 as qoc
    import qoc_b200.standard as std
    from qoc_b200.models import MagnusPolicy
    p = Problem(4, 10, 2, 1, 4, seed=2)
    h_np, _ = _callables(p, False, True)
    res = qoc.grape_schroedinger_discrete(2, 6, [std.TargetStateInfidelity(p.target_states)], 3.0, h_np, p.initial_states, 11,
                                          iteration_count=8, log_iteration_step=0, magnus_policy=MagnusPolicy.M4,
                                          initial_controls=np.full((6, 2), 0.3))
    first = qoc.evolve_schroedinger_discrete(3.0, h_np, p.initial_states, 11, controls=np.full((6, 2), 0.3),
                                             costs=[std.TargetStateInfidelity(p.target_states)], magnus_policy=MagnusPolicy.M4)
    assert res.best_error < first.error
