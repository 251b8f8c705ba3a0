"""GPU: user-defined `Cost` subclasses (qoc/models/cost.py:5-51) - costs that only provide
`cost(controls, states, step)`.  The CUDA path runs the forward pass (qocb_forward), the host evaluates the user cost on the
final states and differentiates it numerically (4-point central differences), and the reverse pass is seeded with that
cotangent (qocb_backward).  Oracle: the same cost written with torch ops under torch.autograd."""
import numpy as np
import pytest

from tests.problems import Problem

pytestmark = pytest.mark.gpu


def rel(a, b):
    return np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel()) / max(np.linalg.norm(np.asarray(b).ravel()), 1e-300)


def _user_costs():
    import torch
    from qoc_b200.models.cost import Cost

    class Leakage(Cost):
        """population of the highest level, quartic in the amplitudes, with an explicit control-energy term"""
        name = "leakage"
        requires_step_evaluation = False

        def __init__(self, cost_multiplier=1.0, with_controls=False):
            super().__init__(cost_multiplier)
            self.with_controls = with_controls

        def cost(self, controls, states, system_eval_step):
            pop = np.abs(states[:, -1, 0]) ** 2
            extra = 0.01 * np.sum(np.abs(controls) ** 2) if self.with_controls else 0.0
            return self.cost_multiplier * (np.sum(pop) + 0.5 * np.sum(pop ** 2) + extra)

    class LeakageT(object):
        requires_step_evaluation = False

        def __init__(self, cost_multiplier=1.0, with_controls=False):
            self.cost_multiplier, self.with_controls = cost_multiplier, with_controls

        def cost(self, controls, states, step):
            pop = torch.abs(states[:, -1, 0]) ** 2
            extra = 0.01 * torch.sum(torch.abs(controls) ** 2) if self.with_controls else 0.0
            return self.cost_multiplier * (torch.sum(pop) + 0.5 * torch.sum(pop ** 2) + extra)
    return Leakage, LeakageT


@pytest.mark.parametrize("case", [(5, 14, 2, 2, False, False), (16, 9, 4, 3, True, True), (64, 6, 4, 4, False, False), (72, 5, 4, 2, True, False)],
                         ids=lambda c: "n%d_M%d_%s" % (c[0], c[2], "c" if c[4] else "r"))
def test_user_cost_vs_oracle(case):
    import qoc_b200.standard as std
    from oracle import qoc_oracle as orc
    from qoc_b200.core.plan import SchroedingerPlan
    from qoc_b200.models import MagnusPolicy
    pol = {2: MagnusPolicy.M2, 4: MagnusPolicy.M4, 6: MagnusPolicy.M6}
    n, slices, order, S, cc, with_controls = case
    Leakage, LeakageT = _user_costs()
    p = Problem(n, slices, 2, S, order, complex_controls=cc, F=1, seed=3 * n)
    costs = p.costs(std) + [Leakage(0.6, with_controls)]
    ocosts = p.costs(orc) + [LeakageT(0.6, with_controls)]
    plan = SchroedingerPlan(p.hamiltonian_numpy(), p.initial_states, costs, p.T, p.N, control_eval_count=p.M, control_count=2,
                            complex_controls=cc, magnus_policy=pol[order])
    err, grads, finals = plan.cost_and_grad(p.controls)
    err_f, finals_f = plan.cost(p.controls)
    plan.close()
    o_err, o_grad, o_fin = orc.schroedinger_cost_and_grad(p.controls, orc.make_hamiltonian(p.h0, p.drives, cc), p.initial_states,
                                                          ocosts, p.T, p.N, order=order)
    assert abs(err - o_err) <= 1e-10 * abs(o_err) and abs(err_f - o_err) <= 1e-10 * abs(o_err)
    assert rel(finals, o_fin) < 1e-10
    assert rel(grads, o_grad) < 1e-8, rel(grads, o_grad)


def test_user_step_cost_is_refused_loudly():
    import qoc_b200.standard as std
    from qoc_b200.core.plan import SchroedingerPlan
    from qoc_b200.models import MagnusPolicy
    from qoc_b200.models.cost import Cost

    class StepCost(Cost):
        requires_step_evaluation = True

        def cost(self, controls, states, step):
            return 0.0
    p = Problem(4, 5, 1, 1, 2)
    with pytest.raises(NotImplementedError):
        SchroedingerPlan(p.hamiltonian_numpy(), p.initial_states, [StepCost()], p.T, p.N, control_eval_count=p.M, control_count=1,
                         magnus_policy=MagnusPolicy.M2)
