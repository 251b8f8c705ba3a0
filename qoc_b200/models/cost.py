"""
cost.py - base class of cost functions (qoc/models/cost.py:5-51).
"""


class Cost(object):
    """
    Fields:
    cost_multiplier :: float - weight of this cost in the total error
    name :: str - identifier
    requires_step_evaluation :: bool - True: evaluated at every cost step of the evolution,
        False: evaluated once on the final states
    """
    name = "parent_cost"
    requires_step_evaluation = False

    def __init__(self, cost_multiplier=1.):
        super().__init__()
        self.cost_multiplier = cost_multiplier

    def __str__(self):
        return self.name

    __repr__ = __str__

    def cost(self, controls, states, system_eval_step):
        raise NotImplementedError("The cost {} has not implemented an evaluation function.".format(self))

    # --- B200 path hooks ---------------------------------------------------------------------------
    def device_terms(self, state_count, hilbert_size):
        """state-dependent costs return a list of (kind, step_cost, weight, vectors[S][F][n], counts[S])
        descriptors for the CUDA reductions; control-only costs return []."""
        raise NotImplementedError("The cost {} has no B200 device descriptor; user-defined state costs are not "
                                  "supported by the CUDA path yet.".format(self))

    def control_value_and_grad(self, controls):
        """control-only costs override this with their analytic (value, d/dx + i d/dy).  The base implementation marks
        "not a control-only cost": the plans accept a cost without device terms only if its class overrides this hook."""
        return 0.0, None
