"""
make_golden_timedep.py - golden vectors from the UNMODIFIED reference forward for hamiltonians that use their `time`
argument (the reference evaluates the callable at every Magnus node,
qoc/core/schroedingerdiscrete.py:483-497).  Same mechanism as make_golden.py (numpy stands in for autograd.numpy; every
number is produced by the reference's own code).  The time-dependent callables are the seeded ones of
tests/problems.py:Problem.hamiltonian_td_numpy.

Run (in the build container only; /root/reference does not exist on the GPU box):
    python tests/golden/make_golden_timedep.py
Outputs: tests/golden/schroedinger_timedep_*.npz (committed).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (installs the stubs and imports the reference)
from tests.problems import Problem  # noqa: E402

CASES = [
    # n, slices, K, S, order, complex_controls, seed
    (4, 11, 2, 2, 2, False, 31),
    (5, 9, 1, 2, 4, True, 32),
    (6, 7, 2, 1, 6, True, 33),
]


def main():
    for i, (n, slices, K, S, order, cc, seed) in enumerate(CASES):
        p = Problem(n, slices, K, S, order, complex_controls=cc, seed=seed)
        ham = p.hamiltonian_td_numpy()
        costs = [mg.TargetStateInfidelity(p.target_states, cost_multiplier=0.8)]

        def forward(c):
            return mg.evolve_schroedinger_discrete(p.T, ham, p.initial_states, p.N, controls=c, costs=costs,
                                                   magnus_policy=mg.POLICIES[order]).error
        res = mg.evolve_schroedinger_discrete(p.T, ham, p.initial_states, p.N, controls=p.controls, costs=costs,
                                              magnus_policy=mg.POLICIES[order])
        mg.save("schroedinger_timedep_%d" % i,
                dict(n=n, slices=slices, K=K, S=S, order=order, complex_controls=cc, seed=seed, controls=p.controls,
                     error=res.error, final_states=res.final_states, fd_grad=mg.fd_grad(forward, p.controls)))


if __name__ == "__main__":
    main()
