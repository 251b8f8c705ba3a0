"""transmon_pi.py - the workload of the reference's examples/0_transmon_pi.py (BASELINE.json configs[0]) on the B200
path: GRAPE on the Schroedinger equation, one transmon, pi pulse |0> -> |1>, one complex control, 11 control points.
`import qoc_b200 as qoc` is the only change a user of the reference makes (HDF5 saving needs h5py and is skipped when
it is absent).  Run: python examples/transmon_pi.py [iterations]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qoc_b200 import grape_schroedinger_discrete  # noqa: E402
from qoc_b200.standard import (Adam, TargetStateInfidelity, get_annihilation_operator, get_creation_operator,  # noqa: E402
                               SIGMA_Z)

HILBERT_SIZE = 2
a, adag = get_annihilation_operator(HILBERT_SIZE), get_creation_operator(HILBERT_SIZE)
hamiltonian = lambda controls, time: SIGMA_Z / 2 + controls[0] * a + np.conjugate(controls[0]) * adag
INITIAL_STATES = np.stack((np.array([[1], [0]]),), axis=0)
TARGET_STATES = np.stack((np.array([[0], [1]]),), axis=0)
COSTS = [TargetStateInfidelity(TARGET_STATES)]
EVOLUTION_TIME = 10
CONTROL_EVAL_COUNT = SYSTEM_EVAL_COUNT = EVOLUTION_TIME + 1


def main(iteration_count=1000, log_iteration_step=100):
    return grape_schroedinger_discrete(1, CONTROL_EVAL_COUNT, COSTS, EVOLUTION_TIME, hamiltonian, INITIAL_STATES,
                                       SYSTEM_EVAL_COUNT, complex_controls=True, iteration_count=iteration_count,
                                       log_iteration_step=log_iteration_step, optimizer=Adam())


if __name__ == "__main__":
    res = main(int(sys.argv[1]) if len(sys.argv) > 1 else 1000)
    print("best error {:.6e} at iteration {}".format(res.best_error, res.best_iteration))
