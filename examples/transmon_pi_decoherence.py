"""transmon_pi_decoherence.py - the workload of the reference's examples/1_transmon_pi_dechoerence.py (BASELINE.json
configs[1]) on the B200 path: GRAPE on the Lindblad master equation with T1 decay, L-BFGS-B on the host.
Run: python examples/transmon_pi_decoherence.py [iterations]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qoc_b200 import grape_lindblad_discrete  # noqa: E402
from qoc_b200.standard import (LBFGSB, TargetDensityInfidelity, conjugate_transpose, get_annihilation_operator,  # noqa: E402
                               get_creation_operator, SIGMA_Z)

a, adag = get_annihilation_operator(2), get_creation_operator(2)
hamiltonian = lambda controls, time: SIGMA_Z / 2 + controls[0] * a + np.conjugate(controls[0]) * adag
T1 = 1e3
lindblad_data = lambda time: (np.stack((1 / T1,)), np.stack((a,)))
INITIAL_STATES = np.stack((np.array([[1], [0]]),), axis=0)
TARGET_STATES = np.stack((np.array([[0], [1]]),), axis=0)
INITIAL_DENSITIES = np.matmul(INITIAL_STATES, conjugate_transpose(INITIAL_STATES))
TARGET_DENSITIES = np.matmul(TARGET_STATES, conjugate_transpose(TARGET_STATES))
COSTS = [TargetDensityInfidelity(TARGET_DENSITIES)]


def main(iteration_count=50, log_iteration_step=5):
    return grape_lindblad_discrete(1, 11, COSTS, 10, INITIAL_DENSITIES, 2, complex_controls=True, hamiltonian=hamiltonian,
                                   iteration_count=iteration_count, lindblad_data=lindblad_data,
                                   log_iteration_step=log_iteration_step, max_control_norms=np.array((5,)),
                                   optimizer=LBFGSB())


if __name__ == "__main__":
    res = main(int(sys.argv[1]) if len(sys.argv) > 1 else 50)
    print("best error {:.6e} at iteration {}".format(res.best_error, res.best_iteration))
