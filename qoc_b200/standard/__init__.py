"""
standard - standard definitions of the qoc API (mirror of qoc.standard): constants, cost classes, functions,
optimisers and utilities.  Plot helpers of the reference (matplotlib/qutip) are out of scope.
"""
from .constants import (get_annihilation_operator, get_creation_operator, get_eij, SIGMA_X, SIGMA_Y, SIGMA_Z,
                        SIGMA_MINUS, SIGMA_PLUS)
from .costs import (ControlArea, ControlBandwidthMax, ControlNorm, ControlVariation, ForbidDensities,
                    ForbidStates, TargetDensityInfidelity, TargetDensityInfidelityTime, TargetStateInfidelity,
                    TargetStateInfidelityTime)
from .functions import (commutator, conjugate_transpose, expm, expm_vjp, krons, matmuls, rms_norm,
                        column_vector_list_to_matrix, matrix_to_column_vector_list)
from .optimizers import Adam, LBFGSB, SGD
from .utils import ans_jacobian, generate_save_file_path, CustomJSONEncoder, make_autograd_primitive

__all__ = [
    "get_annihilation_operator", "get_creation_operator", "get_eij", "SIGMA_X", "SIGMA_Y", "SIGMA_Z",
    "SIGMA_MINUS", "SIGMA_PLUS", "ControlArea", "ControlBandwidthMax", "ControlNorm", "ControlVariation",
    "ForbidDensities", "ForbidStates", "TargetDensityInfidelity", "TargetDensityInfidelityTime",
    "TargetStateInfidelity", "TargetStateInfidelityTime", "commutator", "conjugate_transpose", "expm",
    "expm_vjp", "krons", "rms_norm", "matmuls", "column_vector_list_to_matrix", "matrix_to_column_vector_list",
    "Adam", "LBFGSB", "SGD", "ans_jacobian", "generate_save_file_path", "CustomJSONEncoder",
    "make_autograd_primitive",
]
