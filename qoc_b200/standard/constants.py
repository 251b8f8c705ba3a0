"""
constants.py - Pauli matrices and ladder operators (qoc/standard/constants.py).
"""
import numpy as np

SIGMA_X = np.array(((0, 1), (1, 0)))
SIGMA_Y = np.array(((0, -1j), (1j, 0)))
SIGMA_Z = np.array(((1, 0), (0, -1)))
SIGMA_PLUS = np.array(((0, 1), (0, 0)))
SIGMA_MINUS = np.array(((0, 0), (1, 0)))


def get_creation_operator(size):
    """a^dagger truncated to `size` levels: sqrt(1..size-1) on the first sub-diagonal."""
    return np.diag(np.sqrt(np.arange(1, size)), k=-1)


def get_annihilation_operator(size):
    """a truncated to `size` levels: sqrt(1..size-1) on the first super-diagonal."""
    return np.diag(np.sqrt(np.arange(1, size)), k=1)


def get_eij(i, j, size):
    out = np.zeros((size, size))
    out[i, j] = 1
    return out
