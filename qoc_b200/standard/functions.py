"""
functions.py - host conveniences of qoc.standard.functions plus the device-backed `expm`.

`expm(a)` keeps the reference name (qoc/standard/functions/expm.py:276) and runs the batched Pade-13 CUDA
kernel through the C ABI; there is no NumPy expm in this package.
"""
from functools import reduce

import numpy as np

from qoc_b200 import _lib


def commutator(a, b):
    return np.matmul(a, b) - np.matmul(b, a)


def conjugate_transpose(matrix):
    return np.conjugate(np.swapaxes(matrix, -1, -2))


def krons(*matrices):
    return reduce(np.kron, matrices)


def matmuls(*matrices):
    return reduce(np.matmul, matrices)


def rms_norm(array):
    return np.sqrt(np.sum(array * np.conjugate(array)) / np.prod(np.shape(array)))


def column_vector_list_to_matrix(column_vector_list):
    return np.hstack(column_vector_list)


def matrix_to_column_vector_list(matrix):
    return np.stack([np.vstack(matrix[:, i]) for i in range(matrix.shape[1])])


def expm(a, device=0):
    """matrix exponential of (batch x) n x n complex matrices on the GPU (Pade-13 scaling and squaring)."""
    a = np.ascontiguousarray(a, dtype=np.complex128)
    n = a.shape[-1]
    batch = int(np.prod(a.shape[:-2])) if a.ndim > 2 else 1
    out = np.empty_like(a)
    lib = _lib.load()
    _lib.check(lib.qocb_expm_batched(n, batch, _lib.ptr(a), _lib.ptr(out), device))
    return out


def expm_vjp(a, ubar, device=0):
    """(expm(a), cotangent of a) for an output cotangent `ubar` in autograd's convention."""
    a = np.ascontiguousarray(a, dtype=np.complex128)
    ubar = np.ascontiguousarray(ubar, dtype=np.complex128)
    n = a.shape[-1]
    batch = int(np.prod(a.shape[:-2])) if a.ndim > 2 else 1
    out, abar = np.empty_like(a), np.empty_like(a)
    lib = _lib.load()
    _lib.check(lib.qocb_expm_vjp_batched(n, batch, _lib.ptr(a), _lib.ptr(ubar), _lib.ptr(out), _lib.ptr(abar), device))
    return out, abar
