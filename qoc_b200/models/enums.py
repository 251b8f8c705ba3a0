"""
enums.py - policy enumerations of the qoc API (qoc/models/magnuspolicy.py, interpolationpolicy.py,
programtype.py, operationpolicy.py, performancepolicy.py).  Values and string forms are part of the
on-disk / logging contract, so they match the reference.
"""
from enum import Enum


class _Named(Enum):
    def __str__(self):
        return self._label()

    __repr__ = __str__


class MagnusPolicy(_Named):
    """order of the Magnus expansion used per time slice (https://arxiv.org/abs/1709.06483)."""
    M2 = 1
    M4 = 2
    M6 = 3

    def _label(self):
        return {1: "magnus_m2", 2: "magnus_m4", 3: "magnus_m6"}[self.value]

    @property
    def order(self):
        return 2 * self.value


class InterpolationPolicy(_Named):
    LINEAR = 1

    def _label(self):
        return "interpolation_linear"


class ProgramType(_Named):
    EVOLVE = 1
    GRAPE = 2

    def _label(self):
        return "evolve" if self.value == 1 else "grape"


class OperationPolicy(_Named):
    """kept for signature compatibility (the reference never reads it, qoc/standard/optimizers/adam.py:41)."""
    CPU = 1
    GPU = 2
    CPU_SPARSE = 3
    GPU_SPARSE = 4

    def _label(self):
        return {1: "operation_cpu", 2: "operation_gpu", 3: "operation_cpu_sparse", 4: "operation_gpu_sparse"}[self.value]


class PerformancePolicy(_Named):
    TIME = 1
    MEMORY = 2

    def _label(self):
        return "performance_time" if self.value == 1 else "performance_memory"
