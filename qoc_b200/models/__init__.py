"""
models - data models of the qoc API (mirror of qoc/models): enums, the Cost base class, program state and
result containers.  Field names follow the reference because the hot-path seam reads them
(qoc/core/schroedingerdiscrete.py:371-388).
"""
from .enums import (InterpolationPolicy, MagnusPolicy, OperationPolicy, PerformancePolicy, ProgramType)
from .cost import Cost
from .state import (Dummy, ProgramState, GrapeState,
                    EvolveSchroedingerDiscreteState, EvolveSchroedingerResult,
                    GrapeSchroedingerDiscreteState, GrapeSchroedingerResult,
                    EvolveLindbladDiscreteState, EvolveLindbladResult,
                    GrapeLindbladDiscreteState, GrapeLindbladResult)

__all__ = [
    "Cost", "Dummy", "InterpolationPolicy", "EvolveLindbladDiscreteState", "EvolveLindbladResult",
    "GrapeLindbladDiscreteState", "GrapeLindbladResult", "MagnusPolicy", "OperationPolicy",
    "PerformancePolicy", "ProgramType", "ProgramState", "GrapeState", "EvolveSchroedingerDiscreteState",
    "EvolveSchroedingerResult", "GrapeSchroedingerDiscreteState", "GrapeSchroedingerResult",
]
