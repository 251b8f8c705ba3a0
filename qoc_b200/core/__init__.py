"""
core - the programs of the qoc API (mirror of qoc.core).
"""
from .lindbladdiscrete import evolve_lindblad_discrete, grape_lindblad_discrete
from .schroedingerdiscrete import (evolve_schroedinger_discrete, evaluate_schroedinger_discrete,
                                   grape_schroedinger_discrete)

__all__ = ["evolve_lindblad_discrete", "grape_lindblad_discrete", "evolve_schroedinger_discrete",
           "evaluate_schroedinger_discrete", "grape_schroedinger_discrete"]
