"""
core - the programs of the qoc API (mirror of qoc.core).
"""
from .schroedingerdiscrete import (evolve_schroedinger_discrete, evaluate_schroedinger_discrete,
                                   grape_schroedinger_discrete)

__all__ = ["evolve_schroedinger_discrete", "evaluate_schroedinger_discrete", "grape_schroedinger_discrete"]
