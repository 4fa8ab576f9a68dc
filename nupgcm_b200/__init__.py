"""nupgcm_b200 — nuPGCM's per-timestep solve path on NVIDIA B200 (sm_100a).

Host side (this package) mirrors the reference's Julia API for the path — ``Parameters``,
``Forcings``, ``Mesh``, ``Spaces``, ``FEData``, ``InversionToolkit``, ``EvolutionToolkit``,
``Model``, ``invert!``/``evolve!``/``run!`` (spelled ``invert_`` etc.), the ``CPU()``/``GPU()``
switch — and drives ``libnupgcm_b200.so`` (hand-written CUDA, C ABI in
``include/nupgcm_b200.h``) through ctypes.  There is no CPU fallback.
"""
from .architectures import (CPU, GPU, architecture, on_architecture, print_memory_status,
                            vector_type)
from .dofs import DoFHandler, FEData
from .evolution import EvolutionToolkit, collect_evolution_LHS_
from .inputs import (ConvectionParameterization, EddyParameterization, Forcings, Parameters,
                     SurfaceDirichletBC, SurfaceFluxBC)
from .inversion import InversionToolkit
from .io import save_state, set_state_from_file_
from .iterative_solvers import IterativeSolverToolkit, iterative_solve_
from .meshes import Mesh
from .preconditioners import BlockDiagonalPreconditioner
from .model import Model, State, evolve_, invert_, run_, set_b_, sync_flow_
from .spaces import Spaces
from .timesteppers import BDF1, BDF2, evolution_parameter, update_t_

__all__ = [
    "CPU", "GPU", "architecture", "on_architecture", "print_memory_status", "vector_type",
    "DoFHandler", "FEData", "EvolutionToolkit", "collect_evolution_LHS_", "Forcings",
    "ConvectionParameterization", "EddyParameterization", "BlockDiagonalPreconditioner",
    "Parameters", "SurfaceDirichletBC", "SurfaceFluxBC", "InversionToolkit",
    "IterativeSolverToolkit", "iterative_solve_", "Mesh", "Model", "State", "evolve_",
    "invert_", "run_", "set_b_", "sync_flow_", "Spaces", "BDF1", "BDF2",
    "evolution_parameter", "update_t_", "save_state", "set_state_from_file_",
]
