"""B200-native sweep engine behind ClassicalSpinMC.jl's API (Python host mirror).

The compute path is libcsmc.so (hand-written sm_100a CUDA behind the C-ABI of include/csmc.h); this
package is the host side that plays the role of the reference's Julia layer (the image has no Julia
toolchain; julia/ClassicalSpinMC holds the same layer written against the same header).  Names follow
the reference's exports (src/ClassicalSpinMC.jl:6-35) with the trailing ``!`` dropped.
"""
from .unit_cell import (UnitCell, addBasisSite, addBilinear, addCubic, addOnSite, addQuartic,
                        addZeemanCoupling)
from .bravais import BreathingPyrochlore, FCC, Honeycomb, Pyrochlore, Square, Triangular
from .lattice import Lattice, random_spin_orientation, set_spin, get_spin
from .hamiltonian import energy_density, get_local_field, total_energy
from .observables import Observables, get_magnetization, specific_heat, susceptibility, update_observables
from .metropolis import (Metropolis, MetropolisAdaptive, MetropolisConstraint,
                         MetropolisConstraintAdaptive, MetropolisFixedCone)
from .monte_carlo import (MonteCarlo, MCParamsBuffer, SimulationParameters, deterministic_updates,
                          parallel_tempering, simulated_annealing)
from .reciprocal import get_allowed_wavevectors, get_k_path, get_k_plane, reciprocal
from .spin_correlations import (compute_equal_time_correlations, compute_equal_time_structure_factor,
                                runEqualTimeStructureFactor)
from .hdf5 import (create_params_file, overwrite_keys, read_lattice, read_spin_configuration,
                   write_MC_checkpoint)

__all__ = [
    "UnitCell", "addBasisSite", "addBilinear", "addCubic", "addQuartic", "addZeemanCoupling", "addOnSite",
    "Lattice", "set_spin", "random_spin_orientation", "get_magnetization",
    "overwrite_keys", "write_MC_checkpoint", "create_params_file", "read_lattice", "read_spin_configuration",
    "Metropolis", "MetropolisAdaptive", "MetropolisConstraint", "MetropolisConstraintAdaptive", "MetropolisFixedCone",
    "MonteCarlo", "simulated_annealing", "deterministic_updates", "parallel_tempering",
    "total_energy", "energy_density", "get_local_field",
    "Triangular", "Square", "Honeycomb", "FCC", "Pyrochlore", "BreathingPyrochlore",
    "compute_equal_time_correlations", "runEqualTimeStructureFactor", "compute_equal_time_structure_factor",
    "reciprocal", "get_allowed_wavevectors", "get_k_path", "get_k_plane",
]
