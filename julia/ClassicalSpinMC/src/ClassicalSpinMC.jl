module ClassicalSpinMC

using LinearAlgebra, Random, Printf, Dates
using FunctionWrappers: FunctionWrapper

include("libcsmc.jl")          # ccall bindings of include/csmc.h
include("model.jl")            # UnitCell, add*!, Bravais presets
include("lattice_device.jl")   # Lattice + Hamiltonian evaluation through the library
include("binning.jl")          # Observables (log-binning error propagation)
include("drivers.jl")          # MonteCarlo, simulated_annealing!, deterministic_updates!, parallel_tempering!
include("files.jl")            # HDF5 params / configuration files

export UnitCell, addBasisSite!, addBilinear!, addCubic!, addQuartic!, addZeemanCoupling!, addOnSite!
export Lattice, set_spin!, random_spin_orientation
export get_magnetization
export overwrite_keys!, write_MC_checkpoint, create_params_file, read_lattice, read_spin_configuration!
export Metropolis, MetropolisAdaptive, MetropolisConstraint, MetropolisConstraintAdaptive, MetropolisFixedCone
export MonteCarlo, simulated_annealing!, deterministic_updates!, parallel_tempering!
export total_energy, energy_density, get_local_field
export Triangular, Square, Honeycomb, FCC, Pyrochlore, BreathingPyrochlore
export compute_equal_time_correlations

end
