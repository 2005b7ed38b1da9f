"""Hamiltonian evaluation on a Lattice — mirror of src/hamiltonian.jl's public functions.  Bodies run
on the GPU through libcsmc; the host ``lat.spins`` array is uploaded first (it is the state of a
bare Lattice, test/latticetests.jl:15,29 mutate it directly)."""
from __future__ import annotations


def get_local_field(lattice, point: int):
    """src/hamiltonian.jl:3-67 — returns (Hx - hx, Hy - hy, Hz - hz); ``point`` is 1-based."""
    lattice.upload()
    f = lattice.engine().local_field(int(point))
    return (float(f[0]), float(f[1]), float(f[2]))


def total_energy(lattice) -> float:
    """src/hamiltonian.jl:70-132"""
    lattice.upload()
    return float(lattice.engine().total_energy()[0])


def energy_density(lattice) -> float:
    """src/hamiltonian.jl:134-136"""
    return total_energy(lattice) / lattice.size


def energy(lattice, point: int) -> float:
    """src/hamiltonian.jl:139-196 (internal in the reference)."""
    lattice.upload()
    return float(lattice.engine().site_energy_all()[int(point) - 1])
