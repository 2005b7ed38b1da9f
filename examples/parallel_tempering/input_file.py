"""Inputs of the parallel-tempering example (values of the reference's examples/parallel_tempering/input_file.jl)."""
import numpy as np

# constants
k_B = 1 / 11.6            # meV / K
mu_B = 0.67 * k_B         # K/T -> meV/T

# local z axis of each basis site
z = [np.array(v) / np.sqrt(3) for v in ([1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1])]

# Monte Carlo parameters
t_thermalization = int(1e6)
t_measurement = int(1e6)
probe_rate = 2000
swap_rate = 50
overrelaxation = 10
report_interval = int(1e4)
checkpoint_rate = 1000

# lattice
L = 8
S = 0.5

# Zeeman coupling: g tensor diag(gxx, gyy, gzz) in the local frames, field direction h
g = np.array([0.0, 0.0, 2.18])
h = np.array([1.0, 0.0, 0.0])
h_local = [(h @ zi) * g for zi in z]

# exchange in meV
Jxx = Jzz = 0.043
Jyy = 0.065

# temperature window in meV
Tmin = 0.09 * k_B
Tmax = 14 * k_B
