"""Equal-time spin structure factor — mirror of src/spin_correlations.jl:6-43, computed on the GPU."""
from __future__ import annotations

import numpy as np


def compute_equal_time_correlations(lat, ks):
    """Suv[3u+v, n] = Re(s_u(k_n) conj(s_v(k_n))) / N with s_u(k) = sum_i exp(-i k.r_i) s_i^u; ``ks`` is (D, N_k)
    as in the reference.  Returns a (9, N_k) array."""
    ks = np.asarray(ks, dtype=np.float64)
    if ks.ndim != 2 or ks.shape[0] != lat.unit_cell.D:
        raise ValueError("ks must have shape (D, N_k)")
    lat.upload()
    return lat.engine().structure_factor(lat.unit_cell.lattice_vectors, lat.unit_cell.basis, ks)
