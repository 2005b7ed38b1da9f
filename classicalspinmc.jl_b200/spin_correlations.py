"""Equal-time spin structure factor — mirror of src/spin_correlations.jl:6-43, computed on the GPU."""
from __future__ import annotations

import datetime
import os

import numpy as np

from . import hdf5 as h5
from . import parallel


def compute_equal_time_correlations(lat, ks):
    """Suv[3u+v, n] = Re(s_u(k_n) conj(s_v(k_n))) / N with s_u(k) = sum_i exp(-i k.r_i) s_i^u; ``ks`` is (D, N_k)
    as in the reference.  Returns a (9, N_k) array."""
    ks = np.asarray(ks, dtype=np.float64)
    if ks.ndim != 2 or ks.shape[0] != lat.unit_cell.D:
        raise ValueError("ks must have shape (D, N_k)")
    lat.upload()
    return lat.engine().structure_factor(lat.unit_cell.lattice_vectors, lat.unit_cell.basis, ks)


def runEqualTimeStructureFactor(path, lat, ks, override=False):
    """src/spin_correlations.jl:48-108 — batch runner over the ``IC_<n>.h5`` files of a directory: the
    configurations are split over the ranks (remainder to the first ranks), each one is loaded into ``lat``,
    its structure factor computed on the rank's GPU and written to ``spin_correlations/{SSF, SSF_momentum}``
    of the same file.  Files that already hold the group are skipped unless ``override``."""
    rank, comm_size = parallel.comm_info()
    n_ic = len(os.listdir(path))
    per, rem = divmod(n_ic, comm_size)
    counts = [per + (1 if r < rem else 0) for r in range(comm_size)]        # :66-76
    ic = sum(counts[:rank])
    ks = np.asarray(ks, dtype=np.float64)
    for i in range(counts[rank]):
        file = os.path.join(path, f"IC_{ic}.h5")
        h5.read_spin_configuration(lat, file)
        f = h5._open(file, "r")
        exists = h5._has_group(f, "spin_correlations")
        f.close()
        if exists and not override:
            print(f"Skipping IC_{ic}")
        else:
            print(f"Computing SSF {i + 1}/{counts[rank]} on rank {rank}")
            S = compute_equal_time_correlations(lat, ks)
            print(f"Writing IC {ic} to file on rank {rank}")
            f = h5._open(file, "r+")
            h5.overwrite_keys(f, {"spin_correlations/SSF": S, "spin_correlations/SSF_momentum": ks})
            f.close()
        ic += 1
    print(f"Calculation completed on rank {rank} on", datetime.datetime.now().strftime("%d %b %Y %H:%M:%S"))


def compute_equal_time_structure_factor(path, dest):
    """src/spin_correlations.jl:114-145 — average ``spin_correlations/SSF`` over every file in ``path`` (the
    mean of the reference's LogBinner is the plain mean) and write it with the momenta into ``dest``."""
    files = sorted(os.listdir(path))
    print(f"Initializing LogBinner in {path}")
    total, ks = None, None
    print("Collecting correlations from", len(files), "files")
    for name in files:
        f = h5._open(os.path.join(path, name), "r")
        S = np.asarray(h5._get_jl(f, "spin_correlations/SSF"), dtype=np.float64)
        if ks is None:
            ks = np.asarray(h5._get_jl(f, "spin_correlations/SSF_momentum"))
        f.close()
        total = S.copy() if total is None else total + S
    mean = total / len(files)
    print(f"Writing to {dest}")
    d = h5._open(dest, "r+")
    h5.overwrite_keys(d, {"spin_correlations/SSF": mean, "spin_correlations/SSF_momentum": ks})
    d.close()
    print("Done")
    return mean
