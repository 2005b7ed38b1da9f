"""Output files — mirror of src/hdf5.jl's writers/readers (layout in SURVEY.md section 5).

The files are real HDF5 with the reference's group / dataset / attribute names.  Backend: ``h5py`` when it
can be imported; otherwise ``minih5`` (this package), a pure-Python writer/reader of the HDF5 1.8-compatible
subset the reference's files need (superblock v0, old-style groups, contiguous datasets, root attributes) —
the build image has no HDF5 library at all, see minih5.py for what that implies for validation.  File names
keep the reference's ``.h5`` / ``.h5.params`` suffixes (src/monte_carlo.jl:96-99).  The Julia package
(julia/ClassicalSpinMC) writes the same layout through HDF5.jl.
"""
from __future__ import annotations

import ast
import os

import numpy as np

from . import minih5

try:  # pragma: no cover - not installed in the build image
    import h5py
except Exception:  # noqa: BLE001
    h5py = None


# ---- tiny container abstraction ---------------------------------------------------------------------
def _is_mini(f):
    return isinstance(f, minih5.File)


def _open(filename, mode):
    if h5py is not None:
        return h5py.File(filename, mode)
    return minih5.File(filename, mode)


def _set(f, path, value):
    if not _is_mini(f):
        if path in f:
            del f[path]
        f[path] = value
    else:
        f.data[path] = np.asarray(value)


def _get(f, path):
    if not _is_mini(f):
        return f[path][()]
    return f.data[path]


def _jl(a):
    """Julia <-> on-disk orientation.  HDF5.jl stores a column-major Julia array A[i1, ..., ik] with the dataspace
    dimensions reversed (the bytes stay in place), so an HDF5 reader in C order (h5py, this module) sees
    A.T[ik, ..., i1] (util/load.py:88-93 relies on it for ``spins``).  Every array in the reference's files goes
    through this one function, on write and on read, so a file written by Julia reads back identically here and the
    other way round; it is its own inverse."""
    a = np.asarray(a)
    return np.ascontiguousarray(a.transpose()) if a.ndim >= 2 else a


def _set_jl(f, path, value):
    _set(f, path, _jl(value))


def _get_jl(f, path):
    return _jl(_get(f, path))


def _mkgroup(f, path):
    """create_group(...): the reference creates its groups even when they stay empty (src/hdf5.jl:37-71)
    and its reader iterates over them unconditionally (:92-116)."""
    if not _is_mini(f):
        f.require_group(path)
    else:
        f.groups.add(path.strip("/"))


def _has_group(f, path):
    if not _is_mini(f):
        return path in f
    p = path.strip("/")
    return p in f.groups or any(k.startswith(p + "/") for k in f.data)


def _set_attr(f, name, value):
    if not _is_mini(f):
        f.attrs[name] = value
    else:
        f.data["@attrs/" + name] = value if isinstance(value, (str, bytes)) else np.asarray(value)


def _get_attr(f, name):
    if not _is_mini(f):
        return f.attrs[name]
    v = f.data["@attrs/" + name]
    return v.item() if isinstance(v, np.ndarray) and v.shape == () else v


def _keys(f, group):
    if not _is_mini(f):
        return list(f[group].keys()) if group in f else []
    pre = group.rstrip("/") + "/"
    return sorted({k[len(pre):].split("/")[0] for k in f.data if k.startswith(pre)})


def _julia_tuple(t):
    """string(offset) of a Julia NTuple: '(0, -1)'; 1-tuples print as '(1,)'."""
    t = tuple(int(v) for v in t)
    return "(" + ", ".join(str(v) for v in t) + ("," if len(t) == 1 else "") + ")"


# ---- params file (src/hdf5.jl:36-147) -------------------------------------------------------------------
def dump_unit_cell(f, uc):
    """src/hdf5.jl:36-76"""
    for g in ("field", "onsite", "bilinear", "cubic", "quartic"):
        _mkgroup(f, "unit_cell/" + g)
    # Julia shapes: lattice_vectors D x D with columns a_i (:39), basis n_basis x D (:40), tensors [a, b, (c, (d))]
    _set_jl(f, "unit_cell/lattice_vectors", np.stack(uc.lattice_vectors, axis=1))
    _set_jl(f, "unit_cell/basis", np.stack(uc.basis, axis=0))
    for b, vec in uc.field:
        _set(f, f"unit_cell/field/{b}", vec)
    for b, mat in uc.onsite:
        _set_jl(f, f"unit_cell/onsite/{b}", mat)
    for b1, b2, mat, off in uc.bilinear:
        _set_jl(f, f"unit_cell/bilinear/({b1},{b2}),{_julia_tuple(off)}", mat)           # :60
    for b1, b2, b3, mat, o2, o3 in uc.cubic:
        _set_jl(f, f"unit_cell/cubic/({b1},{b2},{b3}),{_julia_tuple(o2)},{_julia_tuple(o3)}", mat)   # :67
    for b1, b2, b3, b4, mat, o2, o3, o4 in uc.quartic:
        _set_jl(f, f"unit_cell/quartic/({b1},{b2},{b3},{b4}),{_julia_tuple(o2)},{_julia_tuple(o3)},{_julia_tuple(o4)}", mat)  # :74


def dump_metadata(f, mc):
    """src/hdf5.jl:125-137"""
    dump_unit_cell(f, mc.lattice.unit_cell)
    _set(f, "lattice/size", np.array(mc.lattice.shape, dtype=np.int64))
    _set(f, "lattice/S", mc.lattice.S)
    _set(f, "lattice/bc", np.array(mc.lattice.bc))
    for k, v in mc.parameters._asdict().items():
        _set_attr(f, k, v)


def create_params_file(mc, filename):
    """src/hdf5.jl:142-147"""
    f = _open(filename, "w")
    dump_metadata(f, mc)
    f.close()
    return filename


def write_attributes(filename, d):
    """src/hdf5.jl:27-31"""
    f = _open(filename, "r+")
    for k, v in d.items():
        _set_attr(f, k, v)
    f.close()


def read_unit_cell(f):
    """src/hdf5.jl:81-120"""
    from .unit_cell import UnitCell
    lv = np.asarray(_get_jl(f, "unit_cell/lattice_vectors"))
    uc = UnitCell(*[lv[:, i] for i in range(lv.shape[1])])
    for row in np.atleast_2d(_get_jl(f, "unit_cell/basis")):
        uc.basis.append(np.array(row, dtype=np.float64))
    for key in _keys(f, "unit_cell/field"):
        uc.field.append((int(key), np.array(_get(f, f"unit_cell/field/{key}"))))
    for key in _keys(f, "unit_cell/onsite"):
        uc.onsite.append((int(key), np.array(_get_jl(f, f"unit_cell/onsite/{key}"))))
    for key in _keys(f, "unit_cell/bilinear"):
        (b1, b2), off = ast.literal_eval(key)                                    # eval(Meta.parse(key)), :103
        uc.bilinear.append((b1, b2, np.array(_get_jl(f, f"unit_cell/bilinear/{key}")), tuple(off)))
    for key in _keys(f, "unit_cell/cubic"):
        (b1, b2, b3), o2, o3 = ast.literal_eval(key)
        uc.cubic.append((b1, b2, b3, np.array(_get_jl(f, f"unit_cell/cubic/{key}")), tuple(o2), tuple(o3)))
    for key in _keys(f, "unit_cell/quartic"):
        (b1, b2, b3, b4), o2, o3, o4 = ast.literal_eval(key)
        uc.quartic.append((b1, b2, b3, b4, np.array(_get_jl(f, f"unit_cell/quartic/{key}")), tuple(o2), tuple(o3), tuple(o4)))
    return uc


def read_lattice(f):
    """src/hdf5.jl:152-159; ``f`` is an open file object or a filename."""
    from .lattice import Lattice
    own = isinstance(f, (str, os.PathLike))
    fid = _open(f, "r") if own else f
    size = tuple(int(v) for v in _get(fid, "lattice/size"))
    uc = read_unit_cell(fid)
    S = float(_get(fid, "lattice/S"))
    bc = _get(fid, "lattice/bc")
    bc = bc.decode() if isinstance(bc, bytes) else str(bc)
    if own:
        fid.close()
    return Lattice(size, uc, S, bc=bc)


# ---- configuration file (src/hdf5.jl:164-270) ---------------------------------------------------------
def _spins_for_file(spins):
    """Julia writes its 3 x N column-major array; h5py readers see (N, 3) (util/load.py:88-93)."""
    return _jl(spins)


def initialize_hdf5(mc, paramsfile, outpath=None, T=None, spins=None):
    """src/hdf5.jl:164-171"""
    f = _open(outpath or mc.outpath, "w")
    _set_attr(f, "T", float(mc.T if T is None else T))
    _set_attr(f, "paramsfile", paramsfile)
    _set(f, "spins", _spins_for_file(mc.lattice.spins if spins is None else spins))
    _set_jl(f, "site_positions", mc.lattice.site_positions)                      # Julia D x N
    f.close()


def write_MC_checkpoint(mc, outpath=None, spins=None):
    """src/hdf5.jl:176-180: overwrite dataset ``spins`` in the configuration file."""
    f = _open(outpath or mc.outpath, "r+")
    _set(f, "spins", _spins_for_file(mc.lattice.spins if spins is None else spins))
    f.close()


def write_initial_configuration(filename, mc, spins=None, T=None, config_path=None):
    """src/hdf5.jl:185-194"""
    src = _open(config_path or mc.outpath, "r")
    paramsfile = _get_attr(src, "paramsfile")
    src.close()
    f = _open(filename, "w")
    _set_attr(f, "T", float(mc.T if T is None else T))
    _set_attr(f, "paramsfile", paramsfile)
    _set(f, "spins", _spins_for_file(mc.lattice.spins if spins is None else spins))
    f.close()


def write_final_observables(mc, outpath=None, spins=None, observables=None, T=None):
    """src/hdf5.jl:204-239 (without the optional spin_correlations group)."""
    from .observables import _specific_heat, _susceptibility
    obs = mc.observables if observables is None else observables
    T = float(mc.T if T is None else T)
    f = _open(outpath or mc.outpath, "r+")
    _set(f, "spins", _spins_for_file(mc.lattice.spins if spins is None else spins))
    heat, dheat = _specific_heat(obs.energy, T, mc.lattice.size)
    chi, dchi = _susceptibility(obs.magnetization, T, mc.lattice.size)
    vals = {
        "specific_heat": heat, "specific_heat_err": abs(dheat / heat) if heat != 0 else np.nan,      # :220-221
        "susceptibility": chi, "susceptibility_err": abs(dchi / chi) if chi != 0 else np.nan,         # :222-223
        "magnetization": obs.magnetization.mean(1), "magnetization_err": obs.magnetization.std_error(1),
        "energy": obs.energy.mean(1), "energy_err": obs.energy.std_error(1),                          # :224-227
    }
    for k, v in vals.items():
        _set(f, f"observables/{k}", v)
    f.close()
    return vals


def overwrite_keys(fid, d):
    """src/hdf5.jl:244-252.  Arrays are given in the Julia orientation (e.g. SSF 9 x N_k, momenta D x N_k)."""
    for k, v in d.items():
        _set_jl(fid, k, v)


def read_spin_configuration(lat, filename):
    """src/hdf5.jl:266-270"""
    f = _open(filename, "r")
    lat.spins[:, :] = _get_jl(f, "spins")
    f.close()


def read_observables(filename):
    f = _open(filename, "r")
    out = {k: float(np.asarray(_get(f, f"observables/{k}"))) for k in _keys(f, "observables")}
    f.close()
    return out
