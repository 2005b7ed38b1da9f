"""Reciprocal-space helpers — mirror of src/reciprocal.jl: the callers that produce the momentum matrices
``ks`` (D x N_k, one wavevector per column) handed to compute_equal_time_correlations / MonteCarlo(ks=...).
Host-side numpy; no device work."""
from __future__ import annotations

import itertools

import numpy as np


def reciprocal(*a):
    """src/reciprocal.jl:3-16 — reciprocal lattice vectors (b_i . a_j = 2 pi delta_ij) in 2-D or 3-D."""
    a = [np.asarray(v, dtype=np.float64) for v in a]
    if len(a) == 2:
        a1, a2 = a
        mag = 2 * np.pi / (a1[0] * a2[1] - a1[1] * a2[0])
        return mag * np.array([a2[1], -a2[0]]), mag * np.array([-a1[1], a1[0]])
    if len(a) == 3:
        a1, a2, a3 = a
        mag = 2 * np.pi / np.dot(a1, np.cross(a2, a3))
        return mag * np.cross(a2, a3), mag * np.cross(a3, a1), mag * np.cross(a1, a2)
    raise ValueError("reciprocal() takes two or three lattice vectors")


def get_allowed_wavevectors(uc, shape, min=0, max=1):  # noqa: A002 - keyword names of the reference
    """src/reciprocal.jl:23-29 — wavevectors commensurate with a ``shape`` lattice, ``min``..``max`` copies of
    the reciprocal cell per dimension.  Column order follows Iterators.product (first dimension fastest)."""
    B = np.stack(reciprocal(*uc.lattice_vectors), axis=1)                     # columns b_i
    rs = [np.arange(min * dim, max * dim + 1) / dim for dim in shape]
    steps = np.array([t[::-1] for t in itertools.product(*rs[::-1])], dtype=np.float64).T   # D x n, first index fastest
    return B @ steps


def _within(temp, lower, upper):
    keep = np.ones(temp.shape[1], dtype=bool)
    for i in range(temp.shape[0]):
        keep &= (np.round(lower[i], 6) <= np.round(temp[i], 6)) & (np.round(temp[i], 6) <= np.round(upper[i], 6))
    return keep


def get_k_plane(uc, shape, min=-2, max=2):  # noqa: A002
    """src/reciprocal.jl:32-41"""
    ks = get_allowed_wavevectors(uc, shape, min=min, max=max)
    D = ks.shape[0]
    return ks[:, _within(ks, [min * 2 * np.pi] * D, [max * 2 * np.pi] * D)]


def _parallel_to(ks, origin, line):
    """columns of ks whose displacement from ``origin`` is (anti)parallel to the unit vector ``line``
    (src/reciprocal.jl:49-50,77-81; the zero displacement gives NaN there and is dropped)."""
    d = ks - np.asarray(origin, dtype=np.float64)[:, None]
    nrm = np.linalg.norm(d, axis=0)
    with np.errstate(invalid="ignore", divide="ignore"):
        c = np.abs(line @ (d / nrm))
    return np.round(c, 6) == 1.0


def get_k_path(uc, *args, **kw):
    """Two methods, as in the reference:
    ``get_k_path(uc, direction, shape, min=-2, max=2)`` — src/reciprocal.jl:43-60 (the displacement is taken
    from ``direction`` itself, as the reference does);
    ``get_k_path(uc, hsp, path, shape)`` — src/reciprocal.jl:65-103: ``hsp`` maps labels to high-symmetry
    points, ``path`` lists labels; returns (point_count, kpath)."""
    if isinstance(args[0], dict):
        hsp, path, shape = args
        ks = get_allowed_wavevectors(uc, shape)
        kpath = np.asarray(hsp[path[0]], dtype=np.float64)[:, None]
        point_count = np.zeros(len(path), dtype=np.int64)
        for ind in range(len(path) - 1):
            p1 = np.asarray(hsp[path[ind]], dtype=np.float64)
            p2 = np.asarray(hsp[path[ind + 1]], dtype=np.float64)
            line = (p2 - p1) / np.linalg.norm(p2 - p1)
            temp = ks[:, _parallel_to(ks, p1, line)]
            new_path = temp[:, _within(temp, np.minimum(p1, p2), np.maximum(p1, p2))]
            if new_path.shape[1] and not np.array_equal(np.round(new_path[:, -1], 5), np.round(p2, 5)):
                new_path = new_path[:, ::-1]
            kpath = np.hstack([kpath, new_path])
            point_count[ind + 1] = new_path.shape[1] + point_count[ind]
        return point_count, kpath
    direction, shape = args[0], args[1]
    lo = args[2] if len(args) > 2 else kw.get("min", -2)
    hi = args[3] if len(args) > 3 else kw.get("max", 2)
    ks = get_allowed_wavevectors(uc, shape, min=lo, max=hi)
    direction = np.asarray(direction, dtype=np.float64)
    direction = direction / np.linalg.norm(direction)
    temp = ks[:, _parallel_to(ks, direction, direction)]
    D = ks.shape[0]
    return temp[:, _within(temp, [lo * 2 * np.pi] * D, [hi * 2 * np.pi] * D)]
