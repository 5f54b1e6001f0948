"""Structured synthetic meshes in gmsh node order (the benchmark / test inputs of SURVEY.md §8d).

gmsh is not available offline, so the named configs are generated here: a lattice of `order*n + 1` nodes per
direction, x fastest; every cell's nodes are placed from the element's own local-coordinate table
(`elements.local_coords`), so the connectivity can never disagree with the shape functions.  Simplices split
each quad into 2 triangles and each cube into 6 Kuhn tetrahedra (all positively oriented).  Nodes may be
jittered (`jitter` in units of the lattice spacing, seeded) so Jacobians are not constant.
"""
from __future__ import annotations

import itertools

import numpy as np

from . import elements as el

_ORDER = {"TRI3": 1, "TRI6": 2, "QUAD4": 1, "QUAD9": 2, "TETRA4": 1, "TETRA10": 2, "HEXA8": 1, "HEXA27": 2}


def _kuhn_simplices(dim: int):
    """Vertex offsets (in cell units) of the dim! Kuhn simplices of the unit cube, positively oriented."""
    out = []
    for perm in itertools.permutations(range(dim)):
        v = [np.zeros(dim, int)]
        for ax in perm:
            nxt = v[-1].copy()
            nxt[ax] = 1
            v.append(nxt)
        v = np.array(v)
        if np.linalg.det((v[1:] - v[0]).astype(float)) < 0:
            v[[-1, -2]] = v[[-2, -1]]  # swapping two vertices restores a positive Jacobian
        out.append(v)
    return out


def structured_mesh(elemType: str, n, lengths=None, jitter: float = 0.0, seed: int = 0):
    """Returns (coords (Nn,3) float64, connect (Ne,nPe) int64) for `n` cells per direction (int or tuple)."""
    elemType = str(elemType)
    dim = el.elem_dim(elemType)
    order = _ORDER[elemType]
    n = (n,) * dim if np.isscalar(n) else tuple(n)
    lengths = (1.0,) * dim if lengths is None else tuple(lengths)
    npts = [order * k + 1 for k in n]

    # lattice node ids, x fastest
    strides = np.ones(dim, dtype=np.int64)
    for d in range(1, dim):
        strides[d] = strides[d - 1] * npts[d - 1]

    # cell origins (lexicographic, x fastest)
    grids = np.meshgrid(*[np.arange(k, dtype=np.int64) * order for k in n[::-1]], indexing="ij")
    origin = np.stack([g.ravel() for g in grids[::-1]], axis=1)  # (ncell, dim) lattice index of each cell's corner

    loc = el.local_coords(elemType)
    if el._FAMILY[elemType] == "tensor":
        off = np.rint((loc + 1) / 2 * order).astype(np.int64)  # (nPe, dim)
        connect = ((origin[:, None, :] + off[None, :, :]) * strides).sum(-1)
    else:
        bary = np.concatenate([1 - loc.sum(1, keepdims=True), loc], axis=1)  # (nPe, dim+1)
        parts = []
        for verts in _kuhn_simplices(dim):
            off = np.rint(bary @ (verts * order)).astype(np.int64)  # node offsets inside the cell
            parts.append(((origin[:, None, :] + off[None, :, :]) * strides).sum(-1))
        # interleave so the simplices of one cell are consecutive
        connect = np.stack(parts, axis=1).reshape(-1, loc.shape[0])

    axes = [np.arange(k, dtype=float) for k in npts]
    mg = np.meshgrid(*axes[::-1], indexing="ij")
    lattice = np.stack([g.ravel() for g in mg[::-1]], axis=1)  # (Nn, dim) in lattice units
    if jitter:
        rng = np.random.default_rng(seed)
        lattice = lattice + rng.uniform(-jitter, jitter, size=lattice.shape)
    coords = np.zeros((lattice.shape[0], 3))
    for d in range(dim):
        coords[:, d] = lattice[:, d] * (lengths[d] / (npts[d] - 1))
    return coords, np.ascontiguousarray(connect, dtype=np.int64)


def hexa8_slab(n, rank: int, world: int, jitter: float = 0.0, seed: int = 0):
    """Rank `rank`'s view of the `nx x ny x (world*nz)` HEXA8 mesh (unit cells, x fastest) cut into z-slabs of `nz` layers —
    the weak-scaling input of SURVEY.md §8e, built WITHOUT the global mesh.

    Returns (connect (Nc, 8) GLOBAL node ids, elem_ids (Nc,) global and ascending, owner_of(ids) -> rank,
    coords_of(ids) -> (k, 3)) for a superset of the rank's local elements: its own layers plus the first layer of the
    next slab (whose elements touch the interface plane this rank owns — they are the ghost elements).  The jitter of a
    node depends only on (seed, its lattice plane), so ranks sharing a plane see the same coordinates."""
    nx, ny, nz = (n,) * 3 if np.isscalar(n) else tuple(n)
    P, L = (nx + 1) * (ny + 1), nx * ny
    lay0 = rank * nz
    lay1 = min((rank + 1) * nz + 1, world * nz)  # one ghost layer above
    _, c = structured_mesh("HEXA8", (nx, ny, lay1 - lay0))
    connect = c + lay0 * P
    elem_ids = np.arange(lay0 * L, lay1 * L, dtype=np.int64)
    h = 1.0 / max(nx, ny, nz)

    def owner_of(ids):
        plane = np.asarray(ids, dtype=np.int64) // P
        return np.minimum(np.maximum(plane - 1, 0) // nz, world - 1)

    def coords_of(ids):
        ids = np.asarray(ids, dtype=np.int64)
        plane, rem = ids // P, ids % P
        lat = np.stack([rem % (nx + 1), rem // (nx + 1), plane], axis=1).astype(np.float64)
        if jitter:
            for k in np.unique(plane):
                sel = plane == k
                jit = np.random.default_rng([seed, int(k)]).uniform(-jitter, jitter, size=(P, 3))
                lat[sel] += jit[rem[sel]]
        return lat * h

    return connect, elem_ids, owner_of, coords_of
