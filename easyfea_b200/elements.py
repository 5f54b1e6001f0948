"""Reference-element data (node layouts, shape functions, quadrature) for the element types of the hot path.

Host-side, tiny, NumPy only.  These tables are the *data* rows G1 of SURVEY.md §8(a): what the reference evaluates
through Python lambdas in `EasyFEA/FEM/Elems/*.py` (`_N/_dN`) at the points of `EasyFEA/FEM/_gauss.py:363-516`
(`Gauss_factory`).  Here every Lagrange family is generated from the node layout itself (tensor-product or
barycentric Lagrange polynomials), so a table can never disagree with the gmsh node order it was built from;
`tests/test_elements.py` pins each table against fixtures minted from the live reference.

When the package is used as a drop-in behind EasyFEA, tables are read from the reference group object instead
(`groupElem.Get_N_pg / Get_dN_pg / Get_gauss`), so element types not listed here still work there.
"""
from __future__ import annotations

import numpy as np

RIGI = "rigi"
MASS = "mass"

# ---------------------------------------------------------------------------------------------------------
# node layouts in gmsh order (local coordinates), cf. `Get_Local_Coords` of the reference element classes:
# `_tri.py:47,119`, `_quad.py:50,247`, `_tetra.py:63,164`, `_hexa.py:59,842`
# ---------------------------------------------------------------------------------------------------------
_H27_XYZ = (
    "-++--++-0--++0+-0-+000-+000",  # xi   of the 27 nodes: corners, edges, faces, centre
    "--++--++-0-0-+++-00+0-00+00",  # eta
    "----++++--0-0-00++++-0000+0",  # zeta
)


def _signs(text: str, dim: int) -> np.ndarray:
    val = {"-": -1.0, "0": 0.0, "+": 1.0}
    arr = np.array([val[c] for c in text], dtype=float)
    return arr[: (arr.size // dim) * dim].reshape(-1, dim)


_LOCAL = {
    "TRI3": np.array([[0, 0], [1, 0], [0, 1]], float),
    "TRI6": np.array([[0, 0], [1, 0], [0, 1], [0.5, 0], [0.5, 0.5], [0, 0.5]], float),
    "QUAD4": _signs("--+-++-+", 2),
    "QUAD9": _signs("--+-++-+0-+00+-000", 2),
    "TETRA4": np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], float),
    "TETRA10": np.array(
        [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0.5, 0, 0], [0.5, 0.5, 0], [0, 0.5, 0], [0, 0, 0.5],
         [0, 0.5, 0.5], [0.5, 0, 0.5]], float),
    "HEXA8": _signs("---+--++--+---++-++++-++", 3),
    "HEXA27": np.stack([_signs(t, 1)[:, 0] for t in _H27_XYZ], axis=1),
}

_FAMILY = {"TRI3": "simplex", "TRI6": "simplex", "TETRA4": "simplex", "TETRA10": "simplex",
           "QUAD4": "tensor", "QUAD9": "tensor", "HEXA8": "tensor", "HEXA27": "tensor"}

SUPPORTED = tuple(_LOCAL)


def elem_dim(elemType: str) -> int:
    return _LOCAL[str(elemType)].shape[1]


def elem_nPe(elemType: str) -> int:
    return _LOCAL[str(elemType)].shape[0]


def local_coords(elemType: str) -> np.ndarray:
    """(nPe, dim) local coordinates of the nodes, gmsh order."""
    return _LOCAL[str(elemType)].copy()


# ---------------------------------------------------------------------------------------------------------
# shape functions
# ---------------------------------------------------------------------------------------------------------
def _lagrange_1d(nodes: np.ndarray, x: np.ndarray):
    """Values and derivatives of the 1D Lagrange basis on `nodes` at points x -> (len(x), len(nodes)) each."""
    n = nodes.size
    val = np.ones((x.size, n))
    der = np.zeros((x.size, n))
    for a in range(n):
        for b in range(n):
            if b == a:
                continue
            val[:, a] *= (x - nodes[b]) / (nodes[a] - nodes[b])
        for b in range(n):
            if b == a:
                continue
            term = np.full(x.size, 1.0 / (nodes[a] - nodes[b]))
            for c in range(n):
                if c in (a, b):
                    continue
                term *= (x - nodes[c]) / (nodes[a] - nodes[c])
            der[:, a] += term
    return val, der


def _tensor_shape(elemType: str, pts: np.ndarray):
    loc = _LOCAL[elemType]
    nPe, dim = loc.shape
    nodes1d = np.unique(loc)  # (-1, 1) or (-1, 0, 1)
    vals, ders = zip(*[_lagrange_1d(nodes1d, pts[:, d]) for d in range(dim)])
    idx = np.searchsorted(nodes1d, loc)  # (nPe, dim) index of each node's 1D basis function
    N = np.ones((pts.shape[0], nPe))
    dN = np.ones((pts.shape[0], dim, nPe))
    for d in range(dim):
        N *= vals[d][:, idx[:, d]]
        for k in range(dim):
            dN[:, k, :] *= (ders[d] if k == d else vals[d])[:, idx[:, d]]
    return N, dN


def _simplex_shape(elemType: str, pts: np.ndarray):
    loc = _LOCAL[elemType]
    nPe, dim = loc.shape
    # barycentric coordinates l_0 = 1 - sum(x), l_k = x_{k-1}; gradient of l_i wrt x
    def bary(x):
        return np.concatenate([1.0 - x.sum(axis=1, keepdims=True), x], axis=1)

    L = bary(pts)  # (nPts, dim+1)
    gL = np.concatenate([-np.ones((1, dim)), np.eye(dim)], axis=0)  # (dim+1, dim)
    Lnodes = bary(loc)  # (nPe, dim+1)
    N = np.zeros((pts.shape[0], nPe))
    dN = np.zeros((pts.shape[0], dim, nPe))
    order = 1 if nPe == dim + 1 else 2
    for a in range(nPe):
        on = np.flatnonzero(Lnodes[a] > 1e-12)
        if order == 1:
            i = on[0]
            N[:, a] = L[:, i]
            dN[:, :, a] = gL[i][None, :]
        elif on.size == 1:  # vertex: l(2l - 1)
            i = on[0]
            N[:, a] = L[:, i] * (2 * L[:, i] - 1)
            dN[:, :, a] = (4 * L[:, i] - 1)[:, None] * gL[i][None, :]
        else:  # edge midpoint: 4 l_i l_j
            i, j = on
            N[:, a] = 4 * L[:, i] * L[:, j]
            dN[:, :, a] = 4 * (L[:, i, None] * gL[j][None, :] + L[:, j, None] * gL[i][None, :])
    return N, dN


def shape_functions(elemType: str, pts: np.ndarray):
    """N (nPts, nPe) and dN/dxi (nPts, dim, nPe) at local points `pts` (nPts, dim)."""
    elemType = str(elemType)
    pts = np.atleast_2d(np.asarray(pts, float))
    if _FAMILY[elemType] == "tensor":
        return _tensor_shape(elemType, pts)
    return _simplex_shape(elemType, pts)


# ---------------------------------------------------------------------------------------------------------
# quadrature — point ORDER follows `_gauss.py` so per-Gauss-point fields line up with the reference
# ---------------------------------------------------------------------------------------------------------
def _quad_rule(elemType: str, matrixType: str):
    t = elemType
    rigi = str(matrixType) == RIGI
    if t == "TRI3":
        return _triangle(1 if rigi else 3)
    if t == "TRI6":
        return _triangle(3 if rigi else 6)
    if t == "QUAD4":
        a = 1 / np.sqrt(3)
        return _LOCAL["QUAD4"] * a, np.ones(4)
    if t == "QUAD9":
        a = 0.774596669241483  # the reference's truncated sqrt(3/5), `_gauss.py:121`
        loc = _LOCAL["QUAD9"]
        w = np.array([25 / 81, 40 / 81, 64 / 81])[(loc == 0).sum(axis=1)]
        return loc * a, w
    if t == "TETRA4":
        return _tetra(1 if rigi else 4)
    if t == "TETRA10":
        return _tetra(4)
    if t == "HEXA8":
        a = 1 / np.sqrt(3)
        g = np.array([-a, a])
        X, Y, Z = np.meshgrid(g, g, g, indexing="ij")  # x slowest, z fastest
        return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1), np.ones(8)
    if t == "HEXA27":
        a = np.sqrt(3 / 5)
        g = np.array([-a, 0.0, a])
        X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
        pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
        c1, c2 = 5 / 9, 8 / 9
        table = np.array([c1**3, c1**2 * c2, c1 * c2**2, c2**3])
        return pts, table[(pts == 0).sum(axis=1)]
    raise NotImplementedError(f"no quadrature table for {t}")


def _triangle(nPg: int):
    if nPg == 1:
        return np.array([[1 / 3, 1 / 3]]), np.array([1 / 2])
    if nPg == 3:
        return np.array([[1 / 6, 1 / 6], [2 / 3, 1 / 6], [1 / 6, 2 / 3]]), np.full(3, 1 / 6)
    if nPg == 6:  # degree-4 Hammer rule
        a, b = 0.445948490915965, 0.091576213509771
        p1, p2 = 0.11169079483905, 0.0549758718227661
        pts = np.array([[b, b], [1 - 2 * b, b], [b, 1 - 2 * b], [a, 1 - 2 * a], [a, a], [1 - 2 * a, a]])
        return pts, np.array([p2, p2, p2, p1, p1, p1])
    raise NotImplementedError


def _tetra(nPg: int):
    if nPg == 1:
        return np.array([[1 / 4, 1 / 4, 1 / 4]]), np.array([1 / 6])
    if nPg == 4:
        a = (5 - np.sqrt(5)) / 20
        b = (5 + 3 * np.sqrt(5)) / 20
        return np.array([[a, a, a], [a, a, b], [a, b, a], [b, a, a]]), np.full(4, 1 / 24)
    raise NotImplementedError


class GaussTable:
    """Quadrature points/weights plus N and dN evaluated there, for one (elemType, matrixType)."""

    def __init__(self, elemType: str, matrixType: str):
        elemType = str(elemType)
        self.coord, self.weights = _quad_rule(elemType, matrixType)
        self.coord = np.ascontiguousarray(self.coord, dtype=float)
        self.weights = np.ascontiguousarray(self.weights, dtype=float)
        N, dN = shape_functions(elemType, self.coord)
        self.N_pg = np.ascontiguousarray(N[:, None, :])  # (nPg, 1, nPe)   like Get_N_pg
        self.dN_pg = np.ascontiguousarray(dN)  # (nPg, dim, nPe) like Get_dN_pg

    @property
    def nPg(self) -> int:
        return self.weights.size


_cache: dict = {}


def gauss_table(elemType: str, matrixType: str) -> GaussTable:
    key = (str(elemType), str(matrixType))
    if key not in _cache:
        _cache[key] = GaussTable(*key)
    return _cache[key]
