"""Host-side mesh containers with the slice of the `_GroupElem` / `Mesh` interface the hot path reads, plus the
device mirror of a group (connectivity, coordinates and reference-element tables resident in HBM).

`ElemGroup` follows EasyFEA/FEM/_group_elem.py:41-330 (connect with GLOBAL node ids, `coord` = rows of the mesh
coordinates used by the group, `_global_to_local_nodes`) so the operator functions accept either an `ElemGroup` or a
reference `_GroupElem` (duck-typed).  The device mirror is stored inside the group's computed-values cache — the
dict the reference clears in `_InitMatrix` (`Utilities/_cache.py:9,43-47`, `_group_elem.py:115-118`) — so the
reference's own invalidation points (coordinate edits, `Mesh._ResetMatrix`, `simu.mesh = ...`) drop it too.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from . import device as dv
from . import elements as el

_CACHE_ATTR = "__cachedComputedValues"  # same attribute name as the reference's CACH_NAME
_DEV_KEY = ("easyfea_b200.device_group",)


class ElemGroup:
    """One group of same-type elements."""

    def __init__(self, elemType: str, connect, coordinates, all_nodes_used: bool = False):
        self.elemType = str(elemType)
        self.dim = el.elem_dim(self.elemType)
        self.nPe = el.elem_nPe(self.elemType)
        connect = np.ascontiguousarray(connect, dtype=np.int64)
        assert connect.ndim == 2 and connect.shape[1] == self.nPe, "connect must be a (Ne, nPe) array."
        coordinates = np.ascontiguousarray(coordinates, dtype=np.float64)
        assert coordinates.ndim == 2 and coordinates.shape[1] == 3, "Must be a (Ncoords, 3) array."
        self._connect = connect
        self.Ncoords = coordinates.shape[0]
        if all_nodes_used:  # skip the unique() of large generated meshes
            self.nodes = np.arange(self.Ncoords)
            self._coord = coordinates
        else:
            self.nodes = np.unique(connect.ravel())
            self._coord = coordinates[self.nodes]
        self._global_to_local_nodes = np.empty(self.Ncoords, dtype=np.int64)
        self._global_to_local_nodes[self.nodes] = np.arange(self.nodes.size)
        setattr(self, _CACHE_ATTR, {})

    # --- the `_GroupElem` properties the path reads ---
    @property
    def Ne(self) -> int:
        return self._connect.shape[0]

    @property
    def Nn(self) -> int:
        return self.nodes.size

    @property
    def connect(self) -> np.ndarray:
        return self._connect

    @property
    def coord(self) -> np.ndarray:
        return self._coord

    @coord.setter
    def coord(self, coordinates) -> None:
        assert coordinates.shape == (self.Ncoords, 3)
        self._coord = np.ascontiguousarray(coordinates[self.nodes], dtype=np.float64)
        self._InitMatrix()

    @property
    def inDim(self) -> int:
        return self.dim if self.dim == 3 or not np.any(self._coord[:, 2]) else 3

    def _InitMatrix(self) -> None:
        getattr(self, _CACHE_ATTR).clear()

    def Get_gauss(self, matrixType):
        return el.gauss_table(self.elemType, matrixType)

    def Get_N_pg(self, matrixType) -> np.ndarray:
        return el.gauss_table(self.elemType, matrixType).N_pg

    def Get_dN_pg(self, matrixType) -> np.ndarray:
        return el.gauss_table(self.elemType, matrixType).dN_pg

    def Get_weight_pg(self, matrixType) -> np.ndarray:
        return el.gauss_table(self.elemType, matrixType).weights

    def Get_assembly_e(self, dof_n: int) -> np.ndarray:
        c = self._connect
        return (c[:, :, None] * dof_n + np.arange(dof_n)[None, None, :]).reshape(c.shape[0], -1)

    def Locates_sol_e(self, sol, dof_n=None):
        sol = np.asarray(sol)
        if dof_n is None:
            dof_n = sol.shape[0] // self.Ncoords
        return sol[self.Get_assembly_e(int(dof_n))]


class Mesh:
    """A set of element groups over one coordinate array; `Get_list_groupElem()` returns the main groups."""

    def __init__(self, groups, coordinates=None):
        self.groups = list(groups.values()) if isinstance(groups, dict) else list(groups)
        self.dim = max(g.dim for g in self.groups)
        self.Nn = self.groups[0].Ncoords
        self.coord = coordinates

    def Get_list_groupElem(self, dim=None):
        dim = self.dim if dim is None else dim
        return [g for g in self.groups if g.dim == dim]

    @property
    def Ne(self):
        return sum(g.Ne for g in self.Get_list_groupElem())

    def _ResetMatrix(self):
        for g in self.groups:
            g._InitMatrix()


# ---------------------------------------------------------------------------------------------------------
# device mirror
# ---------------------------------------------------------------------------------------------------------
class DeviceGroup:
    """Connectivity (int32), coordinates and per-matrixType tables of one group, resident on the device."""

    def __init__(self, g):
        dim = int(g.dim)
        if dim not in (2, 3):
            raise NotImplementedError(f"element dimension {dim} is outside the hot path (2D/3D only)")
        if int(getattr(g, "inDim", dim)) != dim:
            raise NotImplementedError("elements embedded in a higher-dimensional space (dim != inDim) are out of scope")
        self.dim, self.nPe, self.Ne = dim, int(g.nPe), int(g.Ne)
        connect = np.asarray(g.connect)
        if connect.size and connect.max() >= 2**31:
            raise ValueError("node ids must fit int32")
        self.Ncoords = int(g.Ncoords)
        self.connect_glob = dv.to_device(connect.astype(np.int32, copy=False).reshape(self.Ne, self.nPe))
        g2l = getattr(g, "_global_to_local_nodes", None)
        nodes = np.asarray(g.nodes)
        if g2l is None or (nodes.size == self.Ncoords and (nodes.size == 0 or nodes[-1] == nodes.size - 1)):
            self.connect = self.connect_glob  # every node used: local rows == global ids
        else:
            self.connect = dv.to_device(np.asarray(g2l)[connect].astype(np.int32))
        self.coord = dv.to_device(np.asarray(g.coord, dtype=np.float64))
        self._tables = {}
        self._src = g

    def tables(self, matrixType):
        key = str(matrixType)
        if key not in self._tables:
            g = self._src
            dN = np.ascontiguousarray(np.asarray(g.Get_dN_pg(matrixType), dtype=np.float64))
            N = np.ascontiguousarray(np.asarray(g.Get_N_pg(matrixType), dtype=np.float64)).reshape(dN.shape[0], -1)
            w = np.ascontiguousarray(np.asarray(g.Get_weight_pg(matrixType), dtype=np.float64)).ravel()
            assert dN.shape == (w.size, self.dim, self.nPe) and N.shape == (w.size, self.nPe)
            self._tables[key] = (dv.to_device(dN), dv.to_device(N), dv.to_device(w), int(w.size))
        return self._tables[key]

    def nPg(self, matrixType) -> int:
        return self.tables(matrixType)[3]

    def cstruct(self, matrixType) -> _lib.EfbGroup:
        dN, N, w, nPg = self.tables(matrixType)
        return _lib.EfbGroup(self.dim, self.nPe, nPg, int(self.coord.shape[1]), self.Ne, self.connect.data_ptr(),
                             self.coord.data_ptr(), dN.data_ptr(), N.data_ptr(), w.data_ptr())


def device_group(g) -> DeviceGroup:
    """Device mirror of a group, cached where the reference keeps its per-group computed values."""
    if isinstance(g, DeviceGroup):
        return g
    try:
        cache = getattr(g, _CACHE_ATTR)
    except AttributeError:
        cache = {}
        setattr(g, _CACHE_ATTR, cache)
    if _DEV_KEY not in cache:
        cache[_DEV_KEY] = DeviceGroup(g)
    return cache[_DEV_KEY]
