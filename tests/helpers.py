"""Shared test helpers: seeded meshes, ctypes views of numpy arrays, the reference's own relative-error metric."""
import ctypes

import numpy as np

from easyfea_b200 import elements as el
from easyfea_b200 import meshgen

ELEM_CASES = {"TRI3": (5, 4), "TRI6": (3, 3), "QUAD4": (4, 3), "QUAD9": (3, 2), "TETRA4": (3, 2, 2), "TETRA10": (2, 2, 2),
              "HEXA8": (3, 3, 2), "HEXA27": (2, 2, 2)}


def rel_err(a, b):
    """|a-b|_F / |b|_F, the `Check` metric of the reference's tests (tests/FEM/linalg_test.py:28-34)."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    assert a.shape == b.shape, (a.shape, b.shape)
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


def make_mesh(elemType, n=None, jitter=0.2, seed=1):
    n = ELEM_CASES[elemType] if n is None else n
    return meshgen.structured_mesh(elemType, n, jitter=jitter, seed=seed)


class CGroup(ctypes.Structure):
    """mirror of `efb_group` (include/easyfea_b200.h)"""

    _fields_ = [("dim", ctypes.c_int32), ("nPe", ctypes.c_int32), ("nPg", ctypes.c_int32), ("coord_stride", ctypes.c_int32),
                ("Ne", ctypes.c_int64), ("connect", ctypes.c_void_p), ("coord", ctypes.c_void_p), ("dN_pg", ctypes.c_void_p),
                ("N_pg", ctypes.c_void_p), ("w_pg", ctypes.c_void_p)]


def host_group(elemType, coords, connect, matrixType):
    """efb_group over HOST numpy buffers (for the hostcheck emulation); returns (struct, keepalive list)."""
    tab = el.gauss_table(elemType, matrixType)
    c32 = np.ascontiguousarray(connect, dtype=np.int32)
    co = np.ascontiguousarray(coords, dtype=np.float64)
    dN = np.ascontiguousarray(tab.dN_pg)
    N = np.ascontiguousarray(tab.N_pg.reshape(tab.nPg, -1))
    w = np.ascontiguousarray(tab.weights)
    g = CGroup(el.elem_dim(elemType), el.elem_nPe(elemType), tab.nPg, co.shape[1], c32.shape[0], c32.ctypes.data, co.ctypes.data,
               dN.ctypes.data, N.ctypes.data, w.ctypes.data)
    return g, [c32, co, dN, N, w], tab


def p(a):
    """void* of a numpy array or None"""
    return None if a is None else ctypes.c_void_p(a.ctypes.data)
