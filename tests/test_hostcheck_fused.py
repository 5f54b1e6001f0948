"""Fused element integration + assembly (`efb_assemble_elastic`, csrc/fused_kernels.cuh) on the CPU: the one-time node
schedule (`assembly.FusedSchedule`, torch tensor ops — device-agnostic) and the kernel body through the test-only host
emulation, against the oracle's K_e + np.bincount assembly.  Tolerance 1e-12 relative (north star)."""
import ctypes
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from easyfea_b200 import assembly
from easyfea_b200 import elements as el
from oracle import easyfea_oracle as orc
from tests.helpers import host_group, make_mesh, p, rel_err

TOL = 1e-12


def host_node_graph(connect, Nn, coords, elemType):
    """NumPy construction of what `assembly.NodeGraph` builds on the device (rowptr/qlist, adjptr/adj, pos), as CPU tensors"""
    Ne, nPe = connect.shape
    q = np.arange(Ne * nPe, dtype=np.int64)
    node = connect.ravel()
    order = np.lexsort((q, node))
    qlist = q[order]
    rowptr = np.zeros(Nn + 1, dtype=np.int64)
    np.cumsum(np.bincount(node, minlength=Nn), out=rowptr[1:])
    pairs = np.unique(np.stack([np.repeat(connect, nPe, axis=1).ravel(), np.tile(connect, (1, nPe)).ravel()], 1), axis=0)
    adjptr = np.zeros(Nn + 1, dtype=np.int64)
    np.cumsum(np.bincount(pairs[:, 0], minlength=Nn), out=adjptr[1:])
    adj = pairs[:, 1].astype(np.int32)
    key = pairs[:, 0] * Nn + pairs[:, 1]
    a_nodes, b_nodes = np.repeat(connect, nPe, axis=1).ravel(), np.tile(connect, (1, nPe)).ravel()
    pos = (np.searchsorted(key, a_nodes * Nn + b_nodes) - adjptr[a_nodes]).astype(np.int32)
    t = torch.from_numpy
    c32 = t(np.ascontiguousarray(connect, dtype=np.int32))
    dg = SimpleNamespace(dim=el.elem_dim(elemType), nPe=nPe, Ne=Ne, connect=c32, connect_glob=c32, coord=t(np.ascontiguousarray(coords)))
    return SimpleNamespace(dgs=[dg], Nn=Nn, rowptr=t(rowptr), qlist=t(qlist), adjptr=t(adjptr), adj=t(adj), pos=t(pos),
                           max_deg=int(np.diff(adjptr).max()), nnz_node=int(adjptr[-1]))


CASES = [("HEXA8", (5, 4, 3), None), ("HEXA8", (6, 6, 5), 16), ("HEXA8", (9, 5, 4), 4), ("TETRA4", (3, 3, 2), None), ("TRI3", (6, 5), None),
         ("QUAD4", (5, 4), 8), ("QUAD4", (5, 4), None), ("QUAD9", (3, 3), None), ("TETRA10", (2, 2, 2), None), ("TRI6", (4, 3), None)]


@pytest.mark.parametrize("elemType,n,S", CASES)
@pytest.mark.parametrize("general_C", [False, True])
def test_fused_assembly_matches_two_stage(hostcheck, elemType, n, S, general_C):
    rng = np.random.default_rng(7)
    coords, connect = make_mesh(elemType, n)
    Nn = coords.shape[0]
    g, keep, tab = host_group(elemType, coords, connect, "rigi")
    dim, nPe = g.dim, g.nPe
    ns = 3 if dim == 2 else 6
    C = orc.IsoMaterial(dim, 210000.0, 0.3).C
    if general_C:  # fully populated non-symmetric tensor: every Dt term is exercised
        C = C * rng.uniform(0.5, 2.0, C.shape) + rng.uniform(1e3, 1e4, C.shape)
    graph = host_node_graph(connect, Nn, coords, elemType)
    sched = assembly.FusedSchedule(graph, S=S, nPg=tab.nPg)
    # every (node, element) pair appears exactly once
    assert sched.n_tasks == connect.size
    out = np.full(graph.nnz_node * dim * dim, np.nan)
    arr = lambda t_: np.ascontiguousarray(t_.numpy())  # noqa: E731
    cl_nodes, cl_ne, cl_conn, desc, tpos = map(arr, (sched.cl_nodes, sched.cl_ne, sched.cl_conn, sched.desc, sched.tpos))
    rc = hostcheck.hc_assemble_elastic(ctypes.byref(g), p(np.ascontiguousarray(C)), ctypes.c_double(1.7), sched.n_clusters, sched.S,
                                       sched.cap_e, sched.max_deg, p(cl_nodes), p(cl_ne), p(cl_conn), p(desc), p(tpos), p(out))
    assert rc == 0
    geo = orc.geometry(coords[connect][:, :, :dim], tab.dN_pg, tab.weights)
    Ke = 1.7 * orc.linearized_elasticity(geo, C)
    inv, indices, indptr, nnz = orc.csr_map([connect], dim, Nn * dim, True)
    ref = orc.assemble_replay([Ke], inv, nnz)
    assert out.shape == ref.shape and np.isfinite(out).all()
    assert rel_err(out, ref) < TOL


def test_fused_schedule_clusters_are_compact():
    """on a structured (jittered) HEXA8 mesh the Morton clusters are 4x2x2 node bricks: <= 45 elements each, and every element
    is integrated ~2.8 times per launch (the geometry redundancy reported in DESIGN.md)"""
    coords, connect = make_mesh("HEXA8", (16, 16, 16))
    graph = host_node_graph(connect, coords.shape[0], coords, "HEXA8")
    sched = assembly.FusedSchedule(graph, S=16, nPg=8)
    assert sched.cap_e <= 48
    assert sched.redundancy() < 3.2
    need = assembly.smem_bytes(3, 8, 8, 16, sched.cap_e, sched.max_deg)
    assert need <= 113 * 1024  # two CTAs per SM


def test_fused_schedule_partial_rows_and_orphans():
    """owned-row prefix (sharded runs) and nodes without elements: only the scheduled nodes' blocks are written"""
    coords, connect = make_mesh("HEXA8", (4, 3, 3))
    Nn = coords.shape[0]
    coords2 = np.vstack([coords, [[9.0, 9.0, 9.0]]])  # an orphan node at the end
    graph = host_node_graph(connect, Nn + 1, coords2, "HEXA8")
    sched = assembly.FusedSchedule(graph, n_nodes=Nn // 2, nPg=8)
    ids = sched.cl_nodes[:, 0].numpy()
    assert set(ids[ids >= 0].tolist()) == set(range(Nn // 2))
    full = assembly.FusedSchedule(graph, nPg=8)
    ids = full.cl_nodes[:, 0].numpy()
    assert set(ids[ids >= 0].tolist()) == set(range(Nn))  # the orphan is skipped
