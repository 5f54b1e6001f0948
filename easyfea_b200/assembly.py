"""Global CSR assembly on the device — the drop-in for `_Simu.__Get_csr_map` / `__Assemble_csr`
(EasyFEA/Simulations/_simu.py:989-1102; level 2 of the boundary).

`NodeGraph` holds what depends on connectivity only (node -> element rows, node adjacency, slot positions); it is
built once per tuple of contributing groups and shared by every `dof_n`.  `CsrPattern` is its block expansion for one
`(dof_n, Ndof)`: `indptr/indices` bit-identical to scipy's canonical pattern, index dtype chosen like scipy
(`int32` unless `max(Ndof, n_entries) > 2**31-1`).  `replay_*` sums element entries in ascending entry order, i.e.
bit-identical to the reference's `np.bincount`, with no atomics.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from . import device as dv
from .mesh import device_group

_I32MAX = np.iinfo(np.int32).max


def _group_args(dgs, pointers):
    n = len(dgs)
    ptrs = (ctypes.c_void_p * n)(*[p.data_ptr() if p is not None else 0 for p in pointers])
    Ne = (ctypes.c_int64 * n)(*[dg.Ne for dg in dgs])
    nPe = (ctypes.c_int32 * n)(*[dg.nPe for dg in dgs])
    return n, ptrs, Ne, nPe


def _scan(cnt: torch.Tensor) -> torch.Tensor:
    n = cnt.numel()
    out = dv.empty((n + 1,), torch.int64)
    work = dv.empty((n // 1024 + 2,), torch.int64)
    _lib.call("efb_exclusive_scan_i32", dv.ptr(cnt), n, dv.ptr(out), dv.ptr(work), dv.stream_ptr())
    return out


class NodeGraph:
    """Connectivity-only part of the assembly map for an ordered tuple of element groups (A1, stages 1-6 and 8)."""

    def __init__(self, groups, Nn: int):
        self.dgs = [device_group(g) for g in groups]
        if not 1 <= len(self.dgs) <= 8:
            raise ValueError("between 1 and 8 element groups can contribute to one matrix")
        self.Nn = int(Nn)
        for dg in self.dgs:
            if dg.Ncoords > self.Nn:
                raise ValueError("a group references more coordinates than Nn")
        connects = [dg.connect_glob for dg in self.dgs]
        n, ptrs, Ne, nPe = _group_args(self.dgs, connects)
        st = dv.stream_ptr()
        Nn = self.Nn
        cnt = dv.empty((Nn,), torch.int32)
        _lib.call("efb_csr_count_node_rows", n, ptrs, Ne, nPe, Nn, dv.ptr(cnt), st)
        self.rowptr = _scan(cnt)
        n_rows = int(self.rowptr[-1].item())
        self.qlist = dv.empty((n_rows,), torch.int64)
        _lib.call("efb_csr_fill_node_rows", n, ptrs, Ne, nPe, Nn, dv.ptr(self.rowptr), dv.ptr(cnt), dv.ptr(self.qlist), st)
        deg = cnt  # reuse
        err = torch.zeros(1, dtype=torch.int32, device=cnt.device)
        _lib.call("efb_csr_count_adj", n, ptrs, Ne, nPe, Nn, dv.ptr(self.rowptr), dv.ptr(self.qlist), dv.ptr(deg), dv.ptr(err), st)
        if int(err.item()) != 0:
            raise _lib.EfbError("a node has more than 512 distinct neighbour nodes (kAdjCap)")
        self.max_deg = int(deg.max().item()) if Nn else 0
        self.adjptr = _scan(deg)
        self.nnz_node = int(self.adjptr[-1].item())
        self.adj = dv.empty((self.nnz_node,), torch.int32)
        _lib.call("efb_csr_fill_adj", n, ptrs, Ne, nPe, Nn, dv.ptr(self.rowptr), dv.ptr(self.qlist), dv.ptr(self.adjptr),
                  dv.ptr(self.adj), st)
        n_pos = sum(dg.Ne * dg.nPe * dg.nPe for dg in self.dgs)
        self.pos = dv.empty((n_pos,), torch.int32)
        _lib.call("efb_csr_slot_map", n, ptrs, Ne, nPe, dv.ptr(self.adjptr), dv.ptr(self.adj), dv.ptr(self.pos), st)
        self._connect_args = (n, ptrs, Ne, nPe, connects)

    def n_entries(self, dof_n: int) -> int:
        return sum(dg.Ne * (dg.nPe * dof_n) ** 2 for dg in self.dgs)


class DeviceCsr:
    """CSR matrix resident on the device (`indptr`, `indices`, `data` torch tensors).

    `node_graph = (adjptr, adj, dof_n, max_deg)` is set for matrices assembled here: the dof rows of a node are one contiguous
    block whose column structure is the node adjacency, which the solver's SpMV reads instead of `indices`."""

    def __init__(self, indptr, indices, data, shape, node_graph=None):
        self.indptr, self.indices, self.data, self.shape = indptr, indices, data, tuple(shape)
        self.node_graph = node_graph

    @property
    def nnz(self) -> int:
        return int(self.data.numel())

    @property
    def index_bytes(self) -> int:
        return 4 if self.indices.dtype == torch.int32 else 8

    def to_scipy(self):
        from scipy import sparse

        m = sparse.csr_matrix((dv.to_host(self.data), dv.to_host(self.indices), dv.to_host(self.indptr)), shape=self.shape)
        m.has_canonical_format = True  # canonical by construction, like _simu.py:1058-1059
        return m


class CsrPattern:
    """Canonical CSR pattern of one assembly key `(dof_n, isMatrix, Ndof, groups)` — `__Get_csr_map`, _simu.py:1062-1102."""

    def __init__(self, graph: NodeGraph, dof_n: int, Ndof: int, isMatrix: bool = True):
        self.graph, self.dof_n, self.Ndof, self.isMatrix = graph, int(dof_n), int(Ndof), bool(isMatrix)
        d, Nn = self.dof_n, graph.Nn
        if Ndof < Nn * d:
            raise ValueError("Ndof < Nn*dof_n")
        st = dv.stream_ptr()
        if isMatrix:
            n_entries = graph.n_entries(d)
            self.nnz = graph.nnz_node * d * d
            big = max(Ndof, n_entries) > _I32MAX  # scipy's rule (coo -> csr picks int64 when nnz or shape needs it)
            self.index_dtype = torch.int64 if big else torch.int32
            self.indptr = dv.empty((Ndof + 1,), self.index_dtype)
            self.indices = dv.empty((self.nnz,), self.index_dtype)
            _lib.call("efb_csr_expand", Nn, d, Ndof, graph.nnz_node, dv.ptr(graph.adjptr), dv.ptr(graph.adj),
                      8 if big else 4, dv.ptr(self.indptr), dv.ptr(self.indices), st)
        else:
            has = dv.empty((Ndof,), torch.int32)
            _lib.call("efb_csr_row_has_entry", Nn, d, Ndof, dv.ptr(graph.rowptr), dv.ptr(has), st)
            self.indptr64 = _scan(has)
            self.nnz = int(self.indptr64[-1].item())
            n_entries = sum(dg.Ne * dg.nPe * d for dg in graph.dgs)
            big = max(Ndof, n_entries) > _I32MAX
            self.index_dtype = torch.int64 if big else torch.int32
            self.indptr = self.indptr64.to(self.index_dtype)
            self.indices = torch.zeros(self.nnz, dtype=self.index_dtype, device=has.device)
        self.shape = (Ndof, Ndof) if isMatrix else (Ndof, 1)

    # -- values ------------------------------------------------------------------------------------------------
    def _data_args(self, datas):
        g = self.graph
        if len(datas) != len(g.dgs):
            raise ValueError("one data array per contributing group")
        tens = []
        for dg, X in zip(g.dgs, datas):
            ndof = dg.nPe * self.dof_n
            want = dg.Ne * ndof * (ndof if self.isMatrix else 1)
            t = dv.to_device(X)
            assert t.numel() == want, f"Not enough data to fill a {self.shape} CSR."
            tens.append(t)
        return _group_args(g.dgs, tens) + (tens,)

    def replay(self, datas, out=None, n_nodes=None) -> torch.Tensor:
        """CSR `data` (nnz) from per-group element arrays: np.bincount(inv, weights=concat(data)) bit for bit.
        `n_nodes` restricts the matrix replay to the rows of nodes [0, n_nodes) (owned rows of a sharded run)."""
        g = self.graph
        Nn = g.Nn if n_nodes is None else min(int(n_nodes), g.Nn)
        n, ptrs, Ne, nPe, keep = self._data_args(datas)
        st = dv.stream_ptr()
        if self.isMatrix:
            if out is None:
                out = dv.empty((self.nnz,))
            _lib.call("efb_csr_replay_matrix", n, ptrs, Ne, nPe, self.dof_n, Nn, dv.ptr(g.rowptr), dv.ptr(g.qlist),
                      dv.ptr(g.adjptr), dv.ptr(g.pos), g.max_deg, dv.ptr(out), st)
            return out
        dense = torch.zeros(self.Ndof, dtype=torch.float64, device=g.rowptr.device)
        _lib.call("efb_csr_replay_vector", n, ptrs, Ne, nPe, self.dof_n, g.Nn, dv.ptr(g.rowptr), dv.ptr(g.qlist), dv.ptr(dense), st)
        if out is None:
            out = dv.empty((self.nnz,))
        _lib.call("efb_csr_compact_rows", self.Ndof, dv.ptr(self.indptr64), dv.ptr(dense), dv.ptr(out), st)
        self.last_dense = dense
        return out

    @property
    def node_graph(self):
        """(adjptr, adj, dof_n, max_deg) of a matrix pattern whose rows are all node rows (no Lagrange rows), else None"""
        g = self.graph
        return (g.adjptr, g.adj, self.dof_n, g.max_deg) if self.isMatrix and self.Ndof == g.Nn * self.dof_n else None

    def assemble(self, datas) -> DeviceCsr:
        return DeviceCsr(self.indptr, self.indices, self.replay(datas), self.shape, self.node_graph)

    def inv_map(self) -> torch.Tensor:
        """The reference's `inv` array (int32, entry order k), for API parity / tests."""
        g = self.graph
        n, ptrs, Ne, nPe, _ = g._connect_args
        if self.isMatrix:
            inv = dv.empty((g.n_entries(self.dof_n),), torch.int32)
            _lib.call("efb_csr_inv_map", n, ptrs, Ne, nPe, self.dof_n, dv.ptr(g.adjptr), dv.ptr(g.pos), dv.ptr(inv), dv.stream_ptr())
            return inv
        parts = []
        d = self.dof_n
        for dg in g.dgs:  # inv of a vector entry = slot of its row
            rows = (dg.connect_glob.to(torch.int64)[:, :, None] * d + torch.arange(d, device=dg.connect_glob.device)).reshape(-1)
            parts.append(self.indptr64[rows].to(torch.int32))
        return torch.cat(parts)


class Assembler:
    """Caches node graphs and patterns per assembly key, like `@cache_computed_values` on `__Get_csr_map`."""

    def __init__(self):
        self._graphs = {}
        self._patterns = {}

    def clear(self):
        self._graphs.clear()
        self._patterns.clear()

    def pattern(self, dof_n, isMatrix, Ndof, groups) -> CsrPattern:
        gkey = tuple(id(g) for g in groups)
        d = int(dof_n)
        Nn = max(int(g.Ncoords) for g in groups)
        if (gkey, Nn) not in self._graphs:
            self._graphs[(gkey, Nn)] = (NodeGraph(groups, Nn), tuple(groups))  # keep the groups alive with their ids
        key = (gkey, Nn, d, bool(isMatrix), int(Ndof))
        if key not in self._patterns:
            self._patterns[key] = CsrPattern(self._graphs[(gkey, Nn)][0], d, Ndof, isMatrix)
        return self._patterns[key]

    def Get_csr_map(self, dof_n, isMatrix, Ndof, groups):
        """(inv, indices, indptr, nnz) as NumPy arrays — same tuple as `_Simu.__Get_csr_map`."""
        p = self.pattern(dof_n, isMatrix, Ndof, groups)
        return dv.to_host(p.inv_map()), dv.to_host(p.indices), dv.to_host(p.indptr), p.nnz

    def Assemble_csr(self, dict_group_data: dict, dof_n: int, Ndof: int, isMatrix: bool = True, as_device: bool = False):
        """`_Simu.__Assemble_csr` (_simu.py:989-1060): {group: X_e or None} -> scipy csr (or DeviceCsr)."""
        from scipy import sparse

        shape = (Ndof, Ndof) if isMatrix else (Ndof, 1)
        groups = tuple(g for g, X in dict_group_data.items() if X is not None) if dict_group_data else ()
        if not groups or sum(np.size(dict_group_data[g]) if not isinstance(dict_group_data[g], torch.Tensor)
                             else dict_group_data[g].numel() for g in groups) == 0:
            return sparse.csr_matrix(shape)
        pat = self.pattern(dof_n, isMatrix, Ndof, groups)
        datas = [dict_group_data[g] for g in groups]
        if any(np.iscomplexobj(X) if not isinstance(X, torch.Tensor) else X.dtype.is_complex for X in datas):
            # the reference bincounts the real and imaginary parts separately (_simu.py:1048-1053): two replays
            if as_device:
                raise TypeError("complex systems are assembled to scipy matrices only")
            host = [X.cpu().numpy() if isinstance(X, torch.Tensor) else np.asarray(X) for X in datas]
            re = dv.to_host(pat.replay([np.ascontiguousarray(X.real) for X in host]))
            im = dv.to_host(pat.replay([np.ascontiguousarray(X.imag) for X in host]))
            m = sparse.csr_matrix((re + 1j * im, dv.to_host(pat.indices), dv.to_host(pat.indptr)), shape=pat.shape)
            m.has_canonical_format = True
            return m
        A = pat.assemble(datas)
        return A if as_device else A.to_scipy()


# ---------------------------------------------------------------------------------------------------------
# fused element integration + assembly (homogeneous C): the one-time node schedule of `efb_assemble_elastic`
# ---------------------------------------------------------------------------------------------------------
def morton_order(xyz: torch.Tensor, dim: int, return_codes: bool = False):
    """Permutation that sorts points along a Z-order curve of a point lattice.  The lattice index of a point along an axis
    is its RANK along that axis divided into m_d equal-population slabs, m_d = the number of lattice planes a box with n
    points and one spacing h has along that axis (prod_d (ext_d / h + 1) = n): exact for structured meshes (jittered or
    not: every plane holds the same number of points), population-adaptive for unstructured ones.  Consecutive runs of 2^k
    points are then compact bricks.  `xyz` (n, >= dim) on any device; stable, deterministic."""
    n = xyz.shape[0]
    if n == 0:
        e = torch.empty(0, dtype=torch.int64, device=xyz.device)
        return (e, e) if return_codes else e
    x = xyz[:, :dim].to(torch.float64)
    lo, hi = x.min(0).values, x.max(0).values
    ext = [max(float(v), 0.0) for v in (hi - lo).tolist()]
    big = max(ext) if max(ext) > 0 else 1.0
    h_lo, h_hi = big * 1e-9, big * 2.0
    for _ in range(200):  # bisection on the spacing
        h = 0.5 * (h_lo + h_hi)
        cells = 1.0
        for v in ext:
            cells *= v / h + 1.0
        if cells > n:
            h_lo = h
        else:
            h_hi = h
    h = 0.5 * (h_lo + h_hi)
    nbits = 20 if dim == 3 else 30
    code = torch.zeros(n, dtype=torch.int64, device=xyz.device)
    ar = torch.arange(n, dtype=torch.int64, device=xyz.device)
    for d in range(dim):
        m = max(int(round(ext[d] / h)) + 1, 1)
        rank = torch.empty(n, dtype=torch.int64, device=xyz.device)
        rank[torch.argsort(x[:, d], stable=True)] = ar
        q = torch.clamp((rank * m) // n, 0, 2**nbits - 1)
        for bit in range(nbits):
            code |= ((q >> bit) & 1) << (bit * dim + d)
    order = torch.argsort(code, stable=True)
    return (order, code[order]) if return_codes else order


class FusedSchedule:
    """Node clusters + per-cluster element lists + per-task descriptors for `efb_assemble_elastic` (csrc/fused_kernels.cuh).
    Built once per (group, node graph) with torch tensor ops on the device the graph lives on — preprocessing, like the CSR
    pattern itself; nothing here runs per assembly."""

    def __init__(self, graph: NodeGraph, n_nodes: int = None, S: int = None, smem_budget: int = 110 * 1024, nPg: int = None):
        if len(graph.dgs) != 1:
            raise NotImplementedError("the fused assembly handles one element group")
        dg = graph.dgs[0]
        self.graph, self.dg = graph, dg
        dim, nPe, dev = dg.dim, dg.nPe, graph.rowptr.device
        n_nodes = graph.Nn if n_nodes is None else min(int(n_nodes), graph.Nn)
        self.n_nodes, self.dof_n = n_nodes, dim
        self.nPg = int(dg.nPg("rigi") if nPg is None else nPg)
        G = _lib.load().efb_assemble_elastic_group(dim, nPe, self.nPg)  # nodes per warp of the kernel instantiation
        if G < 1:
            raise NotImplementedError(f"no fused assembly kernel for dim={dim}, nPe={nPe}")
        rowptr, adjptr = graph.rowptr, graph.adjptr
        cnt_all = (rowptr[1:n_nodes + 1] - rowptr[:n_nodes])
        # coordinates per GLOBAL node id (the group stores local rows): scatter through the two connectivities
        xyz = torch.zeros((graph.Nn, 3), dtype=torch.float64, device=dev)
        xyz[dg.connect_glob.reshape(-1).long()] = dg.coord[dg.connect.reshape(-1).long()][:, :3]
        used = torch.nonzero(cnt_all > 0).reshape(-1)
        perm, codes = morton_order(xyz[used], dim, return_codes=True)
        order = used[perm]
        self.order = order
        nn = int(order.numel())
        cnt = cnt_all[order]
        tptr = torch.zeros(nn + 1, dtype=torch.int64, device=dev)
        torch.cumsum(cnt, 0, out=tptr[1:])
        n_tasks = int(tptr[-1].item()) if nn else 0
        knode = torch.repeat_interleave(torch.arange(nn, device=dev), cnt)           # ordered-node index of every task
        s = rowptr[order][knode] + (torch.arange(n_tasks, device=dev) - tptr[knode])  # its index in qlist
        q = graph.qlist[s]
        e, a = q // nPe, q % nPe
        Ne = dg.Ne
        max_deg = max(graph.max_deg, 1)
        ar = torch.arange(nn, device=dev)

        def clusters(Sc):
            """clusters = the Morton cells of log2(Sc) interleaved bits (a compact brick of the lattice each), cut into runs of at
            most Sc nodes where a cell holds more: (cluster id, position inside the cluster) of every ordered node"""
            shift = max(int(Sc).bit_length() - 1, 0)
            gk = codes >> shift
            newg = torch.ones(nn, dtype=torch.bool, device=dev)
            newg[1:] = gk[1:] != gk[:-1]
            gstart = torch.cummax(torch.where(newg, ar, torch.zeros_like(ar)), 0).values
            pig = ar - gstart
            start = newg | (pig % Sc == 0)
            return torch.cumsum(start.to(torch.int64), 0) - 1, pig % Sc

        # cluster size: the largest power-of-two multiple of G (at most 16 warps) whose biggest cluster fits the budget
        cands = [S] if S else [c * G for c in (16, 8, 4, 2, 1) if c * G <= 64]
        chosen = None
        for Sc in cands:
            if Sc & (Sc - 1):
                raise ValueError("the cluster size must be a power of two")
            cl_node, pos_node = clusters(Sc)
            ncl = int(cl_node[-1].item()) + 1 if nn else 0
            cl = cl_node[knode]
            ukey, inv = torch.unique(cl * Ne + e, sorted=True, return_inverse=True)
            cl_eptr = torch.searchsorted(ukey // Ne, torch.arange(ncl + 1, device=dev)).to(torch.int64)
            cap_e = int((cl_eptr[1:] - cl_eptr[:-1]).max().item()) if ncl else 0
            chosen = (Sc, cl_node, pos_node, cl, ukey, inv, ncl, cl_eptr, cap_e)
            if S or self._fits(dim, nPe, self.nPg, Sc, cap_e, max_deg, smem_budget):
                break
        Sc, cl_node, pos_node, cl, ukey, inv, ncl, cl_eptr, cap_e = chosen
        self.S, self.n_clusters, self.cap_e, self.max_deg = Sc, ncl, cap_e, max_deg
        le = inv - cl_eptr[cl]
        assert cap_e < 65536 and nPe < 256
        self.desc = (le | (a << 16)).to(torch.int32)
        self.tpos = graph.pos.reshape(-1, nPe)[q].contiguous()                       # (n_tasks, nPe) int32
        # coordinate rows of the clusters' elements, one fixed-size slab per cluster (-1 = empty): the kernel finds its slab
        # from the CTA index alone, so the gather is a two-level chain (connectivity -> coordinates)
        ucl = ukey // Ne
        conn = torch.full((ncl * cap_e, nPe), -1, dtype=torch.int32, device=dev)
        conn[ucl * cap_e + (torch.arange(ukey.numel(), device=dev) - cl_eptr[ucl])] = dg.connect.reshape(-1, nPe)[ukey % Ne]
        self.cl_conn = conn
        self.cl_ne = (cl_eptr[1:] - cl_eptr[:-1]).to(torch.int32).contiguous()
        self.n_integrated = int(ukey.numel())
        d2 = dim * dim
        nodes = torch.full((ncl * Sc, 4), -1, dtype=torch.int64, device=dev)
        nodes[:, 1:] = 0
        k = cl_node * Sc + pos_node
        nodes[k, 0] = order
        nodes[k, 1] = d2 * adjptr[order]
        nodes[k, 2] = (adjptr[order + 1] - adjptr[order]) | (cnt << 32)
        nodes[k, 3] = tptr[:-1]
        self.cl_nodes = nodes.contiguous()
        self.n_tasks = n_tasks

    @staticmethod
    def _fits(dim, nPe, nPg, S, cap_e, max_deg, budget) -> bool:
        need = smem_bytes(dim, nPe, nPg, S, cap_e, max_deg)
        return need is not None and need <= budget

    def redundancy(self) -> float:
        """elements integrated per launch / elements of the group (each element is integrated once per cluster touching it)"""
        return float(self.n_integrated) / max(self.dg.Ne, 1)


def smem_bytes(dim, nPe, nPg, S, cap_e, max_deg):
    """dynamic shared memory of one CTA of the fused kernel (`efb_assemble_elastic_smem`), None without an instantiation"""
    n = _lib.load().efb_assemble_elastic_smem(int(dim), int(nPe), int(nPg), int(S), int(cap_e), int(max_deg))
    return None if n < 0 else int(n)


def assemble_elastic_fused(sched: FusedSchedule, C, matrixType="rigi", scale: float = 1.0, out: torch.Tensor = None):
    """CSR `data` of K = Assembly(LinearizedElasticity(group, C)) for a homogeneous C, element matrices never materialised
    (`efb_assemble_elastic`).  Rows of the scheduled nodes are written; returns `out` (nnz)."""
    dg, g = sched.dg, sched.graph
    mt = getattr(matrixType, "value", str(matrixType))
    C = np.ascontiguousarray(np.asarray(C, dtype=np.float64))
    ns = 3 if dg.dim == 2 else 6
    if C.shape != (ns, ns):
        raise ValueError(f"the fused assembly needs a homogeneous ({ns}, {ns}) C; got {C.shape}")
    if out is None:
        out = dv.empty((g.nnz_node * dg.dim * dg.dim,))
    w = np.ascontiguousarray(dv.to_host(dg.tables(mt)[2]))
    _lib.call("efb_assemble_elastic", dg.cstruct(mt), ctypes.c_void_p(C.ctypes.data), ctypes.c_void_p(w.ctypes.data), float(scale),
              sched.n_clusters, sched.S, sched.cap_e, sched.max_deg, dv.ptr(sched.cl_nodes), dv.ptr(sched.cl_ne),
              dv.ptr(sched.cl_conn), dv.ptr(sched.desc), dv.ptr(sched.tpos), dv.ptr(out), dv.stream_ptr())
    return out


# ---------------------------------------------------------------------------------------------------------
# HEXA8 / 8 Gauss points on the FP64 MMA instruction: staging slots + per-cluster gather programs of `efb_assemble_elastic_mma`
# ---------------------------------------------------------------------------------------------------------
class MmaSchedule:
    """One-time schedule of `efb_assemble_elastic_mma` (csrc/fused_mma.cu) on top of a `FusedSchedule` with clusters of 16
    nodes: element slabs padded to passes of 4, `rowslot` (staging slot | node swizzle of every (cluster element, local node)
    whose node the cluster owns), round headers and the gather programs (every CSR block of a cluster -> one lane of one
    round, rounds sorted by the number of contributions, each block's contributions in ascending element order).  torch
    tensor ops on the schedule's device — preprocessing."""

    S = 16
    TS = 72            # doubles per staged task (kMmaTS)
    MAX_TRIPS = 255

    @staticmethod
    def supports(dg, nPg) -> bool:
        return dg.dim == 3 and dg.nPe == 8 and int(nPg) == 8

    def __init__(self, graph: NodeGraph, n_nodes: int = None, base: "FusedSchedule" = None, nPg: int = None):
        sched = base if base is not None and base.S == self.S else FusedSchedule(graph, n_nodes=n_nodes, S=self.S, nPg=nPg)
        if not self.supports(sched.dg, sched.nPg):
            raise NotImplementedError("efb_assemble_elastic_mma serves HEXA8 with 8 Gauss points")
        self.base, self.graph, self.dg = sched, sched.graph, sched.dg
        S, nPe, cap_e, ncl = self.S, 8, sched.cap_e, sched.n_clusters
        cap4 = max((cap_e + 3) // 4 * 4, 4)
        dev = sched.cl_nodes.device
        nodes = sched.cl_nodes
        i64 = torch.int64
        ar = lambda n: torch.arange(n, device=dev, dtype=i64)  # noqa: E731
        kidx = torch.nonzero(nodes[:, 0] >= 0).reshape(-1)                     # valid cluster slots, = ordered nodes in order
        cnt = nodes[kidx, 2] >> 32
        deg = nodes[kidx, 2] & 0xffffffff
        n_tasks = sched.n_tasks
        ki_t = torch.repeat_interleave(ar(kidx.numel()), cnt)
        kt = kidx[ki_t]                                                        # cluster slot of every task
        cl, iloc_t = kt // S, kt % S
        cl_task0 = torch.searchsorted(cl, ar(ncl + 1))
        self.t_cap = int((cl_task0[1:] - cl_task0[:-1]).max().item()) if ncl else 0
        # staging slot: tasks of a cluster in order, the tasks of nodes 8..15 swapped pairwise (slot parity then differs
        # between node i and node i + 8: with the column swizzle b ^ (i & 7) the 16 nodes of a cluster sit in 16 distinct
        # 8-byte bank classes, so a gather round whose lanes do the same thing for different nodes is conflict-free)
        node_t0 = torch.zeros(kidx.numel() + 1, dtype=i64, device=dev)
        torch.cumsum(cnt, 0, out=node_t0[1:])
        j = ar(n_tasks) - node_t0[ki_t]
        jswap = torch.where((j ^ 1) < cnt[ki_t], j ^ 1, j)
        j2 = torch.where(((iloc_t >> 3) & 1) == 1, jswap, j)
        t_loc = node_t0[ki_t] + j2 - cl_task0[cl]
        xs_t = iloc_t & 7
        desc = sched.desc.to(i64)
        le, a = desc & 0xffff, desc >> 16
        rowslot = torch.full((ncl * cap4 * nPe,), -1, dtype=torch.int16, device=dev)
        rowslot[(cl * cap4 + le) * nPe + a] = (t_loc | (xs_t << 8)).to(torch.int16)
        # pairs of local nodes (2j, 2j+1) with an owned node: the row tiles the kernel computes
        self.n_row_tiles = int((rowslot.view(torch.int32) != -1).sum().item())
        # ---- blocks of every cluster, sorted by (cluster, contributions descending, slot, node in cluster) ----
        bptr = torch.zeros(kidx.numel() + 1, dtype=i64, device=dev)
        torch.cumsum(deg, 0, out=bptr[1:])
        nb = int(bptr[-1].item())
        bk = torch.repeat_interleave(ar(kidx.numel()), deg)                    # index into kidx of every block
        slot = ar(nb) - bptr[bk]
        node_blk0 = nodes[kidx, 1] // 9                                        # adjptr[node]: global id of the node's first block
        gb = node_blk0[bk] + slot
        gb_c = (node_blk0[ki_t].reshape(-1, 1) + sched.tpos.to(i64)).reshape(-1)   # contributions (task, b) -> global block id
        nnzb = int(graph.nnz_node)
        cntb = torch.bincount(gb_c, minlength=nnzb)
        cb = cntb[gb]
        maxc = int(cb.max().item()) if nb else 0
        if maxc > self.MAX_TRIPS:
            raise NotImplementedError("a CSR block with more than 255 element contributions")
        bcl = kidx[bk] // S
        iloc = kidx[bk] % S
        key = ((bcl * (maxc + 1) + (maxc - cb)) * (sched.max_deg + 1) + slot) * S + iloc
        border = torch.argsort(key, stable=True)
        del key
        bcl_s, cb_s = bcl[border], cb[border]
        cl_blk0 = torch.searchsorted(bcl_s, ar(ncl + 1))
        rank = ar(nb) - cl_blk0[bcl_s]
        R_c = (cl_blk0[1:] - cl_blk0[:-1] + 31) // 32
        self.rmax = max(int(R_c.max().item()) if ncl else 1, 1)
        rptr = torch.zeros(ncl + 1, dtype=i64, device=dev)
        torch.cumsum(R_c, 0, out=rptr[1:])
        n_rounds = int(rptr[-1].item())
        rcl = torch.repeat_interleave(ar(ncl), R_c)                            # cluster of every round
        rin = ar(n_rounds) - rptr[rcl]                                         # round index inside its cluster
        c_round = cb_s[cl_blk0[rcl] + rin * 32]                                # trip count = contributions of its first block
        cpad = torch.where(c_round <= 1, 1, torch.where(c_round <= 2, 2, torch.where(c_round <= 4, 4, (c_round + 7) // 8 * 8)))
        rwords = 32 + 16 * cpad
        cs = torch.zeros(n_rounds + 1, dtype=i64, device=dev)
        torch.cumsum(rwords, 0, out=cs[1:])
        prog_off = cs[rptr]                                                    # (ncl + 1) word offset of every cluster program
        roff_in = cs[:-1] - prog_off[rcl]                                      # offset of the round inside the cluster program
        total = int(cs[-1].item())
        # one record per cluster (streamed into shared memory by one bulk copy): conn | rowslot | nodes | hdr
        o_rs, o_nodes = cap4 * 8, cap4 * 12
        o_hdr = o_nodes + 4 * S
        self.rec_words = (o_hdr + 1 + self.rmax + 3) // 4 * 4
        recs = torch.zeros((ncl, self.rec_words), dtype=torch.int32, device=dev)
        recs[:, :o_rs] = -1
        recs[:, :o_rs].view(ncl, cap4, nPe)[:, :cap_e] = sched.cl_conn.reshape(ncl, cap_e, nPe)
        recs[:, o_rs:o_nodes] = rowslot.view(torch.int32).reshape(ncl, cap4 * 4)
        nrec = torch.stack([torch.where(nodes[:, 0] < 0, torch.full_like(nodes[:, 1], -1), nodes[:, 1]), nodes[:, 2] & 0xffffffff], 1)
        recs[:, o_nodes:o_hdr] = nrec.contiguous().view(torch.int32).reshape(ncl, 4 * S)
        recs[:, o_hdr] = R_c.to(torch.int32)
        recs[rcl, o_hdr + 1 + rin] = ((roff_in << 8) | c_round).to(torch.int32)
        self.recs = recs.contiguous()
        self.pw_max = int((prog_off[1:] - prog_off[:-1]).max().item()) if ncl else 4
        prog = torch.full((total + 4,), -1, dtype=torch.int32, device=dev)     # dest -1 = no block
        # sources of lanes without a contribution: the zero block behind the staged tasks
        zsrc = self.t_cap * self.TS
        zw = torch.tensor(zsrc | (zsrc << 16), dtype=torch.int64, device=dev).to(torch.int32)
        smask = torch.zeros(total + 4, dtype=torch.bool, device=dev)
        sidx = (cs[:-1] + 32).repeat_interleave(16 * cpad) + (ar(int((16 * cpad).sum().item())) - torch.repeat_interleave(
            torch.cumsum(16 * cpad, 0) - 16 * cpad, 16 * cpad))
        smask[sidx] = True
        prog[smask] = zw
        del smask, sidx
        bround = rptr[bcl_s] + rank // 32                                      # global round of every sorted block
        blane = rank % 32
        prog[cs[bround] + blane] = ((iloc[border] << 16) | slot[border]).to(torch.int32)
        bpos = torch.full((nnzb,), -1, dtype=i64, device=dev)                  # position of every global block in the sorted order
        bpos[gb[border]] = ar(nb)
        del border
        # contributions: rank inside their block in ascending task (= element) order
        corder = torch.argsort(gb_c, stable=True)
        bstart = torch.zeros(nnzb + 1, dtype=i64, device=dev)
        torch.cumsum(cntb, 0, out=bstart[1:])
        it = torch.empty_like(gb_c)
        it[corder] = ar(gb_c.numel()) - bstart[gb_c[corder]]
        del corder, bstart
        pos_c = bpos[gb_c]
        rc = bround[pos_c]
        b8 = ar(8)
        src = (t_loc.reshape(-1, 1) * self.TS + (b8.reshape(1, -1) ^ xs_t.reshape(-1, 1))).reshape(-1)
        u16 = prog.view(torch.int16)
        u16[(cs[rc] + 32) * 2 + blane[pos_c] * cpad[rc] + it] = src.to(torch.int16)
        self.prog, self.prog_off = prog, prog_off.contiguous()
        self.n_rounds, self.n_blocks = n_rounds, nb
        self.n_clusters, self.cap_e, self.cap4 = ncl, cap_e, cap4

    def smem_bytes(self) -> int:
        return int(_lib.load().efb_assemble_elastic_mma_smem(int(self.t_cap), int(self.rec_words), int(self.pw_max)))

    def fits(self) -> bool:
        return 0 < self.t_cap <= 256 and self.t_cap * self.TS + 8 < 65535 and self.smem_bytes() <= 227 * 1024

    def redundancy(self) -> float:
        return self.base.redundancy()


def assemble_elastic_mma(ms: MmaSchedule, C, matrixType="rigi", scale: float = 1.0, out: torch.Tensor = None):
    """CSR `data` of K = Assembly(LinearizedElasticity(group, C)) for HEXA8 / 8 Gauss points and a homogeneous C through
    `efb_assemble_elastic_mma` (FP64 MMA form of the fused assembly).  Rows of the scheduled nodes are written."""
    sched, dg, g = ms.base, ms.dg, ms.graph
    mt = getattr(matrixType, "value", str(matrixType))
    C = np.ascontiguousarray(np.asarray(C, dtype=np.float64))
    if C.shape != (6, 6):
        raise ValueError(f"the fused assembly needs a homogeneous (6, 6) C; got {C.shape}")
    if out is None:
        out = dv.empty((g.nnz_node * 9,))
    w = np.ascontiguousarray(dv.to_host(dg.tables(mt)[2]))
    _lib.call("efb_assemble_elastic_mma", dg.cstruct(mt), ctypes.c_void_p(C.ctypes.data), ctypes.c_void_p(w.ctypes.data), float(scale),
              ms.n_clusters, ms.cap4, ms.t_cap, ms.rmax, ms.rec_words, ms.pw_max, dv.ptr(ms.recs), dv.ptr(ms.prog_off), dv.ptr(ms.prog),
              dv.ptr(out), dv.stream_ptr())
    return out
