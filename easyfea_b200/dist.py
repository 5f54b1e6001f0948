"""Element-wise mesh partition across the GPUs of one box, owned-row CSR blocks and the PCG communication pattern
(the reference's scheme is gmsh partitions + one ghost-element layer, `FEM/_mesher.py:2303-2393`, with
PETSc `mpiaij` rows per rank, `Simulations/Solvers.py:732-847`).

One process per GPU.  Elements are split into `world` contiguous chunks; a node belongs to the LOWEST rank whose
chunk touches it (the reference's greedy rule, `_mesher.py:2355-2360`).  Rank r integrates every element that touches
one of its owned nodes — its own chunk's elements plus the interface ("ghost") elements of higher chunks, which are
therefore integrated twice, as in the reference — and assembles only the CSR rows of its owned dofs.  Local node
numbering is `[owned nodes, ascending global id | halo nodes, grouped by owner rank, ascending global id]`, so

  * the owned rows of the locally assembled matrix are a PREFIX of its CSR arrays (no extraction step),
  * every slot is still summed in ascending global element order (local elements stay sorted by global id), i.e. the
    owned rows are bit-identical to the same rows of a single-GPU assembly, up to the column permutation,
  * a halo exchange receives straight into contiguous segments of the local vector; only the send side is packed.

Data-path collectives: none during element integration / assembly.  The Jacobi-PCG consumer needs one halo exchange
(grouped NCCL send/recv) of the search direction and two scalar all-reduces per iteration.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from . import device as dv


# ---------------------------------------------------------------------------------------------------------
# host-side partition (NumPy; unit-tested on CPU)
# ---------------------------------------------------------------------------------------------------------
def chunk_bounds(Ne: int, world: int) -> np.ndarray:
    """Element chunk r = [b[r], b[r+1]): contiguous, sizes differ by at most one."""
    return (np.arange(world + 1, dtype=np.int64) * Ne) // world


def node_owners(connect: np.ndarray, Nn: int, world: int) -> np.ndarray:
    """owner[n] = lowest rank whose element chunk touches node n; -1 for orphan nodes (they have no rows to assemble)."""
    Ne = connect.shape[0]
    b = chunk_bounds(Ne, world)
    erank = np.searchsorted(b, np.arange(Ne), side="right") - 1
    owner = np.full(Nn, world, dtype=np.int64)
    np.minimum.at(owner, connect.ravel(), np.repeat(erank, connect.shape[1]))
    owner[owner == world] = -1
    return owner


@dataclass
class Partition:
    """What rank `rank` of `world` holds.  All index arrays are int64 NumPy arrays on the host."""

    rank: int
    world: int
    elem_ids: np.ndarray        # (Ne_loc,) global ids of the local elements, ascending
    n_own_elems: int            # how many of them belong to this rank's own chunk (the rest are ghosts)
    connect: np.ndarray         # (Ne_loc, nPe) LOCAL node ids
    nodes: np.ndarray           # (n_local,) global id of every local node: [owned | halo]
    n_owned: int
    halo_ranks: np.ndarray      # (k,) owner ranks of the halo segments, ascending
    halo_ptr: np.ndarray        # (k+1,) halo segment of halo_ranks[i] = nodes[n_owned + halo_ptr[i] : n_owned + halo_ptr[i+1]]
    owned_offset: int = 0       # number of nodes owned by lower ranks (rank-contiguous global numbering)
    n_global: int = 0
    send: dict = field(default_factory=dict)  # neighbour rank -> LOCAL ids (< n_owned) of the owned nodes it needs

    @property
    def n_local(self) -> int:
        return int(self.nodes.size)

    @property
    def n_halo(self) -> int:
        return self.n_local - self.n_owned

    # -- constructors ----------------------------------------------------------------------------------------
    @classmethod
    def from_candidates(cls, connect: np.ndarray, elem_ids: np.ndarray, owner_of, rank: int, world: int, own_chunk=None,
                        n_global: int = 0):
        """`connect` (global node ids) and ascending `elem_ids` of a SUPERSET of the rank's local elements;
        `owner_of(ids) -> ranks` gives the owner of global nodes.  Keeps the elements touching an owned node."""
        connect = np.asarray(connect, dtype=np.int64)
        elem_ids = np.asarray(elem_ids, dtype=np.int64)
        assert elem_ids.size < 2 or np.all(np.diff(elem_ids) > 0), "local elements must stay in ascending global order"
        uniq, inv = np.unique(connect.ravel(), return_inverse=True)
        own_u = np.asarray(owner_of(uniq), dtype=np.int64)
        keep = (own_u[inv].reshape(connect.shape) == rank).any(axis=1)
        connect, elem_ids = connect[keep], elem_ids[keep]
        uniq, inv = np.unique(connect.ravel(), return_inverse=True)
        own_u = np.asarray(owner_of(uniq), dtype=np.int64)
        # local order: owned first (ascending id), then halo by (owner, id); `uniq` is already ascending
        key = np.where(own_u == rank, -1, own_u)
        order = np.argsort(key, kind="stable")
        nodes = uniq[order]
        new_of_uniq = np.empty_like(order)
        new_of_uniq[order] = np.arange(order.size)
        n_owned = int((own_u == rank).sum())
        halo_owner = own_u[order][n_owned:]
        halo_ranks, first = np.unique(halo_owner, return_index=True)
        halo_ptr = np.append(first, halo_owner.size).astype(np.int64)
        n_own_elems = int(keep.sum()) if own_chunk is None else int(((elem_ids >= own_chunk[0]) & (elem_ids < own_chunk[1])).sum())
        return cls(rank, world, elem_ids, n_own_elems, new_of_uniq[inv].reshape(connect.shape), nodes, n_owned,
                   halo_ranks.astype(np.int64), halo_ptr, 0, int(n_global))

    @classmethod
    def from_global(cls, connect: np.ndarray, Nn: int, world: int, rank: int):
        """Every rank sees the whole connectivity (small / medium meshes, tests)."""
        connect = np.asarray(connect, dtype=np.int64)
        owner = node_owners(connect, Nn, world)
        b = chunk_bounds(connect.shape[0], world)
        part = cls.from_candidates(connect, np.arange(connect.shape[0]), lambda ids: owner[ids], rank, world,
                                   own_chunk=(b[rank], b[rank + 1]), n_global=Nn)
        part.owned_offset = int((owner[owner >= 0] < rank).sum())
        # with the whole mesh at hand the send lists need no communication
        for q in range(world):
            if q == rank:
                continue
            touched = np.unique(connect[(owner[connect] == q).any(axis=1)].ravel())
            need = touched[owner[touched] == rank]  # my nodes that rank q holds as halo
            if need.size:
                part.send[q] = np.searchsorted(part.nodes[:part.n_owned], need)
        return part

    def halo_ids(self, q: int) -> np.ndarray:
        """global ids of the halo nodes owned by rank q (ascending)"""
        i = int(np.searchsorted(self.halo_ranks, q))
        if i >= self.halo_ranks.size or self.halo_ranks[i] != q:
            return np.empty(0, dtype=np.int64)
        return self.nodes[self.n_owned + self.halo_ptr[i]: self.n_owned + self.halo_ptr[i + 1]]

    def plan_exchange(self, group=None) -> None:
        """Fill `send` / `owned_offset` by talking to the other ranks (each rank only knows its own halo needs)."""
        import torch.distributed as dist

        needs = {int(q): self.halo_ids(int(q)) for q in self.halo_ranks}
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (self.n_owned, needs), group=group)
        self.owned_offset = int(sum(g[0] for g in gathered[:self.rank]))
        self.n_global = max(self.n_global, int(sum(g[0] for g in gathered)))
        self.send = {}
        own = self.nodes[:self.n_owned]
        for q, (_, their) in enumerate(gathered):
            if q != self.rank and self.rank in their and their[self.rank].size:
                loc = np.searchsorted(own, their[self.rank])
                assert np.array_equal(own[loc], their[self.rank]), "a neighbour asks for nodes this rank does not own"
                self.send[q] = loc


# ---------------------------------------------------------------------------------------------------------
# communication
# ---------------------------------------------------------------------------------------------------------
def _pack_cuda(src: torch.Tensor, idx: torch.Tensor, dst: torch.Tensor) -> None:
    """dst[i] = src[idx[i]] on the device (C ABI, no torch indexing on the hot path)"""
    if not src.is_cuda:
        raise _lib.EfbError("RowComm packs on the device; pass `gather=` explicitly for host tensors (tests only)")
    _lib.call("efb_pack_f64", idx.numel(), dv.ptr(idx), dv.ptr(src), dv.ptr(dst), dv.stream_ptr())


class RowComm:
    """Halo exchange + scalar all-reduce for vectors laid out `[owned | halo]` with `dof_n` values per node.

    `gather(src, idx, dst)` fills the send buffer; the default is the CUDA pack kernel.  With backend `nccl` the
    sends/receives of one exchange are issued as one group (`batch_isend_irecv`), over NVLink on one box."""

    def __init__(self, part: Partition, dof_n: int, device=None, group=None, gather=None):
        import torch.distributed as dist

        self.dist, self.group, self.part, self.d = dist, group, part, int(dof_n)
        self.gather = gather or _pack_cuda
        d = self.d
        self.device = torch.device(device) if device is not None else dv.device()
        self.peers_send = sorted(part.send)
        idx, self.send_ptr = [], [0]
        for q in self.peers_send:
            loc = np.asarray(part.send[q], dtype=np.int64)
            idx.append((loc[:, None] * d + np.arange(d)[None, :]).ravel())
            self.send_ptr.append(self.send_ptr[-1] + loc.size * d)
        cat = np.concatenate(idx) if idx else np.empty(0, dtype=np.int64)
        assert cat.size == 0 or cat.max() < 2**31
        self.send_idx = torch.from_numpy(cat.astype(np.int32)).to(self.device)
        self.send_buf = torch.empty(cat.size, dtype=torch.float64, device=self.device)
        self.recv_segments = [(int(q), (part.n_owned + int(part.halo_ptr[i])) * d, (part.n_owned + int(part.halo_ptr[i + 1])) * d)
                              for i, q in enumerate(part.halo_ranks)]
        self.bytes_per_exchange = 8 * (cat.size + part.n_halo * d)

    def halo_exchange(self, x: torch.Tensor) -> None:
        """Refresh the halo part of `x` (n_local*dof_n) from the ranks that own those nodes."""
        dist = self.dist
        assert x.numel() == self.part.n_local * self.d
        if not self.peers_send and not self.recv_segments:
            return
        if self.send_idx.numel():
            self.gather(x, self.send_idx, self.send_buf)
        ops = []
        for i, q in enumerate(self.peers_send):
            ops.append(dist.P2POp(dist.isend, self.send_buf[self.send_ptr[i]:self.send_ptr[i + 1]], q, group=self.group))
        for q, lo, hi in self.recv_segments:
            ops.append(dist.P2POp(dist.irecv, x[lo:hi], q, group=self.group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    # -- peer-memory PCG workspace (efb_pcg_iterate) -------------------------------------------------------------
    def pcg_workspace(self):
        """Control block + the two search-direction buffers of this rank inside a cudaIpc-shared region, mapped on every
        rank of the group, and the push plan of the halo exchange (built once per communicator, collective)."""
        if getattr(self, "_ws", None) is None:
            self._ws = PeerWorkspace(self)
        return self._ws

    def all_reduce_sum(self, t: torch.Tensor) -> None:
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)

    def all_reduce_max(self, t: torch.Tensor) -> None:
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)


# ---------------------------------------------------------------------------------------------------------
# sharded system: local group, owned-row CSR block, distributed solve
# ---------------------------------------------------------------------------------------------------------
class ShardedGroup:
    """The rank's elements as an `ElemGroup` in LOCAL node numbering + the assembly pattern of its owned rows."""

    def __init__(self, elemType: str, part: Partition, coords_local: np.ndarray):
        from .assembly import Assembler
        from .mesh import ElemGroup

        assert coords_local.shape[0] == part.n_local
        self.part = part
        self.group = ElemGroup(elemType, part.connect, coords_local, all_nodes_used=True)
        self.assembler = Assembler()

    def owned_block(self, Xe, dof_n: int, out=None):
        """Assemble the element matrices `Xe` (Ne_loc, ndof, ndof) and return the owned-row block as a DeviceCsr of shape
        (n_owned*dof_n, n_local*dof_n) with LOCAL column ids (views on the full local arrays, no copy)."""
        from .assembly import DeviceCsr

        part, d = self.part, int(dof_n)
        pat = self.assembler.pattern(d, True, part.n_local * d, (self.group,))
        data = pat.replay([Xe], out=out, n_nodes=part.n_owned)
        nrows = part.n_owned * d
        if not hasattr(pat, "_nnz_owned"):
            pat._nnz_owned = int(pat.indptr[nrows].item())
        nz = pat._nnz_owned
        return DeviceCsr(pat.indptr[:nrows + 1], pat.indices[:nz], data[:nz], (nrows, part.n_local * d), pat.node_graph)

    def owned_vector(self, Fe, dof_n: int):
        """dense owned part (n_owned*dof_n) of the assembled element vectors `Fe` (Ne_loc, ndof)"""
        part, d = self.part, int(dof_n)
        pat = self.assembler.pattern(d, False, part.n_local * d, (self.group,))
        pat.replay([Fe])
        return pat.last_dense[:part.n_owned * d]


# ---------------------------------------------------------------------------------------------------------
# peer-memory workspace of the fused PCG iterations
# ---------------------------------------------------------------------------------------------------------
def _pad256(nbytes: int) -> int:
    return (int(nbytes) + 255) // 256 * 256


def peer_push_plan(comm: "RowComm", recv_lo_of_rank):
    """Who stores what where in the peer-memory halo exchange: for every neighbour this rank sends to, the range of
    `comm.send_idx` it gets and the first entry of this rank's segment inside THAT rank's `[owned | halo]` vector
    (`recv_lo_of_rank[q]` = {source rank: first entry of its segment on rank q}, gathered from all ranks).
    Returns (send_rank, send_ptr, send_dst, recv_rank) as Python lists."""
    rank = comm.part.rank
    send_rank = [int(q) for q in comm.peers_send]
    send_ptr = [int(v) for v in comm.send_ptr]
    send_dst = []
    for q in send_rank:
        their = recv_lo_of_rank[q]
        assert rank in their, "a neighbour this rank sends to does not expect a segment from it"
        send_dst.append(int(their[rank]))
    recv_rank = [int(q) for q, _, _ in comm.recv_segments]
    return send_rank, send_ptr, send_dst, recv_rank


class LocalWorkspace:
    """Single-GPU workspace of `efb_pcg_iterate`: control block + two p buffers in ordinary device memory."""

    def __init__(self, n: int):
        import ctypes

        lib = _lib.load()
        self.n = int(n)
        ctrl = lib.efb_pcg_ctrl_bytes()
        lay = (ctypes.c_int32 * 5)()
        lib.efb_pcg_ctrl_layout(lay)
        self.rz_off, self.rr_off, self.err_off, self.iters_off = int(lay[0]), int(lay[1]), int(lay[2]), int(lay[3])
        self.cg2_init_off = int(lay[4])
        self.ctrl_bytes = ctrl
        self.pbuf_off = (ctrl, ctrl + _pad256(self.n * 8))
        self.nbytes = ctrl + 2 * _pad256(self.n * 8)
        self._alloc()
        self.p = [self.f64[o // 8: o // 8 + self.n] for o in self.pbuf_off]
        self.ctrl = self.f64[: ctrl // 8]
        self.peer = _lib.EfbPcgPeer()
        self._fill_peer()

    def _alloc(self):
        self.f64 = torch.zeros(self.nbytes // 8, dtype=torch.float64, device=dv.device())
        self.base = self.f64.data_ptr()

    def _fill_peer(self):
        P = self.peer
        P.world, P.rank, P.n_send, P.n_recv = 1, 0, 0, 0
        P.base[0] = self.base
        P.pbuf_off[0][0], P.pbuf_off[0][1] = self.pbuf_off
        P.send_idx = None
        P.ar_seq, P.halo_seq = 0, 0

    def status(self):
        """(r.r of the last finished iteration, error flag, iterations of the last persistent launch) — one small copy"""
        c = self.ctrl.cpu().numpy()
        return float(c[self.rr_off]), int(c.view(np.uint32)[self.err_off]), int(c.view(np.uint64)[self.iters_off])

    def advance(self, n_iters: int, reductions_per_iter: int = 2):
        self.peer.ar_seq += reductions_per_iter * n_iters
        self.peer.halo_seq += n_iters


class PeerWorkspace(LocalWorkspace):
    """Row-sharded workspace: the region is plain cudaMalloc memory exported with cudaIpc, every rank maps every other
    rank's region, and the iteration kernels store reductions / interface entries straight into the neighbours' regions
    (NVLink).  Construction is collective over the communicator's group."""

    def __init__(self, comm: RowComm):
        self.comm = comm
        super().__init__(comm.part.n_local * comm.d)

    def _alloc(self):
        import ctypes

        p = ctypes.c_void_p()
        _lib.call("efb_peer_alloc", self.nbytes, ctypes.byref(p))
        self.base = int(p.value)
        self.f64 = dv.view_f64(self.base, self.nbytes // 8)

    def _fill_peer(self):
        import ctypes

        comm, part, d = self.comm, self.comm.part, self.comm.d
        dist = comm.dist
        world, rank = part.world, part.rank
        if world > _lib.MAX_RANKS:
            raise _lib.EfbError(f"peer-memory PCG supports up to {_lib.MAX_RANKS} ranks of one box, got {world}")
        handle = (ctypes.c_char * 64)()
        _lib.call("efb_peer_export", ctypes.c_void_p(self.base), handle)
        recv_lo = {int(q): int(lo) for q, lo, _ in comm.recv_segments}  # first entry of the segment received from q
        mine = (bytes(handle.raw), self.pbuf_off, recv_lo)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine, group=comm.group)
        P = self.peer
        P.world, P.rank = world, rank
        self._opened = []
        for q, (h, off, _) in enumerate(gathered):
            if q == rank:
                P.base[q] = self.base
            else:
                ptr = ctypes.c_void_p()
                hb = (ctypes.c_char * 64).from_buffer_copy(h)
                _lib.call("efb_peer_open", hb, ctypes.byref(ptr))
                self._opened.append(int(ptr.value))
                P.base[q] = int(ptr.value)
            P.pbuf_off[q][0], P.pbuf_off[q][1] = int(off[0]), int(off[1])
        send_rank, send_ptr, send_dst, recv_rank = peer_push_plan(comm, [g[2] for g in gathered])
        P.n_send, P.n_recv = len(send_rank), len(recv_rank)
        for i, q in enumerate(send_rank):
            P.send_rank[i], P.send_dst[i] = q, send_dst[i]
        for i, v in enumerate(send_ptr):
            P.send_ptr[i] = v
        for i, q in enumerate(recv_rank):
            P.recv_rank[i] = q
        P.send_idx = comm.send_idx.data_ptr() if comm.send_idx.numel() else None
        P.ar_seq, P.halo_seq = 0, 0
        dist.barrier(group=comm.group)  # every region is mapped before anybody stores into it

    def close(self):
        import ctypes

        for ptr in getattr(self, "_opened", []):
            _lib.call("efb_peer_close", ctypes.c_void_p(ptr))
        self._opened = []
        if getattr(self, "base", None):
            _lib.call("efb_peer_free", ctypes.c_void_p(self.base))
            self.base = None
