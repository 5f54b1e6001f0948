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


def rcb_element_ranks(centroids: np.ndarray, world: int) -> np.ndarray:
    """Recursive coordinate bisection of the elements by their centroids (SURVEY.md section 8e "general: RCB on centroids";
    the reference partitions with gmsh/METIS, `FEM/_mesher.py:2303-2393`): rank of every element, parts of equal size (+-1),
    each cut along the longest extent of the part.  Independent of the element ORDER up to ties in the cut coordinate (broken by
    the element id), so a mesh whose elements are stored in a random order gets the same compact parts as a sorted one.
    Works for any `world` (a part of r ranks is cut r//2 : r - r//2)."""
    c = np.ascontiguousarray(centroids, dtype=np.float64)
    Ne = c.shape[0]
    erank = np.zeros(Ne, dtype=np.int64)
    stack = [(np.arange(Ne, dtype=np.int64), 0, int(world))]
    while stack:
        idx, r0, nr = stack.pop()
        if nr <= 1 or idx.size == 0:
            erank[idx] = r0
            continue
        pts = c[idx]
        axis = int(np.argmax(pts.max(0) - pts.min(0))) if idx.size else 0
        nl = nr // 2
        k = (idx.size * nl) // nr  # elements of the lower part
        order = np.lexsort((idx, pts[:, axis]))  # by coordinate, ties by element id
        stack.append((idx[order[:k]], r0, nl))
        stack.append((idx[order[k:]], r0 + nl, nr - nl))
    return erank


def chunk_element_ranks(Ne: int, world: int) -> np.ndarray:
    """contiguous chunks by element index (structured meshes generated slab by slab)"""
    return np.searchsorted(chunk_bounds(Ne, world), np.arange(Ne), side="right") - 1


def node_owners(connect: np.ndarray, Nn: int, world: int, erank: np.ndarray = None) -> np.ndarray:
    """owner[n] = lowest rank with an element touching node n (the reference's greedy rule, `_mesher.py:2355-2360`); -1 for
    orphan nodes (they have no rows to assemble).  `erank`: rank of every element (default: contiguous chunks)."""
    Ne = connect.shape[0]
    if erank is None:
        erank = chunk_element_ranks(Ne, world)
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
                        n_global: int = 0, erank=None):
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
        if erank is not None:  # rank of every CANDIDATE element
            n_own_elems = int((np.asarray(erank)[keep] == rank).sum())
        elif own_chunk is not None:
            n_own_elems = int(((elem_ids >= own_chunk[0]) & (elem_ids < own_chunk[1])).sum())
        else:
            n_own_elems = int(keep.sum())
        return cls(rank, world, elem_ids, n_own_elems, new_of_uniq[inv].reshape(connect.shape), nodes, n_owned,
                   halo_ranks.astype(np.int64), halo_ptr, 0, int(n_global))

    @classmethod
    def from_global(cls, connect: np.ndarray, Nn: int, world: int, rank: int, erank: np.ndarray = None):
        """Every rank sees the whole connectivity (small / medium meshes, tests).  `erank`: element -> rank (e.g.
        `rcb_element_ranks`); default: contiguous chunks.  See `from_distributed` for the build that never gathers the mesh."""
        connect = np.asarray(connect, dtype=np.int64)
        if erank is None:
            erank = chunk_element_ranks(connect.shape[0], world)
        owner = node_owners(connect, Nn, world, erank)
        part = cls.from_candidates(connect, np.arange(connect.shape[0]), lambda ids: owner[ids], rank, world, n_global=Nn,
                                   erank=erank)
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
# distributed build: no rank ever holds the whole connectivity
# ---------------------------------------------------------------------------------------------------------
def _sortable_u64(x: np.ndarray) -> np.ndarray:
    """float64 -> uint64 keys with the same order (sign-magnitude to biased), for exact k-th value searches by bisection"""
    b = np.ascontiguousarray(x, dtype=np.float64).view(np.uint64)
    neg = (b >> np.uint64(63)).astype(bool)
    return np.where(neg, ~b, b | np.uint64(1) << np.uint64(63))


def _alltoallv(by_dest, group=None, device="cpu", width: int = 1):
    """exchange int64 rows: `by_dest[q]` (k_q, width) goes to rank q; returns the list of arrays received from every rank.
    Sizes travel in one all-gather, payloads point to point (works with gloo on CPU tensors and nccl on CUDA tensors)."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    send = [np.ascontiguousarray(by_dest[q], dtype=np.int64).reshape(-1, width) for q in range(world)]
    sizes = torch.tensor([a.shape[0] for a in send], dtype=torch.int64, device=device)
    allsz = [torch.zeros(world, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(allsz, sizes, group=group)
    recv_n = [int(allsz[q][rank].item()) for q in range(world)]
    recv = [torch.empty((recv_n[q], width), dtype=torch.int64, device=device) for q in range(world)]
    ops, keep = [], []
    for q in range(world):
        if q == rank:
            recv[q] = torch.from_numpy(send[q]).to(device)
            continue
        if send[q].shape[0]:
            t = torch.from_numpy(send[q]).to(device)
            keep.append(t)
            ops.append(dist.P2POp(dist.isend, t, dist.get_global_rank(group, q) if group is not None else q, group=group))
        if recv_n[q]:
            ops.append(dist.P2POp(dist.irecv, recv[q], dist.get_global_rank(group, q) if group is not None else q, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return [r.cpu().numpy() for r in recv]


def rcb_element_ranks_distributed(centroids: np.ndarray, elem_ids: np.ndarray, world_parts: int, group=None, device="cpu") -> np.ndarray:
    """`rcb_element_ranks` for elements spread over the ranks of `group` (each rank passes ITS elements' centroids and global
    ids): the same cuts, found without gathering anything — the k-th smallest coordinate of a part is located by bisection on
    the ordered bit pattern of the doubles (<= 64 all-reduces of one count per active part), ties by bisection on the element id."""
    import torch.distributed as dist

    c = np.ascontiguousarray(centroids, dtype=np.float64)
    ids = np.asarray(elem_ids, dtype=np.int64)
    n = c.shape[0]
    part = np.zeros(n, dtype=np.int64)      # first rank of the part each element currently belongs to
    parts = [(0, int(world_parts))]         # (first rank, number of ranks)

    def allsum(v):
        t = torch.tensor(v, dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.cpu().numpy()

    def allred(v, op):
        t = torch.tensor(v, dtype=torch.float64, device=device)
        dist.all_reduce(t, op=op, group=group)
        return t.cpu().numpy()

    while any(nr > 1 for _, nr in parts):
        act = [(r0, nr) for r0, nr in parts if nr > 1]
        P = len(act)
        members = [np.flatnonzero(part == r0) for r0, _ in act]
        big = np.finfo(np.float64).max
        dim = c.shape[1]
        mn = allred([[c[m, a].min() if m.size else big for a in range(dim)] for m in members], dist.ReduceOp.MIN)
        mx = allred([[c[m, a].max() if m.size else -big for a in range(dim)] for m in members], dist.ReduceOp.MAX)
        axis = np.argmax(mx - mn, axis=1)
        cnt = allsum([m.size for m in members])
        k = np.array([(int(cnt[i]) * (act[i][1] // 2)) // act[i][1] for i in range(P)], dtype=np.int64)  # size of the lower part
        keys = [_sortable_u64(c[members[i], axis[i]]) for i in range(P)]
        # smallest key value v with count(key <= v) >= k (k >= 1); parts with k == 0 send everything up
        lo = np.zeros(P, dtype=np.uint64)
        hi = np.full(P, np.iinfo(np.uint64).max, dtype=np.uint64)
        for _ in range(64):
            mid = lo + (hi - lo) // np.uint64(2)
            le = allsum([int((keys[i] <= mid[i]).sum()) for i in range(P)])
            ok = le >= np.maximum(k, 1)
            hi = np.where(ok, mid, hi)
            lo = np.where(ok, lo, mid + np.uint64(1))
        cut = hi
        lt = allsum([int((keys[i] < cut[i]).sum()) for i in range(P)])
        need = k - lt  # how many of the ties (key == cut) go down: the ones with the smallest element ids
        tie_ids = [ids[members[i]][keys[i] == cut[i]] for i in range(P)]
        ilo = np.zeros(P, dtype=np.int64)
        ihi = np.full(P, np.iinfo(np.int64).max // 2, dtype=np.int64)
        for _ in range(62):  # smallest id bound b with count(tie id < b) >= need
            mid = ilo + (ihi - ilo) // 2
            below = allsum([int((tie_ids[i] < mid[i]).sum()) for i in range(P)])
            ok = below >= need
            ihi = np.where(ok, mid, ihi)
            ilo = np.where(ok, ilo, mid + 1)
        new_parts = [(r0, nr) for r0, nr in parts if nr <= 1]
        for i, (r0, nr) in enumerate(act):
            m = members[i]
            nl = nr // 2
            down = (keys[i] < cut[i]) | ((keys[i] == cut[i]) & (ids[m] < ihi[i])) if k[i] > 0 else np.zeros(m.size, dtype=bool)
            part[m[~down]] = r0 + nl
            new_parts += [(r0, nl), (r0 + nl, nr - nl)]
        parts = new_parts
    return part


def build_partition_distributed(connect_slice: np.ndarray, elem_id0: int, centroids_slice: np.ndarray, Nn: int, group=None,
                                device="cpu", partitioner: str = "rcb"):
    """The rank's `Partition` when the mesh arrives in pieces: rank r passes the connectivity (global node ids) and the
    centroids of ITS contiguous slice of the element list, `[elem_id0, elem_id0 + len)`.  No rank ever holds the whole mesh:
      1. element -> rank by distributed recursive coordinate bisection (or the slices themselves, `partitioner="chunks"`),
      2. elements migrate to their rank,
      3. node owners (lowest rank with an element on the node) are resolved at a directory rank `node % world`,
      4. a rank that holds an element on a node owned by a LOWER rank sends that element there (the ghost layer),
    then `Partition.from_candidates` + `plan_exchange`, exactly as in the gathered build: the result is identical to
    `Partition.from_global(connect, Nn, world, rank, erank=rcb_element_ranks(centroids, world))`."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    conn = np.ascontiguousarray(connect_slice, dtype=np.int64)
    ne, nPe = conn.shape
    ids = elem_id0 + np.arange(ne, dtype=np.int64)
    if partitioner == "rcb":
        erank = rcb_element_ranks_distributed(centroids_slice, ids, world, group, device)
    elif partitioner == "chunks":
        erank = np.full(ne, rank, dtype=np.int64)
    else:
        raise ValueError(partitioner)
    # 2. migration: (id, nodes...) rows to the element's rank
    rows = np.concatenate([ids[:, None], conn], axis=1)
    got = _alltoallv([rows[erank == q] for q in range(world)], group, device, width=nPe + 1)
    mine = np.concatenate(got, axis=0)
    mine = mine[np.argsort(mine[:, 0], kind="stable")]
    my_ids, my_conn = mine[:, 0], mine[:, 1:]
    # 3. owners through the directory
    touched = np.unique(my_conn.ravel())
    asked = _alltoallv([touched[touched % world == q] for q in range(world)], group, device)
    asked = [a.ravel() for a in asked]
    allnodes = np.concatenate(asked) if asked else np.empty(0, dtype=np.int64)
    allranks = np.concatenate([np.full(a.size, q, dtype=np.int64) for q, a in enumerate(asked)]) if asked else allnodes
    u, inv = np.unique(allnodes, return_inverse=True)
    own_u = np.full(u.size, world, dtype=np.int64)
    np.minimum.at(own_u, inv, allranks)
    replies = _alltoallv([np.stack([a, own_u[np.searchsorted(u, a)]], axis=1) if a.size else np.empty((0, 2), dtype=np.int64)
                          for a in asked], group, device, width=2)
    rep = np.concatenate(replies, axis=0)
    rep = rep[np.argsort(rep[:, 0], kind="stable")]
    assert np.array_equal(rep[:, 0], touched)
    owner_touched = rep[:, 1]
    node_owner = owner_touched[np.searchsorted(touched, my_conn)]  # (my elements, nPe)
    # 4. ghosts: my element goes to every LOWER rank that owns one of its nodes, with the owners of its nodes
    out_rows = []
    for q in range(world):
        sel = (node_owner == q).any(axis=1) if q < rank else np.zeros(my_ids.size, dtype=bool)
        out_rows.append(np.concatenate([my_ids[sel, None], my_conn[sel], node_owner[sel]], axis=1))
    ghosts = np.concatenate(_alltoallv(out_rows, group, device, width=1 + 2 * nPe), axis=0)
    cand_ids = np.concatenate([my_ids, ghosts[:, 0]])
    cand_conn = np.concatenate([my_conn, ghosts[:, 1:1 + nPe]], axis=0)
    cand_own = np.concatenate([node_owner, ghosts[:, 1 + nPe:]], axis=0)
    cand_rank = np.concatenate([np.full(my_ids.size, rank, dtype=np.int64), np.full(ghosts.shape[0], -1, dtype=np.int64)])
    order = np.argsort(cand_ids, kind="stable")
    cand_ids, cand_conn, cand_own, cand_rank = cand_ids[order], cand_conn[order], cand_own[order], cand_rank[order]
    known, first = np.unique(cand_conn.ravel(), return_index=True)
    known_owner = cand_own.ravel()[first]

    def owner_of(q):
        return known_owner[np.searchsorted(known, q)]

    part = Partition.from_candidates(cand_conn, cand_ids, owner_of, rank, world, n_global=Nn, erank=cand_rank)
    part.plan_exchange(group)
    return part


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


def row_push_plan(send_idx: np.ndarray, send_ptr, send_dst, n_rows: int):
    """Invert the per-neighbour send lists into a per-row plan: (push_id (n_rows) int32, push_ptr int64, push_nbr int32,
    push_pos int64), or None when nothing is sent.  Entry j of neighbour s goes to position send_dst[s] + (j - send_ptr[s])."""
    send_idx = np.asarray(send_idx, dtype=np.int64)
    if send_idx.size == 0:
        return None
    nbr = np.concatenate([np.full(send_ptr[s + 1] - send_ptr[s], s, dtype=np.int32) for s in range(len(send_ptr) - 1)])
    pos = np.concatenate([send_dst[s] + np.arange(send_ptr[s + 1] - send_ptr[s], dtype=np.int64) for s in range(len(send_ptr) - 1)])
    order = np.argsort(send_idx, kind="stable")
    rows, first, counts = np.unique(send_idx[order], return_index=True, return_counts=True)
    push_id = np.full(n_rows, -1, dtype=np.int32)
    push_id[rows] = np.arange(rows.size, dtype=np.int32)
    push_ptr = np.zeros(rows.size + 1, dtype=np.int64)
    np.cumsum(counts, out=push_ptr[1:])
    return push_id, push_ptr, np.ascontiguousarray(nbr[order]), np.ascontiguousarray(pos[order])


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
        # two p buffers, two z buffers (Chebyshev form): each over [owned | halo]
        self.pbuf_off = tuple(ctrl + k * _pad256(self.n * 8) for k in range(4))
        self.nbytes = ctrl + 4 * _pad256(self.n * 8)
        self._alloc()
        self.p = [self.f64[o // 8: o // 8 + self.n] for o in self.pbuf_off[:2]]
        self.zb = [self.f64[o // 8: o // 8 + self.n] for o in self.pbuf_off[2:]]
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
        for k in range(4):
            P.pbuf_off[0][k] = self.pbuf_off[k]
        P.send_idx = None
        P.ar_seq, P.halo_seq = 0, 0

    def status(self):
        """(r.r of the last finished iteration, error flag, iterations of the last persistent launch) — one small copy"""
        c = self.ctrl.cpu().numpy()
        return float(c[self.rr_off]), int(c.view(np.uint32)[self.err_off]), int(c.view(np.uint64)[self.iters_off])

    def advance(self, n_iters: int, reductions_per_iter: int = 2, halos_per_iter: int = 1):
        self.peer.ar_seq += reductions_per_iter * n_iters
        self.peer.halo_seq += halos_per_iter * n_iters


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
            for k in range(4):
                P.pbuf_off[q][k] = int(off[k])
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
        # the same plan seen from the rows (single-reduction form: the update kernel stores a row into the neighbours when it
        # produces it): push_id[row] = -1 | c, entries push_ptr[c]:push_ptr[c+1] of (neighbour index, entry in its vector)
        self._push = row_push_plan(comm.send_idx.cpu().numpy(), send_ptr, send_dst, part.n_owned * d)
        if self._push is not None:
            self._push = tuple(torch.from_numpy(a).to(comm.device) for a in self._push)
            P.push_id, P.push_ptr, P.push_nbr, P.push_pos = (t.data_ptr() for t in self._push)
        dist.barrier(group=comm.group)  # every region is mapped before anybody stores into it

    def close(self):
        import ctypes

        for ptr in getattr(self, "_opened", []):
            _lib.call("efb_peer_close", ctypes.c_void_p(ptr))
        self._opened = []
        if getattr(self, "base", None):
            _lib.call("efb_peer_free", ctypes.c_void_p(self.base))
            self.base = None
