"""Host-side logic of the multi-GPU path on CPU: element partition, ghost elements, owned-row blocks,
and the PCG communication pattern over `gloo` with world_size 2 and 3.  The numerical kernels are replaced by the NumPy
oracle + scipy here (this is a test of partition/communication logic, not of the CUDA path)."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from easyfea_b200 import dist as efd
from easyfea_b200 import elements as el
from easyfea_b200 import meshgen
from oracle import easyfea_oracle as orc
from tests.helpers import make_mesh


def global_system(elemType, n, seed=3):
    coords, connect = make_mesh(elemType, n, seed=seed)
    dim = el.elem_dim(elemType)
    tab = el.gauss_table(elemType, "rigi")
    C = orc.IsoMaterial(dim, 210000.0, 0.3).C
    Ke = orc.linearized_elasticity(orc.geometry(coords[connect][:, :, :dim], tab.dN_pg, tab.weights), C)
    Nn = coords.shape[0]
    inv, indices, indptr, nnz = orc.csr_map([connect], dim, Nn * dim, True)
    K = sp.csr_matrix((orc.assemble_replay([Ke], inv, nnz), indices, indptr), shape=(Nn * dim, Nn * dim))
    return coords, connect, dim, Ke, K


def local_block(part, Ke_global, dim):
    """owned-row block assembled from the rank's local elements only (oracle replay in local numbering)"""
    Ke = Ke_global[part.elem_ids]
    inv, indices, indptr, nnz = orc.csr_map([part.connect], dim, part.n_local * dim, True)
    A = sp.csr_matrix((orc.assemble_replay([Ke], inv, nnz), indices, indptr), shape=(part.n_local * dim,) * 2)
    return A[: part.n_owned * dim]


@pytest.mark.parametrize("elemType,n", [("HEXA8", (3, 3, 5)), ("TETRA4", (3, 2, 4)), ("TRI3", (6, 7)), ("QUAD9", (3, 5))])
@pytest.mark.parametrize("world", [2, 3, 4])
def test_partition_invariants_and_owned_rows(elemType, n, world):
    coords, connect, dim, Ke, K = global_system(elemType, n)
    Nn = coords.shape[0]
    owner = efd.node_owners(connect, Nn, world)
    assert owner.min() >= 0
    seen = np.zeros(Nn, int)
    offsets = []
    for r in range(world):
        p = efd.Partition.from_global(connect, Nn, world, r)
        offsets.append(p.owned_offset)
        own = p.nodes[: p.n_owned]
        seen[own] += 1
        assert np.array_equal(own, np.flatnonzero(owner == r))  # ascending global ids
        # local elements = exactly those touching an owned node, in ascending global order
        assert np.array_equal(p.elem_ids, np.flatnonzero((owner[connect] == r).any(axis=1)))
        assert np.array_equal(p.nodes[p.connect], connect[p.elem_ids])
        b = efd.chunk_bounds(connect.shape[0], world)
        assert p.n_own_elems == ((p.elem_ids >= b[r]) & (p.elem_ids < b[r + 1])).sum()
        # halo segments are grouped by owner
        for i, q in enumerate(p.halo_ranks):
            seg = p.nodes[p.n_owned + p.halo_ptr[i]: p.n_owned + p.halo_ptr[i + 1]]
            assert np.all(owner[seg] == q) and np.all(np.diff(seg) > 0)
        # owned rows of the local assembly are bit-identical to the global rows (columns mapped back to global ids)
        A = local_block(p, Ke, dim).tocoo()
        gcol = (p.nodes[A.col // dim] * dim + A.col % dim)
        grow = (p.nodes[A.row // dim] * dim + A.row % dim)
        G = sp.csr_matrix((A.data, (grow, gcol)), shape=K.shape)
        rows = (own[:, None] * dim + np.arange(dim)).ravel()
        D = (G[rows] - K[rows]).tocoo()
        assert D.nnz == 0 or np.abs(D.data).max() == 0.0
        assert G[rows].nnz == K[rows].nnz
    assert np.all(seen == 1)
    assert offsets == sorted(offsets) and offsets[0] == 0


def test_slab_candidates_equal_global_partition():
    n, world = (3, 2, 2), 3
    coords, connect = meshgen.structured_mesh("HEXA8", (3, 2, 6))
    for r in range(world):
        c, e, own, cof = meshgen.hexa8_slab(n, r, world, jitter=0.2, seed=3)
        p1 = efd.Partition.from_candidates(c, e, own, r, world)
        p2 = efd.Partition.from_global(connect, coords.shape[0], world, r)
        for f in ("elem_ids", "connect", "nodes", "halo_ptr", "halo_ranks"):
            assert np.array_equal(getattr(p1, f), getattr(p2, f)), f
        assert p1.n_owned == p2.n_owned
    # neighbouring ranks see identical coordinates on the shared planes
    ca = meshgen.hexa8_slab(n, 0, world, jitter=0.2, seed=3)[3]
    cb = meshgen.hexa8_slab(n, 1, world, jitter=0.2, seed=3)[3]
    ids = np.arange(2 * 12, 3 * 12)
    assert np.array_equal(ca(ids), cb(ids))


# ---- multi-process (gloo) ---------------------------------------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cpu_gather(src, idx, dst):
    dst.copy_(src[idx.long()])


def _worker(rank, world, port, elemType, n, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        coords, connect, dim, Ke, K = global_system(elemType, n)
        Nn = coords.shape[0]
        ref = efd.Partition.from_global(connect, Nn, world, rank)
        owner = efd.node_owners(connect, Nn, world)
        part = efd.Partition.from_candidates(connect, np.arange(connect.shape[0]), lambda ids: owner[ids], rank, world)
        part.plan_exchange()  # send lists negotiated over the wire
        assert part.owned_offset == ref.owned_offset and part.n_global == Nn
        assert sorted(part.send) == sorted(ref.send)
        for q in part.send:
            assert np.array_equal(part.send[q], ref.send[q])
        comm = efd.RowComm(part, dim, device="cpu", gather=_cpu_gather)
        # 1. halo exchange: owned entries = f(global dof), halo entries arrive from their owners
        gd = (part.nodes[:, None] * dim + np.arange(dim)).ravel()
        f = np.sin(0.37 * gd) + gd
        x = torch.zeros(part.n_local * dim, dtype=torch.float64)
        x[: part.n_owned * dim] = torch.from_numpy(f[: part.n_owned * dim])
        comm.halo_exchange(x)
        assert np.array_equal(x.numpy(), f)
        # 1b. the same exchange as the peer-memory kernels do it: every rank STORES its interface entries at the planned
        #     offsets of the neighbours' vectors (emulated here by shipping (destination, offset, values) over gloo)
        recv_lo = {int(q): int(lo) for q, lo, _ in comm.recv_segments}
        gathered = [None] * world
        dist.all_gather_object(gathered, recv_lo)
        send_rank, send_ptr, send_dst, recv_rank = efd.peer_push_plan(comm, gathered)
        assert recv_rank == [int(q) for q in part.halo_ranks]
        idx = comm.send_idx.long().numpy()
        stores = [(q, send_dst[i], f[idx[send_ptr[i]:send_ptr[i + 1]]]) for i, q in enumerate(send_rank)]
        all_stores = [None] * world
        dist.all_gather_object(all_stores, stores)
        y = f.copy()
        y[part.n_owned * dim:] = np.nan
        for src, msgs in enumerate(all_stores):
            for q, off, vals in msgs:
                if q == rank:
                    assert src != rank and np.all(np.isnan(y[off:off + vals.size])), "segments overlap"
                    y[off:off + vals.size] = vals
        assert np.array_equal(y, f), "peer stores do not reproduce the halo exchange"
        # 2. distributed matvec == rows of the global matvec
        A = local_block(part, Ke, dim)
        y = A @ x.numpy()
        assert np.allclose(y, (K @ (np.sin(0.37 * np.arange(Nn * dim)) + np.arange(Nn * dim)))[gd[: part.n_owned * dim]], rtol=1e-13)
        # 3. Jacobi-PCG with the solver's communication pattern: halo(p) -> SpMV -> allreduce(pAp) -> allreduce(rz, rr)
        lat = coords[:, 0]
        fixed_nodes = np.flatnonzero(lat < lat.min() + 1e-9 + 0.05)
        fixed = np.zeros(Nn * dim, bool)
        fixed[(fixed_nodes[:, None] * dim + np.arange(dim)).ravel()] = True
        b_glob = np.cos(np.arange(Nn * dim) * 0.1)
        nown = part.n_owned * dim
        free = ~fixed[gd[:nown]]
        b = np.where(free, b_glob[gd[:nown]], 0.0)
        dinv = np.where(free, 1.0 / A.diagonal(), 0.0)
        xo, r = np.zeros(nown), b.copy()
        z = r * dinv
        p = torch.zeros(part.n_local * dim, dtype=torch.float64)
        p[:nown] = torch.from_numpy(z)
        s = torch.tensor([r @ z, r @ r])
        comm.all_reduce_sum(s)
        rz, rr0 = float(s[0]), float(s[1])
        it = 0
        while it < 2000:
            comm.halo_exchange(p)
            Ap = np.where(free, A @ p.numpy(), 0.0)
            t = torch.tensor([p.numpy()[:nown] @ Ap])
            comm.all_reduce_sum(t)
            alpha = rz / float(t[0])
            xo += alpha * p.numpy()[:nown]
            r -= alpha * Ap
            z = r * dinv
            s = torch.tensor([r @ z, r @ r])
            comm.all_reduce_sum(s)
            it += 1
            if float(s[1]) <= 1e-20 * rr0:
                break
            p[:nown] = torch.from_numpy(z + (float(s[0]) / rz) * p.numpy()[:nown])
            rz = float(s[0])
        # compare with the global direct solve
        import scipy.sparse.linalg as spla

        fg = ~fixed
        xg = np.zeros(Nn * dim)
        xg[fg] = spla.spsolve(K[fg][:, fg].tocsc(), b_glob[fg])
        err = np.linalg.norm(xo - xg[gd[:nown]]) / np.linalg.norm(xg)
        assert err < 1e-7, err
        out.put((rank, "ok", it))
    except Exception as exc:  # surface the failure in the parent
        import traceback

        out.put((rank, "fail", traceback.format_exc() + str(exc)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,elemType,n", [(2, "HEXA8", (3, 3, 4)), (3, "TETRA4", (2, 2, 5)), (2, "TRI3", (6, 6))])
def test_gloo_halo_exchange_and_pcg_pattern(world, elemType, n):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, elemType, n, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in results:
        assert status == "ok", f"rank {rank}: {info}"


# ---- RCB partitioning and the distributed build (no rank holds the whole mesh) -------------------------------------------
def test_rcb_parts_are_balanced_compact_and_order_independent():
    coords, connect = make_mesh("HEXA8", (8, 8, 8))
    cent = coords[connect].mean(1)
    Nn = coords.shape[0]
    for world in (2, 3, 4, 8):
        er = efd.rcb_element_ranks(cent, world)
        cnt = np.bincount(er, minlength=world)
        assert cnt.max() - cnt.min() <= 1
        perm = np.random.default_rng(world).permutation(connect.shape[0])
        assert np.array_equal(efd.rcb_element_ranks(cent[perm], world), er[perm])  # same parts whatever the element order
        # a randomly ORDERED mesh: contiguous chunks scatter every rank over the whole domain, RCB does not care
        halo_rcb = max(efd.Partition.from_global(connect[perm], Nn, world, r, erank=er[perm]).n_halo for r in range(world))
        halo_chunks = max(efd.Partition.from_global(connect[perm], Nn, world, r).n_halo for r in range(world))
        assert halo_rcb < 0.5 * halo_chunks
        if world == 8:  # bricks instead of slabs: smaller interfaces even on the sorted mesh
            halo_slab = max(efd.Partition.from_global(connect, Nn, world, r).n_halo for r in range(world))
            assert halo_rcb < halo_slab


def _dist_build_worker(rank, world, port, elemType, n, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        coords, connect = make_mesh(elemType, n)
        perm = np.random.default_rng(11).permutation(connect.shape[0])  # the mesh file stores the elements in a random order
        connect = connect[perm]
        Nn = coords.shape[0]
        b = efd.chunk_bounds(connect.shape[0], world)
        sl = slice(int(b[rank]), int(b[rank + 1]))  # this rank READS only its slice
        part = efd.build_partition_distributed(connect[sl], int(b[rank]), coords[connect[sl]].mean(1), Nn)
        ref = efd.Partition.from_global(connect, Nn, world, rank, erank=efd.rcb_element_ranks(coords[connect].mean(1), world))
        ref.plan_exchange()
        for name in ("elem_ids", "connect", "nodes", "halo_ranks", "halo_ptr"):
            assert np.array_equal(getattr(part, name), getattr(ref, name)), name
        assert (part.n_owned, part.n_own_elems, part.owned_offset, part.n_global) == (ref.n_owned, ref.n_own_elems, ref.owned_offset, ref.n_global)
        assert sorted(part.send) == sorted(ref.send) and all(np.array_equal(part.send[q], ref.send[q]) for q in part.send)
        out.put((rank, "ok", part.n_halo))
    except Exception as exc:
        import traceback

        out.put((rank, "fail", traceback.format_exc() + str(exc)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,elemType,n", [(2, "HEXA8", (4, 4, 4)), (3, "TETRA4", (3, 3, 3)), (4, "TRI3", (8, 8)), (4, "HEXA8", (4, 4, 4))])
def test_distributed_partition_build_equals_gathered_build(world, elemType, n):
    """each rank reads a slice of a randomly ordered element list; distributed RCB + migration + owner directory + ghost
    exchange reproduce `Partition.from_global(..., erank=rcb_element_ranks(...))` field by field (structured lattices: the cut
    coordinate has many ties, resolved by element id in both builds)"""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dist_build_worker, args=(r, world, port, elemType, n, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in results:
        assert status == "ok", f"rank {rank}: {info}"


def test_row_push_plan_inverts_the_send_lists():
    """the single-reduction PCG pushes interface rows from the update kernel: per-row plan == per-neighbour send lists"""
    rng = np.random.default_rng(0)
    n_rows = 50
    lists = [np.sort(rng.choice(n_rows, size=k, replace=False)) for k in (7, 11, 5)]
    send_idx = np.concatenate(lists)
    send_ptr = np.concatenate([[0], np.cumsum([a.size for a in lists])]).tolist()
    send_dst = [100, 300, 40]
    push_id, push_ptr, push_nbr, push_pos = efd.row_push_plan(send_idx, send_ptr, send_dst, n_rows)
    got = set()
    for row in range(n_rows):
        c = push_id[row]
        if c >= 0:
            for t in range(push_ptr[c], push_ptr[c + 1]):
                got.add((row, int(push_nbr[t]), int(push_pos[t])))
    want = {(int(r), s, send_dst[s] + j) for s, a in enumerate(lists) for j, r in enumerate(a)}
    assert got == want
    assert efd.row_push_plan(np.empty(0, dtype=np.int64), [0], [], n_rows) is None
