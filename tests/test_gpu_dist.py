"""Row-sharded assembly + Jacobi-PCG over NCCL on 2 GPUs of one box against the single-GPU result.  -m gpu, skipped with
fewer than 2 devices (run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from easyfea_b200 import dist as efd
        from easyfea_b200 import mesh, meshgen, phasefield, staggered

        n = (6, 5, 4)
        connect, elem_ids, owner_of, coords_of = meshgen.hexa8_slab(n, rank, world, jitter=0.15, seed=2)
        part = efd.Partition.from_candidates(connect, elem_ids, owner_of, rank, world)
        part.plan_exchange()
        g = mesh.ElemGroup("HEXA8", part.connect, coords_of(part.nodes), all_nodes_used=True)
        sysm = staggered.LocalSystem(g, part, lambda p, d: efd.RowComm(p, d))
        mat = phasefield.IsotropicMaterial(3, 210000.0, 0.3)
        es = staggered.ElasticSolve(sysm, mat.C)
        P = (n[0] + 1) * (n[1] + 1)
        plane = part.nodes // P
        loc = np.arange(part.n_local)
        es.bc.add(loc[plane == 0], [0, 0, 0], [0, 1, 2], 3)
        es.bc.add(loc[plane == world * n[2]], [0.05], [2], 3)
        es.pcg_single_reduction = True  # first solve on a fresh communicator: the single-reduction form
        u, info = es.solve(tol=1e-10)  # fused iterations: reductions + halo pushes through peer memory
        assert info["converged"] and info["fused"] and not info["persistent"] and info["single_reduction"], info
        u_sr = u[: part.n_owned * 3].cpu().numpy()
        es.pcg_single_reduction = False  # the classic two-reduction form (the default) from here on, unless stated
        es.u.zero_()
        u, info = es.solve(tol=1e-10)
        assert np.linalg.norm(u_sr - u[: part.n_owned * 3].cpu().numpy()) <= 1e-8 * np.linalg.norm(u_sr)
        assert info["converged"] and not info["single_reduction"], info
        u_peer = u[: part.n_owned * 3].cpu().numpy()
        u2, info2 = es.solve(tol=1e-10)  # second solve on the same communicator (sequence numbers carry on), warm start
        assert info2["converged"], info2
        es.u.zero_()
        es.pcg_persistent = True  # one cooperative kernel per solve, same peer-memory exchanges
        u4, info4 = es.solve(tol=1e-10)
        es.pcg_persistent = False
        assert info4["converged"] and info4["fused"] and info4["persistent"], info4
        assert np.linalg.norm(u4[: part.n_owned * 3].cpu().numpy() - u_peer) <= 1e-8 * np.linalg.norm(u_peer)
        es.u.zero_()
        es.pcg_single_reduction = True  # Chronopoulos-Gear form: one all-reduce per iteration, the halo travels with z
        u5, info5 = es.solve(tol=1e-10)
        es.pcg_single_reduction = False
        assert info5["converged"] and info5["single_reduction"], info5
        assert np.linalg.norm(u5[: part.n_owned * 3].cpu().numpy() - u_peer) <= 1e-8 * np.linalg.norm(u_peer)
        u6, info6 = es.solve(tol=1e-10)  # and back to the classic form on the same communicator (sequence numbers carry on)
        assert info6["converged"], info6
        es.u.zero_()
        es.pcg_fused = False  # NCCL send/recv + all-reduce per iteration: same iterates up to summation order
        es.pcg_precond_degree = info["precond_degree"]  # the same polynomial as the fused solve it is compared with
        u3, info3 = es.solve(tol=1e-10)
        es.pcg_precond_degree = "auto"
        assert info3["converged"] and not info3["fused"], info3
        u_nccl = u3[: part.n_owned * 3].cpu().numpy()
        assert abs(info3["iterations"] - info["iterations"]) <= 25, (info, info3)
        assert np.linalg.norm(u_peer - u_nccl) <= 1e-8 * max(np.linalg.norm(u_nccl), 1e-300), "peer-memory PCG differs from the NCCL loop"
        own = part.nodes[: part.n_owned]
        out.put((rank, "ok", own, u_peer, info["iterations"]))
    except Exception as exc:
        import traceback

        out.put((rank, "fail", traceback.format_exc() + str(exc), None, None))
    finally:
        dist.destroy_process_group()


def test_sharded_elastic_solve_matches_single_gpu():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from easyfea_b200 import mesh, meshgen, phasefield, staggered

    world, n = 2, (6, 5, 4)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    bad = [f"rank {r[0]}: {r[2]}" for r in res if r[1] != "ok"]
    assert not bad, "\n".join(bad)
    # single-GPU solve of the same global mesh (rank-independent jitter per plane -> same coordinates)
    connect, elem_ids, owner_of, coords_of = meshgen.hexa8_slab((n[0], n[1], world * n[2]), 0, 1, jitter=0.15, seed=2)
    Nn = (n[0] + 1) * (n[1] + 1) * (world * n[2] + 1)
    coords = coords_of(np.arange(Nn))
    g = mesh.ElemGroup("HEXA8", connect, coords, all_nodes_used=True)
    es = staggered.ElasticSolve(staggered.LocalSystem(g), phasefield.IsotropicMaterial(3, 210000.0, 0.3).C)
    P = (n[0] + 1) * (n[1] + 1)
    plane = np.arange(Nn) // P
    es.bc.add(np.flatnonzero(plane == 0), [0, 0, 0], [0, 1, 2], 3)
    es.bc.add(np.flatnonzero(plane == world * n[2]), [0.05], [2], 3)
    u1, info = es.solve(tol=1e-10)
    u1 = u1.cpu().numpy().reshape(Nn, 3)
    got = np.zeros_like(u1)
    for _, _, own, ur, _ in res:
        got[own] = ur.reshape(-1, 3)
    assert np.linalg.norm(got - u1) / np.linalg.norm(u1) < 1e-8


# ---------------------------------------------------------------------------------------------------------
# sharded PhaseFieldStaggered and TransientSolve against the single-GPU run (fields, Niter, PCG iteration counts)
# ---------------------------------------------------------------------------------------------------------
def _pf_setup(elemType, n, split, sysm_of, nodes_of):
    """phase-field shear case on a structured mesh (crack as d = 1); `sysm_of(group args)` builds the LocalSystem"""
    from easyfea_b200 import meshgen, phasefield, staggered

    dim = 2 if elemType in ("TRI3", "QUAD4") else 3
    L, l0 = 1e-3, 1e-4
    lengths = (L,) * dim
    lattice, connect = meshgen.structured_mesh(elemType, n, lengths=lengths)
    coords, _ = meshgen.structured_mesh(elemType, n, lengths=lengths, jitter=0.15, seed=5)
    sysm, nodes = sysm_of(elemType, connect, coords)
    pfm = phasefield.PhaseFieldModel(phasefield.IsotropicMaterial(dim, 210e9, 0.3, planeStress=False), split, "AT2", 2.7e3, l0)
    simu = staggered.PhaseFieldStaggered(sysm, pfm, pcg_tol=1e-11)
    x, y = lattice[nodes, 0], lattice[nodes, 1]
    loc = np.arange(nodes.size)
    tol = 1e-12
    sets = (loc[(np.abs(y - L / 2) < tol) & (x <= L / 2 + tol)], loc[np.abs(y - L) < tol], loc[np.abs(y) < tol])
    return simu, sets, dim, nodes


def _pf_run(simu, sets, dim, loads=(4e-6, 8e-6)):
    crack, top, bot = sets
    out = []
    for dep in loads:
        simu.Bc_Init()
        simu.add_dirichlet(crack, [1], [0], problemType="damage")
        simu.add_dirichlet(top, [dep, 0.5 * dep] + [0] * (dim - 2), list(range(dim)))
        simu.add_dirichlet(bot, [0] * dim, list(range(dim)))
        u, d, conv = simu.Solve(1e-3, 50, convOption=0)
        nd = simu.sys.n_owned
        out.append((u[: nd * dim].cpu().numpy().copy(), d[:nd].cpu().numpy().copy(), simu.Niter, simu.info["elastic"]["iterations"],
                    simu.info["damage"]["iterations"], bool(conv)))
        simu.Save_Iter()
    return out


def _transient_run(sysm, nodes, lattice):
    from easyfea_b200 import phasefield, transient

    x = lattice[nodes, 0]
    loc = np.arange(nodes.size)
    xmax = float(lattice[:, 0].max())
    lo, hi = loc[x < 1e-12], loc[x > xmax - 1e-12]
    dyn = transient.TransientSolve.elastodynamic(sysm, phasefield.IsotropicMaterial(3, 210000.0, 0.3).C, 2.0, coefM=0.15, coefK=2e-4)
    dyn.Solver_Set_Hyperbolic_Algorithm(dt=0.05)
    dyn.pcg_tol = 1e-11
    res = []
    for s in range(2):
        dyn.bc = type(dyn.bc)(dyn.bc.n)
        dyn.bc.add(lo, [0.0, 0.0, 0.0], [0, 1, 2], 3)
        dyn.bc.add(hi, [0.01 * (s + 1)], [0], 3)
        dyn.Solve()
    nown = sysm.n_owned * 3
    return dyn.u[:nown].cpu().numpy().copy(), dyn.v[:nown].cpu().numpy().copy(), dyn.a[:nown].cpu().numpy().copy(), dyn.info["iterations"]


def _worker2(rank, world, port, out, partitioner):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from easyfea_b200 import dist as efd
        from easyfea_b200 import mesh, meshgen, staggered

        def sharded(elemType, connect, coords):
            Nn = coords.shape[0]
            erank = efd.rcb_element_ranks(coords[connect].mean(1), world) if partitioner == "rcb" else None
            part = efd.Partition.from_global(connect, Nn, world, rank, erank=erank)
            g = mesh.ElemGroup(elemType, part.connect, coords[part.nodes], all_nodes_used=True)
            return staggered.LocalSystem(g, part, lambda p, d: efd.RowComm(p, d)), part.nodes

        res = {}
        for key, (et, n, split) in {"tri3": ("TRI3", (16, 16), "Miehe"), "tetra4": ("TETRA4", (6, 6, 2), "He")}.items():
            simu, sets, dim, nodes = _pf_setup(et, n, split, sharded, None)
            res[key] = (nodes[: simu.sys.n_owned], _pf_run(simu, sets, dim))
        lattice, connect = meshgen.structured_mesh("HEXA8", (5, 4, 3))
        coords, _ = meshgen.structured_mesh("HEXA8", (5, 4, 3), jitter=0.12, seed=7)
        sysm, nodes = sharded("HEXA8", connect, coords)
        res["transient"] = (nodes[: sysm.n_owned], _transient_run(sysm, nodes, lattice))
        out.put((rank, "ok", res))
    except Exception as exc:
        import traceback

        out.put((rank, "fail", traceback.format_exc() + str(exc)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("partitioner", ["chunks", "rcb"])
def test_sharded_phasefield_and_transient_match_single_gpu(partitioner):
    """2 GPUs vs 1: staggered phase-field loop (fields 1e-7, identical staggered iteration counts, PCG iteration counts within a
    few iterations: the dot products are summed in a different order) and two Newmark steps (u, v, a)"""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from easyfea_b200 import mesh, meshgen, staggered

    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker2, args=(r, world, port, out, partitioner)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=900) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    bad = [f"rank {r[0]}: {r[2]}" for r in res if r[1] != "ok"]
    assert not bad, "\n".join(bad)

    def single(elemType, connect, coords):
        g = mesh.ElemGroup(elemType, connect, coords, all_nodes_used=True)
        return staggered.LocalSystem(g), np.arange(coords.shape[0])

    for key, (et, n, split) in {"tri3": ("TRI3", (16, 16), "Miehe"), "tetra4": ("TETRA4", (6, 6, 2), "He")}.items():
        simu, sets, dim, nodes = _pf_setup(et, n, split, single, None)
        ref = _pf_run(simu, sets, dim)
        Nn = nodes.size
        for k, (u1, d1, Niter, it_u, it_d, conv) in enumerate(ref):
            u, d = np.zeros_like(u1), np.zeros_like(d1)
            for _, _, r in res:
                own, steps = r[key]
                uk, dk, Nk, iu, idd, ck = steps[k]
                u.reshape(Nn, dim)[own] = uk.reshape(-1, dim)
                d[own] = dk
                assert Nk == Niter and ck == conv, (key, k, Nk, Niter)
                assert abs(iu - it_u) <= max(25, it_u // 20) and abs(idd - it_d) <= max(25, it_d // 20), (key, k, iu, it_u, idd, it_d)
            assert np.linalg.norm(u - u1) <= 1e-7 * np.linalg.norm(u1), (key, k)
            assert np.linalg.norm(d - d1) <= 1e-7 * np.linalg.norm(d1), (key, k)
    lattice, connect = meshgen.structured_mesh("HEXA8", (5, 4, 3))
    coords, _ = meshgen.structured_mesh("HEXA8", (5, 4, 3), jitter=0.12, seed=7)
    sysm, nodes = single("HEXA8", connect, coords)
    u1, v1, a1, it1 = _transient_run(sysm, nodes, lattice)
    u, v, a = np.zeros_like(u1), np.zeros_like(v1), np.zeros_like(a1)
    for _, _, r in res:
        own, (uk, vk, ak, itk) = r["transient"]
        for full, part_ in ((u, uk), (v, vk), (a, ak)):
            full.reshape(-1, 3)[own] = part_.reshape(-1, 3)
    for full, ref_ in ((u, u1), (v, v1), (a, a1)):
        assert np.linalg.norm(full - ref_) <= 1e-7 * np.linalg.norm(ref_)
