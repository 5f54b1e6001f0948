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
        u3, info3 = es.solve(tol=1e-10)
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
