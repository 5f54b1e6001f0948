"""`assembly.MmaSchedule` (staging slots + gather programs of `efb_assemble_elastic_mma`, csrc/fused_mma.cu) on the CPU: the
schedule is torch tensor ops, so it is built here on CPU tensors and EXECUTED by a NumPy emulation of the kernel's data flow
(T rows of the owned nodes -> staging -> per-block gather in program order); the summed T blocks must equal the direct sum
over (element, a, b).  The arithmetic of the kernel itself is checked on the GPU (tests/test_gpu_assembly.py)."""
import numpy as np
import pytest

from easyfea_b200 import assembly
from easyfea_b200 import elements as el
from tests.helpers import make_mesh
from tests.test_hostcheck_fused import host_node_graph


def scaled_gradients(X, dN_pg, w):
    """g[p, k, a] = sqrt(w_p |det F|) dN_a/dx_k of one element (F[r][c] = sum_n dN[p][r][n] x[n][c])"""
    F = np.einsum("prn,nc->prc", dN_pg, X)
    Fi = np.linalg.inv(F)
    sw = np.sqrt(w * np.abs(np.linalg.det(F)))
    return np.einsum("p,pdk,pka->pda", sw, Fi, dN_pg)


@pytest.mark.parametrize("n,n_own", [((5, 4, 3), None), ((9, 6, 5), None), ((4, 4, 4), 60)])
def test_mma_schedule_program_sums_every_block(n, n_own):
    coords, connect = make_mesh("HEXA8", n)
    Nn = coords.shape[0]
    tab = el.gauss_table("HEXA8", "rigi")
    graph = host_node_graph(connect, Nn, coords, "HEXA8")
    ms = assembly.MmaSchedule(graph, n_nodes=n_own, nPg=8)
    sched = ms.base
    assert sched.S == 16 and ms.t_cap <= 128
    ncl, cap4, t_cap = ms.n_clusters, ms.cap4, ms.t_cap
    assert cap4 % 4 == 0 and cap4 >= sched.cap_e
    cl_nodes = sched.cl_nodes.numpy().reshape(ncl, 16, 4)
    recs = ms.recs.numpy()
    o_rs, o_nodes, o_hdr = cap4 * 8, cap4 * 12, cap4 * 12 + 64
    assert ms.rec_words % 4 == 0 and ms.rec_words >= o_hdr + 1 + ms.rmax
    conn = np.ascontiguousarray(recs[:, :o_rs]).reshape(ncl, cap4, 8)
    rowslot = np.ascontiguousarray(recs[:, o_rs:o_nodes]).view(np.int16).reshape(ncl, cap4, 8)
    nrec = np.ascontiguousarray(recs[:, o_nodes:o_hdr]).view(np.int64).reshape(ncl, 16, 2)
    hdr = recs[:, o_hdr:]
    assert (np.diff(ms.prog_off.numpy()) <= ms.pw_max).all() and (np.diff(ms.prog_off.numpy()) % 4 == 0).all()
    prog, prog_off = ms.prog.numpy(), ms.prog_off.numpy()
    u16 = prog.view(np.uint16)
    adjptr = graph.adjptr.numpy()
    outT = np.full((graph.nnz_node, 9), np.nan)
    written = np.zeros(graph.nnz_node, dtype=np.int64)

    def cpad_of(c):
        return 1 if c <= 1 else 2 if c <= 2 else 4 if c <= 4 else (c + 7) // 8 * 8

    for c in range(ncl):
        stage = np.full(t_cap * 72 + 72, np.nan)
        stage[t_cap * 72:] = 0.0  # the zero block: source of lanes without a contribution
        for le in range(cap4):
            if conn[c, le, 0] < 0:
                assert (rowslot[c, le] == -1).all()
                continue
            g = scaled_gradients(coords[conn[c, le]][:, :3], tab.dN_pg, tab.weights)
            for a in range(8):
                v = int(rowslot[c, le, a])
                if v >= 0:
                    ts, xs = v & 0xFF, (v >> 8) & 7
                    assert ts < t_cap
                    T = np.einsum("pk,plb->klb", g[:, :, a], g).reshape(9, 8)
                    for b in range(8):
                        dst = ts * 72 + np.arange(9) * 8 + (b ^ xs)
                        assert np.isnan(stage[dst]).all()  # every staging cell is written once
                        stage[dst] = T[:, b]
        R = hdr[c, 0]
        last_c = None
        for r in range(R):
            h = int(hdr[c, 1 + r])
            cnt, o = h & 0xFF, h >> 8
            assert last_c is None or cnt <= last_c  # rounds sorted by trip count
            last_c = cnt
            cpad = cpad_of(cnt)
            base = prog_off[c] + o
            assert base % 4 == 0
            for lane in range(32):
                dest = prog[base + lane]
                acc = np.zeros(9)
                nsrc = 0
                for it in range(cpad):
                    s = int(u16[(base + 32) * 2 + lane * cpad + it])
                    if s != t_cap * 72:
                        assert it < cnt and s < t_cap * 72
                        acc += stage[s + np.arange(9) * 8]
                        nsrc += 1
                if dest < 0:
                    assert nsrc == 0
                    continue
                i, slot = dest >> 16, dest & 0xFFFF
                node = cl_nodes[c, i, 0]
                off, dg_ = nrec[c, i]
                assert node >= 0 and slot < dg_ and nsrc >= 1
                assert off == 9 * adjptr[node] and dg_ == adjptr[node + 1] - adjptr[node]
                outT[adjptr[node] + slot] = acc
                written[adjptr[node] + slot] += 1
    # direct sum over (element, a, b)
    ref = np.zeros((graph.nnz_node, 9))
    pos = graph.pos.numpy().reshape(-1, 8, 8)
    for e in range(connect.shape[0]):
        g = scaled_gradients(coords[connect[e]][:, :3], tab.dN_pg, tab.weights)
        T = np.einsum("pka,plb->abkl", g, g).reshape(8, 8, 9)
        for a in range(8):
            ref[adjptr[connect[e, a]] + pos[e, a]] += T[a]
    n_sched = Nn if n_own is None else n_own
    owned = np.zeros(graph.nnz_node, dtype=bool)
    owned[: adjptr[n_sched]] = True
    assert (written[owned] == 1).all() and (written[~owned] == 0).all()
    assert np.abs(outT[owned] - ref[owned]).max() <= 1e-13 * np.abs(ref).max()
