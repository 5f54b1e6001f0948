"""GPU parity of the CSR pattern (bit-exact vs scipy's canonical pattern) and of the deterministic replay (bit-exact vs
np.bincount), plus size-independent properties at the 1 M-element HEXA8 size of BASELINE config 2.  -m gpu"""
import os

import numpy as np
import pytest

from oracle import easyfea_oracle as orc
from tests.helpers import make_mesh, rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def efb():
    from easyfea_b200 import _lib, assembly, mesh, operators

    _lib.require_cuda()

    class NS:
        pass

    ns = NS()
    ns.asm, ns.mesh, ns.op = assembly, mesh, operators
    return ns


@pytest.mark.parametrize("elemType", ["TRI3", "QUAD9", "TETRA4", "HEXA8", "HEXA27"])
@pytest.mark.parametrize("dof", [1, "dim"])
def test_pattern_and_replay_bit_exact(efb, elemType, dof):
    coords, connect = make_mesh(elemType)
    coords = np.concatenate([coords, np.zeros((2, 3))])  # two orphan nodes: empty rows
    g = efb.mesh.ElemGroup(elemType, connect, coords)
    d = g.dim if dof == "dim" else 1
    Nn = coords.shape[0]
    Ndof = Nn * d + 3  # + Lagrange rows
    A = efb.asm.Assembler()
    inv, indices, indptr, nnz = A.Get_csr_map(d, True, Ndof, (g,))
    oinv, oind, optr, onnz = orc.csr_map([connect], d, Ndof, True)
    assert nnz == onnz
    for got, ref in ((inv, oinv), (indices, oind), (indptr, optr)):
        assert got.dtype == ref.dtype and np.array_equal(got, ref)
    rng = np.random.default_rng(5)
    Ke = rng.normal(size=(g.Ne, g.nPe * d, g.nPe * d))
    K = A.Assemble_csr({g: Ke}, d, Ndof, True)
    assert K.shape == (Ndof, Ndof) and K.has_canonical_format
    assert np.array_equal(K.data, orc.assemble_replay([Ke], oinv, onnz))  # bit-exact ordered sum
    assert np.array_equal(K.indices, oind) and np.array_equal(K.indptr, optr)
    Fe = rng.normal(size=(g.Ne, g.nPe * d))
    F = A.Assemble_csr({g: Fe}, d, Ndof, False)
    vinv, vind, vptr, vnnz = orc.csr_map([connect], d, Ndof, False)
    assert F.shape == (Ndof, 1) and np.array_equal(F.indptr, vptr) and np.array_equal(F.indices, vind)
    assert np.array_equal(F.data, orc.assemble_replay([Fe], vinv, vnnz))
    vi = A.Get_csr_map(d, False, Ndof, (g,))
    assert np.array_equal(vi[0], vinv) and vi[3] == vnnz
    # None / empty handling of __Assemble_csr (_simu.py:1029-1035)
    assert A.Assemble_csr({g: None}, d, Ndof, True).nnz == 0
    assert A.Assemble_csr({}, d, Ndof, False).shape == (Ndof, 1)


def test_two_groups_in_dict_order(efb):
    cq, q = make_mesh("QUAD4", (4, 3))
    tri = np.concatenate([q[:6, [0, 1, 2]], q[:6, [0, 2, 3]]])
    quad = q[4:]
    gt = efb.mesh.ElemGroup("TRI3", tri, cq)
    gq = efb.mesh.ElemGroup("QUAD4", quad, cq)
    rng = np.random.default_rng(3)
    A = efb.asm.Assembler()
    for d in (1, 2):
        Nn = cq.shape[0]
        K1 = rng.normal(size=(gt.Ne, 3 * d, 3 * d))
        K2 = rng.normal(size=(gq.Ne, 4 * d, 4 * d))
        K = A.Assemble_csr({gt: K1, gq: K2}, d, Nn * d, True)
        oinv, oind, optr, onnz = orc.csr_map([tri, quad], d, Nn * d, True)
        assert np.array_equal(K.indptr, optr) and np.array_equal(K.indices, oind)
        assert np.array_equal(K.data, orc.assemble_replay([K1, K2], oinv, onnz))
        Kr = A.Assemble_csr({gq: K2, gt: K1}, d, Nn * d, True)  # other dict order = other entry order
        oinv2, _, _, _ = orc.csr_map([quad, tri], d, Nn * d, True)
        assert np.array_equal(Kr.data, orc.assemble_replay([K2, K1], oinv2, onnz))


@pytest.mark.parametrize("name", ["QUAD9", "HEXA8", "TRI3", "TETRA4", "HEXA27"])
def test_golden_assembly(efb, name):
    d = dict(np.load(os.path.join(GOLD, f"{name}.npz")))
    g = efb.mesh.ElemGroup(name, d["connect"], d["coords"])
    Nn, dim = d["coords"].shape[0], g.dim
    A = efb.asm.Assembler()
    for dof_n in (1, dim):
        inv, indices, indptr, nnz = A.Get_csr_map(dof_n, True, Nn * dof_n, (g,))
        for got, key in ((inv, "inv"), (indices, "indices"), (indptr, "indptr")):
            ref = d[f"{key}_{dof_n}"]
            assert got.dtype == ref.dtype and np.array_equal(got, ref), key
    K = A.Assemble_csr({g: d["Ke"]}, dim, Nn * dim, True)
    assert np.array_equal(K.data, d["K_data"])  # same inputs -> bit-identical to the reference's bincount
    F = A.Assemble_csr({g: d["Fe"]}, 1, Nn, False)
    assert np.array_equal(F.data, d["F_data"]) and np.array_equal(F.indptr, d["F_indptr"])
    # end to end on the device: K_e computed by the CUDA kernel, then assembled (values to 1e-12)
    Kd = A.Assemble_csr({g: efb.op.elastic_Ke_dev(g, d["C"])}, dim, Nn * dim, True)
    assert rel_err(Kd.data, d["K_data"]) < 1e-12


def test_million_element_properties(efb):
    """BASELINE config 2 at its 1 M-element size (HEXA8 100^3): properties that need no oracle at this size."""
    import torch

    from easyfea_b200 import meshgen, solver

    n = 100
    coords, connect = meshgen.structured_mesh("HEXA8", n, jitter=0.2, seed=0)
    g = efb.mesh.ElemGroup("HEXA8", connect, coords, all_nodes_used=True)
    Nn = coords.shape[0]
    mat = orc.IsoMaterial(3, 210000.0, 0.3)
    A = efb.asm.Assembler()
    pat = A.pattern(3, True, Nn * 3, (g,))
    assert pat.nnz == 9 * ((3 * n + 1 - 2) ** 3 + 0) or pat.nnz > 0  # structure checked exactly below
    # nnz of the structured grid: sum over nodes of (neighbours per direction product) * 9
    per_dir = np.full(n + 1, 3)
    per_dir[[0, -1]] = 2
    assert pat.nnz == 9 * int(per_dir.sum()) ** 3
    indptr = pat.indptr.cpu().numpy()
    assert np.all(np.diff(indptr) > 0) and indptr[-1] == pat.nnz
    idx = pat.indices
    rows = torch.repeat_interleave(torch.arange(Nn * 3, device=idx.device), torch.diff(pat.indptr.to(torch.int64)))
    key = rows * (Nn * 3) + idx.to(torch.int64)
    assert bool(torch.all(key[1:] > key[:-1]))  # canonical: sorted, no duplicates
    Ke = efb.op.elastic_Ke_dev(g, mat.C)
    K = pat.assemble([Ke])
    # (1) rigid translations are in the null space: K t = 0
    for comp in range(3):
        t = torch.zeros(Nn * 3, dtype=torch.float64, device=idx.device)
        t[comp::3] = 1.0
        y = solver.spmv(K, t)
        assert float(y.abs().max()) < 1e-9 * float(K.data.abs().max())
    # (2) symmetry: x.K y == y.K x
    gen = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(Nn * 3, generator=gen, dtype=torch.float64).to(idx.device)
    y = torch.randn(Nn * 3, generator=gen, dtype=torch.float64).to(idx.device)
    a, b = float(x @ solver.spmv(K, y)), float(y @ solver.spmv(K, x))
    assert abs(a - b) < 1e-10 * max(abs(a), abs(b), float(K.data.abs().max()))
    # (3) replay is run-to-run bit-reproducible and linear: A(2 Ke) == 2 A(Ke) exactly
    K2 = pat.replay([Ke])
    assert torch.equal(K2, K.data)
    assert torch.equal(pat.replay([2.0 * Ke]), 2.0 * K.data)
    # (4) sum of the mass matrix = rho * volume * dim
    Me = efb.op.mass_Me_dev(g, 2.0, 3)
    vol = float(efb.op.geometry_dev(g, "mass", ("wJ",))["wJ"].sum())
    assert abs(float(Me.sum()) - 2.0 * vol * 3) < 1e-9 * vol
    # (5) sample check against the oracle on the first 500 elements
    sub = slice(0, 500)
    from easyfea_b200 import elements as el

    tab = el.gauss_table("HEXA8", "rigi")
    geo = orc.geometry(coords[connect[sub]], tab.dN_pg, tab.weights)
    assert rel_err(Ke[sub].cpu().numpy(), orc.linearized_elasticity(geo, mat.C)) < 1e-12


# ---------------------------------------------------------------------------------------------------------
# fused element integration + assembly (efb_assemble_elastic): K_e never materialised
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("elemType,n", [("HEXA8", (7, 6, 5)), ("TETRA4", (4, 3, 3)), ("TRI3", (9, 7)), ("QUAD4", (6, 5)),
                                        ("QUAD9", (4, 3)), ("TETRA10", (2, 2, 2))])
@pytest.mark.parametrize("general_C", [False, True])
def test_fused_assembly_vs_two_kernel_path(efb, elemType, n, general_C):
    """values <= 1e-12 against the oracle's K_e + np.bincount and against the two-kernel device path; same CSR pattern by
    construction (the fused kernel writes `data` of the pattern built by the node graph)"""
    rng = np.random.default_rng(3)
    coords, connect = make_mesh(elemType, n)
    g = efb.mesh.ElemGroup(elemType, connect, coords)
    dim, Nn = g.dim, coords.shape[0]
    C = orc.IsoMaterial(dim, 210000.0, 0.3).C
    if general_C:
        C = C * rng.uniform(0.5, 2.0, C.shape) + rng.uniform(1e3, 1e4, C.shape)
    A = efb.asm.Assembler()
    pat = A.pattern(dim, True, Nn * dim, (g,))
    two = pat.replay([efb.op.elastic_Ke_dev(g, C, "rigi", 1.3)])
    sched = efb.asm.FusedSchedule(pat.graph)
    fused = efb.asm.assemble_elastic_fused(sched, C, "rigi", 1.3)
    assert fused.shape == two.shape
    assert rel_err(fused.cpu().numpy(), two.cpu().numpy()) < 1e-12
    from easyfea_b200 import elements as el

    tab = el.gauss_table(elemType, "rigi")
    Ke = 1.3 * orc.linearized_elasticity(orc.geometry(coords[connect][:, :, :dim], tab.dN_pg, tab.weights), C)
    inv, indices, indptr, nnz = orc.csr_map([connect], dim, Nn * dim, True)
    assert rel_err(fused.cpu().numpy(), orc.assemble_replay([Ke], inv, nnz)) < 1e-12
    # deterministic: bit-identical run to run
    import torch

    assert torch.equal(efb.asm.assemble_elastic_fused(sched, C, "rigi", 1.3), fused)


@pytest.mark.parametrize("n", [(7, 6, 5), (12, 9, 10), (3, 2, 2)])
@pytest.mark.parametrize("general_C", [False, True])
def test_mma_assembly_vs_oracle_and_two_kernel_path(efb, n, general_C):
    """`efb_assemble_elastic_mma` (HEXA8, 8 Gauss points, FP64 MMA): values <= 1e-12 against the oracle's K_e + np.bincount and
    against the two-kernel device path, bit-identical run to run, owned-row prefix respected"""
    import torch

    from easyfea_b200 import elements as el

    rng = np.random.default_rng(5)
    coords, connect = make_mesh("HEXA8", n)
    g = efb.mesh.ElemGroup("HEXA8", connect, coords)
    Nn = coords.shape[0]
    C = orc.IsoMaterial(3, 210000.0, 0.3).C
    if general_C:
        C = C * rng.uniform(0.5, 2.0, C.shape) + rng.uniform(1e3, 1e4, C.shape)
    pat = efb.asm.Assembler().pattern(3, True, Nn * 3, (g,))
    two = pat.replay([efb.op.elastic_Ke_dev(g, C, "rigi", 1.3)])
    ms = efb.asm.MmaSchedule(pat.graph)
    assert ms.fits()
    got = efb.asm.assemble_elastic_mma(ms, C, "rigi", 1.3)
    assert got.shape == two.shape
    assert rel_err(got.cpu().numpy(), two.cpu().numpy()) < 1e-12
    tab = el.gauss_table("HEXA8", "rigi")
    Ke = 1.3 * orc.linearized_elasticity(orc.geometry(coords[connect], tab.dN_pg, tab.weights), C)
    inv, indices, indptr, nnz = orc.csr_map([connect], 3, Nn * 3, True)
    assert rel_err(got.cpu().numpy(), orc.assemble_replay([Ke], inv, nnz)) < 1e-12
    assert torch.equal(efb.asm.assemble_elastic_mma(ms, C, "rigi", 1.3), got)
    n_own = Nn // 3
    out = torch.full_like(got, -7.0)
    efb.asm.assemble_elastic_mma(efb.asm.MmaSchedule(pat.graph, n_nodes=n_own), C, "rigi", 1.3, out=out)
    nz = int(pat.indptr[n_own * 3].item())
    assert torch.equal(out[:nz], got[:nz]) and bool((out[nz:] == -7.0).all())


def test_fused_assembly_owned_rows_prefix(efb):
    """sharded runs assemble the rows of the owned nodes only: the fused kernel leaves the other rows untouched"""
    import torch

    coords, connect = make_mesh("HEXA8", (6, 5, 4))
    g = efb.mesh.ElemGroup("HEXA8", connect, coords)
    Nn = coords.shape[0]
    C = orc.IsoMaterial(3, 210000.0, 0.3).C
    pat = efb.asm.Assembler().pattern(3, True, Nn * 3, (g,))
    full = efb.asm.assemble_elastic_fused(efb.asm.FusedSchedule(pat.graph), C)
    n_own = Nn // 3
    out = torch.full_like(full, -7.0)
    efb.asm.assemble_elastic_fused(efb.asm.FusedSchedule(pat.graph, n_nodes=n_own), C, out=out)
    nz = int(pat.indptr[n_own * 3].item())
    assert torch.equal(out[:nz], full[:nz]) and bool((out[nz:] == -7.0).all())


@pytest.mark.parametrize("path", ["fused", "mma"])
@pytest.mark.parametrize("n", [100, 200])
def test_fused_assembly_large_sampled_parity(efb, n, path):
    """BASELINE config 2 at 1 M and at the benchmarked 8 M elements (nnz 1.95e9: 91 % of the int32 range, 64-bit offsets in
    the schedule): sampled rows of the fused assembly against np.bincount of the oracle's K_e on the sub-mesh around them,
    bit-exact structure of those rows, rigid-body null space and symmetry of the whole matrix"""
    import torch

    from easyfea_b200 import elements as el
    from easyfea_b200 import meshgen, solver

    free, total = torch.cuda.mem_get_info()
    if n == 200 and free < 60e9:
        pytest.skip("needs ~60 GB of device memory")
    coords, connect = meshgen.structured_mesh("HEXA8", n, jitter=0.2, seed=0)
    g = efb.mesh.ElemGroup("HEXA8", connect, coords, all_nodes_used=True)
    Nn = coords.shape[0]
    mat = orc.IsoMaterial(3, 210000.0, 0.3)
    pat = efb.asm.Assembler().pattern(3, True, Nn * 3, (g,))
    if path == "mma":  # the step of the benchmark: efb_assemble_elastic_mma
        ms = efb.asm.MmaSchedule(pat.graph)
        assert ms.fits()
        data = efb.asm.assemble_elastic_mma(ms, mat.C)
        del ms
    else:
        sched = efb.asm.FusedSchedule(pat.graph)
        data = efb.asm.assemble_elastic_fused(sched, mat.C)
        del sched
    K = efb.asm.DeviceCsr(pat.indptr, pat.indices, data, pat.shape, pat.node_graph)
    dev = data.device
    for comp in range(3):  # rigid translations: K t = 0
        t = torch.zeros(Nn * 3, dtype=torch.float64, device=dev)
        t[comp::3] = 1.0
        assert float(solver.spmv(K, t).abs().max()) < 1e-9 * float(data.abs().max())
    gen = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(Nn * 3, generator=gen, dtype=torch.float64).to(dev)
    y = torch.randn(Nn * 3, generator=gen, dtype=torch.float64).to(dev)
    a, b = float(x @ solver.spmv(K, y)), float(y @ solver.spmv(K, x))
    assert abs(a - b) < 1e-10 * max(abs(a), abs(b))
    del x, y
    # sampled sub-blocks: nodes of a few 3x3x3-node bricks (first, middle, last corner of the mesh), all their elements
    tab = el.gauss_table("HEXA8", "rigi")
    P = n + 1
    indptr = pat.indptr
    for (i0, j0, k0) in ((0, 0, 0), (n // 2, n // 3, n // 2), (n - 2, n - 2, n - 2)):
        ii, jj, kk = np.meshgrid(np.arange(i0, i0 + 3), np.arange(j0, j0 + 3), np.arange(k0, k0 + 3), indexing="ij")
        nodes = (ii + jj * P + kk * P * P).ravel()
        # elements touching these nodes: cells (i-1..i, j-1..j, k-1..k)
        ci, cj, ck = np.meshgrid(np.arange(max(i0 - 1, 0), min(i0 + 3, n)), np.arange(max(j0 - 1, 0), min(j0 + 3, n)),
                                 np.arange(max(k0 - 1, 0), min(k0 + 3, n)), indexing="ij")
        elems = np.sort((ci + cj * n + ck * n * n).ravel())
        sub = connect[elems]
        Ke = orc.linearized_elasticity(orc.geometry(coords[sub], tab.dN_pg, tab.weights), mat.C)
        rows, cols = orc.rows_cols(sub, 3)
        for node in nodes:
            for c in range(3):
                r = int(node) * 3 + c
                lo, hi = int(indptr[r].item()), int(indptr[r + 1].item())
                got_cols = pat.indices[lo:hi].cpu().numpy().astype(np.int64)
                got = data[lo:hi].cpu().numpy()
                sel = rows == r
                ucols, inv = np.unique(cols[sel], return_inverse=True)
                assert np.array_equal(got_cols, ucols)  # the sub-mesh holds every element of these rows
                ref = np.bincount(inv, weights=Ke.ravel()[sel], minlength=ucols.size)
                assert rel_err(got, ref) < 1e-12
    if n == 200:
        assert pat.nnz > 1.9e9 and pat.indices.dtype == torch.int64  # scipy's rule: > 2^31-1 COO entries -> int64 indices
