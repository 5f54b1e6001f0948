"""TEST / BASELINE INFRASTRUCTURE: meshes and simulations of the LIVE reference (imported through oracle/ref_import.py) built
from the synthetic generators of `easyfea_b200.meshgen`.  gmsh is stubbed, so meshes are hand-built:
`GroupElemFactory.Create(ElemType.X, connect, coords)` + `Mesh({ElemType.X: group})`.
Users: tests/ and the reference legs of bench.py — never the product."""
import numpy as np

from easyfea_b200 import meshgen


def ref_mesh(EasyFEA, elemType, coords, connect, boundary=None):
    """reference `Mesh` of one main group (+ optional boundary group `(elemType, connect)` needed by add_surfLoad/add_lineLoad)"""
    from EasyFEA.FEM import ElemType, GroupElemFactory, Mesh

    groups = {}
    if boundary is not None:
        bt, bc = boundary
        groups[ElemType(bt)] = GroupElemFactory.Create(ElemType(bt), np.asarray(bc, dtype=int), coords)
    groups[ElemType(elemType)] = GroupElemFactory.Create(ElemType(elemType), np.asarray(connect, dtype=int), coords)
    return Mesh(groups)


def quad9_boundary_seg3(nx: int, ny: int):
    """SEG3 elements (gmsh node order: end, end, middle) on the four sides of the structured QUAD9 lattice
    (2nx+1) x (2ny+1), node id = i + j*(2nx+1)"""
    W, H = 2 * nx + 1, 2 * ny + 1
    nid = lambda i, j: i + j * W  # noqa: E731
    segs = []
    for r in range(ny):
        segs.append([nid(W - 1, 2 * r), nid(W - 1, 2 * r + 2), nid(W - 1, 2 * r + 1)])  # x = L
        segs.append([nid(0, 2 * r + 2), nid(0, 2 * r), nid(0, 2 * r + 1)])              # x = 0
    for c in range(nx):
        segs.append([nid(2 * c, 0), nid(2 * c + 2, 0), nid(2 * c + 1, 0)])              # y = 0
        segs.append([nid(2 * c + 2, H - 1), nid(2 * c, H - 1), nid(2 * c + 1, H - 1)])  # y = h
    return np.array(segs, dtype=np.int64)


def readme_cantilever(EasyFEA, nx=28, ny=3, L=120.0, h=13.0):
    """BASELINE config 1, `README.md:33-73`: 2D plane-stress beam, organised QUAD9 mesh (meshSize h/3 -> 28 x 3 elements,
    57 x 7 nodes), clamped at x = 0, surface load F/h/h on x = L.  Returns (simu, mesh, nodesX0, nodesXL)."""
    from EasyFEA import Models, Simulations

    coords, connect = meshgen.structured_mesh("QUAD9", (nx, ny), lengths=(L, h))
    mesh = ref_mesh(EasyFEA, "QUAD9", coords, connect, boundary=("SEG3", quad9_boundary_seg3(nx, ny)))
    mat = Models.Elastic.Isotropic(2, 210000, 0.3, planeStress=True, thickness=h)
    simu = Simulations.Elastic(mesh, mat)
    nodesX0 = mesh.Nodes_Conditions(lambda x, y, z: x == 0)
    nodesXL = mesh.Nodes_Conditions(lambda x, y, z: x == L)
    simu.add_dirichlet(nodesX0, [0, 0], ["x", "y"])
    simu.add_surfLoad(nodesXL, [-800 / h / h], ["y"])
    return simu, mesh, nodesX0, nodesXL


def shear_sets(lattice, L):
    """node sets of the shear test from the un-jittered lattice: crack {y = L/2, x <= L/2}, top, bottom"""
    x, y = lattice[:, 0], lattice[:, 1]
    tol = 1e-12 * max(L, 1.0)
    crack = np.flatnonzero((np.abs(y - L / 2) < tol) & (x <= L / 2 + tol))
    return crack, np.flatnonzero(np.abs(y - L) < tol), np.flatnonzero(np.abs(y) < tol)


def phasefield_case(EasyFEA, elemType, n, split, regu="AT2", L=1e-3, l0=1e-4, E=210e9, v=0.3, Gc=2.7e3, jitter=0.15, seed=5,
                    solver="History"):
    """small version of BASELINE configs 3/4 (shear test, crack as d = 1 Dirichlet); returns (simu, sets, dim)"""
    from EasyFEA import Models, Simulations

    dim = 2 if elemType in ("TRI3", "TRI6", "QUAD4", "QUAD9") else 3
    n = tuple(n)
    lengths = (L, L) if dim == 2 else (L, L, L * n[2] / n[0])
    lattice, connect = meshgen.structured_mesh(elemType, n, lengths=lengths)
    coords, _ = meshgen.structured_mesh(elemType, n, lengths=lengths, jitter=jitter, seed=seed)
    mesh = ref_mesh(EasyFEA, elemType, coords, connect)
    mat = Models.Elastic.Isotropic(dim, E=E, v=v, planeStress=False, thickness=1.0)
    pfm = Models.PhaseField(mat, split, regu, Gc, l0, solver=solver)
    simu = Simulations.PhaseField(mesh, pfm)
    return simu, shear_sets(lattice, L), dim


def apply_shear(simu, sets, dim, dep):
    crack, top, bot = sets
    simu.Bc_Init()
    simu.add_dirichlet(crack, [1], ["d"], problemType="damage")
    simu.add_dirichlet(top, [dep, 0.5 * dep] + [0] * (dim - 2), simu.Get_unknowns()[:dim])
    simu.add_dirichlet(bot, [0] * dim, simu.Get_unknowns())


def hexa8_elastic(EasyFEA, n, E=210000.0, v=0.3, jitter=0.2, seed=0):
    """BASELINE config 2 at `n` cells per side: (simu, group, C, Ndof) of the reference, jittered unit cube"""
    from EasyFEA import Models, Simulations

    coords, connect = meshgen.structured_mesh("HEXA8", n, jitter=jitter, seed=seed)
    mesh = ref_mesh(EasyFEA, "HEXA8", coords, connect)
    simu = Simulations.Elastic(mesh, Models.Elastic.Isotropic(3, E=E, v=v))
    g = mesh.Get_list_groupElem()[0]
    return simu, g, np.asarray(simu.material.C), mesh.Nn * 3, coords, connect
