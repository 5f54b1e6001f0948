"""Mint staggered phase-field fixtures from the LIVE reference (EasyFEA v3.5.1 at /root/reference, gmsh stubbed).

Run in the authoring container only:  python tests/golden/make_golden_staggered.py
Small versions of BASELINE configs 3 and 4: shear test on a structured square / cube, crack modelled as `d = 1` Dirichlet on
{y = L/2, x <= L/2} (the `openCrack=False` branch of examples/PhaseField/Shear.py), `Simulations.PhaseField.Solve` with
the reference's default direct solver, a few load steps with `Save_Iter` in between (history update)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference  # noqa: E402

import_reference()
from EasyFEA import Models, Simulations  # noqa: E402
from EasyFEA.FEM import ElemType, GroupElemFactory, Mesh  # noqa: E402

from easyfea_b200 import meshgen  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
L, l0, E, v, Gc = 1e-3, 1e-4, 210e9, 0.3, 2.7e3
CASES = {"TRI3_Miehe": ("TRI3", (16, 16), "Miehe", "AT2", [4e-6, 8e-6, 1.0e-5]),
         "TETRA4_He": ("TETRA4", (6, 6, 2), "He", "AT2", [4e-6, 8e-6]),
         "QUAD9_Amor": ("QUAD9", (6, 6), "Amor", "AT2", [4e-6, 8e-6])}


def bc_sets(coords, dim, n):
    """node sets from the lattice (exact, no float compares): crack, top, bottom, left, right"""
    x, y = coords[:, 0], coords[:, 1]
    tol = 1e-12
    crack = np.flatnonzero((np.abs(y - L / 2) < tol) & (x <= L / 2 + tol))
    top = np.flatnonzero(np.abs(y - L) < tol)
    bot = np.flatnonzero(np.abs(y) < tol)
    left = np.flatnonzero((np.abs(x) < tol) & (y > tol) & (y < L - tol))
    right = np.flatnonzero((np.abs(x - L) < tol) & (y > tol) & (y < L - tol))
    return crack, top, bot, left, right


def main():
    for name, (et, n, split, regu, loads) in CASES.items():
        dim = 2 if et in ("TRI3", "QUAD9") else 3
        lengths = (L, L) if dim == 2 else (L, L, L * n[2] / n[0])
        lattice, connect = meshgen.structured_mesh(et, n, lengths=lengths)
        # geometry: jittered nodes (generic strain states: on a regular mesh under pure shear tr(eps) is rounding noise and
        # the sign/Heaviside switches of the splits amplify 1e-16 differences to O(1)); node sets come from the lattice
        coords, _ = meshgen.structured_mesh(et, n, lengths=lengths, jitter=0.15, seed=5)
        g = GroupElemFactory.Create(ElemType(et), connect, coords)
        mesh = Mesh({ElemType(et): g})
        mat = Models.Elastic.Isotropic(dim, E=E, v=v, planeStress=False, thickness=1.0)
        pfm = Models.PhaseField(mat, split, regu, Gc, l0)
        simu = Simulations.PhaseField(mesh, pfm)
        crack, top, bot, left, right = bc_sets(lattice, dim, n)
        d = {"coords": coords, "connect": connect, "crack": crack, "top": top, "bot": bot, "left": left, "right": right,
             "loads": np.array(loads), "params": np.array([L, l0, E, v, Gc])}
        for k, dep in enumerate(loads):
            simu.Bc_Init()
            simu.add_dirichlet(crack, [1], ["d"], problemType="damage")
            simu.add_dirichlet(top, [dep, 0.5 * dep] + [0] * (dim - 2), simu.Get_unknowns()[:dim])  # shear + tension
            simu.add_dirichlet(bot, [0] * dim, simu.Get_unknowns())
            u, dmg, conv = simu.Solve(1e-3, 50, convOption=0)
            d[f"u_{k}"], d[f"d_{k}"] = np.array(u), np.array(dmg)
            d[f"Niter_{k}"] = simu._PhaseField__Niter
            d[f"psiP_{k}"] = np.asarray(simu._PhaseField__psiP_e_pg)
            simu.Save_Iter()
            print(name, k, "Niter", simu._PhaseField__Niter, "max d (free)", float(np.delete(dmg, crack).max()), "conv", conv)
        np.savez_compressed(os.path.join(OUT, f"staggered_{name}.npz"), **d)


def conv_options():
    """the energy / relative-increment convergence tests of `Simulations.PhaseField.Solve` (convOption 1, 2, 3,
    Simulations/_phasefield.py:354-397) on the TRI3 Miehe and TETRA4 He cases, two load steps each"""
    for name in ("TRI3_Miehe", "TETRA4_He"):
        et, n, split, regu, loads = CASES[name]
        dim = 2 if et in ("TRI3", "QUAD9") else 3
        lengths = (L, L) if dim == 2 else (L, L, L * n[2] / n[0])
        lattice, connect = meshgen.structured_mesh(et, n, lengths=lengths)
        coords, _ = meshgen.structured_mesh(et, n, lengths=lengths, jitter=0.15, seed=5)
        crack, top, bot, left, right = bc_sets(lattice, dim, n)
        for opt, tol in ((1, 1e-2), (2, 1e-3), (3, 5e-2)):
            g = GroupElemFactory.Create(ElemType(et), connect, coords)
            mesh = Mesh({ElemType(et): g})
            mat = Models.Elastic.Isotropic(dim, E=E, v=v, planeStress=False, thickness=1.0)
            simu = Simulations.PhaseField(mesh, Models.PhaseField(mat, split, regu, Gc, l0))
            d = {"coords": coords, "connect": connect, "crack": crack, "top": top, "bot": bot, "loads": np.array(loads[:2]),
                 "params": np.array([L, l0, E, v, Gc]), "tolConv": np.array(tol), "convOption": np.array(opt)}
            for k, dep in enumerate(loads[:2]):
                simu.Bc_Init()
                simu.add_dirichlet(crack, [1], ["d"], problemType="damage")
                simu.add_dirichlet(top, [dep, 0.5 * dep] + [0] * (dim - 2), simu.Get_unknowns()[:dim])
                simu.add_dirichlet(bot, [0] * dim, simu.Get_unknowns())
                u, dmg, conv = simu.Solve(tol, 60, convOption=opt)
                d[f"u_{k}"], d[f"d_{k}"] = np.array(u), np.array(dmg)
                d[f"Niter_{k}"], d[f"convIter_{k}"] = simu._PhaseField__Niter, simu._PhaseField__convIter
                d[f"psiP_{k}"] = np.asarray(simu._PhaseField__psiP_e_pg)
                d[f"Psi_Crack_{k}"], d[f"Psi_Elas_{k}"] = simu._Calc_Psi_Crack(), simu._Calc_Psi_Elas()
                simu.Save_Iter()
                print(name, "convOption", opt, "step", k, "Niter", d[f"Niter_{k}"], "convIter", d[f"convIter_{k}"], "conv", conv)
            np.savez_compressed(os.path.join(OUT, f"staggered_conv{opt}_{name}.npz"), **d)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "conv":
        conv_options()
    else:
        main()
