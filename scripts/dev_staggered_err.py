"""dev tool: print field errors of the device staggered loop vs the reference fixtures for every case/step and PCG tolerance"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easyfea_b200 import mesh, phasefield, staggered
from tests.helpers import rel_err
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for tol in (1e-9, 1e-11, 1e-13):
    for name, et, split in [("TRI3_Miehe", "TRI3", "Miehe"), ("TETRA4_He", "TETRA4", "He"), ("QUAD9_Amor", "QUAD9", "Amor")]:
        d = dict(np.load(os.path.join(GOLD, f"staggered_{name}.npz")))
        L, l0, E, v, Gc = d["params"]
        g = mesh.ElemGroup(et, d["connect"], d["coords"])
        dim = g.dim
        pfm = phasefield.PhaseFieldModel(phasefield.IsotropicMaterial(dim, E, v, False, 1.0), split, "AT2", Gc, l0)
        simu = staggered.PhaseFieldStaggered(staggered.LocalSystem(g), pfm, pcg_tol=tol)
        for k, dep in enumerate(d["loads"]):
            simu.Bc_Init()
            simu.add_dirichlet(d["crack"], [1], [0], problemType="damage")
            simu.add_dirichlet(d["top"], [dep, 0.5 * dep] + [0] * (dim - 2), list(range(dim)))
            simu.add_dirichlet(d["bot"], [0] * dim, list(range(dim)))
            u, dmg, conv = simu.Solve(1e-3, 50, convOption=0)
            print(tol, name, k, "Niter", simu.Niter, int(d[f"Niter_{k}"]), "err d %.2e u %.2e psi %.2e" % (
                rel_err(dmg.cpu().numpy(), d[f"d_{k}"]), rel_err(u.cpu().numpy(), d[f"u_{k}"]), rel_err(simu.psiP.cpu().numpy(), d[f"psiP_{k}"])),
                "pcg it", simu.info["damage"]["iterations"], simu.info["elastic"]["iterations"], flush=True)
            simu.Save_Iter()
