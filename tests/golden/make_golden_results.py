"""Mint post-processing fixtures from the LIVE reference (EasyFEA v3.5.1 at /root/reference, gmsh stubbed).

Run in the authoring container only:  python tests/golden/make_golden_results.py
`Simulations.Elastic.Result(...)` on small jittered meshes with a smooth + random displacement set through `_Set_solutions`:
per-element strain / stress components, von Mises, full Strain / Stress fields, nodal averages (`Mesh.Get_Node_Values`) and
the deformation energy `Wdef`, `Wdef_e` (`_Calc_Psi_Elas`)."""
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
CASES = {"TRI3": ((7, 6), dict(planeStress=True, thickness=0.7)), "QUAD9": ((4, 3), dict(planeStress=False, thickness=1.3)),
         "TETRA4": ((3, 3, 2), {}), "HEXA8": ((4, 3, 3), {}), "HEXA27": ((2, 2, 2), {})}


def main():
    rng = np.random.default_rng(11)
    for et, (n, kw) in CASES.items():
        dim = len(n)
        coords, connect = meshgen.structured_mesh(et, n, jitter=0.12, seed=4)
        g = GroupElemFactory.Create(ElemType(et), connect, coords)
        mesh = Mesh({ElemType(et): g})
        mat = Models.Elastic.Isotropic(dim, E=210000.0, v=0.3, **kw)
        simu = Simulations.Elastic(mesh, mat)
        Nn = coords.shape[0]
        x = coords[:, :dim]
        u = (0.01 * np.sin(2.0 * x + 0.3) + 0.002 * rng.standard_normal((Nn, dim))).ravel()
        simu._Set_solutions(simu.problemType, u)
        d = {"coords": coords, "connect": connect, "u": u, "C": np.asarray(mat.C), "thickness": np.array(kw.get("thickness", 1.0))}
        comps = ["xx", "yy", "xy"] if dim == 2 else ["xx", "yy", "zz", "yz", "xz", "xy"]
        for res in ["S" + c for c in comps] + ["E" + c for c in comps] + ["Svm", "Evm", "Stress", "Strain"]:
            d[res + "_e"] = np.asarray(simu.Result(res, nodeValues=False))
            d[res + "_n"] = np.asarray(simu.Result(res, nodeValues=True))
        d["Wdef"] = np.array(simu.Result("Wdef"))
        d["Wdef_e"] = np.asarray(simu.Result("Wdef_e", nodeValues=False))
        np.savez_compressed(os.path.join(OUT, f"results_{et}.npz"), **d)
        print(et, "Wdef", float(d["Wdef"]), "Svm max", float(d["Svm_e"].max()), d["Stress_e"].shape, d["Stress_n"].shape)


if __name__ == "__main__":
    main()
