#!/usr/bin/env python
"""dev tool (GPU box): measured relative errors of the S3/S4 builders and the degenerate-state identity vs the oracle"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easyfea_b200 import elements as el, mesh, phasefield as pf
from oracle import easyfea_oracle as orc
from tests.helpers import make_mesh, rel_err

for elemType, split, regu in [("TRI3", "Miehe", "AT2"), ("TETRA4", "He", "AT2"), ("QUAD9", "Amor", "AT1"), ("HEXA8", "Stress", "AT1"),
                              ("HEXA8", "Miehe", "AT2"), ("TRI3", "He", "AT2"), ("TETRA4", "Stress", "AT2")]:
    rng = np.random.default_rng(8)
    coords, connect = make_mesh(elemType)
    g = mesh.ElemGroup(elemType, connect, coords)
    dim, Nn = g.dim, coords.shape[0]
    th = 0.5 if dim == 2 else 1.0
    om = orc.IsoMaterial(dim, 210e9, 0.3, False)
    pfm = pf.PhaseFieldModel(pf.IsotropicMaterial(dim, 210e9, 0.3, False, th), split, regu, 2.7e3, 1e-2)
    u = rng.normal(size=Nn * dim) * 1e-5
    dmg = rng.uniform(0, 0.9, Nn)
    u_e = orc.locate_sol_e(u, connect, dim)
    tr, tm = el.gauss_table(elemType, "rigi"), el.gauss_table(elemType, "mass")
    geo_r = orc.geometry(coords[connect][:, :, :dim], tr.dN_pg, tr.weights)
    geo_m = orc.geometry(coords[connect][:, :, :dim], tm.dN_pg, tm.weights)
    Ke = pfm.elastic_Ke_dev(g, u, dmg).cpu().numpy()
    ref = th * orc.pf_elastic_Ke(geo_r, tr.N_pg, om, split, u_e, dmg[connect], clamp=True)
    old = rng.uniform(0, 1, (g.Ne, tm.nPg)) * float(np.median(orc.calc_psi(om, split, orc.strain(geo_m, u_e), True)[0]))
    Kd, Fd, psiP = pfm.damage_system_dev(g, u, old)
    rK, rF, rpsi = orc.pf_damage_system(geo_m, tm.N_pg, om, split, regu, 2.7e3, 1e-2, u_e, old, clamp=True)
    print(elemType, split, "S3 Ke %.2e  S4 psiP %.2e Kd %.2e Fd %.2e" % (rel_err(Ke, ref), rel_err(psiP.cpu().numpy(), rpsi),
          rel_err(Kd.cpu().numpy(), th * rK), rel_err(Fd.cpu().numpy(), th * rF[..., 0])))
