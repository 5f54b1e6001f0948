"""Phase-field law on the device — the drop-in for `Models.PhaseField.Calc_C / Calc_psi_e_pg / Get_g_e_pg /
Get_r_e_pg / Get_f_e_pg` (EasyFEA/Models/_phasefield.py:253-431; level 3 of the boundary) and for the simulation-level
builders `__Construct_Elastic_Matrix` / `__Calc_psiPlus_e_pg` / `__Construct_Damage_Matrix`
(EasyFEA/Simulations/_phasefield.py:444-573).

Splits on the path: Bourdin, Amor, Miehe, Stress, He with an isotropic material.  Degenerate states
(repeated eigenvalues) follow the reference wherever it returns numbers and are repaired per Gauss point where it
returns NaN (DESIGN.md "degenerate states").
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from . import device as dv
from . import operators as op
from .mesh import device_group

SPLIT_IDS = {"Bourdin": 0, "Amor": 1, "Miehe": 2, "Stress": 3, "He": 4}
REGU_IDS = {"AT1": 1, "AT2": 2}


def _name(x) -> str:
    return getattr(x, "value", str(x))


class IsotropicMaterial:
    """Isotropic Hooke law in Kelvin–Mandel form (`Models/Elastic/_laws.py:333-505`): C, λ, μ, bulk, C^{±1/2}."""

    def __init__(self, dim: int, E: float = 210000.0, v: float = 0.3, planeStress: bool = True, thickness: float = 1.0):
        assert dim in (2, 3)
        self.dim, self.E, self.v, self.thickness = dim, float(E), float(v), float(thickness)
        self.planeStress = bool(planeStress) if dim == 2 else False
        self.mu = self.E / (2 * (1 + self.v))
        lam = self.E * self.v / ((1 + self.v) * (1 - 2 * self.v))
        if dim == 2 and self.planeStress:
            lam = self.E * self.v / (1 - self.v**2)
        self.lam = lam
        self.bulk = lam + 2 * self.mu / dim
        ns = 3 if dim == 2 else 6
        I = np.zeros(ns)
        I[:dim] = 1.0
        self.C = lam * np.outer(I, I) + 2 * self.mu * np.eye(ns)
        w, Q = np.linalg.eigh(self.C)
        self.sqrtC = (Q * np.sqrt(w)) @ Q.T
        self.inv_sqrtC = (Q / np.sqrt(w)) @ Q.T

    @classmethod
    def from_reference(cls, material):
        """Read the parameters of a reference `Models.Elastic.Isotropic` object (homogeneous only)."""
        if not hasattr(material, "get_lambda") or np.ndim(material.E) != 0:
            raise NotImplementedError("the device phase-field law needs a homogeneous isotropic material")
        self = cls.__new__(cls)
        self.dim, self.E, self.v = int(material.dim), float(material.E), float(material.v)
        self.planeStress = bool(getattr(material, "planeStress", False)) if self.dim == 2 else False
        self.thickness = float(getattr(material, "thickness", 1.0))
        self.mu, self.lam, self.bulk = float(material.get_mu()), float(material.get_lambda()), float(material.get_bulk())
        self.C = np.array(material.C, dtype=np.float64)
        sC, isC = material.Get_sqrt_C_S()
        self.sqrtC, self.inv_sqrtC = np.array(sC, dtype=np.float64), np.array(isC, dtype=np.float64)
        return self


class PhaseFieldModel:
    """Parameters + device evaluation of the phase-field law.  Mirrors the reference constructor
    `PhaseField(material, split, regularization, Gc, l0, solver="History", A=None)` (:162-213)."""

    def __init__(self, material, split, regularization, Gc: float, l0: float, solver="History", A=None):
        self.material = material if isinstance(material, IsotropicMaterial) else IsotropicMaterial.from_reference(material)
        self.split, self.regularization, self.solver = _name(split), _name(regularization), _name(solver)
        if self.solver not in ("History", "HistoryDamage"):
            # BoundConstrain is a bound-constrained least-squares solve (scipy lsq_linear, Solvers.py:376-384): not a CG system
            raise NotImplementedError(f"phase-field solver {self.solver} is outside the device path (History, HistoryDamage)")
        if self.split not in SPLIT_IDS:
            raise NotImplementedError(f"split {self.split} is outside the hot path (Bourdin, Amor, Miehe, Stress, He)")
        assert self.regularization in REGU_IDS, "regu error"
        if np.ndim(Gc) != 0:
            raise NotImplementedError("heterogeneous Gc is out of scope")
        self.Gc, self.l0 = float(Gc), float(l0)
        self.A = np.eye(self.dim) if A is None else np.asarray(A, dtype=np.float64)

    @classmethod
    def from_reference(cls, pfm):
        return cls(pfm.material, pfm.split, pfm.regularization, pfm.Gc, pfm.l0, pfm.solver, np.asarray(pfm.A))

    @property
    def dim(self) -> int:
        return self.material.dim

    @property
    def thickness(self) -> float:
        return self.material.thickness

    @property
    def k(self) -> float:
        """diffusion term, :236-251"""
        return 3 / 4 * self.Gc * self.l0 if self.regularization == "AT1" else self.Gc * self.l0

    def _cmat(self) -> _lib.EfbPfMaterial:
        m = self.material
        c = _lib.EfbPfMaterial(m.dim, SPLIT_IDS[self.split], int(m.planeStress), 0, m.E, m.v, m.lam, m.mu, m.bulk)
        n = m.C.size
        c.C[:n] = m.C.ravel().tolist()
        c.sqrtC[:n] = m.sqrtC.ravel().tolist()
        c.inv_sqrtC[:n] = m.inv_sqrtC.ravel().tolist()
        return c

    # -- device level ------------------------------------------------------------------------------------------
    def split_dev(self, eps, want=("cP", "cM"), g_e_pg=None):
        """eps (Ne,nPg,ns) device/host -> dict with any of cP, cM, psiP, psiM, Cdeg (= g cP + cM)."""
        eps = dv.to_device(eps)
        Ne, nPg, ns = eps.shape
        assert ns == (3 if self.dim == 2 else 6)
        shapes = {"cP": (Ne, nPg, ns, ns), "cM": (Ne, nPg, ns, ns), "psiP": (Ne, nPg), "psiM": (Ne, nPg), "Cdeg": (Ne, nPg, ns, ns)}
        out = {k: dv.empty(shapes[k]) for k in want}
        gd = None
        if "Cdeg" in want:
            gd = dv.to_device(g_e_pg)
            assert tuple(gd.shape) == (Ne, nPg)
        bits = dv.empty((Ne,), torch.int32) if self.dim == 3 else None
        _lib.call("efb_pf_split", self._cmat(), dv.ptr(eps), Ne, nPg, dv.ptr(bits), dv.ptr(out.get("cP")), dv.ptr(out.get("cM")),
                  dv.ptr(out.get("psiP")), dv.ptr(out.get("psiM")), dv.ptr(gd), dv.ptr(out.get("Cdeg")), dv.stream_ptr())
        return out

    def sigma_dev(self, eps):
        """(SigmaP, SigmaM) (Ne,nPg,ns) = (cP eps, cM eps) on the device, Models/_phasefield.py:360-394"""
        eps = dv.to_device(eps)
        Ne, nPg, ns = eps.shape
        c = self.split_dev(eps, ("cP", "cM"))
        out = []
        for key in ("cP", "cM"):
            sig = dv.empty((Ne, nPg, ns))
            _lib.call("efb_hooke", Ne, nPg, ns, dv.ptr(eps), dv.ptr(c[key]), 2, dv.ptr(sig), dv.stream_ptr())
            out.append(sig)
        return out[0], out[1]

    def degradation_dev(self, d_n, groupElem, matrixType, k_res=1e-12):
        dg = device_group(groupElem)
        mt = op._mt(matrixType)
        dd = dv.to_device(d_n)
        assert dd.numel() == dg.Ncoords, "Dimension problem."
        out = dv.empty((dg.Ne, dg.nPg(mt)))
        _lib.call("efb_pf_degradation", dg.cstruct(mt), dv.ptr(dg.connect_glob), dv.ptr(dd), float(k_res), dv.ptr(out), dv.stream_ptr())
        return out

    def history_rf_dev(self, psiP, psiP_old=None, want_r=True, want_f=True):
        """In place psiP <- max(psiP, psiP_old); returns (r, f) of the damage problem."""
        n = psiP.numel()
        r = dv.empty(tuple(psiP.shape)) if want_r else None
        f = dv.empty(tuple(psiP.shape)) if want_f else None
        _lib.call("efb_pf_history_rf", dv.ptr(psiP), dv.ptr(psiP_old), n, REGU_IDS[self.regularization], self.Gc, self.l0,
                  dv.ptr(r), dv.ptr(f), dv.stream_ptr())
        return r, f

    # -- reference-shaped API (NumPy) ------------------------------------------------------------------------------
    def Calc_C(self, Epsilon_e_pg, verif=False):
        """(cP, cM) (Ne,nPg,ns,ns); replaces `PhaseField.Calc_C`, :396-431."""
        out = self.split_dev(np.asarray(Epsilon_e_pg), ("cP", "cM"))
        cP, cM = dv.to_host(out["cP"]), dv.to_host(out["cM"])
        if verif:  # the reference's own check, tests/Models/phasefield_test.py:120-124
            C = self.material.C
            assert np.linalg.norm(cP + cM - C) / np.linalg.norm(np.broadcast_to(C, cP.shape)) < 1e-12
        return cP, cM

    def Calc_psi_e_pg(self, Epsilon_e_pg):
        """(psiP, psiM) (Ne,nPg); replaces `PhaseField.Calc_psi_e_pg`, :335-358."""
        out = self.split_dev(np.asarray(Epsilon_e_pg), ("psiP", "psiM"))
        return dv.to_host(out["psiP"]), dv.to_host(out["psiM"])

    def Calc_Sigma_e_pg(self, Epsilon_e_pg):
        """(SigmaP, SigmaM) (Ne,nPg,ns); replaces :360-394."""
        sP, sM = self.sigma_dev(np.asarray(Epsilon_e_pg))
        return dv.to_host(sP), dv.to_host(sM)

    def Get_g_e_pg(self, d_n, groupElem, matrixType, k_res=1e-12):
        """g = (1 - N d)^2 + k_res (Ne,nPg); replaces :295-317."""
        return dv.to_host(self.degradation_dev(d_n, groupElem, matrixType, k_res))

    def Get_r_e_pg(self, PsiP_e_pg):
        psi = dv.to_device(np.asarray(PsiP_e_pg)).clone()
        return dv.to_host(self.history_rf_dev(psi, None, True, False)[0])

    def Get_f_e_pg(self, PsiP_e_pg):
        psi = dv.to_device(np.asarray(PsiP_e_pg)).clone()
        return dv.to_host(self.history_rf_dev(psi, None, False, True)[1])

    # -- simulation-level builders, device resident ---------------------------------------------------------------
    def elastic_Ke_dev(self, groupElem, u, d, matrixType=op.RIGI, fused: bool = True):
        """S3: K_e of the displacement sub-problem = LinearizedElasticity(g(d) cP + cM), x thickness in 2D.  One-point simplex
        elements (TRI3, TETRA4) take the one-pass kernel `efb_pf_elastic_Ke`; the others the four-kernel composition."""
        dg = device_group(groupElem)
        mt = op._mt(matrixType)
        scale = self.thickness if self.dim == 2 else 1.0
        if fused and dg.nPg(mt) == 1 and dg.nPe == dg.dim + 1:
            ud, dd = dv.to_device(u), dv.to_device(d)
            if ud.numel() != dg.Ncoords * dg.dim or dd.numel() != dg.Ncoords:
                raise ValueError("Wrong dimension")
            ndof = dg.nPe * dg.dim
            Ke = dv.empty((dg.Ne, ndof, ndof))
            _lib.call("efb_pf_elastic_Ke", self._cmat(), dg.cstruct(mt), dv.ptr(dg.connect_glob), dv.ptr(ud), dv.ptr(dd), 1e-12,
                      float(scale), dv.ptr(Ke), dv.stream_ptr())
            return Ke
        eps = op.strain_dev(groupElem, u, matrixType)
        g = self.degradation_dev(d, groupElem, matrixType)
        C = self.split_dev(eps, ("Cdeg",), g)["Cdeg"]
        return op.elastic_Ke_dev(groupElem, C, matrixType, scale)

    def damage_system_dev(self, groupElem, u, psiP_old=None):
        """S4: (K_e, F_e, psiP) of the damage sub-problem (psi+ at MASS points, history max, R + D, F), x thickness in 2D."""
        dg = device_group(groupElem)
        mt = op.MASS
        eps = op.strain_dev(groupElem, u, mt)
        psiP = self.split_dev(eps, ("psiP",))["psiP"]
        use_hist = self.solver == "History" and psiP_old is not None and tuple(psiP_old.shape) == tuple(psiP.shape)
        r, f = self.history_rf_dev(psiP, dv.to_device(psiP_old) if use_hist else None)
        Ke = dv.empty((dg.Ne, dg.nPe, dg.nPe))
        Fe = dv.empty((dg.Ne, dg.nPe))
        scale = self.thickness if self.dim == 2 else 1.0
        A = dv.to_device(self.A)
        _lib.call("efb_pf_damage_Ke_Fe", dg.cstruct(mt), dv.ptr(r), dv.ptr(f), dv.ptr(A), float(self.k), float(scale),
                  dv.ptr(Ke), dv.ptr(Fe), dv.stream_ptr())
        return Ke, Fe, psiP
