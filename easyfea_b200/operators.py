"""Element-level operators on the device — the drop-in for `EasyFEA.FEM.Operators.Bilinear / Linear` (level 1 of the
boundary) and for the `_GroupElem.Get_*_e_pg` geometry getters.

Same names, argument meaning and broadcast rules as the reference (`Operators/Bilinear.py:25-79,229-249`,
`Operators/Linear.py:18-52`, `FeArray.broadcast` `_linalg.py:426-476`): NumPy in, a fresh writable C-contiguous float64
ndarray out.  `*_dev` variants take and return torch CUDA tensors for device-resident pipelines (assembly, bench).
`groupElem` is an `easyfea_b200.mesh.ElemGroup` or a reference `_GroupElem`.
"""
from __future__ import annotations

import ctypes
import numbers

import numpy as np
import torch

from . import _lib
from . import device as dv
from .mesh import device_group

RIGI, MASS = "rigi", "mass"


def _mt(matrixType) -> str:
    return getattr(matrixType, "value", str(matrixType))


# ---------------------------------------------------------------------------------------------------------
# argument normalisation (host logic, unit-tested on CPU)
# ---------------------------------------------------------------------------------------------------------
def coef_mode(coef, Ne: int, nPg: int):
    """FeArray.broadcast rules for a scalar-valued coefficient -> (array or None, EFB_COEF_* mode, scalar)."""
    if isinstance(coef, numbers.Number) or isinstance(coef, (np.floating, np.integer)):
        return None, 0, float(coef)
    arr = coef.detach() if isinstance(coef, torch.Tensor) else np.asarray(coef, dtype=np.float64)
    shape = tuple(arr.shape)
    if len(shape) == 0:
        return None, 0, float(arr)
    if len(shape) == 2 and shape == (Ne, nPg):
        return arr, 3, 0.0
    if len(shape) == 1 and shape[0] == Ne:  # (Ne,) wins over (nPg,) exactly as in the reference
        return arr, 1, 0.0
    if len(shape) == 1 and shape[0] == nPg:
        return arr, 2, 0.0
    raise ValueError(f"coefficient of shape {shape} is not scalar, (Ne,)={Ne}, (nPg,)={nPg} or (Ne, nPg)")


def tensor_mode(T, Ne: int, nPg: int, n: int):
    """Leading-axis rule of FeArray.broadcast(tensor_ndim=2) -> (array, EFB_TENSOR_* mode)."""
    arr = T.detach() if isinstance(T, torch.Tensor) else np.asarray(T, dtype=np.float64)
    shape = tuple(arr.shape)
    if shape[-2:] != (n, n):
        raise ValueError(f"tensor coefficient must end with ({n}, {n}); got {shape}")
    lead = shape[:-2]
    if lead == (Ne, nPg):
        return arr, 2
    if lead == (Ne,):
        return arr, 1
    if lead == ():
        return arr, 0
    raise ValueError(f"With tensor_ndim=2, leading axes must be (), (Ne,), or (Ne, nPg); got {lead}.")


# ---------------------------------------------------------------------------------------------------------
# device-level API
# ---------------------------------------------------------------------------------------------------------
def geometry_dev(groupElem, matrixType, want=("jac",)):
    """Per-Gauss-point geometry on the device; `want` ⊂ {F, detF, jac, wJ, invF, dN, B} -> dict of tensors."""
    dg = device_group(groupElem)
    mt = _mt(matrixType)
    nPg, dim, nPe, Ne = dg.nPg(mt), dg.dim, dg.nPe, dg.Ne
    ns = 3 if dim == 2 else 6
    shapes = {"F": (Ne, nPg, dim, dim), "detF": (Ne, nPg), "jac": (Ne, nPg), "wJ": (Ne, nPg), "invF": (Ne, nPg, dim, dim),
              "dN": (Ne, nPg, dim, nPe), "B": (Ne, nPg, ns, nPe * dim)}
    out = {k: dv.empty(shapes[k]) for k in want}
    cs = dg.cstruct(mt)
    _lib.call("efb_geometry", cs, *[dv.ptr(out.get(k)) for k in ("F", "detF", "jac", "wJ", "invF", "dN", "B")], dv.stream_ptr())
    return out


def geometry_parts_dev(groupElem, matrixType, want=("leftDisp",), dof_n: int = 1):
    """The cached per-Gauss-point factors G8-G10 on the device; `want` ⊂ {leftDisp, reaction, diffuse, source}."""
    dg = device_group(groupElem)
    mt = _mt(matrixType)
    nPg, dim, nPe, Ne = dg.nPg(mt), dg.dim, dg.nPe, dg.Ne
    ns = 3 if dim == 2 else 6
    dof_n = int(dof_n)
    nd = nPe * dof_n
    shapes = {"leftDisp": (Ne, nPg, nPe * dim, ns), "reaction": (Ne, nPg, nd, nd), "diffuse": (Ne, nPg, nPe, dim),
              "source": (Ne, nPg, nd, dof_n)}
    out = {k: dv.empty(shapes[k]) for k in want}
    _lib.call("efb_geometry_parts", dg.cstruct(mt), dof_n, *[dv.ptr(out.get(k)) for k in ("leftDisp", "reaction", "diffuse", "source")],
              dv.stream_ptr())
    return out


def elastic_Ke_dev(groupElem, C, matrixType=RIGI, scale=1.0, out=None):
    dg = device_group(groupElem)
    mt = _mt(matrixType)
    ns = 3 if dg.dim == 2 else 6
    Carr, mode = tensor_mode(C, dg.Ne, dg.nPg(mt), ns)
    ndof = dg.nPe * dg.dim
    if out is None:
        out = dv.empty((dg.Ne, ndof, ndof))
    if mode == 0:  # homogeneous C: by value through the kernel arguments, no device copy needed
        Ch = np.ascontiguousarray(Carr.cpu().numpy() if isinstance(Carr, torch.Tensor) else Carr, dtype=np.float64)
        _lib.call("efb_elastic_Ke", dg.cstruct(mt), None, ctypes.c_void_p(Ch.ctypes.data), 0, float(scale), dv.ptr(out), dv.stream_ptr())
    else:
        Cd = dv.to_device(Carr)
        _lib.call("efb_elastic_Ke", dg.cstruct(mt), dv.ptr(Cd), None, mode, float(scale), dv.ptr(out), dv.stream_ptr())
    return out


def mass_Me_dev(groupElem, coef=1.0, dof_n=1, matrixType=MASS, scale=1.0, out=None):
    dg = device_group(groupElem)
    mt = _mt(matrixType)
    arr, mode, sc = coef_mode(coef, dg.Ne, dg.nPg(mt))
    cd = None if arr is None else dv.to_device(arr)
    ndof = dg.nPe * int(dof_n)
    if out is None:
        out = dv.empty((dg.Ne, ndof, ndof))
    _lib.call("efb_mass_Me", dg.cstruct(mt), dv.ptr(cd), mode, sc, int(dof_n), float(scale), dv.ptr(out), dv.stream_ptr())
    return out


def diffusion_Ke_dev(groupElem, A=None, coef=1.0, matrixType=RIGI, scale=1.0, out=None):
    dg = device_group(groupElem)
    mt = _mt(matrixType)
    nPg = dg.nPg(mt)
    arr, mode, sc = coef_mode(coef, dg.Ne, nPg)
    cd = None if arr is None else dv.to_device(arr)
    Ad, A_mode = None, 0
    if A is not None:
        Aarr, A_mode = tensor_mode(A, dg.Ne, nPg, dg.dim)
        Ad = dv.to_device(Aarr)
    if out is None:
        out = dv.empty((dg.Ne, dg.nPe, dg.nPe))
    _lib.call("efb_diffusion_Ke", dg.cstruct(mt), dv.ptr(Ad), A_mode, dv.ptr(cd), mode, sc, float(scale), dv.ptr(out),
              dv.stream_ptr())
    return out


def source_Fe_dev(groupElem, f=1.0, dof_n=1, matrixType=MASS, scale=1.0, out=None):
    dg = device_group(groupElem)
    mt = _mt(matrixType)
    arr, mode, sc = coef_mode(f, dg.Ne, dg.nPg(mt))
    fd = None if arr is None else dv.to_device(arr)
    dof_n = int(dof_n)
    if out is None:
        out = dv.empty((dg.Ne, dg.nPe * dof_n, dof_n))
    _lib.call("efb_source_Fe", dg.cstruct(mt), dv.ptr(fd), mode, sc, dof_n, float(scale), dv.ptr(out), dv.stream_ptr())
    return out


def internal_force_dev(groupElem, sigma_e_pg, matrixType=RIGI, out=None):
    dg = device_group(groupElem)
    mt = _mt(matrixType)
    ns = 3 if dg.dim == 2 else 6
    sd = dv.to_device(sigma_e_pg)
    if tuple(sd.shape) != (dg.Ne, dg.nPg(mt), ns):
        raise ValueError(f"sigma_e_pg must be (Ne, nPg, {ns}); got {tuple(sd.shape)}")
    if out is None:
        out = dv.empty((dg.Ne, dg.nPe * dg.dim))
    _lib.call("efb_internal_force", dg.cstruct(mt), dv.ptr(sd), dv.ptr(out), dv.stream_ptr())
    return out


def strain_dev(groupElem, u, matrixType=RIGI, out=None):
    """eps (Ne,nPg,ns) from the nodal displacement vector u (Ncoords*dim), `_laws.py:127-157`."""
    dg = device_group(groupElem)
    mt = _mt(matrixType)
    ns = 3 if dg.dim == 2 else 6
    ud = dv.to_device(u)
    if ud.numel() != dg.Ncoords * dg.dim:
        raise ValueError("Wrong dimension")  # Locates_sol_e, _group_elem.py:1783
    if out is None:
        out = dv.empty((dg.Ne, dg.nPg(mt), ns))
    _lib.call("efb_strain", dg.cstruct(mt), dv.ptr(dg.connect_glob), dv.ptr(ud), dv.ptr(out), dv.stream_ptr())
    return out


def hyperelastic_Ke_Re_dev(groupElem, u, dWde, d2Wde, matrixType=RIGI, scale=1.0, want=("Ke", "Re")):
    """Tangent and residual of a hyperelastic law at the state `u` from the law's dW/de (Ne,nPg,ns) and d2W/de2 (Ne,nPg,ns,ns)
    (`__second_piola_block`, Operators/NonLinear.py:124-141): device tensors (K_e (Ne,ndof,ndof), R_e (Ne,ndof)), interleaved dofs."""
    dg = device_group(groupElem)
    mt = _mt(matrixType)
    ns = 3 if dg.dim == 2 else 6
    nPg, ndof = dg.nPg(mt), dg.nPe * dg.dim
    ud, dW, d2W = dv.to_device(u), dv.to_device(dWde), dv.to_device(d2Wde)
    if ud.numel() != dg.Ncoords * dg.dim:
        raise ValueError("wrong displacement field dimension")  # HyperElasticState._CheckFormat, _state.py:26-32
    if tuple(dW.shape) != (dg.Ne, nPg, ns) or tuple(d2W.shape) != (dg.Ne, nPg, ns, ns):
        raise ValueError(f"dWde / d2Wde must be (Ne, nPg, {ns}) / (Ne, nPg, {ns}, {ns}); got {tuple(dW.shape)} / {tuple(d2W.shape)}")
    Ke = dv.empty((dg.Ne, ndof, ndof)) if "Ke" in want else None
    Re = dv.empty((dg.Ne, ndof)) if "Re" in want else None
    _lib.call("efb_hyperelastic_Ke_Re", dg.cstruct(mt), dv.ptr(dg.connect_glob), dv.ptr(ud), dv.ptr(dW), dv.ptr(d2W), float(scale),
              dv.ptr(Ke), dv.ptr(Re), dv.stream_ptr())
    return Ke, Re


# ---------------------------------------------------------------------------------------------------------
# reference-shaped API (NumPy in / NumPy out)
# ---------------------------------------------------------------------------------------------------------
def LinearizedElasticity(groupElem, C, matrixType=RIGI) -> np.ndarray:
    """``∫ ε(u):C:ε(v)`` -> (Ne, nPe·dim, nPe·dim); replaces Operators/Bilinear.py:62-79."""
    return dv.to_host(elastic_Ke_dev(groupElem, C, matrixType))


def UV(groupElem, coef=1.0, dof_n: int = 1, matrixType=MASS) -> np.ndarray:
    """``∫ coef u v`` -> (Ne, nPe·dof_n, nPe·dof_n); replaces Bilinear.py:42-59."""
    return dv.to_host(mass_Me_dev(groupElem, coef, dof_n, matrixType))


def GradUGradV(groupElem, coef=1.0, matrixType=RIGI) -> np.ndarray:
    """``∫ coef ∇u·∇v`` -> (Ne, nPe, nPe); replaces Bilinear.py:25-39."""
    return dv.to_host(diffusion_Ke_dev(groupElem, None, coef, matrixType))


def GradU_A_GradV(groupElem, A, coef=1.0, matrixType=RIGI) -> np.ndarray:
    """``∫ coef ∇u·A·∇v`` -> (Ne, nPe, nPe); replaces Bilinear.py:229-249."""
    return dv.to_host(diffusion_Ke_dev(groupElem, A, coef, matrixType))


def V(groupElem, f=1.0, dof_n: int = 1, matrixType=MASS) -> np.ndarray:
    """``∫ f v`` -> (Ne, nPe·dof_n, dof_n) (the reference keeps the trailing axis); replaces Linear.py:18-35."""
    return dv.to_host(source_Fe_dev(groupElem, f, dof_n, matrixType))


def InternalForce(groupElem, sigma_e_pg, matrixType=RIGI) -> np.ndarray:
    """``∫ σ:ε(v)`` -> (Ne, nPe·dim); replaces Linear.py:38-52."""
    return dv.to_host(internal_force_dev(groupElem, sigma_e_pg, matrixType))


def SecondPiolaKirchhoffStressTensor(material, state):
    """``(K_e, R_e)`` of a hyperelastic constitutive law at `state`, dofs ``(x1,y1,z1,...,xn,yn,zn)``; replaces
    `Operators.NonLinear.SecondPiolaKirchhoffStressTensor` (Operators/NonLinear.py:144-201).  `material` supplies
    `Compute_dWde(state)` / `Compute_d2Wde(state)` (Kelvin-Mandel) and `thickness`; `state` is a `HyperElasticState`
    (`groupElem`, `displacement`, `matrixType`).  The law itself stays the reference's; the kinematic operator, the material and
    geometric tangents and the residual are integrated on the device."""
    g = state.groupElem
    scale = float(material.thickness) if int(g.dim) == 2 else 1.0
    Ke, Re = hyperelastic_Ke_Re_dev(g, np.asarray(state.displacement), np.asarray(material.Compute_dWde(state)),
                                    np.asarray(material.Compute_d2Wde(state)), state.matrixType, scale)
    return dv.to_host(Ke), dv.to_host(Re)


def Calc_Epsilon_e_pg(groupElem, sol, matrixType=RIGI) -> np.ndarray:
    """ε = B·u_e -> (Ne, nPg, ns); replaces `_Elastic.Calc_Epsilon_e_pg`, Models/Elastic/_laws.py:127-157."""
    return dv.to_host(strain_dev(groupElem, sol, matrixType))


_GETTERS = {"Get_F_e_pg": "F", "Get_invF_e_pg": "invF", "Get_dN_e_pg": "dN", "Get_B_e_pg": "B",
            "Get_weightedJacobian_e_pg": "wJ"}


def Get_jacobian_e_pg(groupElem, matrixType, absoluteValues=True) -> np.ndarray:
    """|det F| (or signed) -> (Ne, nPg); replaces _group_elem.py:871-888."""
    key = "jac" if absoluteValues else "detF"
    return dv.to_host(geometry_dev(groupElem, matrixType, (key,))[key])


def _make_getter(name, key):
    def getter(groupElem, matrixType) -> np.ndarray:
        return dv.to_host(geometry_dev(groupElem, matrixType, (key,))[key])

    getter.__name__ = name
    getter.__doc__ = f"`_GroupElem.{name}` on the device (EasyFEA/FEM/_group_elem.py:832-1333)."
    return getter


for _n, _k in _GETTERS.items():
    globals()[_n] = _make_getter(_n, _k)


def Get_leftDispPart_e_pg(groupElem, matrixType) -> np.ndarray:
    """wJ·Bᵀ -> (Ne, nPg, nPe·dim, ns); replaces _group_elem.py:1314-1333."""
    return dv.to_host(geometry_parts_dev(groupElem, matrixType, ("leftDisp",))["leftDisp"])


def Get_ReactionPart_e_pg(groupElem, matrixType, dof_n: int = 1) -> np.ndarray:
    """wJ·NᵀN (block-diagonal N for dof_n > 1) -> (Ne, nPg, nPe·dof_n, nPe·dof_n); replaces _group_elem.py:1337-1360."""
    return dv.to_host(geometry_parts_dev(groupElem, matrixType, ("reaction",), dof_n)["reaction"])


def Get_DiffusePart_e_pg(groupElem, matrixType) -> np.ndarray:
    """wJ·∇Nᵀ -> (Ne, nPg, nPe, dim); replaces _group_elem.py:1362-1380."""
    return dv.to_host(geometry_parts_dev(groupElem, matrixType, ("diffuse",))["diffuse"])


def Get_SourcePart_e_pg(groupElem, matrixType, dof_n: int = 1) -> np.ndarray:
    """wJ·Nᵀ -> (Ne, nPg, nPe·dof_n, dof_n); replaces _group_elem.py:1382-1407."""
    return dv.to_host(geometry_parts_dev(groupElem, matrixType, ("source",), dof_n)["source"])
