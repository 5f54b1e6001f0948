"""TEST INFRASTRUCTURE — NumPy restatement ("port") of EasyFEA's element-integration + assembly hot path.

This file is the parity ORACLE for the CUDA path.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; nothing in `easyfea_b200/` does.

Parity pinning: the reference ships no golden vectors for this path.  The oracle is pinned
instead against the LIVE reference (v3.5.1) in the authoring container by `tests/test_oracle_vs_reference.py`
(skipped where `/root/reference` is absent) and against fixtures minted from the live reference by
`tests/golden/make_golden.py` (`tests/golden/*.npz`, checked everywhere).

Every function cites the reference lines it restates (paths relative to the reference root).  Arrays are
plain ndarrays: element axis `e`, Gauss-point axis `p`, Kelvin-Mandel strain axis of size ns = 3 (2D) / 6 (3D).
"""
from __future__ import annotations

import numpy as np

SQRT2 = np.sqrt(2.0)


# =========================================================================================================
# G2-G10: Gauss-point geometry                                         EasyFEA/FEM/_group_elem.py:832-1407
# =========================================================================================================
def det_small(A):
    """Closed-form determinant, EasyFEA/FEM/_linalg.py:533-573 (same association of products)."""
    n = A.shape[-1]
    if n == 1:
        return A[..., 0, 0]
    if n == 2:
        return A[..., 0, 0] * A[..., 1, 1] - A[..., 1, 0] * A[..., 0, 1]
    a, b, c = A[..., 0, 0], A[..., 0, 1], A[..., 0, 2]
    d, e, f = A[..., 1, 0], A[..., 1, 1], A[..., 1, 2]
    g, h, i = A[..., 2, 0], A[..., 2, 1], A[..., 2, 2]
    return a * (e * i - h * f) - b * (d * i - g * f) + c * (d * h - g * e)


def inv_small(A):
    """Adjugate / determinant inverse, EasyFEA/FEM/_linalg.py:576-656."""
    n = A.shape[-1]
    det = det_small(A)
    out = np.empty_like(A, dtype=float)
    if n == 1:
        return 1.0 / A
    if n == 2:
        out[..., 0, 0] = A[..., 1, 1]
        out[..., 0, 1] = -A[..., 0, 1]
        out[..., 1, 0] = -A[..., 1, 0]
        out[..., 1, 1] = A[..., 0, 0]
    else:
        for r in range(3):
            for c in range(3):
                # cofactor of entry (c, r) -> adjugate entry (r, c)
                rows = [k for k in range(3) if k != c]
                cols = [k for k in range(3) if k != r]
                minor = A[..., rows[0], cols[0]] * A[..., rows[1], cols[1]] - A[..., rows[1], cols[0]] * A[..., rows[0], cols[1]]
                out[..., r, c] = minor if (r + c) % 2 == 0 else -minor
    return out * (1.0 / det)[..., None, None]


def geometry(coord_e, dN_pg, w_pg):
    """All per-Gauss-point geometric factors of one element group.

    coord_e (Ne, nPe, dim) nodal coordinates gathered per element; dN_pg (nPg, dim, nPe); w_pg (nPg,).
    Returns dict with
      F     (Ne,nPg,dim,dim)  F[i,j] = sum_n dN[p,i,n] x[e,n,j]            _group_elem.py:832-869
      detF  (Ne,nPg)          signed determinant                            :871-888 (absoluteValues=False)
      jac   (Ne,nPg)          |det F|                                       :883-886
      wJ    (Ne,nPg)          jac * w_p                                     :890-900
      invF  (Ne,nPg,dim,dim)                                                :902-915
      dN    (Ne,nPg,dim,nPe)  invF @ dN_pg                                  :1083-1105
    """
    F = np.einsum("pin,enj->epij", dN_pg, coord_e)
    detF = det_small(F)
    jac = np.abs(detF)
    invF = inv_small(F)
    dN = np.einsum("epij,pjn->epin", invF, dN_pg)
    return {"F": F, "detF": detF, "jac": jac, "wJ": jac * w_pg[None, :], "invF": invF, "dN": dN}


def B_matrix(dN):
    """Kelvin-Mandel strain-displacement operator (Ne,nPg,ns,nPe*dim), _group_elem.py:1241-1312."""
    Ne, nPg, dim, nPe = dN.shape
    c = 1.0 / SQRT2
    ns = 3 if dim == 2 else 6
    B = np.zeros((Ne, nPg, ns, nPe * dim))
    X, Y, Z = (slice(k, None, dim) for k in range(3))
    gx, gy = dN[:, :, 0], dN[:, :, 1]
    if dim == 2:
        B[:, :, 0, X] = gx
        B[:, :, 1, Y] = gy
        B[:, :, 2, X] = gy * c
        B[:, :, 2, Y] = gx * c
    else:
        gz = dN[:, :, 2]
        B[:, :, 0, X] = gx
        B[:, :, 1, Y] = gy
        B[:, :, 2, Z] = gz
        B[:, :, 3, Y], B[:, :, 3, Z] = gz * c, gy * c
        B[:, :, 4, X], B[:, :, 4, Z] = gz * c, gx * c
        B[:, :, 5, X], B[:, :, 5, Y] = gy * c, gx * c
    return B


def N_rep(N_pg, dof_n):
    """Block-diagonal shape functions (nPg, dof_n, dof_n*nPe), _group_elem.py:1007-1046."""
    N = N_pg.reshape(N_pg.shape[0], -1)
    nPg, nPe = N.shape
    out = np.zeros((nPg, dof_n, dof_n * nPe))
    for r in range(dof_n):
        out[:, r, r::dof_n] = N
    return out


def geometry_parts(geo, N_pg, dof_n=1):
    """The cached per-Gauss-point factors of the reference, _group_elem.py:1314-1407:
    leftDisp = wJ B^T (:1315-1333), reaction = (wJ N^T) N (:1338-1360), diffuse = wJ dN^T (:1363-1380),
    source = wJ N^T (:1383-1407), N block-diagonal for dof_n > 1."""
    wJ = geo["wJ"][:, :, None, None]
    B = B_matrix(geo["dN"])
    Nr = N_rep(N_pg, dof_n)[None]  # (1, nPg, dof_n, ndof)
    NrT = np.swapaxes(Nr, -1, -2)
    return {"leftDisp": wJ * np.swapaxes(B, -1, -2), "reaction": (wJ * NrT) @ Nr,
            "diffuse": wJ * np.swapaxes(geo["dN"], -1, -2), "source": wJ * NrT}


def _lead(coef, Ne, nPg, tail=0):
    """Broadcast rule of FeArray.broadcast, EasyFEA/FEM/_linalg.py:426-476 -> array broadcastable to (Ne,nPg,...)."""
    a = np.asarray(coef, dtype=float)
    if tail:
        lead = a.shape[: a.ndim - tail]
        if lead == (Ne, nPg):
            return a
        if lead == (Ne,):
            return a[:, None]
        if lead == ():
            return a[None, None]
        raise ValueError(f"leading axes {lead}")
    if a.ndim == 0:
        return a
    if a.shape[:2] == (Ne, nPg):
        return a
    if a.ndim == 1 and a.shape[0] == Ne:
        return a[:, None]
    if a.ndim == 1 and a.shape[0] == nPg:
        return a[None, :]
    raise ValueError(f"cannot broadcast {a.shape} to ({Ne},{nPg})")


# =========================================================================================================
# O1-O4: operators                                 EasyFEA/FEM/Operators/Bilinear.py, Linear.py
# =========================================================================================================
def linearized_elasticity(geo, C):
    """K_e = sum_p (wJ B^T) C B, Bilinear.py:62-79 -> (Ne, ndof, ndof)."""
    B = B_matrix(geo["dN"])
    Ne, nPg = B.shape[:2]
    C = np.broadcast_to(_lead(C, Ne, nPg, tail=2), (Ne, nPg) + np.shape(C)[-2:])
    left = geo["wJ"][:, :, None, None] * np.swapaxes(B, -1, -2)
    return np.einsum("epij,epjk->eik", left @ C, B)


def uv(geo, N_pg, coef=1.0, dof_n=1):
    """M_e = sum_p coef wJ N^T N, Bilinear.py:42-59 -> (Ne, nPe*dof_n, nPe*dof_n)."""
    Ne, nPg = geo["wJ"].shape
    Nr = N_rep(N_pg, dof_n)
    NtN = np.einsum("pri,prj->pij", Nr, Nr)
    cw = np.broadcast_to(_lead(coef, Ne, nPg), (Ne, nPg)) * geo["wJ"]
    return np.einsum("ep,pij->eij", cw, NtN)


def grad_u_a_grad_v(geo, A=None, coef=1.0):
    """sum_p coef wJ dN^T (A) dN, Bilinear.py:25-39 and :229-249 -> (Ne, nPe, nPe)."""
    dN = geo["dN"]
    Ne, nPg = dN.shape[:2]
    cw = np.broadcast_to(_lead(coef, Ne, nPg), (Ne, nPg)) * geo["wJ"]
    if A is None:
        AdN = dN
    else:
        A = np.broadcast_to(_lead(A, Ne, nPg, tail=2), (Ne, nPg) + np.shape(A)[-2:])
        AdN = A @ dN
    return np.einsum("ep,epia,epib->eab", cw, dN, AdN)


def source_v(geo, N_pg, f=1.0, dof_n=1):
    """F_e = sum_p f wJ N^T, Linear.py:18-35 -> (Ne, nPe*dof_n, dof_n) (the reference keeps the trailing axis)."""
    Ne, nPg = geo["wJ"].shape
    Nr = N_rep(N_pg, dof_n)  # (nPg, dof_n, ndof)
    fw = np.broadcast_to(_lead(f, Ne, nPg), (Ne, nPg)) * geo["wJ"]
    return np.einsum("ep,pri->eir", fw, Nr)


def internal_force(geo, sigma_e_pg):
    """sum_p wJ B^T sigma, Linear.py:38-52 -> (Ne, ndof)."""
    B = B_matrix(geo["dN"])
    return np.einsum("ep,epsi,eps->ei", geo["wJ"], B, sigma_e_pg)


def locate_sol_e(sol, connect, dof_n):
    """u_e (Ne, nPe*dof_n): dof of node n, component i is n*dof_n + i, _group_elem.py:346-380, 1769-1788."""
    asm = assembly_e(connect, dof_n)
    return np.asarray(sol)[asm]


def strain(geo, u_e):
    """eps = B u_e, EasyFEA/Models/Elastic/_laws.py:127-157 -> (Ne, nPg, ns)."""
    return np.einsum("epsi,ei->eps", B_matrix(geo["dN"]), u_e)


# =========================================================================================================
# P2-P7: phase-field law                                             EasyFEA/Models/_phasefield.py
# =========================================================================================================
class IsoMaterial:
    """Isotropic Hooke law in Kelvin-Mandel form, EasyFEA/Models/Elastic/_laws.py:333-505."""

    def __init__(self, dim, E, v, planeStress=False, thickness=1.0):
        self.dim, self.E, self.v, self.planeStress, self.thickness = dim, float(E), float(v), bool(planeStress), thickness
        self.mu = E / (2 * (1 + v))  # :387-395
        lam = E * v / ((1 + v) * (1 - 2 * v))  # :376-385
        if dim == 2 and planeStress:
            lam = E * v / (1 - v**2)
        self.lam = lam
        self.bulk = lam + 2 * self.mu / dim  # :397-405
        ns = 3 if dim == 2 else 6
        I = np.zeros(ns)
        I[:dim] = 1.0
        # C = lam I(x)I + 2 mu Id in Kelvin-Mandel notation (what _Behavior :407-482 evaluates to)
        self.C = lam * np.outer(I, I) + 2 * self.mu * np.eye(ns)
        self.S = np.linalg.inv(self.C)
        lamC, Q = np.linalg.eigh(self.C)  # Get_sqrt_C_S :223-247
        self.sqrtC = (Q * np.sqrt(lamC)) @ Q.T
        self.inv_sqrtC = (Q / np.sqrt(lamC)) @ Q.T


def _vec_to_mat(v):
    """Kelvin-Mandel vector -> symmetric matrix, EasyFEA/Models/_utils.py:155-188."""
    ns = v.shape[-1]
    dim = 2 if ns == 3 else 3
    M = np.zeros(v.shape[:-1] + (dim, dim))
    for d in range(dim):
        M[..., d, d] = v[..., d]
    if dim == 2:
        M[..., 0, 1] = M[..., 1, 0] = v[..., 2] / SQRT2
    else:
        M[..., 1, 2] = M[..., 2, 1] = v[..., 3] / SQRT2
        M[..., 0, 2] = M[..., 2, 0] = v[..., 4] / SQRT2
        M[..., 0, 1] = M[..., 1, 0] = v[..., 5] / SQRT2
    return M


def _mat_to_vec(M):
    """symmetric matrix -> Kelvin-Mandel vector, EasyFEA/Models/_utils.py:191-222."""
    dim = M.shape[-1]
    if dim == 2:
        return np.stack([M[..., 0, 0], M[..., 1, 1], M[..., 0, 1] * SQRT2], axis=-1)
    return np.stack([M[..., 0, 0], M[..., 1, 1], M[..., 2, 2], M[..., 1, 2] * SQRT2, M[..., 0, 2] * SQRT2,
                     M[..., 0, 1] * SQRT2], axis=-1)


def _eig3(M, I, c2, c3, c1, arg):
    """3D eigenvalues/projectors for given case masks (broadcastable to (Ne,nPg)); Models/_phasefield.py:836-948."""
    I1 = M[..., 0, 0] + M[..., 1, 1] + M[..., 2, 2]
    MM = M @ M
    I2 = 0.5 * (I1**2 - (MM[..., 0, 0] + MM[..., 1, 1] + MM[..., 2, 2]))
    g = I1**2 - 3 * I2
    sg = np.sqrt(g)
    theta = 1 / 3 * np.arccos(arg)  # :833
    v1 = I1 / 3  # case 4 (triple eigenvalue) initialisation, :839-849
    v2 = I1 / 3
    v3 = I1 / 3
    M1 = np.zeros_like(M)
    M1[..., 0, 0] = 1.0
    M3 = np.zeros_like(M)
    M3[..., 2, 2] = 1.0
    Irg = (1 / 3 * (I1 - sg))[..., None, None] * I  # :855
    gm12 = g ** (-1 / 2)
    m = lambda c: np.broadcast_to(c, g.shape)  # noqa: E731
    mm = lambda c: m(c)[..., None, None]  # noqa: E731

    # case 2: two maximum eigenvalues, :861-874
    v1 = np.where(m(c2), v1 - 2 / 3 * sg, v1)
    v2 = np.where(m(c2), v2 + 1 / 3 * sg, v2)
    v3 = np.where(m(c2), v3 + 1 / 3 * sg, v3)
    M1 = np.where(mm(c2), gm12[..., None, None] * (Irg - M), M1)
    M3 = np.where(mm(c2), 0.5 * (I - M1), M3)
    # case 3: two minimum eigenvalues, :882-895 (applied on top of case 2, as the reference does)
    v1 = np.where(m(c3), v1 - 1 / 3 * sg, v1)
    v2 = np.where(m(c3), v2 - 1 / 3 * sg, v2)
    v3 = np.where(m(c3), v3 + 2 / 3 * sg, v3)
    M3 = np.where(mm(c3), gm12[..., None, None] * (M - Irg), M3)
    M1 = np.where(mm(c3), 0.5 * (I - M3), M1)
    # case 1: three distinct eigenvalues, :902-934
    v1 = np.where(m(c1), v1 + 2 / 3 * (sg * np.cos(2 * np.pi / 3 + theta)), v1)
    v2 = np.where(m(c1), v2 + 2 / 3 * (sg * np.cos(2 * np.pi / 3 - theta)), v2)
    v3 = np.where(m(c1), v3 + 2 / 3 * (sg * np.cos(theta)), v3)
    A1 = (M - v2[..., None, None] * I) @ (M - v3[..., None, None] * I) / ((v1 - v2) * (v1 - v3))[..., None, None]
    A3 = (M - v1[..., None, None] * I) @ (M - v2[..., None, None] * I) / ((v3 - v1) * (v3 - v2))[..., None, None]
    M1 = np.where(mm(c1), A1, M1)
    M3 = np.where(mm(c1), A3, M3)

    M1 = M1 / np.sqrt((M1**2).sum(axis=(-1, -2)))[..., None, None]  # :944-946
    M3 = M3 / np.sqrt((M3**2).sum(axis=(-1, -2)))[..., None, None]
    return np.stack([v1, v2, v3], axis=-1), M1, M3


def eigen_projectors(v, clamp=False):
    """Eigenvalues (ascending) and eigenprojectors of the tensor held in KM vector v (Ne,nPg,ns).

    Restates `_Eigen_values_vectors_projectors`, Models/_phasefield.py:751-948, including its quirk that the
    3D repeated-eigenvalue cases are selected PER ELEMENT (`np.where(test)[0]`, :863,884,904-906).

    clamp=True adds the build's documented repair policy: wherever the reference formulas
    give a non-finite decomposition at a Gauss point (2D: negative discriminant; 3D: Lode argument outside
    [-1,1], or g == 0 inside a non-degenerate element), that point alone is recomputed with the discriminant
    clamped at 0 / the argument clamped to [-1,1] and the case chosen per point.  Points where the reference is
    finite are untouched, so clamp=True == clamp=False wherever the reference returns numbers.
    Returns vals (Ne,nPg,dim), [M_i (Ne,nPg,dim,dim)].
    """
    M = _vec_to_mat(v)
    dim = M.shape[-1]
    I = np.broadcast_to(np.eye(dim), M.shape)
    with np.errstate(all="ignore"):
        if dim == 2:
            det = det_small(M)
            tr = M[..., 0, 0] + M[..., 1, 1]
            delta = tr**2 - 4 * det  # :784
            if clamp:
                delta = np.where(delta < 0, 0.0, delta)
            root = np.sqrt(delta)
            vals = np.stack([(tr - root) / 2, (tr + root) / 2], axis=-1)  # :786-788
            dv = vals[..., 0] - vals[..., 1]
            distinct = vals[..., 0] != vals[..., 1]
            dv_safe = np.where(dv == 0, 1.0, dv)
            M1 = np.zeros_like(M)
            M1[..., 0, 0] = 1.0
            cand = (M - vals[..., 1, None, None] * I) / dv_safe[..., None, None]  # :803
            M1[distinct] = cand[distinct]
            return vals, [M1, I - M1]

        I1 = M[..., 0, 0] + M[..., 1, 1] + M[..., 2, 2]
        MM = M @ M
        I2 = 0.5 * (I1**2 - (MM[..., 0, 0] + MM[..., 1, 1] + MM[..., 2, 2]))  # :813
        I3 = det_small(M)
        g = I1**2 - 3 * I2
        gnz = g != 0
        arg = 0.5 * (2 * I1**3 - 9 * I1 * I2 + 27 * I3)  # :822
        arg = np.where(gnz, arg / np.where(gnz, g ** (3 / 2), 1.0), arg)
        theta = 1 / 3 * np.arccos(arg)
        t2 = gnz & (theta == np.pi / 3)
        t3 = gnz & (theta == 0)
        t1 = gnz & (theta != 0) & (theta != np.pi / 3)
        e2 = t2.any(axis=1)[:, None]  # element-level selection, :863
        e3 = t3.any(axis=1)[:, None]  # :884
        e1 = t1.any(axis=1)[:, None] & ~(e2 | e3)  # :904-906
        vals, M1, M3 = _eig3(M, I, e2, e3, e1, arg)
        if clamp:
            bad = ~(np.isfinite(vals).all(-1) & np.isfinite(M1).all((-1, -2)) & np.isfinite(M3).all((-1, -2)))
            gpos = g > 0  # g <= 0 or NaN counts as the triple-eigenvalue case
            argc = np.where(gpos, np.clip(np.nan_to_num(arg, nan=0.0), -1.0, 1.0), 0.0)
            thc = 1 / 3 * np.arccos(argc)
            p2 = gpos & (thc == np.pi / 3)
            p3 = gpos & (thc == 0)
            p1 = gpos & ~p2 & ~p3
            rv, rM1, rM3 = _eig3(M, I, p2, p3, p1, argc)
            vals = np.where(bad[..., None], rv, vals)
            M1 = np.where(bad[..., None, None], rM1, M1)
            M3 = np.where(bad[..., None, None], rM3, M3)
        M2 = I - (M1 + M3)
        return vals, [M1, M2, M3]


_PAIR_I = np.array([0, 1, 2, 1, 0, 0])
_PAIR_J = np.array([0, 1, 2, 2, 2, 1])


def spectral_projectors(v, clamp=False):
    """projP, projM (Ne,nPg,ns,ns) with vP = projP v; `__Spectral_Decomposition`, Models/_phasefield.py:1043-1243."""
    vals, Ms = eigen_projectors(v, clamp)
    ns = v.shape[-1]
    with np.errstate(all="ignore"):
        valp = (vals + np.abs(vals)) / 2  # :1077
        H = np.heaviside(vals, 0.5)  # :1081
        ms = [_mat_to_vec(Mi) for Mi in Ms]

        def theta(a, b, half):
            den = vals[..., a] - vals[..., b]
            den = np.where(den == 0, 1.0, den)
            return (valp[..., a] - valp[..., b]) / (half * den)

        if ns == 3:
            beta = theta(0, 1, 1.0)  # :1092
            gam = H - beta[..., None]
            projP = beta[..., None, None] * np.eye(3)
            for a in range(2):
                projP = projP + gam[..., a, None, None] * (ms[a][..., :, None] * ms[a][..., None, :])
            return projP, np.eye(3) - projP

        projP = np.zeros(v.shape[:2] + (6, 6))
        for a in range(3):
            projP = projP + H[..., a, None, None] * (ms[a][..., :, None] * ms[a][..., None, :])  # :1150-1160
        scale = np.ones((6, 6))
        scale[3:, :3] = scale[:3, 3:] = SQRT2
        scale[3:, 3:] = 2.0
        i, j = _PAIR_I[:, None], _PAIR_J[:, None]
        k, l = _PAIR_I[None, :], _PAIR_J[None, :]
        for (a, b) in ((0, 1), (0, 2), (1, 2)):
            A, Bm = Ms[a], Ms[b]
            G = (A[..., i, k] * Bm[..., j, l] + A[..., i, l] * Bm[..., j, k]
                 + Bm[..., i, k] * A[..., j, l] + Bm[..., i, l] * A[..., j, k]) * scale  # :1170-1196
            projP = projP + theta(a, b, 2.0)[..., None, None] * G
        return projP, np.eye(6) - projP


def _Rp_Rm(v, dim):
    """R+- = (1 +- sign(trace))/2 with the in-plane trace only in 2D, Models/_phasefield.py:485-501."""
    tr = v[..., 0] + v[..., 1]
    if dim == 3:
        tr = tr + v[..., 2]
    return (1 + np.sign(tr)) / 2, (1 + np.sign(-tr)) / 2


def calc_C(mat: IsoMaterial, split: str, eps, clamp=False):
    """(cP, cM) (Ne,nPg,ns,ns) for a strain field eps (Ne,nPg,ns); `Calc_C`, Models/_phasefield.py:396-431."""
    dim = mat.dim
    ns = eps.shape[-1]
    Ivec = np.zeros(ns)
    Ivec[:dim] = 1.0
    IxI = np.outer(Ivec, Ivec)  # :221-231
    Id = np.eye(ns)
    shape = eps.shape[:2] + (ns, ns)
    if split == "Bourdin":  # :433-449
        return np.broadcast_to(mat.C, shape).copy(), np.zeros(shape)
    if split == "Amor":  # :451-483
        Rp, Rm = _Rp_Rm(eps, dim)
        cP = mat.bulk * (Rp[..., None, None] * IxI) + 2 * mat.mu * (Id - 1 / dim * IxI)
        cM = mat.bulk * (Rm[..., None, None] * IxI)
        return cP, cM
    if split == "Miehe":  # :515-537
        pP, pM = spectral_projectors(eps, clamp)
        Rp, Rm = _Rp_Rm(eps, dim)
        cP = mat.lam * (Rp[..., None, None] * IxI) + 2 * mat.mu * pP
        cM = mat.lam * (Rm[..., None, None] * IxI) + 2 * mat.mu * pM
        return cP, cM
    if split == "Stress":  # :573-632
        sig = eps @ mat.C.T
        pP, pM = spectral_projectors(sig, clamp)
        Rp, Rm = _Rp_Rm(sig, dim)
        E, v, mu = mat.E, mat.v, mat.mu
        if dim == 2:
            a = (1 + v) / E
            b = v / E if mat.planeStress else v * (1 + v) / E
        else:
            a = 1 / (2 * mu)
            b = v / E
        sP = a * pP - b * Rp[..., None, None] * IxI
        sM = a * pM - b * Rm[..., None, None] * IxI
        return mat.C.T @ sP @ mat.C, mat.C.T @ sM @ mat.C
    if split == "He":  # :680-749
        epst = eps @ mat.sqrtC.T
        pPt, pMt = spectral_projectors(epst, clamp)
        pP = mat.inv_sqrtC @ pPt @ mat.sqrtC
        pM = mat.inv_sqrtC @ pMt @ mat.sqrtC
        return mat.C @ pP, mat.C @ pM
    raise ValueError(split)


def calc_psi(mat, split, eps, clamp=False):
    """psi+- = 1/2 eps . (c+- eps), Models/_phasefield.py:335-394 -> two (Ne,nPg)."""
    cP, cM = calc_C(mat, split, eps, clamp)
    sP = np.einsum("epij,epj->epi", cP, eps)
    sM = np.einsum("epij,epj->epi", cM, eps)
    return np.sum(0.5 * eps * sP, -1), np.sum(0.5 * eps * sM, -1)


def degradation(d_e, N_pg, k_res=1e-12):
    """g = (1 - N d_e)^2 + k_res, Models/_phasefield.py:295-317; d_e (Ne,nPe) -> (Ne,nPg)."""
    N = N_pg.reshape(N_pg.shape[0], -1)
    return (1 - d_e @ N.T) ** 2 + k_res


def pf_k(regu, Gc, l0):
    """diffusion coefficient, Models/_phasefield.py:236-251."""
    return 3 / 4 * Gc * l0 if regu == "AT1" else Gc * l0


def pf_r(regu, Gc, l0, psiP):
    """reaction term, Models/_phasefield.py:253-271."""
    return 2 * psiP if regu == "AT1" else 2 * psiP + Gc / l0


def pf_f(regu, Gc, l0, psiP):
    """source term, Models/_phasefield.py:273-293."""
    if regu == "AT1":
        f = 2 * psiP - (3 * Gc) / (8 * l0)
        return (f + np.abs(f)) / 2
    return 2 * psiP


def pf_elastic_Ke(geo, N_pg_rigi, mat, split, u_e, d_e, clamp=False):
    """S3: K_e of the displacement sub-problem, Simulations/_phasefield.py:444-482 (without thickness)."""
    eps = strain(geo, u_e)
    cP, cM = calc_C(mat, split, eps, clamp)
    g = degradation(d_e, N_pg_rigi)
    return linearized_elasticity(geo, g[..., None, None] * cP + cM)


def pf_damage_system(geo_mass, N_pg_mass, mat, split, regu, Gc, l0, u_e, psiP_old=None, A=None, clamp=False):
    """S4: (K_e, F_e, psiP) of the damage sub-problem, Simulations/_phasefield.py:494-573 (without thickness)."""
    eps = strain(geo_mass, u_e)
    psiP, _ = calc_psi(mat, split, eps, clamp)
    if psiP_old is not None:
        psiP = np.where(psiP - psiP_old < 0, psiP_old, psiP)  # history, :513-530
    dim = mat.dim
    A = np.eye(dim) if A is None else A
    K = uv(geo_mass, N_pg_mass, pf_r(regu, Gc, l0, psiP)) + grad_u_a_grad_v(geo_mass, A, pf_k(regu, Gc, l0))
    F = source_v(geo_mass, N_pg_mass, pf_f(regu, Gc, l0, psiP))
    return K, F, psiP


# =========================================================================================================
# A1-A3: assembly                                         EasyFEA/Simulations/_simu.py:989-1144
# =========================================================================================================
def assembly_e(connect, dof_n):
    """(Ne, nPe*dof_n) global dof of each local dof, _group_elem.py:346-380."""
    connect = np.asarray(connect, dtype=np.int64)
    return (connect[:, :, None] * dof_n + np.arange(dof_n)[None, None, :]).reshape(connect.shape[0], -1)


def rows_cols(connect, dof_n):
    """Flat COO coordinates of all element entries in k = (e, i, j) order, _group_elem.py:382-402."""
    asm = assembly_e(connect, dof_n)
    ndof = asm.shape[1]
    return np.repeat(asm, ndof, axis=1).ravel(), np.tile(asm, (1, ndof)).ravel()


def csr_map(connects, dof_n, Ndof, isMatrix=True):
    """(inv int32, indices, indptr, nnz) — the canonical sorted/unique pattern scipy's COO->CSR produces and the
    entry->slot map; `__Get_csr_map`, _simu.py:1062-1102.  Index dtype follows scipy: int32 unless too large."""
    rows, cols = [], []
    for c in connects:
        if isMatrix:
            r, cc = rows_cols(c, dof_n)
        else:
            r = assembly_e(c, dof_n).ravel()
            cc = np.zeros_like(r)
        rows.append(r)
        cols.append(cc)
    rows = np.concatenate(rows)
    cols = np.concatenate(cols)
    ncol = Ndof if isMatrix else 1
    key = rows * ncol + cols
    canon = np.unique(key)
    inv = np.searchsorted(canon, key).astype(np.int32)
    urow = canon // ncol
    idx_dtype = np.int32 if max(Ndof, key.size) <= np.iinfo(np.int32).max else np.int64
    indices = (canon % ncol).astype(idx_dtype)
    indptr = np.zeros(Ndof + 1, dtype=idx_dtype)
    np.cumsum(np.bincount(urow, minlength=Ndof), out=indptr[1:])
    return inv, indices, indptr, canon.size


def assemble_replay(datas, inv, nnz):
    """csr.data = bincount(inv, weights=concat(data)), `__Assemble_csr`, _simu.py:1037,1055 (ordered sum)."""
    data = np.concatenate([np.asarray(d).ravel() for d in datas])
    return np.bincount(inv, weights=data, minlength=nnz)


# =========================================================================================================
# consumers: Dirichlet solve and the staggered phase-field loop
#                          EasyFEA/Simulations/Solvers.py:502-553, Simulations/_phasefield.py:300-432, 444-578
# =========================================================================================================
def solve_dirichlet(A, b, dofs_known, x_known):
    """x with x[known] prescribed: A_uu x_u = b_u - A_uk x_k (`__Solver_1`, Solvers.py:502-553), direct solve like the
    reference's default `scipy` solver (`spsolve`)."""
    import scipy.sparse.linalg as spla

    n = A.shape[0]
    known = np.zeros(n, bool)
    known[dofs_known] = True
    x = np.zeros(n)
    x[dofs_known] = x_known
    free = ~known
    rhs = b[free] - (A[free] @ x)
    x[free] = spla.spsolve(A[free][:, free].tocsc(), rhs)
    return x


class StaggeredOracle:
    """`Simulations.PhaseField` reduced to what the staggered loop does on ONE element group (History solver, AT1/AT2)."""

    def __init__(self, coords, connect, dN_rigi, w_rigi, N_rigi, dN_mass, w_mass, N_mass, mat, split, regu, Gc, l0, thickness=1.0):
        import scipy.sparse as sp

        self.sp = sp
        dim = mat.dim
        self.dim, self.mat, self.split, self.regu, self.Gc, self.l0 = dim, mat, split, regu, Gc, l0
        self.connect, self.Nn = connect, coords.shape[0]
        self.geo_r = geometry(coords[connect][:, :, :dim], dN_rigi, w_rigi)
        self.geo_m = geometry(coords[connect][:, :, :dim], dN_mass, w_mass)
        self.N_r, self.N_m = N_rigi, N_mass
        self.thickness = thickness if dim == 2 else 1.0
        self.u, self.d = np.zeros(self.Nn * dim), np.zeros(self.Nn)
        self.psiP, self.psiP_old = None, None
        self.map_u = csr_map([connect], dim, self.Nn * dim, True)
        self.map_d = csr_map([connect], 1, self.Nn, True)
        self.bc_u, self.bc_d = ([], []), ([], [])

    def Bc_Init(self):
        self.bc_u, self.bc_d = ([], []), ([], [])

    def add_dirichlet(self, nodes, values, components, problemType="elastic"):
        dofs, vals = self.bc_d if problemType == "damage" else self.bc_u
        dn = 1 if problemType == "damage" else self.dim
        for v, c in zip(values, components):
            dofs.append(np.asarray(nodes) * dn + c)
            vals.append(np.full(len(nodes), float(v)))

    def _csr(self, Xe, m, n):
        inv, indices, indptr, nnz = m
        return self.sp.csr_matrix((assemble_replay([Xe], inv, nnz), indices, indptr), shape=(n, n))

    # -- matrices with the reference's "assemble only when the other field changed" flags --------------------------------
    # `Get_K_C_M_F` (Simulations/_phasefield.py:248-269): Kd is rebuilt when the displacement changed since its last
    # assembly (`__updatedDamage` False), Ku when the damage changed (`__updatedDisplacement` False).
    def _get_Kd(self):
        if not getattr(self, "_updatedDamage", False):
            u_e = locate_sol_e(self.u, self.connect, self.dim)
            Ke, Fe, self.psiP = pf_damage_system(self.geo_m, self.N_m, self.mat, self.split, self.regu, self.Gc, self.l0, u_e,
                                                 self.psiP_old)
            self._Kd = self._csr(Ke * self.thickness, self.map_d, self.Nn)
            self._Fd = np.bincount(self.connect.ravel(), weights=(Fe[..., 0] * self.thickness).ravel(), minlength=self.Nn)
            self._updatedDamage = True
        return self._Kd, self._Fd

    def _get_Ku(self):
        if not getattr(self, "_updatedDisplacement", False):
            u_e = locate_sol_e(self.u, self.connect, self.dim)
            Ke = pf_elastic_Ke(self.geo_r, self.N_r, self.mat, self.split, u_e, self.d[self.connect]) * self.thickness
            self._Ku = self._csr(Ke, self.map_u, self.Nn * self.dim)
            self._updatedDisplacement = True
        return self._Ku

    def Psi_Crack(self):
        """`_Calc_Psi_Crack` (:802-823): 1/2 d^T Kd d with the CURRENT damage matrix"""
        Kd, _ = self._get_Kd()
        return 0.0 if np.linalg.norm(self.d) == 0 else float(0.5 * self.d @ (Kd @ self.d))

    def Psi_Elas(self):
        """`_Calc_Psi_Elas` (:779-800): 1/2 u^T Ku u"""
        Ku = self._get_Ku()
        return 0.0 if np.linalg.norm(self.u) == 0 else float(0.5 * self.u @ (Ku @ self.u))

    def solve_damage(self):
        Kd, Fd = self._get_Kd()
        self.d = solve_dirichlet(Kd, Fd, np.concatenate(self.bc_d[0]), np.concatenate(self.bc_d[1]))
        self._updatedDisplacement = False  # new damage -> new displacement matrices (:367-368)
        return self.d

    def solve_elastic(self):
        Ku = self._get_Ku()
        self.u = solve_dirichlet(Ku, np.zeros(self.Nn * self.dim), np.concatenate(self.bc_u[0]), np.concatenate(self.bc_u[1]))
        self._updatedDamage = False  # new displacement -> new damage matrices (:372-373)
        return self.u

    def iterate(self):
        d_n = self.d
        self.solve_damage()
        self.solve_elastic()
        return np.max(np.abs(self.d - d_n))

    def Solve(self, tolConv=1.0, maxIter=500, convOption=0):
        """`Simulations.PhaseField.Solve` (:300-432): convOption 0 max|d_np1 - d_n|, 1 crack energy, 2 total energy (no
        external work: this driver has no Neumann loads), 3 summed relative increments of u and d (Pech 2022)."""
        Niter, converged = 0, False
        while not converged and Niter < maxIter:
            Niter += 1
            d_n, u_n = self.d, self.u
            if convOption == 1:
                E_n = self.Psi_Crack()
            elif convOption == 2:
                E_n = self.Psi_Crack() + self.Psi_Elas()
            d1 = self.solve_damage()
            u1 = self.solve_elastic()
            if convOption == 0:
                conv = np.max(np.abs(d1 - d_n))
            elif convOption in (1, 2):
                E1 = self.Psi_Crack()
                if convOption == 2:
                    E1 += self.Psi_Elas()
                conv = abs(E_n - E1) if E1 == 0 else abs((E_n - E1) / E1)
            else:
                diffU = np.abs(u1 - u_n)
                diffU[u1 != 0] *= 1 / np.abs(u1[u1 != 0])
                diffD = np.abs(d1 - d_n)
                diffD[d1 != 0] *= 1 / np.abs(d1[d1 != 0])
                convU, convD = np.sum(diffU), np.sum(diffD)
                conv = max(convD, convU)
            if tolConv == 1 or self.d.max() == 0:
                converged = True
            elif convOption == 3:
                converged = bool(convD <= tolConv and convU <= tolConv * 0.999)
            else:
                converged = bool(conv <= tolConv)
        self.Niter, self.convIter = Niter, conv
        return self.u, self.d, converged

    def Save_Iter(self):
        self.psiP_old = self.psiP


# =========================================================================================================
# consumers: time-scheme system build        EasyFEA/Simulations/_simu.py:1399-1455, 1552-1657, 1758-1894
# =========================================================================================================
PARABOLIC = "parabolic"
HYPERBOLIC = ("newmark", "hht", "midpoint", "hht_newmark", "euler_implicit", "euler_explicit")


def time_scheme_coefs(algo, dt, beta=0.25, gamma=0.5, alpha=0.5):
    """(coefK, coefC, coefM) of `A = coefK K + coefC C + coefM M`, `_Solver_Get_K_C_M_coefs_for_time_scheme` (_simu.py:1399-1455)."""
    if algo == "newmark":
        return 1.0, gamma / (beta * dt), 1 / (beta * dt**2)
    if algo == "hht":
        return 1 - alpha, (1 - alpha) * gamma / (beta * dt), (1 - alpha) / (beta * dt**2)
    if algo == "midpoint":
        return 0.5, 1 / dt, 2 / dt**2
    if algo == "hht_newmark":
        return 1 - alpha, gamma / (beta * dt), 1 / (beta * dt**2)
    if algo == "parabolic":
        return 1.0, 1 / (alpha * dt), 0.0
    if algo == "euler_implicit":
        return 1.0, 1 / dt, 1 / dt**2
    if algo == "euler_explicit":
        return 0.0, 0.0, 1.0
    raise NotImplementedError(algo)


def hht_newmark_params(alpha):
    """beta, gamma imposed by `Solver_Set_Hyperbolic_Algorithm` for hht_newmark (_simu.py:1266-1276)"""
    return 0.25 * (1 + alpha) ** 2, 0.5 + alpha


class TransientOracle:
    """One linear problem `K u + C v + M a = F` stepped by the reference's time schemes: right-hand side of
    `_Solver_Apply_Neumann` (_simu.py:1777-1853), system matrix of `_Solver_Apply_Dirichlet` (:1890-1894), elimination solve of
    `__Solver_1` (Solvers.py:502-553) and the corrector of `_Solver_Update_solutions` (_simu.py:1552-1657)."""

    def __init__(self, K, C=None, M=None):
        import scipy.sparse as sp

        n = K.shape[0]
        zero = sp.csr_matrix((n, n))
        self.K, self.C, self.M = K.tocsr(), (zero if C is None else C.tocsr()), (zero if M is None else M.tocsr())
        self.n = n
        self.u, self.v, self.a = np.zeros(n), np.zeros(n), np.zeros(n)
        self.algo = "elliptic"

    def set_parabolic(self, dt, alpha=0.5):
        self.algo, self.dt, self.alpha = "parabolic", dt, alpha

    def set_hyperbolic(self, dt, algo="newmark", beta=0.25, gamma=0.5, alpha=0.5):
        assert algo in HYPERBOLIC
        if algo == "hht_newmark":
            beta, gamma = hht_newmark_params(alpha)
        self.algo, self.dt, self.beta, self.gamma, self.alpha = algo, dt, beta, gamma, alpha

    def rhs(self, F):
        """b of `_Solver_Apply_Neumann` for the current (u_n, v_n, a_n)"""
        K, C, M, u, v, a, algo = self.K, self.C, self.M, self.u, self.v, self.a, self.algo
        b = np.array(F, dtype=float)
        if algo == "elliptic":
            return b
        dt = self.dt
        if algo == "parabolic":
            al = self.alpha
            ut = u + (1 - al) * dt * v
            return b + 1 / (al * dt) * (C @ ut)
        be, ga, al = self.beta, self.gamma, self.alpha
        if algo in ("newmark", "hht_newmark"):
            ut = u + dt * v + dt**2 / 2 * (1 - 2 * be) * a
            vt = v + dt * (1 - ga) * a
            b = b + (ga / (be * dt) * C + 1 / (be * dt**2) * M) @ ut - C @ vt
            if algo == "hht_newmark":
                b = b - al * (K @ u)
            return b
        if algo == "midpoint":
            return b + (2 / dt**2 * M + 1 / dt * C - 0.5 * K) @ u + 2 / dt * (M @ v)
        if algo == "hht":
            cM, cC = 1 / (be * dt**2), ga / (be * dt)
            b = b - ((al - 1) * (cM * M + cC * C) + al * K) @ u
            b = b - ((al - 1) / (be * dt) * M + ((al - 1) * (ga / be) + 1) * C) @ v
            b = b - (((al - 1) / (2 * be) + 1) * M + dt * (al - 1) * (ga / (2 * be) - 1) * C) @ a
            return b
        if algo == "euler_implicit":
            return b + (1 / dt**2 * M + 1 / dt * C) @ u + (1 / dt) * (M @ v)
        if algo == "euler_explicit":
            return b - K @ u - C @ v
        raise NotImplementedError(algo)

    def matrix(self):
        if self.algo == "elliptic":
            return self.K
        cK, cC, cM = time_scheme_coefs(self.algo, self.dt, getattr(self, "beta", 0.25), getattr(self, "gamma", 0.5), self.alpha)
        return (cK * self.K + cC * self.C + cM * self.M).tocsr()

    def update(self, x):
        """(u, v, a)_{n+1} from the solved unknown, `_Solver_Update_solutions`"""
        u, v, a, algo = self.u, self.v, self.a, self.algo
        if algo == "elliptic":
            return x, v, a
        dt = self.dt
        if algo == "parabolic":
            vt = u + (1 - self.alpha) * dt * v
            return x, (x - vt) / (self.alpha * dt), a
        be, ga = self.beta, self.gamma
        if algo in ("newmark", "hht_newmark"):
            ut = u + dt * v + dt**2 / 2 * (1 - 2 * be) * a
            vt = v + dt * (1 - ga) * a
            a1 = (x - ut) / (be * dt**2)
            return x, vt + ga * dt * a1, a1
        if algo == "midpoint":
            v1 = 2 / dt * (x - u) - v
            return x, v1, 2 / dt * (v1 - v) - a
        if algo == "hht":
            a1 = 1 / (be * dt) * ((x - u) / dt - v) + (1 - 1 / (2 * be)) * a
            return x, dt * ((1 - ga) * a + ga * a1) + v, a1
        if algo == "euler_implicit":
            v1 = (x - u) / dt
            return x, v1, (v1 - v) / dt
        if algo == "euler_explicit":
            return u + dt * v, v + dt * x, x
        raise NotImplementedError(algo)

    def Solve(self, F, dofs_known, x_known):
        """one step: Dirichlet values are those of the solved unknown (zero accelerations for euler_explicit, _simu.py:1903-1905)"""
        x_known = np.zeros(len(dofs_known)) if self.algo == "euler_explicit" else np.asarray(x_known, dtype=float)
        x = solve_dirichlet(self.matrix(), self.rhs(F), np.asarray(dofs_known), x_known)
        self.u, self.v, self.a = self.update(x)
        return self.u


# =========================================================================================================
# post-processing fields
#                 EasyFEA/Models/_utils.py:302-430, Models/Elastic/_laws.py:159-215, Simulations/_elastic.py:323-396, FEM/_mesh.py:822-873
# =========================================================================================================
def hooke(eps, C):
    """sigma = C eps at every Gauss point (`Calc_Sigma_e_pg`, _laws.py:159-185); C (ns,ns), (Ne,ns,ns) or (Ne,nPg,ns,ns)"""
    C = np.asarray(C)
    if C.ndim == 2:
        return np.einsum("ij,epj->epi", C, eps)
    if C.ndim == 3:
        return np.einsum("eij,epj->epi", C, eps)
    return np.einsum("epij,epj->epi", C, eps)


def field_result_e(field_e_pg, result, coef=SQRT2):
    """`Result_strain_or_stress_field_e` for one group (_utils.py:302-430): shear components / coef, then component, von Mises
    or the whole field, averaged over the Gauss points.  `result` like "Sxx", "Eyz", "Svm", "Stress", "Strain"."""
    f = np.array(field_e_pg, dtype=float)
    ns = f.shape[2]
    dim = 2 if ns == 3 else 3
    f[:, :, dim:] *= 1 / coef
    if dim == 2:
        xx, yy, xy = (f[:, :, i] for i in range(3))
        comp = {"xx": xx, "yy": yy, "xy": xy}
        vm = np.sqrt(xx**2 + yy**2 - xx * yy + 3 * xy**2)
    else:
        xx, yy, zz, yz, xz, xy = (f[:, :, i] for i in range(6))
        comp = {"xx": xx, "yy": yy, "zz": zz, "yz": yz, "xz": xz, "xy": xy}
        vm = np.sqrt(0.5 * ((xx - yy) ** 2 + (yy - zz) ** 2 + (zz - xx) ** 2 + 6 * (xy**2 + yz**2 + xz**2)))
    if result in ("Strain", "Stress"):
        return f.mean(1)
    if "vm" in result:
        return vm.mean(1)
    for key, val in comp.items():
        if key in result:
            return val.mean(1)
    raise ValueError(result)


def psi_elas_e(eps, C, wJ, thickness=1.0):
    """Wdef_e = int 1/2 sigma:eps (`_Calc_Psi_Elas`, _elastic.py:323-396, raw element stresses)"""
    psi = 0.5 * np.einsum("epi,epi->ep", hooke(eps, C), eps)
    return (thickness * wJ * psi).sum(1)


def node_values(connect, Nn, result_e):
    """`Mesh.Get_Node_Values` (_mesh.py:822-873): average of the values of the elements around each node"""
    import scipy.sparse as sp

    Ne, nPe = connect.shape
    r = np.asarray(result_e, dtype=float)
    is1d = r.ndim == 1
    r = r.reshape(Ne, -1)
    M = sp.csr_matrix((np.ones(Ne * nPe), (connect.ravel(), np.repeat(np.arange(Ne), nPe))), shape=(Nn, Ne))
    cnt = np.asarray(M.sum(axis=1)).reshape(-1, 1)
    out = (M @ r) * 1 / np.where(cnt == 0, 1, cnt)
    return out.ravel() if is1d else out


# =========================================================================================================
# section 8f rank 3: hyperelastic tangent / residual          EasyFEA/FEM/Operators/NonLinear.py:37-201
# =========================================================================================================
def hyper_deformation_gradient(geo, u_e, dim):
    """F = I + grad u at the Gauss points, (Ne,nPg,dim,dim), F[i][m] = delta_im + sum_a u_a,i dN_a/dx_m
    (`HyperElasticState.Compute_F`, Models/HyperElastic/_state.py:88-143, un-padded)."""
    ua = u_e.reshape(u_e.shape[0], -1, dim)  # (Ne, nPe, dim)
    return np.eye(dim)[None, None] + np.einsum("eai,epma->epim", ua, geo["dN"])


def hyper_De(G):
    """`__Build_De` (_state.py:320-366): rows = Kelvin-Mandel components of sym(G^T . d), columns = flat(d) = [d_x u_x, d_y u_x, ...]
    (component-major, derivative fastest)."""
    Ne, nPg, dim, _ = G.shape
    c = 1.0 / SQRT2
    if dim == 2:
        D = np.zeros((Ne, nPg, 3, 4))
        g = lambda i, j: G[:, :, i, j]  # noqa: E731
        D[:, :, 0, 0], D[:, :, 0, 2] = g(0, 0), g(1, 0)
        D[:, :, 1, 1], D[:, :, 1, 3] = g(0, 1), g(1, 1)
        D[:, :, 2, 0], D[:, :, 2, 1], D[:, :, 2, 2], D[:, :, 2, 3] = c * g(0, 1), c * g(0, 0), c * g(1, 1), c * g(1, 0)
        return D
    D = np.zeros((Ne, nPg, 6, 9))
    g = lambda i, j: G[:, :, i, j]  # noqa: E731
    rows = {0: [(0, (0, 0)), (3, (1, 0)), (6, (2, 0))], 1: [(1, (0, 1)), (4, (1, 1)), (7, (2, 1))],
            2: [(2, (0, 2)), (5, (1, 2)), (8, (2, 2))],
            3: [(1, (0, 2)), (2, (0, 1)), (4, (1, 2)), (5, (1, 1)), (7, (2, 2)), (8, (2, 1))],
            4: [(0, (0, 2)), (2, (0, 0)), (3, (1, 2)), (5, (1, 0)), (6, (2, 2)), (8, (2, 0))],
            5: [(0, (0, 1)), (1, (0, 0)), (3, (1, 1)), (4, (1, 0)), (6, (2, 1)), (7, (2, 0))]}
    for r, entries in rows.items():
        for col, (i, j) in entries:
            D[:, :, r, col] = (c if r >= 3 else 1.0) * g(i, j)
    return D


def hyper_Ke_Re(geo, u_e, dW, d2W, dim, thickness=1.0):
    """(K_e, R_e) of `SecondPiolaKirchhoffStressTensor` (NonLinear.py:144-201): B = De(u) grad, K_e = sum_p wJ (B^T d2W B) +
    geometric tangent `g (x) I` with g = sum_p wJ dN^T S dN (:75-96), R_e = sum_p wJ B^T dW; built component-major like the
    reference and returned in the interleaved dof order of its final reorder (:99-121)."""
    dN, wJ = geo["dN"], geo["wJ"]
    Ne, nPg, _, nPe = dN.shape
    F = hyper_deformation_gradient(geo, u_e, dim)
    De = hyper_De(F)
    grad = np.zeros((Ne, nPg, dim * dim, dim * nPe))  # flat(grad v) row (i, k) <- dof (component i, node a): dN[k][a]
    for i in range(dim):
        grad[:, :, i * dim:(i + 1) * dim, i * nPe:(i + 1) * nPe] = dN
    B = De @ grad
    A_lin = np.einsum("ep,epji,epjk,epkl->eil", wJ, B, d2W, B, optimize=True)
    S = _vec_to_mat(dW) if dim == 3 else np.stack([np.stack([dW[..., 0], dW[..., 2] / SQRT2], -1),
                                                   np.stack([dW[..., 2] / SQRT2, dW[..., 1]], -1)], -2)
    gmat = np.einsum("ep,epab,epac,epcd->ebd", wJ, dN, S, dN, optimize=True)
    A_geo = np.einsum("eab,jk->ejakb", gmat, np.eye(dim)).reshape(Ne, dim * nPe, dim * nPe)
    R = np.einsum("ep,epi,epij->ej", wJ, dW, B, optimize=True)
    K = (A_lin + A_geo) * thickness
    R = R * thickness
    perm = np.arange(nPe * dim).reshape(-1, nPe).T.ravel()
    return K[:, perm[:, None], perm[None, :]], R[:, perm]
