"""Host logic of the PCG consumer that needs no GPU: the Chebyshev recurrence coefficients handed to `efb_pcg_iterate_cheb` /
`efb_pcg_cheb_update` (solver.cheb_coefficients) against a NumPy restatement of the polynomial preconditioner, and the
lanes-per-node rules.  The device kernels are checked in tests/test_gpu_phasefield_solver.py and tests/test_gpu_dist.py."""
import numpy as np
import scipy.sparse as sp

from easyfea_b200 import solver


def _spd_system(n=400, seed=0):
    rng = np.random.default_rng(seed)
    A = sp.random(n, n, density=0.02, random_state=seed, format="csr")
    A = (A + A.T) * 0.5
    A = A + sp.diags(np.abs(A).sum(1).A1 + rng.uniform(0.1, 1.0, n))  # diagonally dominant: SPD
    return A.tocsr(), rng.standard_normal(n)


def _cheb_apply(A, invD, r, degree, lmin, lmax):
    """z = q(D^-1 A) D^-1 r with the coefficients the device kernels receive"""
    theta, coefs = solver.cheb_coefficients(degree, lmin, lmax)
    d = invD * r / theta
    z = d.copy()
    for c1, c2 in coefs:
        d = c1 * d + c2 * invD * (r - A @ z)
        z = z + d
    return z


def test_cheb_coefficients_reproduce_the_chebyshev_polynomial():
    """the recurrence is the classical Chebyshev semi-iteration for D^-1 A z = D^-1 r started from zero: after m steps the
    error polynomial is T_m((theta - x) / delta) / T_m(theta / delta) on the spectrum"""
    A, r = _spd_system()
    D = A.diagonal()
    invD = 1.0 / D
    S = (sp.diags(invD ** 0.5) @ A @ sp.diags(invD ** 0.5)).toarray()  # similar to D^-1 A, symmetric
    lam, V = np.linalg.eigh(S)
    lmin, lmax = 0.9 * lam[0], 1.05 * lam[-1]
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    for degree in (2, 3, 4, 6):
        z = _cheb_apply(A, invD, r, degree, lmin, lmax)
        # exact solution of A z* = r; error e = z* - z = p_m(D^-1 A) z*  with p_m(x) = T_m((theta - x)/delta) / T_m(theta/delta)
        zs = np.linalg.solve(A.toarray(), r)
        Tm = np.polynomial.chebyshev.Chebyshev.basis(degree)
        pm = Tm((theta - lam) / delta) / Tm(theta / delta)
        # in the symmetrised basis: D^(1/2) e = V diag(pm) V^T D^(1/2) z*
        e_ref = (V * pm) @ (V.T @ (D ** 0.5 * zs)) / D ** 0.5
        assert np.linalg.norm((zs - z) - e_ref) <= 1e-10 * np.linalg.norm(zs)


def test_polynomial_preconditioner_cuts_pcg_iterations():
    A, b = _spd_system(n=600, seed=3)
    # make it ill-conditioned: scale a block of rows/columns
    s = np.ones(A.shape[0])
    s[:50] = 1e-3
    A = (sp.diags(s) @ A @ sp.diags(s) + 1e-9 * sp.identity(A.shape[0])).tocsr()
    invD = 1.0 / A.diagonal()

    def pcg(prec, tol=1e-8):
        x = np.zeros_like(b)
        r = b.copy()
        z = prec(r)
        p = z.copy()
        rz = r @ z
        it = 0
        while np.linalg.norm(r) > tol * np.linalg.norm(b) and it < 5000:
            Ap = A @ p
            a = rz / (p @ Ap)
            x += a * p
            r -= a * Ap
            z = prec(r)
            rzn = r @ z
            p = z + (rzn / rz) * p
            rz = rzn
            it += 1
        return x, it

    xj, itj = pcg(lambda r: invD * r)
    v = np.ones(A.shape[0])
    for _ in range(solver.CHEB_POWER_ITERS):
        v = invD * (A @ v)
        lam = np.linalg.norm(v)
        v /= lam
    lmax = solver.CHEB_SAFETY * lam
    xc, itc = pcg(lambda r: _cheb_apply(A, invD, r, solver.CHEB_DEGREE, lmax / solver.CHEB_RATIO, lmax))
    assert itc < itj
    assert np.linalg.norm(xc - xj) <= 1e-6 * np.linalg.norm(xj)


def test_lanes_rules():
    # (nnz, nrows, dof_n) of the benchmark systems: TRI3 d=2 (7 neighbours), TETRA4 d=3 (15), HEXA8 d=3 (27), HEXA27 d=3 (~100)
    assert solver.lanes_per_node(28 * 10**6, 2 * 10**6, 2) == 2
    assert solver.lanes_per_node(45 * 3 * 10**6, 3 * 10**6, 3) == 4
    assert solver.lanes_per_node(81 * 3 * 10**6, 3 * 10**6, 3) == 16
    assert solver.cheb_lanes_per_node(28 * 10**6, 2 * 10**6, 2) == 2
    assert solver.cheb_lanes_per_node(45 * 3 * 10**6, 3 * 10**6, 3) == 4
    assert solver.cheb_lanes_per_node(81 * 3 * 10**6, 3 * 10**6, 3) == 4
    assert solver.cheb_lanes_per_node(300 * 3 * 10**6, 3 * 10**6, 3) == 8
