"""Host logic of the device time stepping (easyfea_b200.transient) against the NumPy oracle, no GPU: the history-weight table
`b = F + sum_X X @ (x_u u + x_v v + x_a a)`, the system-matrix coefficients and the corrector formulas of every scheme
(_simu.py:1399-1455, 1552-1657, 1777-1853), on random sparse SPD-like matrices and random states."""
import numpy as np
import pytest
import scipy.sparse as sp

from easyfea_b200 import transient as tr
from oracle import easyfea_oracle as orc

SCHEMES = [("parabolic", {}), ("newmark", {}), ("newmark", {"beta": 0.3, "gamma": 0.6}), ("hht", {"alpha": 0.1}), ("midpoint", {}),
           ("hht_newmark", {"alpha": 1 / 6}), ("euler_implicit", {}), ("euler_explicit", {})]


def random_system(n=40, seed=0):
    rng = np.random.default_rng(seed)

    def spd():
        A = sp.random(n, n, density=0.2, random_state=rng.integers(1 << 30), format="csr")
        return (A @ A.T + sp.identity(n)).tocsr()

    return spd(), spd(), spd(), rng


def setup(algo, kw, K, C, M):
    o = orc.TransientOracle(K, C, M)
    dt = 0.07
    if algo == "parabolic":
        o.set_parabolic(dt, kw.get("alpha", 0.5))
        beta, gamma, alpha = 0.25, 0.5, kw.get("alpha", 0.5)
    else:
        o.set_hyperbolic(dt, algo, **kw)
        beta, gamma, alpha = o.beta, o.gamma, o.alpha
    return o, dt, beta, gamma, alpha


@pytest.mark.parametrize("algo,kw", SCHEMES)
def test_history_weights_match_reference_rhs(algo, kw):
    K, C, M, rng = random_system()
    o, dt, beta, gamma, alpha = setup(algo, kw, K, C, M)
    o.u, o.v, o.a = rng.standard_normal((3, K.shape[0]))
    F = rng.standard_normal(K.shape[0])
    b = F.copy()
    for name, (xu, xv, xa) in tr._history_weights(algo, dt, beta, gamma, alpha).items():
        b += {"K": K, "C": C, "M": M}[name] @ (xu * o.u + xv * o.v + xa * o.a)
    ref = o.rhs(F)
    assert np.linalg.norm(b - ref) <= 1e-12 * np.linalg.norm(ref)


@pytest.mark.parametrize("algo,kw", SCHEMES)
def test_coefs_match_oracle(algo, kw):
    K, C, M, _ = random_system(seed=1)
    o, dt, beta, gamma, alpha = setup(algo, kw, K, C, M)
    cK, cC, cM = tr.time_scheme_coefs(algo, dt, beta, gamma, alpha)
    A = cK * K + cC * C + cM * M
    assert abs(A - o.matrix()).max() <= 1e-12 * abs(A).max()


def test_hht_newmark_parameters_are_imposed():
    """`Solver_Set_Hyperbolic_Algorithm` overrides beta, gamma for hht_newmark (_simu.py:1266-1276)"""
    assert orc.hht_newmark_params(1 / 6) == (0.25 * (1 + 1 / 6) ** 2, 0.5 + 1 / 6)
    with pytest.raises(NotImplementedError):
        tr.time_scheme_coefs("unknown", 0.1)
    with pytest.raises(NotImplementedError):
        tr._history_weights("unknown", 0.1, 0.25, 0.5, 0.5)
