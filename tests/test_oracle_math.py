"""T0 KATs: the restated math substrate of the oracle.

The only in-tree tests of the reference that touch this path are Sophus' property tests
(thirdparty/Sophus/sophus/test_se3.cpp:43-60 sample transforms + tests.hpp group laws); they pin exp/log/Adj here.
Dense factorizations (Eigen, not vendored, version unpinned) are pinned against numpy at 1e-10.
"""
import numpy as np
import pytest
import orc


def _so3_exp(w):
    th = np.linalg.norm(w)
    Kx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + Kx
    return np.eye(3) + np.sin(th) / th * Kx + (1 - np.cos(th)) / th ** 2 * Kx @ Kx


def sophus_samples():
    """test_se3.cpp:43-60: SE3(SO3::exp(w), t) sample set."""
    pi = np.pi
    S = [((0.2, 0.5, 0.0), (0, 0, 0)), ((0.2, 0.5, -1.0), (10, 0, 0)), ((0, 0, 0), (0, 100, 5)), ((0, 0, 0.00001), (0, 0, 0)),
         ((0, 0, 0.00001), (0, -0.00000001, 0.0000000001)), ((0, 0, 0.00001), (0.01, 0, 0)), ((pi, 0, 0), (4, -5, 0))]
    out = [orc.se3_from_rt(_so3_exp(np.array(w, float)), np.array(t, float)) for w, t in S]
    a = orc.se3_from_rt(_so3_exp(np.array([0.2, 0.5, 0.0])), np.zeros(3))
    b = orc.se3_from_rt(_so3_exp(np.array([pi, 0, 0.0])), np.zeros(3))
    c = orc.se3_from_rt(_so3_exp(np.array([-0.2, -0.5, 0.0])), np.zeros(3))
    out.append(orc.se3_mul(orc.se3_mul(a, b), c))
    a2 = orc.se3_from_rt(_so3_exp(np.array([0.3, 0.5, 0.1])), np.array([2, 0, -7.0]))
    c2 = orc.se3_from_rt(_so3_exp(np.array([-0.3, -0.5, -0.1])), np.array([0, 6, 0.0]))
    out.append(orc.se3_mul(orc.se3_mul(a2, b), c2))
    return out


def T44(T7):
    M = np.eye(4); M[:3, :3] = orc.se3_rot(T7); M[:3, 3] = T7[4:]; return M


@pytest.mark.parametrize("i", range(9))
def test_exp_log_roundtrip(i):
    T = sophus_samples()[i]
    T2 = orc.se3_exp(orc.se3_log(T))
    assert np.allclose(T44(T), T44(T2), atol=1e-9)          # tests.hpp: expLogTest


def test_group_laws_and_adjoint():
    S = sophus_samples()
    for A in S:
        assert np.allclose(T44(orc.se3_mul(A, orc.se3_inv(A))), np.eye(4), atol=1e-10)
        for B in S[:4]:
            assert np.allclose(T44(orc.se3_mul(A, B)), T44(A) @ T44(B), atol=1e-9)
    # Adj: A exp(x) A^-1 == exp(Adj_A x)   (tests.hpp adjointTest)
    x = np.array([0.1, -0.2, 0.3, 0.02, -0.01, 0.03])
    for A in S:
        lhs = orc.se3_mul(orc.se3_mul(A, orc.se3_exp(x)), orc.se3_inv(A))
        rhs = orc.se3_exp(orc.se3_adj(A) @ x)
        assert np.allclose(T44(lhs), T44(rhs), atol=1e-8)


def test_exp_matches_matrix_exponential():
    from scipy.linalg import expm
    rng = np.random.default_rng(0)
    for _ in range(20):
        a = rng.normal(0, 0.5, 6)
        M = np.zeros((4, 4)); w = a[3:]
        M[:3, :3] = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]); M[:3, 3] = a[:3]
        assert np.allclose(T44(orc.se3_exp(a)), expm(M), atol=1e-12)
    # small-angle branch (so3.hpp:356-361)
    a = np.array([1.0, 2.0, 3.0, 1e-12, -2e-12, 1e-13])
    assert np.allclose(T44(orc.se3_exp(a))[:3, 3], a[:3], atol=1e-9)


def test_ldlt_against_numpy():
    rng = np.random.default_rng(1)
    for n in (6, 7, 8, 46):
        for _ in range(5):
            J = rng.normal(size=(n + 5, n)); A = J.T @ J + 1e-3 * np.eye(n)
            s = np.exp(rng.uniform(-3, 3, n)); A = A * s[:, None] * s[None, :]       # badly scaled like the SCALE_* system
            b = rng.normal(size=n)
            x = orc.ldlt_solve(A, b)
            assert np.allclose(x, np.linalg.solve(A, b), rtol=1e-8, atol=1e-10)
    # semi-definite: zero row/col -> D^+ gives 0 for that variable (Eigen LDLT::solve tolerance branch)
    A = np.diag([2.0, 0.0, 3.0]); x = orc.ldlt_solve(A, np.array([2.0, 5.0, 3.0]))
    assert np.allclose(x, [1.0, 0.0, 1.0])


def test_aff_light_and_pyr_levels():
    out = np.zeros(2)
    orc.lib().orc_aff_from_to(0.5, 2.0, 0.1, 3.0, 0.3, -1.0, out)       # NumType.h:149-158
    a = np.exp(0.3 - 0.1) * 2.0 / 0.5
    assert np.allclose(out, [a, -1.0 - a * 3.0])
    orc.lib().orc_aff_from_to(0.0, 2.0, 0.1, 3.0, 0.3, -1.0, out)       # zero exposure -> both forced to 1
    assert np.allclose(out, [np.exp(0.2), -1.0 - np.exp(0.2) * 3.0])
    L = orc.lib().orc_pyr_levels
    assert L(1200, 360) == 4 and L(1400, 360) == 4 and L(1920, 1200) == 5 and L(640, 192) == 4   # SURVEY §8 sizes
