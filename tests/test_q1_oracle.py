"""The oracle's Q1 element (oracle.c etype 1: multilinear basis on axis-aligned cells, 2-point Gauss
rule per axis) pinned by closed forms.  Q1 is BASELINE configs[3]'s element and NOT a reference
element type, so there is no reference vector to pin it with: what is checked is
the textbook element matrices, the analytic Jacobian against differences, and the reference's
gauss assertion (test/gauss.ini:53-55) for the Q1 discretisation.  CPU only."""
import numpy as np
import pytest
import scipy.sparse as sp

import cases as K

ORC, INI, OMESH = K.ORC, K.INI, K.OMESH


def _matrix(om, form, x=None, t=0.0):
    rp, ci = om.pattern()
    vals = np.zeros(ci.size)
    om.jacobian(form, t, 1.0, np.zeros(om.ndofs) if x is None else x, rp, ci, vals)
    return sp.csr_matrix((vals, ci, rp), shape=(om.ndofs, om.ndofs)).toarray()


@pytest.mark.parametrize("dim,h", [(2, (0.5, 0.25)), (3, (0.5, 0.25, 2.0))])
def test_single_cell_mass_and_stiffness(dim, h):
    mesh = OMESH.structured(dim, [1] * dim, [0.0] * dim, list(h), element="cube")
    om = ORC.Model(INI.parse_ini(K.GAUSS), mesh)      # storage 1, D = 0.005
    vol = float(np.prod(h))
    M = _matrix(om, 1)
    Kd = _matrix(om, 0) / 0.005
    n = 1 << dim
    for a in range(n):
        for b in range(n):
            diff = a ^ b
            m1 = [(1 / 6 if (diff >> k) & 1 else 1 / 3) for k in range(dim)]
            assert abs(M[a, b] - vol * np.prod(m1)) <= 1e-15
            kab = 0.0
            for k in range(dim):
                f = (-1.0 if (diff >> k) & 1 else 1.0) / h[k] ** 2
                for l in range(dim):
                    if l != k:
                        f *= m1[l]
                kab += f
            assert abs(Kd[a, b] - vol * kab) <= 1e-13 * vol * max(1.0, abs(kab))
    assert np.allclose(Kd.sum(axis=1), 0.0, atol=1e-13)


@pytest.mark.parametrize("name", ["grayscott2d_q1", "grayscott3d_q1", "mitchell_schaefer_q1"])
def test_analytic_jacobian_is_the_derivative(name):
    case = K.Q1_CASES[name]
    om = case.oracle()
    x = K.rand_state(om.ndofs, 1)
    z = K.rand_state(om.ndofs, 2, -1.0, 1.0)
    rp, ci = om.pattern()
    for form in (0, 1):
        va, vn = np.zeros(ci.size), np.zeros(ci.size)
        om.jacobian(form, 0.1, 1.0, x, rp, ci, va)
        om.jacobian(form, 0.1, 1.0, x, rp, ci, vn, numerical=True)
        assert np.abs(va - vn).max() <= 1e-6 * np.abs(va).max()
        y = np.zeros(om.ndofs)
        om.jacobian_apply(form, 0.1, 1.0, x, z, y)
        A = sp.csr_matrix((va, ci, rp), shape=(om.ndofs, om.ndofs))
        assert np.abs(A @ z - y).max() <= 1e-13 * np.abs(y).max()
        eps = 1e-6
        rp_, rm_ = np.zeros(om.ndofs), np.zeros(om.ndofs)
        om.residual(form, 0.1, 1.0, x + eps * z, rp_)
        om.residual(form, 0.1, 1.0, x - eps * z, rm_)
        assert np.abs((rp_ - rm_) / (2 * eps) - y).max() <= 1e-7 * np.abs(y).max()


def test_pattern_is_the_27_point_stencil():
    om = K.Q1_CASES["grayscott3d_q1"].oracle()
    rp, ci = om.pattern()
    assert np.diff(rp).max() == 27 * 2 and np.diff(rp).min() == 8 * 2


def test_gauss_kat_on_cubes():
    """test/gauss.ini:38-55 on a 32^2 Q1 lattice: L2 error at t = 1.2 below 0.5, maximum principle."""
    cfg = INI.parse_ini(K.GAUSS)
    INI.set_key(cfg, "model.time_step_operator.type", "Alexander2")
    om = ORC.Model(cfg, OMESH.structured(2, [32, 32], [-1, -1], [2, 2], element="cube"))
    S = ORC.StepOperator(om)
    u, t = om.initial(1.0), 1.0
    for _ in range(4):
        u, ok = S.apply(u, t, 0.05)
        assert ok
        t += 0.05
    X = om.mesh.coords
    exact = np.exp(-(X ** 2).sum(axis=1) / (4 * t * 0.005)) / (4 * np.pi * t * 0.005)
    Mm = _matrix(om, 1) if om.ndofs < 1500 else None
    e = u - exact
    l2 = np.sqrt(e @ (Mm @ e)) if Mm is not None else np.sqrt((2 / 32) ** 2 * (e ** 2).sum())
    assert l2 <= 0.5
    assert u.max() <= 1 / (4 * np.pi * 0.005) and u.min() >= -1e-2


def test_poisson_kat_on_cubes():
    """test/poisson.ini:27-32: u = |x|^2 with Dirichlet data; L2 error below the reference's warn level."""
    case = K.Q1_CASES["poisson_q1"]
    om = case.oracle()
    S = ORC.StepOperator(om)
    u, ok = S.apply(om.initial(0.0), 0.0, 0.1)
    assert ok
    X = om.mesh.coords
    e = u - (X ** 2).sum(axis=1)
    assert np.sqrt((1 / 16) ** 2 * (e ** 2).sum()) <= 1e-2


def test_reduce_functionals_on_cubes():
    """[model.reduce] on Q1 cells in the oracle (3-point Gauss rule per axis, exact to degree 5): volume,
    a polynomial of degree (4, 5), a gradient functional of a bilinear field, the points-per-cell count.
    (The product does not build reduce on cubes yet and says so: DESIGN.md section 7.)"""
    cfg = INI.parse_ini(K.GAUSS + """
[model.reduce]
vol.evaluation.expression = integration_factor
poly.evaluation.expression = position_x^4 * position_y^5 * integration_factor
ugrad.evaluation.expression = (grad_u_x + 2*grad_u_y) * integration_factor
cells.evaluation.expression = integration_factor / entity_volume
""")
    mesh = OMESH.structured(2, [3, 5], [0, 0], [1.5, 2.0], element="cube")
    om = ORC.Model(cfg, mesh)
    X = mesh.coords
    u = 3 * X[:, 0] + 0.5 * X[:, 1] + X[:, 0] * X[:, 1]
    vals, status = ORC.reduce(om, u, 0.0)
    assert vals["vol"] == pytest.approx(3.0, rel=1e-14)
    assert vals["poly"] == pytest.approx(1.5 ** 5 / 5 * 2.0 ** 6 / 6, rel=1e-13)
    # d/dx = 3 + y, d/dy = 0.5 + x integrated over [0,1.5] x [0,2]
    assert vals["ugrad"] == pytest.approx((3 * 3.0 + 1.5 * 2.0) + 2 * (0.5 * 3.0 + 2.0 * 1.125), rel=1e-13)
    assert vals["cells"] == pytest.approx(15.0, rel=1e-13)


@pytest.mark.parametrize("dim,n", [(2, 32), (3, 16)])
def test_gauss_assertions_on_cubes_through_reduce(dim, n):
    """test/gauss.ini:38-55 in the reference's own reduce vocabulary, Q1 discretisation: no error fires."""
    cfg = INI.parse_ini(K.GAUSS + K.REDUCE["gauss"])
    INI.set_key(cfg, "model.time_step_operator.type", "Alexander2")
    om = ORC.Model(cfg, OMESH.structured(dim, [n] * dim, [-1] * dim, [2] * dim, element="cube"))
    S = ORC.StepOperator(om)
    u, t, _ = ORC.evolve(S, om.initial(1.0), 1.0, 1.2, 0.1)
    vals, status = ORC.reduce(om, u, t)
    assert max(status.values()) < 2 and vals["u_error"] <= 0.5
