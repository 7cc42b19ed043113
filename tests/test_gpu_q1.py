"""GPU parity of the Q1 (cube cell) kernels -- BASELINE configs[3]'s element -- against the oracle's
own Q1 element (oracle.c etype 1; Q1 is not a reference element type).

Same bars as tests/test_gpu_parity.py: operator level <= 1e-12 relative L2, fields after time
stepping <= 1e-10, integer structures bit exact (tests/test_host_parity.py covers those on CPU).
"""
import numpy as np
import pytest

import cases as K

pytestmark = pytest.mark.gpu

OP_TOL = 1e-12
FIELD_TOL = 1e-10
ALL = list(K.Q1_CASES)


def rel(a, b):
    d = np.linalg.norm(a - b)
    n = np.linalg.norm(b)
    return d / n if n > 0 else d


def make(name, **over):
    import dune_copasi_b200 as D
    case = K.Q1_CASES[name]
    om = case.oracle(**over)
    cfg, model, grid = K.product_objects(case, **over)
    return case, om, cfg, model, grid, D.Operator(model, grid)


@pytest.mark.parametrize("name", ALL)
def test_residual(name):
    case, om, cfg, model, grid, op = make(name)
    x = K.rand_state(om.ndofs, 1)
    t = case.t0 + 0.3
    for wM, wA in ((1.0, 0.0), (0.0, 1.0), (-1.0, 0.0), (0.7, 0.3 * case.dt)):
        ref = np.zeros(om.ndofs)
        if wM:
            om.residual(1, t, wM, x, ref)
        if wA:
            om.residual(0, t, wA, x, ref)
        got = op.residual(t, wM, wA, x)
        assert rel(got, ref) <= OP_TOL, (name, wM, wA, rel(got, ref))
    base = K.rand_state(om.ndofs, 2)
    got = op.residual(t, 1.0, 0.5, x, base.copy())      # additive semantics
    ref = base.copy()
    om.residual(1, t, 1.0, x, ref)
    om.residual(0, t, 0.5, x, ref)
    assert rel(got, ref) <= OP_TOL


@pytest.mark.parametrize("name", ALL)
def test_jacobian_csr_and_apply(name):
    import scipy.sparse as sp
    case, om, cfg, model, grid, op = make(name)
    x = K.rand_state(om.ndofs, 3)
    t = case.t0 + 0.1
    rp, ci = om.pattern()
    assert op.nnz == ci.size
    wM, wA = 1.0, 0.25 * case.dt
    ref = np.zeros(ci.size)
    om.jacobian(1, t, wM, x, rp, ci, ref)
    om.jacobian(0, t, wA, x, rp, ci, ref)
    got = op.jacobian(t, wM, wA, x)
    assert rel(got, ref) <= OP_TOL, (name, rel(got, ref))
    # matrix-free apply = assembled matrix times z (constrained entries of z act as zero)
    z = K.rand_state(om.ndofs, 5, -1.0, 1.0)
    z2 = z.copy()
    cd, _ = om.constraints()
    z2[cd] = 0.0
    yref = np.zeros(om.ndofs)
    om.jacobian_apply(1, t, wM, x, z2, yref)
    om.jacobian_apply(0, t, wA, x, z2, yref)
    y = op.jacobian_apply(t, wM, wA, x, z)
    assert rel(y, yref) <= OP_TOL, (name, rel(y, yref))
    A = sp.csr_matrix((got, ci, rp), shape=(om.ndofs, om.ndofs))
    assert rel(y, A @ z2) <= OP_TOL


@pytest.mark.parametrize("name", ["grayscott3d_q1", "grayscott2d_q1", "gauss3d_q1", "mitchell_schaefer_q1"])
def test_block_diagonal(name):
    import scipy.sparse as sp
    case, om, cfg, model, grid, op = make(name)
    x = K.rand_state(om.ndofs, 6)
    t, wM, wA = case.t0, 1.0, 0.5 * case.dt
    rp, ci = om.pattern()
    vals = np.zeros(ci.size)
    om.jacobian(1, t, wM, x, rp, ci, vals)
    om.jacobian(0, t, wA, x, rp, ci, vals)
    A = sp.csr_matrix((vals, ci, rp), shape=(om.ndofs, om.ndofs))
    ns = om.comp_nspec[0]
    got = op.block_diagonal(t, wM, wA, x, om.ndofs * ns)
    ref = np.zeros(om.ndofs * ns)
    for b in range(om.ndofs // ns):
        ref[b * ns * ns:(b + 1) * ns * ns] = A[b * ns:(b + 1) * ns, b * ns:(b + 1) * ns].toarray().ravel()
    assert rel(got, ref) <= OP_TOL, (name, rel(got, ref))


@pytest.mark.parametrize("matrix_free", [False, True])
@pytest.mark.parametrize("prec", ["Jacobi", "BlockJacobi"])
@pytest.mark.parametrize("name", ["grayscott3d_q1", "gauss2d_q1", "poisson_q1"])
def test_linear_solve(name, prec, matrix_free):
    """Jacobi exercises the scalar-diagonal kernel (matrix free) / the CSR diagonal (matrix based)."""
    import dune_copasi_b200 as D
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    case, om, cfg, model, grid, op = make(name)
    x = K.rand_state(om.ndofs, 7)
    t, wM, wA = case.t0, (0.0 if name == "poisson_q1" else 1.0), 0.5 * case.dt
    lcfg = D.Config(f"type = BiCGSTAB\npreconditioner.type = {prec}\nmatrix_free = {'true' if matrix_free else 'false'}\n"
                    "convergence_condition.iteration_range = 1 2000\n")
    solver = D.Solver(op, lcfg)
    solver.linearize(t, wM, wA, x)
    b = K.rand_state(om.ndofs, 8, -1.0, 1.0)
    cd, _ = om.constraints()
    b[cd] = 0.0
    z, res = solver.solve(b, 1e-12)
    assert res.converged, (name, prec, matrix_free, res.reduction, res.iterations)
    S = K.ORC.StepOperator(om)
    vals = S._stage_jacobian(x, t, wM, wA)
    A = sp.csr_matrix((vals, S.colidx, S.rowptr), shape=(om.ndofs, om.ndofs))
    assert rel(z, spl.spsolve(A.tocsc(), b)) <= 1e-8
    v = K.rand_state(om.ndofs, 9, -1.0, 1.0)
    assert rel(solver.apply_operator(v), A @ v) <= OP_TOL
    # same Krylov recurrence as the oracle's dune-istl restatement on the Q1 matrix
    zo, ro = K.ORC.linear_solve(S.rowptr, S.colidx, vals, b,
                                {"type": "BiCGSTAB", "preconditioner": {"type": prec, "block_size": int(om.comp_nspec[0])},
                                 "convergence_condition": {"iteration_range": "1 2000"}}, 1e-12)
    assert ro.converged and abs(res.half_iterations - ro.iterations_x2) <= max(2, ro.iterations_x2 // 10)


STEP_CASES = [("gauss2d_q1", "Alexander2", 2), ("gauss3d_q1", "ImplicitEuler", 2), ("poisson_q1", "ImplicitEuler", 1),
              ("grayscott2d_q1", "Alexander2", 3), ("grayscott3d_q1", "Alexander2", 2),
              ("grayscott3d_q1", "ImplicitEuler", 2), ("mitchell_schaefer_q1", "Alexander2", 3),
              ("grayscott2d_q1", "RungeKutta4", 2)]


@pytest.mark.parametrize("matrix_free", [False, True])
@pytest.mark.parametrize("name,rk,nsteps", STEP_CASES)
def test_time_steps_match_oracle(name, rk, nsteps, matrix_free):
    import dune_copasi_b200 as D
    over = {"model.time_step_operator.type": rk,
            "model.time_step_operator.linear_solver.matrix_free": "true" if matrix_free else "false"}
    case, om, cfg, model, grid, op = make(name, **over)
    S = K.ORC.StepOperator(om)
    u = om.initial(case.t0)
    st = D.Stepper(op, cfg)
    st.set_state(grid.interpolate(model, case.t0), case.t0)
    t = case.t0
    for _ in range(nsteps):
        u, ok = S.apply(u, t, case.dt)
        assert ok
        assert st.step(case.dt)
        t += case.dt
    got, tg = st.get_state()
    assert abs(tg - t) < 1e-12
    assert rel(got, u) <= FIELD_TOL, (name, rk, matrix_free, rel(got, u))
    assert st.stats()["kernel_launches"] > 0


def test_gauss_kat_on_cubes():
    """The reference's gauss assertion (test/gauss.ini:53-55: L2 error <= 0.5 at t = 1.2, maximum
    below the initial peak) holds for the Q1 discretisation of the same problem."""
    import dune_copasi_b200 as D
    over = {"model.time_step_operator.type": "Alexander2"}
    case = K.Case("gauss_q1_kat", K.GAUSS, 2, None, t0=1.0, structured=([32, 32], [-1, -1], [2, 2]), element="cube")
    cfg, model, grid = K.product_objects_structured(case, **over)
    op = D.Operator(model, grid)
    st = D.Stepper(op, cfg)
    st.set_state(grid.interpolate(model, 1.0), 1.0)
    for _ in range(4):
        assert st.step(0.05)
    u, t = st.get_state()
    assert abs(t - 1.2) < 1e-12
    X = grid.coords()
    Dc = 0.005
    exact = np.exp(-(X ** 2).sum(axis=1) / (4 * t * Dc)) / (4 * np.pi * t * Dc)
    h2 = (2.0 / 32) ** 2
    assert np.sqrt(h2 * ((u - exact) ** 2).sum()) <= 0.5          # nodal (lumped) L2 norm
    assert u.max() <= 1.0 / (4 * np.pi * Dc) and u.min() >= -1e-2


@pytest.mark.parametrize("name", ["gauss2d_q1", "gauss3d_q1", "poisson_q1", "mitchell_schaefer_q1"])
def test_reduce_matches_oracle(name):
    """[model.reduce] functionals on cube cells (3-point Gauss rule per axis, kernels/reduce.cuh
    dc_reduce_q1_kernel) against the oracle's Q1 branch of reduce (oracle/core.py)."""
    import dune_copasi_b200 as D
    case, om, cfg, model, grid, op = make(name)
    red = D.Reducer(op, cfg)
    assert red.keys
    for seed, t in ((None, case.t0), (41, case.t0 + 0.37)):
        x = om.initial(t) if seed is None else K.rand_state(om.ndofs, seed, -0.5, 1.5)
        ref, ref_status = K.ORC.reduce(om, x, t)
        got = red.apply(t, x, raise_on_error=False)
        assert list(got) == list(ref)
        for key in ref:
            scale = max(abs(ref[key]), 1e-300)
            # (+ 1e-25: the gradient energy of a constant field is rounding noise on cubes, exactly 0 on simplices)
            assert abs(got[key] - ref[key]) <= 1e-11 * scale + 1e-25, (key, got[key], ref[key])
            assert red.status[key] == ref_status[key], (key, red.status[key], ref_status[key])


def test_unsupported_configurations_fail_loudly():
    import dune_copasi_b200 as D
    case = K.Q1_CASES["grayscott2d_q1"]
    for over, what in (({"model.assembly.b200.scheme": "atomic"}, "scheme"),
                       ({"model.jacobian.type": "numerical"}, "numerical")):
        cfg, model, grid = K.product_objects(case, **over)
        with pytest.raises(D.DcbError):
            D.Operator(model, grid)
    # transmission / outflow terms need facets, which cube grids do not carry
    cell = K.Case("cell_q1", K.CELL, 3, None, structured=([4, 4, 4], [0, 0, 0], [1, 1, 1]), element="cube")
    with pytest.raises(D.DcbError):
        K.product_objects_structured(cell)


@pytest.mark.parametrize("n", [256])
def test_fullsize_properties(n):
    """BASELINE configs[3] at full size (256^3 Q1 cells, 33.9 M DOFs): the oracle cannot run here,
    so size-independent properties: the apply is linear and is the derivative of the residual, the
    diffusion operator is symmetric with zero row sums, the mass operator sums to the volume."""
    import dune_copasi_b200 as D
    case = K.Case("gs_q1_full", K.GRAY_SCOTT, 3, None, dt=1.0, structured=([n] * 3, [0, 0, 0], [1, 1, 1]), element="cube")
    cfg, model, grid = K.product_objects_structured(case)
    op = D.Operator(model, grid)
    nd = grid.ndofs
    assert nd == 2 * (n + 1) ** 3
    rng = np.random.default_rng(0)
    x = rng.uniform(0.1, 1.0, nd)
    z = rng.uniform(-1.0, 1.0, nd)
    w = rng.uniform(-1.0, 1.0, nd)
    t, wM, wA = 0.0, 1.0, 0.5
    Jz = op.jacobian_apply(t, wM, wA, x, z)
    Jw = op.jacobian_apply(t, wM, wA, x, w)
    Jzw = op.jacobian_apply(t, wM, wA, x, 2.0 * z - 3.0 * w)
    assert rel(Jzw, 2.0 * Jz - 3.0 * Jw) <= 1e-12
    eps = 1e-6
    r0 = op.residual(t, wM, wA, x)
    r1 = op.residual(t, wM, wA, x + eps * z)
    assert rel((r1 - r0) / eps, Jz) <= 1e-5
    # mass form: 1^T M z = integral of the interpolant of z; with z = 1: the volume per species
    ones = np.ones(nd)
    M1 = op.jacobian_apply(t, 1.0, 0.0, x, ones)
    assert abs(M1.sum() - 2.0) <= 1e-9
    # reaction-free part of the stiffness form is symmetric with zero row sums: use x = 0, where the
    # Gray-Scott reaction Jacobian is diagonal (-F, -(F+k)) times the mass matrix
    x0 = np.zeros(nd)
    Kz = op.jacobian_apply(t, 0.0, 1.0, x0, z)
    Kw = op.jacobian_apply(t, 0.0, 1.0, x0, w)
    assert abs(w @ Kz - z @ Kw) <= 1e-10 * abs(w @ Kz)
    K1 = op.jacobian_apply(t, 0.0, 1.0, x0, ones)
    Mu = op.jacobian_apply(t, 1.0, 0.0, x0, ones)
    coef = np.tile([0.042, 0.042 + 0.061], nd // 2)
    assert rel(K1, coef * Mu) <= 1e-10


MARCH_CASES = ["gauss2d", "gauss3d", "poisson", "grayscott2d", "grayscott3d", "mitchell_schaefer"] + ALL


@pytest.mark.parametrize("march", [0, 1, 3, 8])
@pytest.mark.parametrize("name", MARCH_CASES)
def test_marching_kernels(name, march):
    """Residual / apply through the marching drivers (threads walk `march` cells up the last axis,
    warp shuffles along x) on lattices whose rows do not align with warps; march = 0 is the
    one-thread-per-cell driver.  struct_march_fill = 0 keeps the columns long on these small grids."""
    import dune_copasi_b200 as D
    case = K.ALL_CASES[name]
    over = {"model.assembly.b200.struct_march": march, "model.assembly.b200.struct_march_apply": march,
            "model.assembly.b200.struct_march_fill": 0}
    om = case.oracle(**over)
    cfg, model, grid = K.product_objects(case, **over)
    op = D.Operator(model, grid)
    x = K.rand_state(om.ndofs, 21)
    z = K.rand_state(om.ndofs, 22, -1.0, 1.0)
    t, wM, wA = case.t0 + 0.2, 0.9, 0.4 * case.dt
    ref = np.zeros(om.ndofs)
    om.residual(1, t, wM, x, ref)
    om.residual(0, t, wA, x, ref)
    assert rel(op.residual(t, wM, wA, x), ref) <= OP_TOL, (name, march)
    cd, _ = om.constraints()
    z2 = z.copy()
    z2[cd] = 0.0
    ref = np.zeros(om.ndofs)
    om.jacobian_apply(1, t, wM, x, z2, ref)
    om.jacobian_apply(0, t, wA, x, z2, ref)
    assert rel(op.jacobian_apply(t, wM, wA, x, z), ref) <= OP_TOL, (name, march)
