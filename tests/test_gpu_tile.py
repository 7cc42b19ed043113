"""GPU parity of the tile-marching drivers (kernels/assembly_tile.cuh): owner-computes residual and
Jacobian apply with plain stores, slots + fix-up pass for the vertices on tile edges / chunk boundary
planes, and the BiCGSTAB whose vector updates and dot products are fused into the apply sweeps.

Every case runs with the production tile shape and with deliberately small tiles (2 warps of one
cell row / one warp of 3 rows, chunks of 3 / 2 layers) so that the small lattices of the suite have cut vertices along every
axis, corners shared by 8 contributors included.  Reference for the numbers: the CPU oracle;
reference for "nothing changed": the per-cell kernels (model.assembly.b200.tile = false) and the
unfused BiCGSTAB (linear_solver.b200.fused = false).
"""
import numpy as np
import pytest

import cases as K

pytestmark = pytest.mark.gpu

OP_TOL = 1e-12
FIELD_TOL = 1e-10
SMALL = {"model.assembly.b200.tile_w": 2, "model.assembly.b200.tile_r": 1, "model.assembly.b200.tile_lz": 3}
ODD = {"model.assembly.b200.tile_w": 1, "model.assembly.b200.tile_r": 3, "model.assembly.b200.tile_lz": 2}
SHAPES = {"default": {}, "small": SMALL, "odd": ODD}
for _s in SHAPES.values():
    _s["model.assembly.b200.tile"] = "true"     # the tile drivers are an option (off by default, see operator.cpp)
P1 = ["gauss2d", "gauss3d", "exp", "poisson", "grayscott2d", "grayscott3d", "mitchell_schaefer", "grayscott3d_aniso",
      "gauss3d_aniso", "grayscott2d_aniso", "poisson_aniso", "grayscott3d_wide", "grayscott2d_wide"]
Q1 = ["gauss2d_q1", "gauss3d_q1", "poisson_q1", "grayscott2d_q1", "grayscott3d_q1", "mitchell_schaefer_q1"]


def rel(a, b):
    d = np.linalg.norm(a - b)
    n = np.linalg.norm(b)
    return d / n if n > 0 else d


def make(name, **over):
    import dune_copasi_b200 as D
    case = K.ALL_CASES[name]
    om = case.oracle(**over)
    cfg, model, grid = K.product_objects(case, **over)
    return case, om, cfg, model, grid, D.Operator(model, grid)


@pytest.mark.parametrize("shape", list(SHAPES))
@pytest.mark.parametrize("name", P1 + Q1)
def test_tile_residual_and_apply(name, shape):
    case, om, cfg, model, grid, op = make(name, **SHAPES[shape])
    assert op.uses_tiles
    x = K.rand_state(om.ndofs, 41)
    z = K.rand_state(om.ndofs, 42, -1.0, 1.0)
    base = K.rand_state(om.ndofs, 43)
    t = case.t0 + 0.2
    for wM, wA in ((1.0, 0.0), (0.0, 1.0), (0.7, 0.3 * case.dt)):
        ref = base.copy()
        if wM:
            om.residual(1, t, wM, x, ref)
        if wA:
            om.residual(0, t, wA, x, ref)
        got = op.residual(t, wM, wA, x, base.copy())      # r += F(x)
        assert rel(got, ref) <= OP_TOL, (name, shape, "residual", wM, wA, rel(got, ref))
    wM, wA = 1.0, 0.25 * case.dt
    z2 = z.copy()
    cd, _ = om.constraints()
    if cd.size:
        z2[cd] = 0.0        # the product treats constrained entries of z as zero
    ref = base.copy()
    om.jacobian_apply(1, t, wM, x, z2, ref)
    om.jacobian_apply(0, t, wA, x, z2, ref)
    got = op.jacobian_apply(t, wM, wA, x, z, base.copy())  # y += J z
    assert rel(got, ref) <= OP_TOL, (name, shape, "apply", rel(got, ref))
    # run to run identical: no atomics anywhere in the sweep
    again = op.jacobian_apply(t, wM, wA, x, z, base.copy())
    assert np.array_equal(got, again)


@pytest.mark.parametrize("name", ["grayscott3d_aniso", "grayscott2d_wide", "poisson_aniso", "grayscott3d_q1"])
def test_tile_equals_per_cell_kernels(name):
    """Same cell integrals behind both drivers: the results differ by the summation order only."""
    case, om, cfg, model, grid, op = make(name, **SMALL)
    _, _, _, _, _, old = make(name)
    assert op.uses_tiles and not old.uses_tiles
    x = K.rand_state(om.ndofs, 44)
    z = K.rand_state(om.ndofs, 45, -1.0, 1.0)
    t, wM, wA = case.t0, 1.0, 0.5 * case.dt
    assert rel(op.residual(t, wM, wA, x), old.residual(t, wM, wA, x)) <= 1e-14
    assert rel(op.jacobian_apply(t, wM, wA, x, z), old.jacobian_apply(t, wM, wA, x, z)) <= 1e-14


@pytest.mark.parametrize("shape", list(SHAPES))
@pytest.mark.parametrize("name", ["grayscott3d", "grayscott3d_aniso", "grayscott3d_wide", "grayscott2d_wide", "gauss3d_aniso",
                                  "mitchell_schaefer", "grayscott3d_q1", "poisson_aniso"])
def test_fused_bicgstab_matches_oracle_and_unfused(name, shape):
    """BiCGSTAB + Jacobi, matrix free: the fused sweeps take the half iterations of the oracle's
    dune-istl restatement and of the unfused device loop, and reach the same solution.  (poisson has
    Dirichlet rows: the solver then keeps the unfused loop on top of the tile apply.)"""
    import dune_copasi_b200 as D
    case, om, cfg, model, grid, op = make(name, **SHAPES[shape])
    x = K.rand_state(om.ndofs, 46)
    b = K.rand_state(om.ndofs, 47, -1.0, 1.0)
    t, wM, wA = case.t0, (0.0 if name.startswith("poisson") else 1.0), 0.5 * case.dt
    cd, _ = om.constraints()
    b[cd] = 0.0
    text = "type = BiCGSTAB\npreconditioner.type = Jacobi\nmatrix_free = true\nconvergence_condition.iteration_range = 1 2000\n"
    out = {}
    for fused in ("true", "false"):
        sol = D.Solver(op, D.Config(text + f"b200.fused = {fused}\n"))
        assert sol.fused == (fused == "true" and cd.size == 0)
        sol.linearize(t, wM, wA, x)
        out[fused] = sol.solve(b, 1e-10)
    S = K.ORC.StepOperator(om)
    vals = S._stage_jacobian(x, t, wM, wA)
    zo, ro = K.ORC.linear_solve(S.rowptr, S.colidx, vals, b, {"type": "BiCGSTAB", "preconditioner": {"type": "Jacobi"}}, 1e-10)
    for fused, (zz, res) in out.items():
        assert res.converged and ro.converged
        # long, erratic BiCGSTAB runs: rounding (here the summation order of the apply) moves the
        # stopping iteration by a few half steps; the short runs must agree exactly
        slack = 0 if ro.iterations_x2 < 60 else ro.iterations_x2 // 10
        assert abs(res.half_iterations - ro.iterations_x2) <= slack, (name, shape, fused, res.half_iterations, ro.iterations_x2)
        assert rel(zz, zo) <= 1e-8, (name, shape, fused, rel(zz, zo))
    assert rel(out["true"][0], out["false"][0]) <= 1e-9


@pytest.mark.parametrize("shape", ["default", "small"])
@pytest.mark.parametrize("name,rk,nsteps", [("grayscott3d_aniso", "Alexander2", 2), ("grayscott2d_wide", "ImplicitEuler", 2),
                                            ("grayscott3d_wide", "Alexander2", 1), ("gauss3d_aniso", "Alexander2", 2),
                                            ("poisson_aniso", "ImplicitEuler", 1), ("grayscott3d_q1", "Alexander2", 2),
                                            ("mitchell_schaefer", "Alexander2", 3)])
def test_tile_time_steps_match_oracle(name, rk, nsteps, shape):
    import dune_copasi_b200 as D
    over = dict(SHAPES[shape], **{"model.time_step_operator.type": rk,
                                  "model.time_step_operator.linear_solver.matrix_free": "true"})
    case, om, cfg, model, grid, op = make(name, **over)
    S = K.ORC.StepOperator(om)
    u = om.initial(case.t0)
    st = D.Stepper(op, cfg)
    st.set_state(grid.interpolate(model, case.t0), case.t0)
    t = case.t0
    for _ in range(nsteps):
        u, ok = S.apply(u, t, case.dt)
        assert ok and st.step(case.dt)
        t += case.dt
    got, _ = st.get_state()
    assert rel(got, u) <= FIELD_TOL, (name, rk, shape, rel(got, u))
