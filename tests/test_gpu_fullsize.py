"""Parity at BASELINE.json's full size (256^3 lattice, 33.9 M DOFs, 100.7 M tetrahedra), where the
oracle cannot run: size-independent properties of the operators, and agreement of two independent
kernel families (implicit-geometry vs element-per-thread) on the same inputs.  Through the C ABI."""
import numpy as np
import pytest

import cases as K

pytestmark = pytest.mark.gpu
N = 256
TOL = 1e-12


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def build(case_name, scheme="auto", **over):
    import dune_copasi_b200 as D
    case = K.CASES[case_name]
    over = dict(over, **{"model.assembly.b200.scheme": scheme})
    cfg = D.Config(case.ini_with(**over))
    model = D.Model(cfg, 3, [])
    if case_name.startswith("gauss"):
        grid = D.Grid.structured(3, [N] * 3, [-1.0] * 3, [2.0] * 3)
    else:
        grid = D.Grid.structured(3, [N] * 3, [0.0] * 3, [1.0] * 3)
    grid.bind(model)
    return case, cfg, model, grid, D.Operator(model, grid)


@pytest.fixture(scope="module")
def grayscott():
    case, cfg, model, grid, op = build("grayscott3d")
    assert op.ndofs == 2 * (N + 1) ** 3
    rng = np.random.default_rng(0)
    u = rng.uniform(0.1, 1.0, op.ndofs)
    z = rng.uniform(-1.0, 1.0, op.ndofs)
    return case, cfg, model, grid, op, u, z


def test_fullsize_kernel_families_agree(grayscott):
    """Structured (one thread per lattice cell, cell-level diffusion) and element-per-thread kernels
    (explicit connectivity and coordinates, atomics) are separate code: same residual and apply."""
    case, cfg, model, grid, op, u, z = grayscott
    _, _, _, _, op2 = build("grayscott3d", scheme="atomic")
    t, wM, wA = 0.3, 1.0, 0.25
    assert rel(op.residual(t, wM, wA, u), op2.residual(t, wM, wA, u)) <= TOL
    assert rel(op.jacobian_apply(t, wM, wA, u, z), op2.jacobian_apply(t, wM, wA, u, z)) <= TOL


def test_fullsize_apply_is_linear_and_is_the_derivative_of_the_residual(grayscott):
    case, cfg, model, grid, op, u, z = grayscott
    t, wM, wA = 0.0, 1.0, 0.5
    z2 = np.random.default_rng(1).uniform(-1.0, 1.0, op.ndofs)
    jz, jz2 = op.jacobian_apply(t, wM, wA, u, z), op.jacobian_apply(t, wM, wA, u, z2)
    assert rel(op.jacobian_apply(t, wM, wA, u, 0.3 * z - 1.7 * z2), 0.3 * jz - 1.7 * jz2) <= TOL
    eps = 1e-6
    fd = (op.residual(t, wM, wA, u + eps * z) - op.residual(t, wM, wA, u - eps * z)) / (2 * eps)
    assert rel(jz, fd) <= 1e-8        # central difference of a cubic reaction: O(eps^2) + rounding / eps


def test_fullsize_diffusion_operator_is_symmetric_and_conservative():
    """Pure diffusion + mass (test/gauss.ini at 256^3): J = wM M + wA K with M, K symmetric, K 1 = 0."""
    case, cfg, model, grid, op = build("gauss3d")
    rng = np.random.default_rng(2)
    u = np.zeros(op.ndofs)
    z, w = rng.uniform(-1, 1, op.ndofs), rng.uniform(-1, 1, op.ndofs)
    jz, jw = op.jacobian_apply(1.0, 1.0, 0.1, u, z), op.jacobian_apply(1.0, 1.0, 0.1, u, w)
    assert abs(w @ jz - z @ jw) <= 1e-11 * abs(w @ jz)
    kz = op.jacobian_apply(1.0, 0.0, 1.0, u, z)                 # stiffness form only
    assert abs(kz.sum()) <= 1e-9 * np.abs(kz).sum()             # 1^T K z = 0
    one = np.ones(op.ndofs)
    mass = op.jacobian_apply(1.0, 1.0, 0.0, u, one)             # M 1 = lumped cell volumes
    assert abs(mass.sum() - 8.0) <= 1e-10 * 8.0                 # |[-1,1]^3| = 8


def test_fullsize_implicit_step_conserves_mass_and_keeps_the_known_answer():
    """One implicit Euler step of the heat kernel at 256^3 (natural boundary conditions): the total mass
    is conserved to the linear tolerance, and the reference's assertion (test/gauss.ini:43-55) holds."""
    import dune_copasi_b200 as D
    over = {"model.time_step_operator.type": "ImplicitEuler",
            "model.time_step_operator.linear_solver.matrix_free": "true",
            "model.reduce.u_mass.evaluation.expression": "u * integration_factor"}
    case, cfg, model, grid, op = build("gauss3d", **over)
    st = D.Stepper(op, cfg)
    st.set_state(grid.interpolate(model, 1.0), 1.0)
    red = D.Reducer(op, cfg)
    before = red.apply_dev(st.time, st.state_dev(), raise_on_error=False)   # u_max sits on its bound at t = 1
    assert st.step(0.1)
    after = red.apply_dev(st.time, st.state_dev())       # raises if an error.expression fires
    assert abs(after["u_mass"] - before["u_mass"]) <= 1e-9 * abs(before["u_mass"])
    # the ini's `gauss` is normalised for 2-D; in 3-D it integrates to sqrt(4 pi t D)
    assert abs(before["u_mass"] - np.sqrt(4 * np.pi * 0.005)) <= 2e-2 * np.sqrt(4 * np.pi * 0.005)
    assert after["u_error"] <= 0.50 and after["u_max"] <= before["u_max"]
