"""The reference's system tests run by the oracle FROM THE REFERENCE'S OWN FILES (test/gauss.ini,
exp.ini, poisson.ini with the arguments test/CMakeLists.txt:88-135 passes), assertions included: the
`[model.reduce]` sections of those files are evaluated by the oracle's restatement of reduce.hh and no
`error` expression may fire -- the same pass criterion as the reference's CTest (src/dune_copasi.cc
turns a firing error expression into a failed run).  Only the linear solver is pinned, to the
reference's iterative default preconditioner (SSOR) under BiCGSTAB, because its default solver UMFPack
is a direct method outside this path.  Runs where /root/reference is mounted (this container)."""
import os

import numpy as np
import pytest

import cases as K

ORC, INI, OMESH = K.ORC, K.INI, K.OMESH
REF = "/root/reference/test"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")


def run_reference_ini(name, dim, overrides=(), pin_solver=True):
    path = name if os.path.isabs(name) else os.path.join(REF, name)
    cfg = INI.parse_ini(open(path).read())
    for k, v in overrides:
        INI.set_key(cfg, k, v)
    g = INI.sub(cfg, "grid")
    # structured simplex grid + global refinement (grid/make_multi_domain_grid.hh:76-100): `cells` per
    # axis (default 1) doubled refinement_level times
    def vec(key, default):
        raw = str(g.get(key, "")).split()
        return [float(x) for x in raw][:dim] if raw else [default] * dim
    cells = [int(c) * 2 ** int(g.get("refinement_level", 0)) for c in vec("cells", 1)]
    mesh = OMESH.structured(dim, cells, vec("origin", 0.0), vec("extensions", 1.0))
    if pin_solver:
        INI.set_key(cfg, "model.time_step_operator.linear_solver.type", "BiCGSTAB")
        INI.set_key(cfg, "model.time_step_operator.linear_solver.preconditioner.type", "SSOR")
    om = ORC.Model(cfg, mesh)
    S = ORC.StepOperator(om)
    ts = INI.sub(INI.sub(cfg, "model"), "time_step_operator")
    t0, t_end = float(ts.get("time_begin", 0.0)), float(ts["time_end"])
    dt_max = float(ts["time_step_max"]) if "time_step_max" in ts else None
    dt0 = float(ts.get("time_step_initial", dt_max if dt_max is not None else 0.1))   # src/dune_copasi.cc:397
    u, t, n = ORC.evolve(S, om.initial(t0), t0, t_end, dt0, dt_max=dt_max)
    assert abs(t - t_end) <= 1e-12 * max(1.0, abs(t_end))
    values, status = ORC.reduce(om, u, t)
    return om, u, n, values, status


@pytest.mark.parametrize("dim", [2, 3])
def test_gauss_from_the_reference_file(dim):
    ext, org = " ".join(["2"] * dim), " ".join(["-1"] * dim)
    over = [("grid.dimension", str(dim)), ("grid.extensions", ext), ("grid.origin", org)]
    # (the file's own refinement_level = 5, i.e. 32 cells per axis: at 16^3 the narrow Gaussian undershoots
    # below the file's u_min error threshold, at 32^3 it only warns -- the level the reference chose)
    om, u, n, values, status = run_reference_ini("gauss.ini", dim, over)
    assert set(values) == {"u_max", "u_min", "u_error"}
    assert max(status.values()) < 2, (values, status)       # no `error` expression fired
    if dim == 2:
        assert values["u_error"] <= 0.5 and values["u_max"] <= 1 / (4 * 3.14159265359 * 0.005)


@pytest.mark.parametrize("dim", [2, 3])
def test_exp_from_the_reference_file(dim):
    om, u, n, values, status = run_reference_ini("exp.ini", dim, [("grid.dimension", str(dim))])
    # 99 full steps + snap_to_time, which splits a remainder a few ulp above dt in two (stepper.hh:213-218)
    assert n in (100, 101) and max(status.values()) < 2, (values, status)
    assert values["u_error"] <= 5e-3
    # u_mass_analytic has no integration_factor in the file: a plain sum over the quadrature points
    assert values["u_mass_analytic"] == pytest.approx(np.exp(-20.0) * (values["u_mass_analytic"] / np.exp(-20.0)).round())


@pytest.mark.parametrize("dim", [2, 3])
def test_poisson_from_the_reference_file(dim):
    over = [("grid.dimension", str(dim)), ("parser_context.dim.value", str(dim))]
    om, u, n, values, status = run_reference_ini("poisson.ini", dim, over)
    assert max(status.values()) < 2, (values, status)
    assert values["u_error"] <= 2.0


DOCS = "/root/reference/doc/docusaurus/static/ini/next"


def test_documented_gray_scott_with_its_own_solver_settings():
    """doc/docusaurus/static/ini/next/grey_scott.ini as written -- RestartedGMRes + Jacobi, Newton at 1e-8,
    128^2 lattice (refinement_level 7) -- for the first steps of its adaptive run: every setting of the
    file is inside the data-parallel path, nothing is pinned."""
    om, u, n, values, status = run_reference_ini(os.path.join(DOCS, "grey_scott.ini"), 2,
                                                 [("model.time_step_operator.time_end", "0.6")], pin_solver=False)
    assert n == 5 and om.ndofs == 2 * 129 * 129       # dt = 0.1, 0.11, 0.121, 0.1331, then the snap to 0.6
    U, V = u[0::2], u[1::2]
    assert np.isfinite(u).all() and V.min() > -1e-6 and 0.3 < V.max() < 0.7 and 0.5 < U.min() < 0.7
    # away from the bumps V ~ 0 and U follows U' = F (1 - U): U = 1 - 0.3 exp(-F t), F = 0.042
    assert U.max() == pytest.approx(1.0 - 0.3 * np.exp(-0.042 * 0.6), abs=2e-5)


def test_documented_lotka_volterra_and_heat_files_run():
    """volka_terra.ini (two triangles: an ODE system through the PDE machinery) and heat.ini, a few steps."""
    om, u, n, values, status = run_reference_ini(os.path.join(DOCS, "volka_terra.ini"), 2,
                                                 [("model.time_step_operator.time_end", "1.0")])
    assert n >= 4 and np.isfinite(u).all() and u.min() > 0
    # identical states at the four vertices and no diffusion: the fields stay spatially constant up to the
    # file's (default) solver tolerances -- Newton stops at 1e-8 on the *squared* defect norm
    assert np.ptp(u[0::2]) <= 1e-3 * abs(u[0]) and np.ptp(u[1::2]) <= 1e-3 * abs(u[1])
    om, u, n, values, status = run_reference_ini(os.path.join(DOCS, "heat.ini"), 2,
                                                 [("model.time_step_operator.time_end", "0.3"), ("grid.refinement_level", "2")])
    assert n >= 1 and np.isfinite(u).all()
