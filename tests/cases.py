"""Test problems shared by the CPU (host logic / oracle) and GPU (parity) suites.

Each case restates one of the reference's configurations (BASELINE.json `configs`, SURVEY.md
App. B) as ini text in the reference's own key vocabulary, with the solver pinned to the
data-parallel subset -- no file under /root/reference is read at run time.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import core as ORC  # noqa: E402
from oracle import ini as INI  # noqa: E402
from oracle import mesh as OMESH  # noqa: E402
from dune_copasi_b200.workloads import CELL, CELL10, CELL10_NESTED, GRAY_SCOTT, SOLVER  # noqa: E402  (the bench workloads live with the package)
from dune_copasi_b200 import meshgen as MESHGEN  # noqa: E402


# test/gauss.ini: single compartment, linear diffusion of a Gaussian, D = 0.005, t in [1, 1.2]
GAUSS = """
[compartments.domain]
type = expression
expression = 1
[parser_context.diffusion]
type = constant
value = 0.005
[parser_context.gauss]
type = function
expression = x, y, z, t: exp(-(x^2+y^2+z^2)/(4*t*diffusion)) / (4*3.14159265359*t*diffusion)
[model]
is_linear = true
[model.scalar_field.u]
compartment = domain
[model.scalar_field.u.cross_diffusion.u]
expression = diffusion
[model.scalar_field.u.initial]
expression = gauss(position_x, position_y, position_z, time)
[model.scalar_field.u.storage]
expression = 1
[model.time_step_operator]
time_begin = 1
time_end = 1.2
""" + SOLVER

# test/exp.ini: pure reaction u' = -2u through the non-linear (Newton) path
EXP = """
[compartments.domain]
type = expression
expression = 1
[parser_context.grow_rate]
type = constant
value = -2.
[model.scalar_field.u]
compartment = domain
initial.expression = 1
storage.expression = 1
reaction.expression = grow_rate*u
reaction.jacobian.u.expression = grow_rate
[model.time_step_operator]
time_step_max = 0.1
time_end = 10
""" + SOLVER

# tabulated parser_context functions (src/dune/copasi/parser/context.cc:72-97 `type = interpolation`,
# :237-283 `interpolate = true`): a saturating uptake whose rate and slope are tables, a tabulated source profile
TABLES = """
[compartments.domain]
type = expression
expression = 1
[parser_context.rate]
type = interpolation
domain = 0 0.25 0.5 1 2
range = 0 0.4 0.65 0.9 1
[parser_context.slope]
type = interpolation
domain = 0 0.25 0.2500001 0.5 0.5000001 1 1.0000001 2
range = 1.6 1.6 1 1 0.5 0.5 0.1 0.1
[parser_context.profile]
type = function
expression = x: exp(-8*(x-0.5)^2)
interpolate = true
interpolation.intervals = 64
interpolation.domain.x = 0 1
interpolation.out_of_bounds = clamp
[parser_context.bump]
type = function
expression = s: 0.5 + 0.25*sin(6*s)
interpolate = true
interpolation.intervals = 20
interpolation.domain.s = -1 3
[model.scalar_field.u]
compartment = domain
initial.expression = 0.2 + 0.6*position_x
storage.expression = 1
cross_diffusion.u.expression = 0.01*bump(1)
reaction.expression = profile(position_x) - rate(u)*bump(u)
reaction.jacobian.u.expression = -slope(u)*bump(u)
[model.time_step_operator]
time_step_max = 0.1
time_end = 10
""" + SOLVER

# test/poisson.ini: -lap u = -2 dim with Dirichlet data |x|^2 (constraints)
POISSON = """
[parser_context]
dim.type = constant
dim.value = 2
[compartments.domain]
type = expression
expression = 1
[model]
is_linear = true
parser_type = ExprTk
[model.time_step_operator]
time_end = 0.1
[model.scalar_field.u]
compartment = domain
cross_diffusion.u.expression = 1
reaction.expression = -2*dim
constrain.boundary.expression = in_boundary ? position_x^2+position_y^2+position_z^2 : no_value
initial.expression = 0
""" + SOLVER


# test/mitchell_schaefer.ini with rng == 0 as the reference's CTest does (test/CMakeLists.txt:78-79)
MITCHELL_SCHAEFER = """
[parser_context]
tau_in.type = constant
tau_in.value = 0.1
tau_out.type = constant
tau_out.value = 1
tau_open.type = constant
tau_open.value = 80
tau_close.type = constant
tau_close.value = 60
D.type = constant
D.value = 1e-4
A_m.type = constant
A_m.value = 20e-2
u_0.type = constant
u_0.value = 0.13
gauss.type = function
gauss.expression = x, y, z: exp(-(x^2+y^2+z^2)/(4*0.0001)) / (4*3.14159265359*0.0001)/400
pulse.type = function
pulse.expression = t, t0, dt: sqrt((t-t0)^2)< dt ? 1 : 0
periodic.type = function
periodic.expression = t : cos(t) > 0.95 ? 1 : 0
rng.type = function
rng.expression = x, y: 0
[compartments]
domain.type = expression
domain.expression = 1
[model]
parser_type = ExprTk
[model.scalar_field.u]
compartment = domain
storage.expression = 1
reaction.expression = A_m*(z*u^2*(1-u)/tau_in - u/tau_out) + periodic( 2 * 3.14 * time/200)*gauss(position_x - 0.05, position_y - 0.05, position_z) + 0.85*pulse(time, 470, 20)*gauss(position_x - 0.675, position_y - 0.15, position_z)
reaction.jacobian.u.expression = A_m*(u*z*(2-3*u)/tau_in-1/tau_out)
reaction.jacobian.z.expression = A_m*(1-u)*u^2/tau_in
cross_diffusion.u.expression = D - 0.15 * D * rng(position_x,position_y)
initial.expression = 0.01
[model.scalar_field.z]
compartment = domain
storage.expression = 1
reaction.expression = (u <= u_0) ? ( (1 - z)/(tau_open*(1+0.25*rng(position_x,position_y))) ): (-z/tau_close)
reaction.jacobian.z.expression = (u <= u_0) ? ( -1/(tau_open*(1+0.25*rng(position_x,position_y))) ) : (-1/tau_close)
initial.expression = 1
[model.time_step_operator]
time_end = 5000
time_step_max = 10
""" + SOLVER

# test/two_disks.ini: two compartments, interface flux phi (u_in - u_out), Dirichlet on the rim
TWO_DISKS = """
[compartments]
outer.expression = (gmsh_id == 1)
inner.expression = (gmsh_id == 2)
[parser_context]
phi.type = constant
phi.value = 1
[model.scalar_field.u_in]
compartment = inner
cross_diffusion.u_in.expression = 1
outflow.outer.expression = phi*(u_in - u_out)
outflow.outer.jacobian.u_in.expression = phi
outflow.outer.jacobian.u_out.expression = -phi
[model.scalar_field.u_out]
compartment = outer
cross_diffusion.u_out.expression = 1
storage.expression = 0
constrain.boundary.expression = position_x
outflow.inner.expression = phi*(u_out - u_in)
outflow.inner.jacobian.u_out.expression = phi
outflow.inner.jacobian.u_in.expression = -phi
[model]
is_linear = true
parser_type = ExprTk
[model.time_step_operator]
time_step_max = 1
time_end = 1
""" + SOLVER

# test/two_disks_cell_data.ini: the two-disks problem twice -- (u_in, u_out) with a diffusion coefficient
# read from per-cell grid data (`sigma`, "sigma := position_x" in the reference's comment; its data
# file is a git-LFS pointer, so sigma is generated here as the cell centre's x), (v_in, v_out) with the
# analytic 1 + position_x^2 -- and a reduce functional comparing the two.
TWO_DISKS_CELL_DATA = """
[parser_context]
phi.type = constant
phi.value = 1
[compartments]
outer.expression = (gmsh_id == 1)
inner.expression = (gmsh_id == 2)
[model.scalar_field.u_in]
compartment = inner
cross_diffusion.u_in.expression = (gmsh_id == 2) ? 1 : 0
outflow.outer.expression = phi*(u_in - u_out)
outflow.outer.jacobian.u_in.expression = phi
outflow.outer.jacobian.u_out.expression = -phi
[model.scalar_field.u_out]
compartment = outer
cross_diffusion.u_out.expression = 1 + sigma^2
storage.expression = 0
constrain.boundary.expression = position_x
outflow.inner.expression = phi*(u_out - u_in)
outflow.inner.jacobian.u_out.expression = phi
outflow.inner.jacobian.u_in.expression = -phi
[model.scalar_field.v_in]
compartment = inner
cross_diffusion.v_in.expression = 1
outflow.outer.expression = phi*(v_in - v_out)
outflow.outer.jacobian.v_in.expression = phi
outflow.outer.jacobian.v_out.expression = -phi
[model.scalar_field.v_out]
compartment = outer
cross_diffusion.v_out.expression = 1 + position_x^2
storage.expression = 0
constrain.boundary.expression = position_x
outflow.inner.expression = phi*(v_out - v_in)
outflow.inner.jacobian.v_out.expression = phi
outflow.inner.jacobian.v_in.expression = -phi
[model]
is_linear = true
parser_type = ExprTk
[model.time_step_operator]
time_step_max = 1
time_end = 1
[model.reduce]
u_error.evaluation.expression = ((u_in - v_in)^2 + (u_out - v_out)^2) * integration_factor
u_error.transformation.expression = arg: sqrt(arg)
""" + SOLVER


def two_disks_with_sigma():
    m = OMESH.two_disks(6, 6, 32)
    sigma = m.coords[m.elems][:, :, 0].mean(axis=1)
    m.cell_keys = ["gmsh_id", "sigma"]
    m.cell_data = np.ascontiguousarray(np.stack([m.cell_data[0], sigma]))
    return m




assert CELL10 != CELL


# Synthetic: every volume term of local_operator.hh:417-707 the inis above do not reach -- advection
# (velocity.{x,y,z} with a jacobian), tensor diffusion, solution dependent diffusion with its
# jacobian entries (scalar and tensor), a storage coefficient with a jacobian.
ADVECTION = """
[compartments.domain]
type = expression
expression = 1
[model.scalar_field.u]
compartment = domain
initial.expression = 0.5 + 0.3*sin(3*position_x)*cos(2*position_y)
storage.expression = 1 + 0.1*v
storage.jacobian.v.expression = 0.1
reaction.expression = -u*v + 0.1
reaction.jacobian.u.expression = -v
reaction.jacobian.v.expression = -u
velocity.x.expression = 0.3*v
velocity.y.expression = -0.2 + 0.1*position_x
velocity.jacobian.v.x.expression = 0.3
[model.scalar_field.u.cross_diffusion.u]
type = tensor
xx.expression = 0.01*(1 + u)
xy.expression = 0.002
yx.expression = 0.001*u
yy.expression = 0.02
zz.expression = 0.015
jacobian.u.type = tensor
jacobian.u.xx.expression = 0.01
jacobian.u.yx.expression = 0.001
[model.scalar_field.u.cross_diffusion.v]
expression = 0.003
[model.scalar_field.v]
compartment = domain
initial.expression = 0.4 + 0.2*cos(2*position_x + position_y)
storage.expression = 1
reaction.expression = u*v - 0.3*v
reaction.jacobian.u.expression = v
reaction.jacobian.v.expression = u - 0.3
velocity.x.expression = 0.1
velocity.z.expression = 0.05*u
velocity.jacobian.u.z.expression = 0.05
cross_diffusion.v.expression = 0.01*(1 + u^2)
cross_diffusion.v.jacobian.u.expression = 0.02*u
[model.time_step_operator]
time_end = 1
""" + SOLVER


# [model.reduce] sections: the assertions of the reference's system tests in its own vocabulary
# (test/gauss.ini:38-55, exp.ini:25-35, poisson.ini:27-32, two_disks.ini:45-60, mitchell_schaefer.ini:84-104)
REDUCE = {
    "gauss": """
[model.reduce]
u_max.evaluation.expression = u
u_max.reduction.expression = init, val: max(init, val)
u_max.initial.value = -1e100
u_max.error.expression = arg: arg > 1/(4*3.14159265359*diffusion)
u_min.evaluation.expression = u
u_min.reduction.expression = init, val: min(init, val)
u_min.initial.value = 1e100
u_min.warn.expression = arg: arg < 0
u_min.error.expression = arg: arg < -1e-2
u_error.evaluation.expression = (u - gauss(position_x, position_y, position_z, time))^2 * integration_factor
u_error.transformation.expression = arg: sqrt(arg)
u_error.error.expression = arg: arg > 0.50
""",
    "exp": """
[model.reduce]
u_mass_analytic.evaluation.expression = exp(grow_rate*time)
u_max.evaluation.expression = u
u_max.reduction.expression = init, val: max(init, val)
u_max.initial.value = -1e100
u_error.evaluation.expression = (u - exp(grow_rate*time))^2 * integration_factor
u_error.transformation.expression = arg: sqrt(arg)
u_error.error.expression = arg: arg > 5e-3
""",
    "poisson": """
[model.reduce]
u_error.evaluation.expression = (u - (position_x^2+position_y^2+position_z^2))^2 * integration_factor
u_error.transformation.expression = arg: sqrt(arg)
u_error.warn.expression = arg: arg > 1e-2
u_error.error.expression = arg: arg > 2e-0
""",
    "two_disks": """
[parser_context.u_in_analytic]
type = function
expression = x, y: 8*phi*x/(8*phi + 5)
[parser_context.u_out_analytic]
type = function
expression = x, y: 4*((2*phi + 1) + 1/(x^2 + y^2))*x/(8*phi + 5)
[model.reduce]
u_max.evaluation.expression = max(u_in, u_out)
u_max.reduction.expression = init, val: max(init, val)
u_max.initial.value = -1e100
u_max.error.expression = arg: arg > 2
u_min.evaluation.expression = min(u_in, u_out)
u_min.reduction.expression = init, val: min(init, val)
u_min.initial.value = 1e100
u_min.error.expression = arg: arg < -2
u_error.evaluation.expression = ((sqrt(position_x^2+position_y^2) < 1) ? (u_in - u_in_analytic(position_x, position_y))^2 : (u_out - u_out_analytic(position_x, position_y))^2) * integration_factor
u_error.transformation.expression = arg: sqrt(arg)
u_error.warn.expression = arg: arg > 1e-3
""",
    "mitchell_schaefer": """
[model.reduce]
u_max.evaluation.expression = u
u_max.reduction.expression = init, val: max(init, val)
u_max.initial.value = -1e100
u_max.warn.expression = arg: arg > 1
u_min.evaluation.expression = u
u_min.reduction.expression = init, val: min(init, val)
u_min.initial.value = 1e100
u_min.warn.expression = arg: arg < 0
z_max.evaluation.expression = z
z_max.reduction.expression = init, val: max(init, val)
z_max.initial.value = -1e100
z_max.warn.expression = arg: arg > 1
u_mass.evaluation.expression = u * integration_factor
grad_energy.evaluation.expression = (grad_u_x^2 + grad_u_y^2) * integration_factor
""",
    "cell3d": """
[model.reduce]
c1_mass.evaluation.expression = c1 * integration_factor
n1_max.evaluation.expression = n1
n1_max.reduction.expression = a, b: max(a, b)
n1_max.initial.value = -1e100
volume.evaluation.expression = integration_factor
cells.evaluation.expression = integration_factor / entity_volume
""",
}


class Case:
    def __init__(self, name, ini, dim, mesh_fn, t0=0.0, dt=0.1, structured=None, element="simplex"):
        self.name, self.ini, self.dim, self.mesh_fn, self.t0, self.dt = name, ini, dim, mesh_fn, t0, dt
        self.structured = structured   # (cells, origin, extent) when the mesh is a structured grid
        self.element = element         # "cube": the lattice cells as Q1 elements (BASELINE configs[3])

    def oracle(self, **overrides):
        cfg = INI.parse_ini(self.ini)
        for k, v in overrides.items():
            INI.set_key(cfg, k, str(v))
        mesh = self.mesh_fn()
        return ORC.Model(cfg, mesh)

    def ini_with(self, **overrides):
        cfg = INI.parse_ini(self.ini)
        for k, v in overrides.items():
            INI.set_key(cfg, k, str(v))
        return INI.to_text(cfg)


def _s(dim, n, origin=None, extent=None, element="simplex"):
    cells = [n] * dim if np.isscalar(n) else list(n)
    return lambda: OMESH.structured(dim, cells, origin, extent, element)


CASES = {
    "gauss2d": Case("gauss2d", GAUSS + REDUCE["gauss"], 2, _s(2, 32, [-1, -1], [2, 2]), t0=1.0, structured=([32, 32], [-1, -1], [2, 2])),
    "gauss3d": Case("gauss3d", GAUSS + REDUCE["gauss"], 3, _s(3, 8, [-1, -1, -1], [2, 2, 2]), t0=1.0, structured=([8, 8, 8], [-1, -1, -1], [2, 2, 2])),
    "exp": Case("exp", EXP + REDUCE["exp"], 2, _s(2, 2), structured=([2, 2], [0, 0], [1, 1])),
    "tables": Case("tables", TABLES, 2, _s(2, 6), dt=0.05, structured=([6, 6], [0, 0], [1, 1])),
    # constrain.skeleton binds vertices off the boundary (constraints.hh:128-156), constrain.volume no P1 dof at all (:93-112)
    "poisson_pinned": Case("poisson_pinned", POISSON.replace(
        "initial.expression = 0", "initial.expression = 0\nconstrain.skeleton.expression = "
        "(abs(position_x - 0.5) < 1e-9 and in_skeleton) ? 0.1 + position_y : no_value\nconstrain.volume.expression = 7"),
        2, _s(2, 16), structured=([16, 16], [0, 0], [1, 1])),
    "poisson": Case("poisson", POISSON + REDUCE["poisson"], 2, _s(2, 16), structured=([16, 16], [0, 0], [1, 1])),
    "grayscott2d": Case("grayscott2d", GRAY_SCOTT, 2, _s(2, 32), dt=1.0, structured=([32, 32], [0, 0], [1, 1])),
    "grayscott3d": Case("grayscott3d", GRAY_SCOTT, 3, _s(3, 10), dt=1.0, structured=([10, 10, 10], [0, 0, 0], [1, 1, 1])),
    "mitchell_schaefer": Case("mitchell_schaefer", MITCHELL_SCHAEFER + REDUCE["mitchell_schaefer"], 2, _s(2, 16), dt=0.01, structured=([16, 16], [0, 0], [1, 1])),
    "two_disks": Case("two_disks", TWO_DISKS + REDUCE["two_disks"], 2, lambda: OMESH.two_disks(6, 6, 32), dt=1.0),
    "two_disks_cell_data": Case("two_disks_cell_data", TWO_DISKS_CELL_DATA, 2, two_disks_with_sigma, dt=1.0),
    "cell3d": Case("cell3d", CELL + REDUCE["cell3d"], 3, _s(3, 8), dt=0.05, structured=([8, 8, 8], [0, 0, 0], [1, 1, 1])),
    "cell3d_10": Case("cell3d_10", CELL10, 3, _s(3, 6), dt=0.05, structured=([6, 6, 6], [0, 0, 0], [1, 1, 1])),
    "advection2d": Case("advection2d", ADVECTION, 2, _s(2, 12), dt=0.05, structured=([12, 12], [0, 0], [1, 1])),
    "advection3d": Case("advection3d", ADVECTION, 3, _s(3, 5), dt=0.05, structured=([5, 5, 5], [0, 0, 0], [1, 1, 1])),
}


def nested_mesh(n=8, order="morton"):
    """BASELINE configs[4] in miniature: the unstructured tetrahedral mesh of three nested compartments"""
    coords, elems, keys, data = MESHGEN.nested_compartments(n, order=order)
    m = OMESH.Mesh(dim=3, coords=coords, elems=elems)
    m.cell_keys, m.cell_data = keys, data
    return m


CASES["cell10_nested"] = Case("cell10_nested", CELL10_NESTED, 3, nested_mesh, dt=0.05)


def _p1(name, ini, dim, cells, origin, extent, **kw):
    return Case(name, ini, dim, _s(dim, cells, origin, extent), structured=(cells, origin, extent), **kw)


# P1 lattices that a transposed index decode or swapped per-axis weights cannot survive: different
# cell counts and different mesh widths per axis; the "wide" ones span several tiles of the
# tile-marching kernels along x
ANISO_CASES = {
    "grayscott3d_aniso": _p1("grayscott3d_aniso", GRAY_SCOTT, 3, [10, 9, 8], [0, 0, 0], [1, 0.9, 0.8], dt=1.0),
    "gauss3d_aniso": _p1("gauss3d_aniso", GAUSS + REDUCE["gauss"], 3, [8, 6, 7], [-1, -1, -1], [2, 1.8, 2.2], t0=1.0),
    "grayscott2d_aniso": _p1("grayscott2d_aniso", GRAY_SCOTT, 2, [24, 20], [0, 0], [1, 0.8], dt=1.0),
    "poisson_aniso": _p1("poisson_aniso", POISSON + REDUCE["poisson"], 2, [14, 18], [0, 0], [1, 1.2]),
    "grayscott3d_wide": _p1("grayscott3d_wide", GRAY_SCOTT, 3, [40, 5, 7], [0, 0, 0], [2, 0.3, 0.4], dt=1.0),
    "grayscott2d_wide": _p1("grayscott2d_wide", GRAY_SCOTT, 2, [70, 9], [0, 0], [2, 0.3], dt=1.0),
    "advection3d_aniso": _p1("advection3d_aniso", ADVECTION, 3, [5, 4, 6], [0, 0, 0], [1, 0.7, 1.3], dt=0.05),
}
CASES.update(ANISO_CASES)


def _q1(name, ini, dim, cells, origin, extent, **kw):
    return Case(name, ini, dim, _s(dim, cells, origin, extent, "cube"), structured=(cells, origin, extent),
                element="cube", **kw)


# The same models on the lattice cells as Q1 elements (BASELINE configs[3] "Q1 on a structured grid").
# Not a reference element type: parity is product vs the oracle's own Q1 element.
Q1_CASES = {
    "gauss2d_q1": _q1("gauss2d_q1", GAUSS + REDUCE["gauss"], 2, [24, 20], [-1, -1], [2, 2], t0=1.0),
    "gauss3d_q1": _q1("gauss3d_q1", GAUSS + REDUCE["gauss"], 3, [8, 6, 7], [-1, -1, -1], [2, 2, 2], t0=1.0),
    "poisson_q1": _q1("poisson_q1", POISSON + REDUCE["poisson"], 2, [16, 16], [0, 0], [1, 1]),
    "grayscott2d_q1": _q1("grayscott2d_q1", GRAY_SCOTT, 2, [32, 32], [0, 0], [1, 1], dt=1.0),
    "grayscott3d_q1": _q1("grayscott3d_q1", GRAY_SCOTT, 3, [10, 9, 8], [0, 0, 0], [1, 0.9, 0.8], dt=1.0),
    "mitchell_schaefer_q1": _q1("mitchell_schaefer_q1", MITCHELL_SCHAEFER + REDUCE["mitchell_schaefer"], 2, [16, 16], [0, 0], [1, 1], dt=0.01),
}
ALL_CASES = {**CASES, **Q1_CASES}


def product_objects(case: Case, **overrides):
    """-> (Config, Model, Grid bound) through the C ABI, fed with the oracle-side mesh arrays."""
    import dune_copasi_b200 as D
    mesh = case.mesh_fn()
    cfg = D.Config(case.ini_with(**overrides))
    model = D.Model(cfg, case.dim, mesh.cell_keys)
    grid = product_grid(case, mesh)
    grid.bind(model)
    return cfg, model, grid


def product_objects_structured(case: Case, **overrides):
    """Like product_objects for a structured case, without building the oracle-side mesh (full-size runs)."""
    import dune_copasi_b200 as D
    cfg = D.Config(case.ini_with(**overrides))
    model = D.Model(cfg, case.dim, [])
    grid = D.Grid.structured(case.dim, *case.structured, element=case.element)
    grid.bind(model)
    return cfg, model, grid


def product_grid(case: Case, mesh=None):
    """Structured cases go through the product's own generator (bit-identical arrays, see
    tests/test_host_parity.py) so that the implicit-geometry kernels are eligible."""
    import dune_copasi_b200 as D
    if case.structured:
        return D.Grid.structured(case.dim, *case.structured, element=case.element)
    mesh = mesh or case.mesh_fn()
    return D.Grid.from_arrays(case.dim, mesh.coords, mesh.elems, mesh.cell_keys, mesh.cell_data)


def rand_state(n, seed=0, lo=0.1, hi=1.0):
    return np.random.default_rng(seed).uniform(lo, hi, n)
