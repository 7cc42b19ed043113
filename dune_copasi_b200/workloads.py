"""Workloads of the benchmark and the smoke test, in the reference's own ini vocabulary (the files under
doc/docusaurus/static/ini/next/ and test/ of the reference are the templates): the 3-D two-species
Gray-Scott system of BASELINE.json configs[3] and the three-compartment cell model of configs[4].
tests/cases.py builds its parity cases on the same texts."""
from __future__ import annotations

from .capi import Config

SOLVER = """
[model.time_step_operator.linear_solver]
type = BiCGSTAB
preconditioner.type = Jacobi
convergence_condition.relative_tolerance = 1e-12
[model.time_step_operator.nonlinear_solver]
convergence_condition.relative_tolerance = 1e-20
dx_inverse_fixed_tolerance = true
"""


# doc/docusaurus/static/ini/next/grey_scott.ini: 2 species, cubic reaction; bumps in 2-D / 3-D
GRAY_SCOTT = """
[parser_context.bump]
type = function
expression = x, y, z: 0.5*exp(-100*(x^2 + y^2 + z^2))
[parser_context.F]
type = constant
value = 0.0420
[parser_context.k]
type = constant
value = 0.0610
[parser_context.D]
type = constant
value = 1e-5
[model]
order = 1
parser_type = ExprTk
[model.time_step_operator]
time_begin = 0
time_end = 10000
time_step_initial = 0.1
time_step_max = 50
[compartments]
compartment.expression = 1
[model.scalar_field.U]
compartment = compartment
initial.expression = 0.7
storage.expression = 1
reaction.expression = F*(1-U) - U*V^2
reaction.jacobian.U.expression = -F - V^2
reaction.jacobian.V.expression = -2*U*V
cross_diffusion.U.expression = D*2
[model.scalar_field.V]
compartment = compartment
initial.expression = bump(0.25-position_x, 0.25-position_y, 0.25-position_z) + bump(0.25-position_x, 0.75-position_y, 0.75-position_z) + bump(0.75-position_x, 0.25-position_y, 0.75-position_z) + bump(0.75-position_x, 0.75-position_y, 0.25-position_z)
storage.expression = 1
reaction.expression = -(F+k)*V + U*V^2
reaction.jacobian.U.expression = V^2
reaction.jacobian.V.expression = -(F+k) + 2*U*V
cross_diffusion.V.expression = D
""" + SOLVER


# BASELINE config 5 in miniature: cytosol / nucleus / extracellular space as nested regions of a
# structured tet mesh, several species per compartment, constant cross-diffusion, non-linear
# reactions, membrane fluxes between touching compartments and an outflow boundary condition.
CELL = """
[compartments]
ecs.expression = (max(max(abs(position_x-0.5), abs(position_y-0.5)), abs(position_z-0.5)) > 0.375)
cytosol.expression = (max(max(abs(position_x-0.5), abs(position_y-0.5)), abs(position_z-0.5)) < 0.375) and (max(max(abs(position_x-0.5), abs(position_y-0.5)), abs(position_z-0.5)) > 0.125)
nucleus.expression = (max(max(abs(position_x-0.5), abs(position_y-0.5)), abs(position_z-0.5)) < 0.125)
[parser_context]
k1.type = constant
k1.value = 0.7
k2.type = constant
k2.value = 0.3
perm.type = constant
perm.value = 0.5
hill.type = function
hill.expression = s, K: s^2/(K^2 + s^2)
[model.scalar_field.e1]
compartment = ecs
storage.expression = 1
cross_diffusion.e1.expression = 0.02
reaction.expression = -k2*e1
reaction.jacobian.e1.expression = -k2
initial.expression = 1 + 0.5*position_x
outflow.cytosol.expression = perm*(e1 - c1)
outflow.cytosol.jacobian.e1.expression = perm
outflow.cytosol.jacobian.c1.expression = -perm
outflow.ecs.expression = 0.1*e1
outflow.ecs.jacobian.e1.expression = 0.1
[model.scalar_field.c1]
compartment = cytosol
storage.expression = 1
cross_diffusion.c1.expression = 0.01
cross_diffusion.c2.expression = 0.002
reaction.expression = -k1*c1*c2 + k2*c3
reaction.jacobian.c1.expression = -k1*c2
reaction.jacobian.c2.expression = -k1*c1
reaction.jacobian.c3.expression = k2
initial.expression = 0.2 + 0.1*position_y
outflow.ecs.expression = perm*(c1 - e1)
outflow.ecs.jacobian.c1.expression = perm
outflow.ecs.jacobian.e1.expression = -perm
outflow.nucleus.expression = perm*hill(c1, 0.5) - 0.2*n1
outflow.nucleus.jacobian.c1.expression = perm*2*c1*0.25/((0.25 + c1^2)^2)
outflow.nucleus.jacobian.n1.expression = -0.2
[model.scalar_field.c2]
compartment = cytosol
storage.expression = 1 + 0.5*position_z
cross_diffusion.c2.expression = 0.015
reaction.expression = -k1*c1*c2 + k2*c3
reaction.jacobian.c1.expression = -k1*c2
reaction.jacobian.c2.expression = -k1*c1
reaction.jacobian.c3.expression = k2
initial.expression = 0.5
[model.scalar_field.c3]
compartment = cytosol
storage.expression = 1
cross_diffusion.c3.expression = 0.005*(1 + position_x)
reaction.expression = k1*c1*c2 - k2*c3
reaction.jacobian.c1.expression = k1*c2
reaction.jacobian.c2.expression = k1*c1
reaction.jacobian.c3.expression = -k2
initial.expression = 0.1
[model.scalar_field.n1]
compartment = nucleus
storage.expression = 1
cross_diffusion.n1.expression = 0.01
reaction.expression = -0.05*n1*n2
reaction.jacobian.n1.expression = -0.05*n2
reaction.jacobian.n2.expression = -0.05*n1
initial.expression = 0.3
outflow.cytosol.expression = 0.2*n1 - perm*hill(c1, 0.5)
outflow.cytosol.jacobian.n1.expression = 0.2
outflow.cytosol.jacobian.c1.expression = -perm*2*c1*0.25/((0.25 + c1^2)^2)
[model.scalar_field.n2]
compartment = nucleus
storage.expression = 1
cross_diffusion.n2.expression = 0.01
reaction.expression = 0.05*n1*n2 - 0.01*n2
reaction.jacobian.n1.expression = 0.05*n2
reaction.jacobian.n2.expression = 0.05*n1 - 0.01
initial.expression = 0.05
[model.time_step_operator]
time_end = 1
""" + SOLVER


# BASELINE config 5 names 10 species: the cell model above with four more (a second extracellular
# messenger, a fourth cytosolic species, two more nuclear ones), cross-diffusion between them and a
# second transmission condition cytosol <-> nucleus.  2 + 4 + 4 species in the three compartments.
CELL10 = CELL.replace("[model.time_step_operator]\ntime_end = 1\n", """
[model.scalar_field.e2]
compartment = ecs
storage.expression = 1
cross_diffusion.e2.expression = 0.015
cross_diffusion.e1.expression = 0.001
reaction.expression = k2*e1 - 0.1*e2
reaction.jacobian.e1.expression = k2
reaction.jacobian.e2.expression = -0.1
initial.expression = 0.2
[model.scalar_field.c4]
compartment = cytosol
storage.expression = 1
cross_diffusion.c4.expression = 0.008
cross_diffusion.c1.expression = 0.001
reaction.expression = 0.2*c3 - 0.3*c4*c1
reaction.jacobian.c3.expression = 0.2
reaction.jacobian.c4.expression = -0.3*c1
reaction.jacobian.c1.expression = -0.3*c4
initial.expression = 0.05 + 0.05*position_x
outflow.nucleus.expression = 0.1*(c4 - n3)
outflow.nucleus.jacobian.c4.expression = 0.1
outflow.nucleus.jacobian.n3.expression = -0.1
[model.scalar_field.n3]
compartment = nucleus
storage.expression = 1
cross_diffusion.n3.expression = 0.01
reaction.expression = -0.02*n3 + 0.01*n1
reaction.jacobian.n3.expression = -0.02
reaction.jacobian.n1.expression = 0.01
initial.expression = 0.02
outflow.cytosol.expression = 0.1*(n3 - c4)
outflow.cytosol.jacobian.n3.expression = 0.1
outflow.cytosol.jacobian.c4.expression = -0.1
[model.scalar_field.n4]
compartment = nucleus
storage.expression = 1
cross_diffusion.n4.expression = 0.005
cross_diffusion.n3.expression = 0.002
reaction.expression = 0.02*n3*n4 - 0.01*n4
reaction.jacobian.n3.expression = 0.02*n4
reaction.jacobian.n4.expression = 0.02*n3 - 0.01
initial.expression = 0.1
[model.time_step_operator]
time_end = 1
""")




def _on_gmsh_ids(text: str) -> str:
    """the same model with the compartments marked by the mesh's cell datum, as the reference's Gmsh
    workflows do (`compartments.<name>.expression = (gmsh_id == k)`, e.g. test/two_disks.ini)"""
    head, tail = text.split("[parser_context]", 1)
    return ("\n[compartments]\necs.expression = (gmsh_id == 3)\ncytosol.expression = (gmsh_id == 2)\n"
            "nucleus.expression = (gmsh_id == 1)\n[parser_context]" + tail)


# BASELINE configs[4]: the 10-species model on the nested-compartment tetrahedral mesh of meshgen.nested_compartments
CELL10_NESTED = _on_gmsh_ids(CELL10)
CELL_NESTED = _on_gmsh_ids(CELL)

WORKLOADS = {"grayscott": GRAY_SCOTT, "cell": CELL, "cell10": CELL10, "cell_nested": CELL_NESTED,
             "cell10_nested": CELL10_NESTED}


def config(name: str, **overrides) -> Config:
    """The workload's configuration with ini keys overridden (`model.time_step_operator.type = ...`)."""
    cfg = Config(WORKLOADS[name])
    for k, v in overrides.items():
        cfg.set(k, v)
    return cfg


def ini_text(name: str, **overrides) -> str:
    return config(name, **overrides).dump()
