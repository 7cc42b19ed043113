"""model.jacobian.type = symbolic: the Jacobian entries derived by the product's own differentiator
(csrc/expr.cpp) equal the entries the reference's inis write by hand (`*.jacobian.<wrt>.expression`,
local_equations.hh:553-579), term by term -- reaction, storage, scalar and tensor diffusion,
velocity, outflow -- and produce the same sparsity pattern.  CPU only: the generated model source
is plain C++ once the CUDA qualifiers are defined away (as in test_host_parity.py)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import cases as K

PRELUDE = ("#include <cmath>\n#include <cstdio>\nusing namespace std;\n#define __device__\n#define __host__\n"
           "#define __forceinline__ inline\n#define __noinline__\n")


def _program(model, om, dim, seed):
    """main() that prints every Jacobian coefficient function of the generated source at random points"""
    rng = np.random.default_rng(seed)
    src = model.cuda_source().split("// Argument blocks shared")[0]
    body = [PRELUDE, src, "template <class T> void show(const T* p, int n){ for(int i=0;i<n;++i) printf(\"%.17g\\n\", ((const double*)p)[i]); }\n",
            "int main(){ DcCtx c{}; double us[16], ut[16], gs[16][DC_DIM], gt[16][DC_DIM];\n"]
    pos = rng.uniform(0.1, 0.9, 3)
    body.append(f"c.time=0.7; c.in_volume=1; c.pos[0]={float(pos[0])!r}; c.pos[1]={float(pos[1])!r}; c.pos[2]={float(pos[2]) if dim == 3 else 0.0!r};\n")
    if om.mesh.cell_keys:
        for k in range(len(om.mesh.cell_keys)):
            body.append(f"c.cell[{k}]={float(rng.uniform(0.2, 1.5))!r};\n")
    for s in range(16):
        body.append(f"us[{s}]={float(rng.uniform(0.1, 1.0))!r}; ut[{s}]={float(rng.uniform(0.1, 1.0))!r};")
        for k in range(dim):
            body.append(f"gs[{s}][{k}]={float(rng.uniform(-1, 1))!r}; gt[{s}][{k}]={float(rng.uniform(-1, 1))!r};")
    body.append("\n")
    for comp in range(om.ncomp):
        if om.comp_nspec[comp] == 0:
            continue
        M = f"DcComp<{comp}>"
        body.append(f"{{ double jm[{M}::NS][{M}::NS]; {M}::jac_mass(c,us,gs,0.7,0.3,jm); show(jm,{M}::NS*{M}::NS);\n"
                    f"  double W[{M}::NS][{M}::NS][DC_DIM]; {M}::jac_ext(c,us,gs,0.3,W); show(W,{M}::NS*{M}::NS*DC_DIM);\n"
                    f"  double DT[{M}::NS][{M}::NS][DC_DIM][DC_DIM]; {M}::jac_diff_t(c,us,gs,0.3,DT); show(DT,{M}::NS*{M}::NS*DC_DIM*DC_DIM);\n"
                    f"  printf(\"%d\\n\", (int){M}::HAS_EXT);"
                    f"  for(int i=0;i<{M}::NS;++i)for(int j=0;j<{M}::NS;++j) printf(\"%d\\n\", (int){M}::pair(i,j)); }}\n")
    npairs = src.count("template <> struct DcOutflow<")
    for p in range(npairs):
        O = f"DcOutflow<{p}>"
        body.append(f"{{ double js[{O}::NSS][{O}::NSS], jt[{O}::NSS][{O}::NST]; c.in_volume=0; c.in_skeleton=1; c.nrm[0]=0.6; c.nrm[1]=0.8;\n"
                    f"  {O}::jacobian(c,us,gs,ut,gt,js,jt); show(js,{O}::NSS*{O}::NSS); show(jt,{O}::NSS*{O}::NST); }}\n")
    body.append("return 0; }\n")
    return "".join(body)


def _run(text):
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "m.cpp"), "w").write(text)
        subprocess.check_call(["g++", "-std=c++17", "-O0", "-o", os.path.join(td, "m"), os.path.join(td, "m.cpp")])
        return np.array(subprocess.check_output([os.path.join(td, "m")], text=True).split(), dtype=float)


@pytest.mark.parametrize("name", ["grayscott3d", "mitchell_schaefer", "cell3d", "cell3d_10", "two_disks",
                                  "two_disks_cell_data", "advection2d", "advection3d", "exp"])
def test_derived_entries_equal_the_handwritten_ones(name):
    case = K.CASES[name]
    om = case.oracle()
    outs, patterns = [], []
    for jt in ("analytical", "symbolic"):
        cfg, model, grid = K.product_objects(case, **{"model.jacobian.type": jt})
        outs.append(_run(_program(model, om, case.dim, 7)))
        patterns.append(grid.pattern(model))
    a, s = outs
    assert a.size == s.size and a.size > 0
    scale = np.maximum(np.abs(a), 1e-300)
    assert np.all(np.abs(a - s) <= 1e-12 * np.maximum(scale, np.abs(a).max() * 1e-3)), (name, np.abs(a - s).max())
    assert np.array_equal(patterns[0][0], patterns[1][0]) and np.array_equal(patterns[0][1], patterns[1][1])


def test_the_ini_jacobian_sections_are_not_read_in_symbolic_mode():
    """A deliberately wrong hand-written entry changes the analytical model and leaves the symbolic one alone."""
    import dune_copasi_b200 as D
    case = K.CASES["grayscott2d"]
    wrong = {"model.scalar_field.U.reaction.jacobian.V.expression": "-3*U*V"}
    srcs = {}
    for jt in ("analytical", "symbolic"):
        for tag, over in (("ok", {}), ("wrong", wrong)):
            m = D.Model(D.Config(case.ini_with(**{"model.jacobian.type": jt, **over})), 2)
            srcs[jt, tag] = m.cuda_source()
    assert srcs["analytical", "ok"] != srcs["analytical", "wrong"]
    assert srcs["symbolic", "ok"] == srcs["symbolic", "wrong"]


def test_differentiator_rules_against_differences():
    """Every rule of the differentiator once (functions, powers, quotients, selections, min/max) through
    a one-species model: d(reaction)/du from symbolic mode against a central difference of the
    generated reaction itself."""
    import dune_copasi_b200 as D
    exprs = ["sqrt(1+u^2) + exp(-u)*sin(3*u) - cos(u)/(2+u)", "log(1+u)^2 + tanh(u) - atan(2*u) + u^2.5",
             "(u > 0.4) ? u^3 : 2*u", "min(u, 0.5)*max(u^2, 0.1) + abs(u-0.3)", "pow(u, 3) / (1 + pow(2, u)) + 2^u",
             "sinh(u) - cosh(2*u) + asin(u/2) + acos(u/3) + tan(u/2) + log10(1+u) + log2(2+u) + exp2(u)",
             "hill(u, 0.5) + position_x*u", "atan2(u, 1+u)"]
    pts = [0.23, 0.61, 0.87]
    for e in exprs:
        ini = ("[compartments.domain]\nexpression = 1\n[parser_context.hill]\ntype = function\nexpression = s, K: s^2/(K^2 + s^2)\n"
               "[model]\njacobian.type = symbolic\n[model.scalar_field.u]\ncompartment = domain\nstorage.expression = 1\n"
               f"reaction.expression = {e}\n")
        src = D.Model(D.Config(ini), 2).cuda_source().split("// Argument blocks shared")[0]
        body = [PRELUDE, src, "int main(){ DcCtx c{}; c.pos[0]=0.3; double u[1], g[1][DC_DIM]={{0,0}}, sc[1], jm[1][1];\n"]
        for x in pts:
            body.append(f"u[0]={x!r}; DcComp<0>::jac_mass(c,u,g,0.0,1.0,jm); printf(\"%.17g\\n\", jm[0][0]);\n")
            body.append(f"u[0]={x + 1e-6!r}; DcComp<0>::scalar(c,u,g,0.0,1.0,sc); printf(\"%.17g\\n\", sc[0]);\n")
            body.append(f"u[0]={x - 1e-6!r}; DcComp<0>::scalar(c,u,g,0.0,1.0,sc); printf(\"%.17g\\n\", sc[0]);\n")
        body.append("return 0; }\n")
        out = _run("".join(body)).reshape(-1, 3)
        for (jac, up, dn), x in zip(out, pts):
            fd = (up - dn) / 2e-6       # both are -R: jm = -dR/du
            assert abs(jac - fd) <= 1e-7 * max(1.0, abs(fd)), (e, x, jac, fd)


def test_differentiator_against_sympy():
    """The derived entries against sympy (the Python sibling of SymEngine, the reference's symbolic backend) evaluated
    with 30 digits: agreement to rounding, a much tighter pin than the central differences above."""
    sympy = pytest.importorskip("sympy")
    import dune_copasi_b200 as D
    u, x = sympy.symbols("u position_x")
    fn = {"log10": lambda a: sympy.log(a, 10), "log2": lambda a: sympy.log(a, 2), "exp2": lambda a: 2 ** a,
          "pow": lambda a, b: a ** b, "hill": lambda s, K_: s ** 2 / (K_ ** 2 + s ** 2), "ln": sympy.log}
    exprs = ["sqrt(1+u^2) + exp(-u)*sin(3*u) - cos(u)/(2+u)", "log(1+u)^2 + tanh(u) - atan(2*u) + u^2.5",
             "pow(u, 3) / (1 + pow(2, u)) + 2^u",
             "sinh(u) - cosh(2*u) + asin(u/2) + acos(u/3) + tan(u/2) + log10(1+u) + log2(2+u) + exp2(u)",
             "hill(u, 0.5) + position_x*u", "atan2(u, 1+u)",
             "0.042*(1-u) - u*(0.3+position_x)^2",                                   # Gray-Scott's U equation, V frozen
             "0.2*((0.5+position_x)*u^2*(1-u)/0.1 - u/60)"]                          # Mitchell-Schaefer's u equation, z frozen
    pts = [0.23, 0.61, 0.87]
    for e in exprs:
        ini = ("[compartments.domain]\nexpression = 1\n[parser_context.hill]\ntype = function\nexpression = s, K: s^2/(K^2 + s^2)\n"
               "[model]\njacobian.type = symbolic\n[model.scalar_field.u]\ncompartment = domain\nstorage.expression = 1\n"
               f"reaction.expression = {e}\n")
        src = D.Model(D.Config(ini), 2).cuda_source().split("// Argument blocks shared")[0]
        body = [PRELUDE, src, "int main(){ DcCtx c{}; c.pos[0]=0.3; double u[1], g[1][DC_DIM]={{0,0}}, sc[1], jm[1][1];\n"]
        for p in pts:
            body.append(f"u[0]={p!r}; DcComp<0>::jac_mass(c,u,g,0.0,1.0,jm); printf(\"%.17g\\n\", jm[0][0]);\n")
        body.append("return 0; }\n")
        got = _run("".join(body))
        d = sympy.diff(sympy.sympify(e.replace("^", "**"), locals=dict(fn, u=u, position_x=x)), u)
        for jac, p in zip(got, pts):
            want = -float(d.evalf(30, subs={u: sympy.Float(repr(p), 30), x: sympy.Float("0.3", 30)}))     # jm = -dR/du
            assert abs(jac - want) <= 2e-14 * max(1.0, abs(want)), (e, p, jac, want)
