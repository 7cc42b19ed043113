"""Host logic of the product (through the C ABI) against the oracle -- bit exact for every integer
structure: mesh numbering, compartment marking, DOF numbering, interface/boundary facets, sparsity
pattern, constraint sets; initial values to the last bit.  CPU only (no compute entry point)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import cases as K

ALL = list(K.ALL_CASES)     # incl. the Q1 cube grids (non-reference element, own oracle)


@pytest.mark.parametrize("name", ALL)
def test_mesh_dofmap_pattern_bit_exact(name):
    import dune_copasi_b200 as D
    case = K.ALL_CASES[name]
    om = case.oracle()
    cfg, model, grid = K.product_objects(case)
    m = om.mesh
    if case.structured:
        g2 = D.Grid.structured(case.dim, *case.structured, element=case.element)
        assert np.array_equal(g2.coords(), m.coords)          # same doubles, same order
        assert np.array_equal(g2.elements(), m.elems)
    assert [s for s, _ in model.species()] == om.names
    assert [c for _, c in model.species()] == [s.comp for s in om.species]
    assert np.array_equal(grid.elem_compartment(), m.elem_comp)
    assert grid.ndofs == m.ndofs
    assert np.array_equal(grid.elem_dof(), m.elem_dof)
    fi, fo, li, lo = grid.facets()
    if len(fi) or om.has_outflow:
        for a, b in zip((fi, fo, li, lo), (m.f_in, m.f_out, m.f_lin, m.f_lout)):
            assert np.array_equal(a, b)
    rp, ci = grid.pattern(model)
    orp, oci = om.pattern()
    assert np.array_equal(rp, orp) and np.array_equal(ci, oci)
    assert np.all(np.diff(rp) > 0)
    for r in (0, len(rp) // 2, len(rp) - 2):                  # sorted columns, diagonal present
        cols = ci[rp[r]:rp[r + 1]]
        assert np.all(np.diff(cols) > 0) and r in cols
    assert np.array_equal(grid.interpolate(model, case.t0), om.initial(case.t0))
    d, v = grid.constraints(model)
    od, ov = om.constraints()
    o1, o2 = np.argsort(d), np.argsort(od)
    assert np.array_equal(d[o1], od[o2]) and np.array_equal(v[o1], ov[o2])


def test_ragged_and_empty_compartments():
    """A compartment without cells, cells without a compartment, a species-free compartment."""
    import dune_copasi_b200 as D
    ini = """
[compartments]
left.expression = position_x < 0.3
nowhere.expression = position_x > 5
empty_species.expression = position_x > 0.8
[model.scalar_field.a]
compartment = left
storage.expression = 1
cross_diffusion.a.expression = 1
[model.scalar_field.b]
compartment = nowhere
storage.expression = 1
"""
    case = K.Case("ragged", ini, 2, lambda: K.OMESH.structured(2, [10, 10]), structured=([10, 10], [0, 0], [1, 1]))
    om = case.oracle()
    cfg, model, grid = K.product_objects(case)
    assert model.ncomp == 3 and model.nspec == 2
    ec = grid.elem_compartment()
    assert np.array_equal(ec, om.mesh.elem_comp)
    assert (ec == -1).any() and (ec == 0).any() and not (ec == 1).any() and (ec == 2).any()
    assert grid.ndofs == om.ndofs > 0
    rp, ci = grid.pattern(model)
    orp, oci = om.pattern()
    assert np.array_equal(rp, orp) and np.array_equal(ci, oci)


def test_configuration_errors_are_loud():
    import dune_copasi_b200 as D
    base = K.CASES["exp"].ini
    with pytest.raises(D.DcbError, match="unknown symbol"):
        D.Model(D.Config(base.replace("grow_rate*u", "grow_rate*undefined_name")), 2).cuda_source()
    with pytest.raises(D.DcbError, match="repeated|compartment"):
        D.Model(D.Config("[model.scalar_field.u]\ncompartment = x\n"), 2)
    with pytest.raises(D.DcbError, match="not known type"):
        D.Model(D.Config(base + "\n[model.scalar_field.u.cross_diffusion.u]\ntype = matrix\nexpression = 1\n"), 2)
    # overlapping compartments are refused at bind time
    cfg = D.Config("[compartments]\na.expression = 1\nb.expression = 1\n[model.scalar_field.u]\ncompartment = a\nstorage.expression = 1\n")
    model = D.Model(cfg, 2)
    with pytest.raises(D.DcbError, match="overlapping"):
        D.Grid.structured(2, [2, 2]).bind(model)


def test_vtk_output_round_trips(tmp_path):
    """Model::write_vtk layout (model_multi_compartment.impl.hh:218-300): one .vtu per compartment and
    stamp with a vertex array per species, a .pvd per compartment; values read back exactly."""
    import xml.etree.ElementTree as ET
    case = K.CASES["two_disks"]
    om = case.oracle()
    cfg, model, grid = K.product_objects(case)
    out = tmp_path / "run" / "two_disks"
    u = K.rand_state(om.ndofs, 3)
    grid.write_vtk(model, u, 0.0, out, append=False)
    grid.write_vtk(model, 2.0 * u, 0.5, out, append=True)
    names = sorted(p.name for p in out.iterdir())
    assert names == ["two_disks-inner-00000.vtu", "two_disks-inner-00001.vtu", "two_disks-inner.pvd",
                     "two_disks-outer-00000.vtu", "two_disks-outer-00001.vtu", "two_disks-outer.pvd"]
    m = om.mesh
    for c, cname in enumerate(om.comp_names):
        piece = ET.parse(out / f"two_disks-{cname}-00001.vtu").getroot().find("UnstructuredGrid/Piece")
        verts = m.comp_vertices[c]
        assert int(piece.get("NumberOfPoints")) == len(verts)
        assert int(piece.get("NumberOfCells")) == int((m.elem_comp == c).sum())
        arrays = {a.get("Name"): np.array(a.text.split(), dtype=float) for a in piece.iter("DataArray")}
        sp = [s for s in om.species if s.comp == c]
        for s in sp:
            expect = 2.0 * u[m.comp_offset[c] + np.arange(len(verts)) * len(sp) + s.local]
            assert np.array_equal(arrays[s.name], expect)
        assert np.array_equal(arrays["Coordinates"].reshape(-1, 3)[:, :2], m.coords[verts])
        conn = arrays["connectivity"].astype(int).reshape(-1, 3)
        assert np.array_equal(np.asarray(verts)[conn], m.elems[m.elem_comp == c])
        assert set(arrays["types"]) == {5.0}
        stamps = [float(d.get("timestep")) for d in ET.parse(out / f"two_disks-{cname}.pvd").getroot().iter("DataSet")]
        assert stamps == [0.0, 0.5]
    grid.write_vtk(model, u, 1.0, out, append=False)      # restart of the sequence
    assert len(list(ET.parse(out / "two_disks-inner.pvd").getroot().iter("DataSet"))) == 1


def test_reduce_kernels_compile_without_a_gpu_and_errors_are_loud():
    import dune_copasi_b200 as D
    for name in ("gauss2d", "two_disks", "cell3d"):
        cfg, model, grid = K.product_objects(K.CASES[name])
        model.precompile(reduce_config=cfg)       # NVRTC -> sm_100a cubin in the JIT cache
    bad = K.CASES["exp"].ini_with(**{"model.reduce.u_max.reduction.expression": "a, b, c: max(a, b)"})
    cfg = D.Config(bad)
    with pytest.raises(D.DcbError, match="Reduction arguments must be exactly 2"):
        D.Model(cfg, 2).precompile(reduce_config=cfg)
    bad = K.CASES["exp"].ini_with(**{"model.reduce.u_max.evaluation.expression": "u + nonsense"})
    cfg = D.Config(bad)
    with pytest.raises(D.DcbError, match="unknown symbol"):
        D.Model(cfg, 2).precompile(reduce_config=cfg)


def test_generated_cuda_matches_oracle_vm():
    """The expression lowering (Model::cuda_source) is plain C++ once the CUDA qualifiers are
    defined away: compile it with g++ and compare the point functions of every case with the
    oracle's byte-code interpreter on random arguments."""
    import dune_copasi_b200 as D
    from oracle import core as ORC, expr as E
    rng = np.random.default_rng(0)
    for name in ("grayscott3d", "mitchell_schaefer", "cell3d", "gauss2d"):
        case = K.CASES[name]
        om = case.oracle()
        cfg, model, grid = K.product_objects(case)
        src = model.cuda_source().split("// Argument blocks shared")[0]
        dim = case.dim
        body = ["#include <cmath>\n#include <cstdio>\nusing namespace std;\n#define __device__\n#define __host__\n#define __forceinline__ inline\n#define __noinline__\n", src,
                "int main(){ DcCtx c{}; double u[16], g[16][DC_DIM], sc[16], jm[16][16];\n"]
        # one sample point per compartment
        samples = []
        for comp in range(om.ncomp):
            ns = om.comp_nspec[comp]
            if ns == 0:
                continue
            uu = rng.uniform(0.1, 1.0, ns)
            pos = rng.uniform(0, 1, 3)
            t = 0.7
            samples.append((comp, uu, pos, t))
            body.append(f"c.time={t}; c.pos[0]={float(pos[0])!r}; c.pos[1]={float(pos[1])!r}; c.pos[2]={float(pos[2]) if dim == 3 else 0.0!r};\n")
            for s in range(ns):
                body.append(f"u[{s}]={float(uu[s])!r};")
            body.append(f"for(int i=0;i<16;++i)for(int k=0;k<DC_DIM;++k)g[i][k]=0;\n")
            body.append(f"DcComp<{comp}>::scalar(c,u,g,0.0,1.0,sc); for(int i=0;i<{ns};++i) printf(\"%.17g\\n\", sc[i]);\n")
            body.append(f"DcComp<{comp}>::scalar(c,u,g,1.0,0.0,sc); for(int i=0;i<{ns};++i) printf(\"%.17g\\n\", sc[i]);\n")
            body.append(f"{{ double (*J)[DcComp<{comp}>::NS] = (double(*)[DcComp<{comp}>::NS])jm; DcComp<{comp}>::jac_mass(c,u,g,0.0,1.0,J);"
                        f" for(int i=0;i<{ns};++i)for(int j=0;j<{ns};++j) printf(\"%.17g\\n\", J[i][j]); }}\n")
        body.append("return 0; }\n")
        with tempfile.TemporaryDirectory() as td:
            open(os.path.join(td, "m.cpp"), "w").write("".join(body))
            subprocess.check_call(["g++", "-std=c++17", "-O0", "-o", os.path.join(td, "m"), os.path.join(td, "m.cpp")])
            out = subprocess.check_output([os.path.join(td, "m")], text=True).split()
        got = iter(float(x) for x in out)
        for comp, uu, pos, t in samples:
            ns = om.comp_nspec[comp]
            g0 = int(om._comp_ptr[comp])
            ctx = np.zeros((1, om.sym.nslots))
            ctx[0, E.SLOT_TIME] = t
            ctx[0, E.SLOT_INVOL] = 1
            ctx[0, E.SLOT_POS:E.SLOT_POS + dim] = pos[:dim]
            for s in range(ns):
                ctx[0, om.sym.value_slot(g0 + s)] = uu[s]

            def term(kind, i, j=-1):
                for k, ti, tj, tk, prog in om.terms:
                    if k == kind and ti == i and (j < 0 or tj == j):
                        c_, k_ = om.progs[prog]
                        return ORC.eval_program(c_, k_, ctx)[0]
                return None
            for s in range(ns):       # stiffness scalar = -R
                r = term(ORC.K_REACTION, g0 + s)
                assert next(got) == pytest.approx(-(r or 0.0), rel=1e-14, abs=1e-300), (name, comp, s)
            for s in range(ns):       # mass scalar = u * storage
                st = term(ORC.K_STORAGE, g0 + s)
                assert next(got) == pytest.approx(uu[s] * (st or 0.0), rel=1e-14, abs=1e-300)
            for i in range(ns):
                for j in range(ns):
                    jr = term(ORC.K_REACTION_JAC, g0 + i, g0 + j) if term(ORC.K_REACTION, g0 + i) is not None else None
                    assert next(got) == pytest.approx(-(jr or 0.0), rel=1e-14, abs=1e-300), (name, comp, i, j)


def test_nvrtc_compiles_every_case_without_gpu():
    """NVRTC cross-compiles the per-model kernels for sm_100a on the CPU box."""
    for name in ("exp", "two_disks"):
        cfg, model, grid = K.product_objects(K.CASES[name])
        cubin = model.compile()
        assert cubin[:4] == b"\x7fELF" and len(cubin) > 10000
        assert b"dc_k_patch_residual_0" in cubin and b"dc_k_jacobian_volume_0" in cubin


@pytest.mark.parametrize("name,vtk_type", [("grayscott2d_q1", 8.0), ("grayscott3d_q1", 11.0)])
def test_vtk_output_of_q1_lattices(tmp_path, name, vtk_type):
    """Q1 cells are written as VTK_PIXEL / VTK_VOXEL, whose corner order is the bit pattern used here."""
    import xml.etree.ElementTree as ET
    case = K.Q1_CASES[name]
    om = case.oracle()
    cfg, model, grid = K.product_objects(case)
    u = K.rand_state(om.ndofs, 4)
    out = tmp_path / "q1"
    grid.write_vtk(model, u, 0.0, out, append=False)
    piece = ET.parse(out / "q1-compartment-00000.vtu").getroot().find(".//Piece")
    assert int(piece.get("NumberOfPoints")) == om.mesh.nv and int(piece.get("NumberOfCells")) == om.mesh.ne
    arrays = {a.get("Name"): np.array(a.text.split(), dtype=float) for a in piece.iter("DataArray")}
    nd = om.mesh.elems.shape[1]
    assert set(arrays["types"]) == {vtk_type}
    assert np.array_equal(arrays["connectivity"].astype(int).reshape(-1, nd), om.mesh.elems)
    assert np.array_equal(arrays["offsets"].astype(int), nd * np.arange(1, om.mesh.ne + 1))
    assert np.array_equal(arrays["U"], u[0::2]) and np.array_equal(arrays["V"], u[1::2])
    assert np.array_equal(arrays["Coordinates"].reshape(-1, 3)[:, :case.dim], om.mesh.coords)


def test_blocked_layout_flags_change_the_container_nesting_only():
    """model.blocked_layout.{scalar_fields, compartments} (factory.hh:74-75): EntityGrouping<.., Blocked> /
    Lexicographic<Blocked> nest the dune-istl containers; the flat scalar order the C ABI uses is the same."""
    base = K.CASES["cell3d"]
    ref = None
    for sf in ("false", "true"):
        for cb in ("false", "true"):
            cfg, model, grid = K.product_objects(base, **{"model.blocked_layout.scalar_fields": sf,
                                                          "model.blocked_layout.compartments": cb})
            got = (grid.ndofs, grid.elem_dof().tobytes(), grid.pattern(model)[0].tobytes(), grid.pattern(model)[1].tobytes())
            ref = ref or got
            assert got == ref
