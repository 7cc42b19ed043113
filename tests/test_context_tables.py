"""Tabulated parser_context functions: `type = interpolation` (src/dune/copasi/parser/context.cc:72-97) and
`type = function` with `interpolate = true` (:237-283), restated literally in the oracle (Python evaluator and C
byte-code VM) and in the product's front-end (host evaluation, generated CUDA source compiled as host C++)."""
import math
import os
import subprocess
import tempfile

import numpy as np
import pytest

import cases as K
from oracle import core as ORC, expr as E, ini as INI

CTX = """
[parser_context.tab]
type = interpolation
domain = -1 0 0.5 2
range = 3 1 -2 4
[parser_context.sq]
type = function
expression = x: c*x^2
interpolate = true
interpolation.intervals = 8
interpolation.domain.x = 0 2
interpolation.out_of_bounds = clamp
[parser_context.strict]
type = function
expression = t: 1 + t
interpolate = true
interpolation.intervals = 4
interpolation.domain.t = 0 1
[parser_context.c]
type = constant
value = 3
"""


def _ctx():
    return E.Context.from_config(INI.sub(INI.parse_ini(CTX), "parser_context"))


def test_interpolation_table_is_lower_bound_plus_lerp():
    t = _ctx().tables["tab"]
    assert t(-5) == 3 and t(-1) == 3          # lower_bound == begin: front
    assert t(2.5) == 4                        # lower_bound == end: back
    assert t(0) == 1 and t(0.5) == -2 and t(2) == 4   # nodes are hit exactly
    assert t(-0.5) == pytest.approx(2.0) and t(0.25) == pytest.approx(-0.5) and t(1.25) == pytest.approx(1.0)
    assert E.lerp_std(1.0, 5.0, 1.0) == 5.0 and E.lerp_std(-1.0, 1.0, 0.5) == 0.0


def test_sampled_function_is_the_references_literal_piecewise_constant():
    ctx = _ctx()
    sq = ctx.tables["sq"]
    assert len(sq.range) == 9 and sq.range[4] == 3.0      # c x^2 at x = 1
    # context.cc:262-264: i and j are the same interval number -> the left sample of the interval
    assert sq(1.0) == 3.0 and sq(1.2) == 3.0 and sq(1.249) == 3.0 and sq(1.25) == pytest.approx(3 * 1.25 ** 2)
    assert sq(2.0) == 12.0
    assert sq(-4.0) == 0.0                                # clamp = max(domain[0], pos)
    assert sq(7.0) == 12.0                                # past the table: last sample (the reference reads out of bounds)
    strict = ctx.tables["strict"]
    assert strict(0.5) == 1.5
    with pytest.raises(E.ExprError, match="out of bounds"):
        strict(1.5)


def test_c_vm_and_python_evaluator_agree():
    ctx = _ctx()
    sym = E.Symbols(2, ["u"])
    text = "tab(u) + 2*sq(u + position_x) - strict(0.25*abs(u))"
    code, consts = E.compile_expr(text, sym, ctx)
    ast = E.resolve(E.Parser(text).parse(), ctx)
    rows = np.zeros((50, sym.nslots))
    rng = np.random.default_rng(5)
    rows[:, sym.value_slot(0)] = rng.uniform(-1.5, 2.5, 50)
    rows[:, E.SLOT_POS] = rng.uniform(0, 1, 50)
    got = ORC.eval_program(code, consts, rows)
    for k in range(50):
        want = E.py_eval(ast, {"u": rows[k, sym.value_slot(0)], "position_x": rows[k, E.SLOT_POS]})
        assert got[k] == want
    # out of bounds under `error`: NaN in the VM
    code, consts = E.compile_expr("strict(u)", sym, ctx)
    rows[0, sym.value_slot(0)] = 2.0
    assert math.isnan(ORC.eval_program(code, consts, rows[:1])[0])


def test_generated_source_matches_the_oracle():
    """Model::cuda_source with the tables as constant arrays, compiled as host C++"""
    case = K.CASES["tables"]
    om = case.oracle()
    cfg, model, grid = K.product_objects(case)
    src = model.cuda_source().split("// Argument blocks shared")[0]
    assert "dc_tab_rate_" in src and "dc_tab_profile_" in src
    rng = np.random.default_rng(2)
    pts = [(float(rng.uniform(0.0, 2.2)), float(rng.uniform(-0.2, 1.2))) for _ in range(40)]
    body = ["#include <cmath>\n#include <cstdio>\nusing namespace std;\n#define __device__\n#define __host__\n"
            "#define __forceinline__ inline\n#define __noinline__\n", src,
            "int main(){ DcCtx c{}; double u[4], g[4][DC_DIM] = {}, sc[4], jm[1][1]; c.in_volume = 1;\n"]
    for u, x in pts:
        body.append(f"u[0]={u!r}; c.pos[0]={x!r}; DcComp<0>::scalar(c,u,g,0.0,1.0,sc); printf(\"%.17g\\n\", sc[0]);"
                    f" DcComp<0>::jac_mass(c,u,g,0.0,1.0,jm); printf(\"%.17g\\n\", jm[0][0]);\n")
    body.append("return 0; }\n")
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "m.cpp"), "w").write("".join(body))
        subprocess.check_call(["g++", "-std=c++17", "-O0", "-o", os.path.join(td, "m"), os.path.join(td, "m.cpp")])
        out = [float(v) for v in subprocess.check_output([os.path.join(td, "m")], text=True).split()]
    progs = {k: om.progs[prog] for k, ti, tj, tk, prog in om.terms if k in (ORC.K_REACTION, ORC.K_REACTION_JAC)}
    for n, (u, x) in enumerate(pts):
        ctx = np.zeros((1, om.sym.nslots))
        ctx[0, E.SLOT_INVOL] = 1
        ctx[0, E.SLOT_POS] = x
        ctx[0, om.sym.value_slot(0)] = u
        r = ORC.eval_program(*progs[ORC.K_REACTION], ctx)[0]
        j = ORC.eval_program(*progs[ORC.K_REACTION_JAC], ctx)[0]
        assert out[2 * n] == pytest.approx(-r, rel=1e-14, abs=1e-300)
        assert out[2 * n + 1] == pytest.approx(-j, rel=1e-14, abs=1e-300)


def test_product_host_evaluation_and_errors():
    import dune_copasi_b200 as D
    case = K.CASES["tables"]
    om = case.oracle()
    cfg, model, grid = K.product_objects(case)
    # initial values and the constant-folded diffusion coefficient go through the host evaluator
    u0 = grid.interpolate(model, 0.0)
    assert np.allclose(u0, om.initial(0.0), rtol=0, atol=1e-15)
    model.precompile()                        # NVRTC accepts the generated tables (sm_100a, no GPU needed)
    bad = case.ini_with(**{"parser_context.rate.domain": "0 1 0.5 2 3"})
    with pytest.raises(D.DcbError, match="must be sorted"):
        D.Model(D.Config(bad), 2)
    bad = case.ini_with(**{"parser_context.bump.interpolation.intervals": "0"})
    with pytest.raises(D.DcbError, match="at least one interval"):
        D.Model(D.Config(bad), 2)
    bad = case.ini_with(**{"parser_context.img.type": "tiff", "parser_context.img.path": "no_such_file.tif"})
    with pytest.raises(D.DcbError, match="does not exists"):
        D.Model(D.Config(bad), 2)
    bad = case.ini_with(**{"parser_context.f.type": "random_field"})
    with pytest.raises(D.DcbError, match="not supported"):
        D.Model(D.Config(bad), 2)


TIFF_INI = """
[compartments.domain]
type = expression
expression = 1
[parser_context.img]
type = tiff
path = {a}
[parser_context.mask]
type = tiff
path = {b}
[model.scalar_field.u]
compartment = domain
initial.expression = img(position_x, position_y) + 0.5*mask(position_x + 0.25, position_y)
storage.expression = 1
cross_diffusion.u.expression = 0.01
reaction.expression = -u*mask(position_x, position_y)
reaction.jacobian.u.expression = -mask(position_x, position_y)
[model.time_step_operator]
time_step_max = 0.1
time_end = 10
""" + K.SOLVER


def test_tiff_images_as_context_functions(tmp_path):
    """`type = tiff` (context.cc:66-71, tiff_grayscale.cc:91-105): the product's reader (csrc/tiff.cpp) and the oracle's
    (oracle/tiff.py) agree on every layout they read, host evaluation and the generated device code reproduce the
    oracle's pixel lookup -- inside the image, on its edges and outside (clamped / wrapped as in the reference)."""
    import dune_copasi_b200 as D
    from oracle import tiff as TIFF
    rng = np.random.default_rng(9)
    a, b = str(tmp_path / "a.tif"), str(tmp_path / "b.tif")
    pa = rng.integers(0, 256, (12, 20))
    pb = rng.integers(0, 65536, (33, 17))
    TIFF.write(a, pa, bits=8, x_res=(20, 1), y_res=(12, 1))                                   # unit square, MinIsBlack
    TIFF.write(b, pb, bits=16, x_res=(34, 3), y_res=(33, 2), x_off=(1, 4), y_off=(1, 8), photometric=0, packbits=True,
               big_endian=True, rows_per_strip=5)                                             # offsets, MinIsWhite, strips
    ia, ib = TIFF.read(a), TIFF.read(b)
    assert np.array_equal(ia.values, pa / 256.0) and np.array_equal(ib.values, (65536.0 - pb) / 65536.0)
    assert ia(0.0, 0.0) == pa[11, 0] / 256.0 and ia(0.999, 0.999) == pa[0, 19] / 256.0       # y runs upwards from the last line
    assert ia(5.0, 0.5) == pa[5, 19] / 256.0 and ia(0.5, 7.0) == pa[11, 10] / 256.0            # right of / above: clamp, wrap
    text = TIFF_INI.format(a=a, b=b)
    case = K.Case("tiff", text, 2, lambda: K.OMESH.structured(2, [9, 7]), dt=0.05, structured=([9, 7], [0, 0], [1, 1]))
    om = case.oracle()
    cfg, model, grid = K.product_objects(case)
    assert np.array_equal(grid.interpolate(model, 0.0), om.initial(0.0))                      # both readers, both evaluators
    src = model.cuda_source().split("// Argument blocks shared")[0]
    pts = [(float(rng.uniform(0.1, 1.0)), float(rng.uniform(-0.3, 1.4)), float(rng.uniform(-0.3, 1.4))) for _ in range(60)]
    body = ["#include <cmath>\n#include <cstdio>\nusing namespace std;\n#define __device__\n#define __host__\n"
            "#define __forceinline__ inline\n#define __noinline__\n", src,
            "int main(){ DcCtx c{}; double u[4], g[4][DC_DIM] = {}, sc[4]; c.in_volume = 1;\n"]
    for u, x, y in pts:
        body.append(f"u[0]={u!r}; c.pos[0]={x!r}; c.pos[1]={y!r}; DcComp<0>::scalar(c,u,g,0.0,1.0,sc); printf(\"%.17g\\n\", sc[0]);\n")
    body.append("return 0; }\n")
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "m.cpp"), "w").write("".join(body))
        subprocess.check_call(["g++", "-std=c++17", "-O0", "-o", os.path.join(td, "m"), os.path.join(td, "m.cpp")])
        out = [float(v) for v in subprocess.check_output([os.path.join(td, "m")], text=True).split()]
    prog = {k: om.progs[p] for k, ti, tj, tk, p in om.terms if k == ORC.K_REACTION}[ORC.K_REACTION]
    for n, (u, x, y) in enumerate(pts):
        ctx = np.zeros((1, om.sym.nslots))
        ctx[0, E.SLOT_INVOL] = 1
        ctx[0, E.SLOT_POS], ctx[0, E.SLOT_POS + 1] = x, y
        ctx[0, om.sym.value_slot(0)] = u
        assert out[n] == pytest.approx(-ORC.eval_program(*prog, ctx)[0], rel=1e-15, abs=1e-300)
    model.precompile()
    # formats the own reader does not cover fail loudly, as do images too large for device code
    bad = str(tmp_path / "rgb.tif")
    raw = bytearray(open(a, "rb").read())
    raw[raw.index(b"\x06\x01\x03\x00") + 8] = 2            # PhotometricInterpretation = RGB
    open(bad, "wb").write(bytes(raw))
    with pytest.raises(D.DcbError, match="must be in grayscale"):
        D.Model(D.Config(TIFF_INI.format(a=bad, b=b)), 2)
    big = str(tmp_path / "big.tif")
    TIFF.write(big, rng.integers(0, 256, (600, 600)), bits=8, x_res=(600, 1), y_res=(600, 1))
    big_model = D.Model(D.Config(TIFF_INI.format(a=a, b=big)), 2)      # fine on the host ...
    with pytest.raises(D.DcbError, match="too large for device code"):
        big_model.cuda_source()                                        # ... refused where a kernel would need it


def test_reference_tiff_unit_test_restated(tmp_path):
    """test/dune/copasi/common/tiff_grayscale.cc (the reference's own unit test; its flower-minisblack-{04,08,16}.tif
    are git-LFS pointers here, so the same picture is synthesised at 72 dpi): a 4-bit image is refused
    (UnsupportedEncoding), the 8-bit and the 16-bit version of one picture agree within 2 / 255 pixel by pixel and
    position by position (Compare16vs8Bits) -- through the oracle's reader and the product's."""
    import dune_copasi_b200 as D
    from oracle import tiff as TIFF
    rng = np.random.default_rng(4)
    rows, cols = 43, 73
    yy, xx = np.mgrid[0:rows, 0:cols]
    pic8 = np.clip(128 + 100 * np.sin(xx / 9.0) * np.cos(yy / 7.0) + rng.integers(-5, 6, (rows, cols)), 0, 255).astype(np.int64)
    p8, p16, p4 = (str(tmp_path / n) for n in ("f08.tif", "f16.tif", "f04.tif"))
    TIFF.write(p8, pic8, bits=8, x_res=(72, 1), y_res=(72, 1))
    TIFF.write(p16, pic8 * 257, bits=16, x_res=(72, 1), y_res=(72, 1))
    raw = bytearray(open(p8, "rb").read())
    raw[raw.index(b"\x02\x01\x03\x00") + 8] = 4          # BitsPerSample = 4
    open(p4, "wb").write(bytes(raw))
    with pytest.raises(TIFF.TiffError, match="4 bits not implemented"):
        TIFF.read(p4)
    i8, i16 = TIFF.read(p8), TIFF.read(p16)
    assert (i8.rows, i8.cols) == (i16.rows, i16.cols) == (rows, cols)
    thr = 2.0 / 255
    assert np.abs(i8.values - i16.values).max() <= thr
    res_unit = 72
    for i in range(0, rows, 3):
        for j in range(0, cols, 5):
            assert abs(i8(j / res_unit, i / res_unit) - i16(j / res_unit, i / res_unit)) <= thr
    # the product's reader: the same three files behind parser_context functions
    ini = K.EXP.replace("initial.expression = 1", "initial.expression = a(position_x, position_y) - b(position_x, position_y)")
    ini += "\n[parser_context.a]\ntype = tiff\npath = {}\n[parser_context.b]\ntype = tiff\npath = {}\n"
    case = K.Case("tiff_unit", ini.format(p8, p16), 2, lambda: K.OMESH.structured(2, [8, 8]), structured=([8, 8], [0, 0], [1, 1]))
    cfg, model, grid = K.product_objects(case)
    u0 = grid.interpolate(model, 0.0)
    assert np.abs(u0).max() <= thr and np.array_equal(u0, case.oracle().initial(0.0))
    with pytest.raises(D.DcbError, match="4 bits not implemented"):
        D.Model(D.Config(ini.format(p4, p16)), 2)


def test_tiff_readers_against_libtiff_written_files(tmp_path):
    """A pin against the library the reference reads its images with: Pillow bundles libtiff, so files written by it --
    uncompressed, PackBits, LZW, Deflate, with and without the horizontal predictor, 8 and 16 bit, 72 dpi -- are real
    libtiff output.  The oracle's reader must return exactly the pixels that went in, and the product's reader must
    agree with it at every pixel (sampled through `img(position_x, position_y)` on a lattice of pixel centres)."""
    PIL = pytest.importorskip("PIL.Image")
    import dune_copasi_b200 as D
    from oracle import tiff as TIFF
    rng = np.random.default_rng(1)
    rows, cols = 37, 53
    yy, xx = np.mgrid[0:rows, 0:cols]
    base = (128 + 90 * np.sin(xx / 8.0) * np.cos(yy / 5.0)).astype(np.int64) + rng.integers(0, 12, (rows, cols))
    ini = K.EXP.replace("initial.expression = 1", "initial.expression = img(position_x, position_y)")
    ini += "\n[parser_context.img]\ntype = tiff\npath = {}\n"
    # vertices at the pixel centres: x = (j + 1/2) / 72, y counted upwards from the last scanline
    origin, extent = [0.5 / 72, 0.5 / 72], [(cols - 1) / 72, (rows - 1) / 72]
    mesh = K.OMESH.structured(2, [cols - 1, rows - 1], origin, extent)
    for bits, arr in ((8, base.astype(np.uint8)), (16, (base * 200).astype(np.uint16))):
        for comp in ("raw", "packbits", "tiff_lzw", "tiff_adobe_deflate"):
            for pred in (1, 2):
                if pred == 2 and comp in ("raw", "packbits"):
                    continue
                path = str(tmp_path / f"p{bits}_{comp}_{pred}.tif")
                kw = {"compression": comp, "dpi": (72, 72)}
                if pred == 2:
                    kw["tiffinfo"] = {317: 2}
                PIL.fromarray(arr).save(path, **kw)
                im = TIFF.read(path)
                assert np.array_equal(im.values * 2.0 ** bits, arr.astype(np.float64)), (bits, comp, pred)
                assert float(im.x_res) == 72.0 and float(im.y_res) == 72.0
                cfg = D.Config(ini.format(path))
                model = D.Model(cfg, 2)
                grid = D.Grid.structured(2, [cols - 1, rows - 1], origin, extent)
                grid.bind(model)
                u0 = grid.interpolate(model, 0.0).reshape(rows, cols)           # vertex (j, i): x fastest
                assert np.array_equal(u0[::-1], arr / 2.0 ** bits), (bits, comp, pred)  # y upwards = scanlines reversed
