"""The reference's own ini files, unmodified (plus the command-line overrides its CTest passes), through
the product's configuration / model front-end: they parse, lower to CUDA source, and their hand-written
Jacobian entries equal the symbolic derivatives of their own reaction / storage / outflow expressions.
Runs only where /root/reference is mounted (this container); nothing here is needed on the GPU box."""
import glob
import os
import types

import numpy as np
import pytest

from test_symbolic_jacobian import _program, _run

REF = "/root/reference"
FILES = sorted(glob.glob(f"{REF}/test/*.ini") + glob.glob(f"{REF}/doc/docusaurus/static/ini/next/*.ini"))
# test/CMakeLists.txt:70-80: the mitchell-schaefer system test replaces the random field by a function
OVERRIDES = {"mitchell_schaefer.ini": {"parser_context.rng.type": "function", "parser_context.rng.expression": "x,y:0"}}
# time_snap.ini reads two TIFF images whose files are git-LFS pointers in the reference tree: synthetic images of
# the same names are written next to the test (the file's relative paths resolve against the working directory)
TIFFS = {"time_snap.ini": ("data/tiff/A_initialConcentration.tif", "data/tiff/B_initialConcentration.tif")}
UNSUPPORTED = {}

pytestmark = pytest.mark.skipif(not FILES, reason="reference tree not mounted")


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_reference_ini(path, tmp_path, monkeypatch):
    import dune_copasi_b200 as D
    from oracle import tiff as TIFF
    name = os.path.basename(path)
    if name in TIFFS:
        monkeypatch.chdir(tmp_path)
        rng = np.random.default_rng(3)
        for k, rel in enumerate(TIFFS[name]):
            os.makedirs(os.path.dirname(rel), exist_ok=True)
            TIFF.write(rel, rng.integers(0, 65536, (40, 50)), bits=16, x_res=(50, 1), y_res=(40, 1), packbits=bool(k))
    text = open(path).read()
    keys = [k for k in ("gmsh_id", "sigma") if k in text]
    dim = 2
    outs = []
    for jt in ("analytical", "symbolic"):
        cfg = D.Config(text, **OVERRIDES.get(name, {}))
        cfg.set("model.jacobian.type", jt)
        if name in UNSUPPORTED:
            with pytest.raises(D.DcbError, match=UNSUPPORTED[name]):
                D.Model(cfg, dim, keys)
            return
        model = D.Model(cfg, dim, keys)
        comps = [c for _, c in model.species()]
        shim = types.SimpleNamespace(ncomp=model.ncomp, comp_nspec=[comps.count(c) for c in range(model.ncomp)],
                                     mesh=types.SimpleNamespace(cell_keys=keys))
        outs.append(_run(_program(model, shim, dim, 11)))
    a, s = outs
    assert a.size == s.size and a.size > 0
    tol = 1e-12 * np.maximum(np.abs(a), np.abs(a).max() * 1e-3)
    assert np.all(np.abs(a - s) <= tol), (name, np.abs(a - s).max())


def test_restated_cases_equal_the_reference_files():
    """tests/cases.py restates the reference's test inis as text (the GPU box has no /root/reference): the
    physics sections -- [compartments], [model.scalar_field.*] and the [parser_context] entries they use --
    are the files', entry for entry.  Differences that are allowed, and why:
    rng (replaced by a function exactly as test/CMakeLists.txt:78-79 does), Heaviside / parser_type (unused
    by the equations), u_*_analytic (helpers of the [model.reduce] section, kept with the restated reduce)."""
    import cases as K
    INI = K.INI
    pairs = {"gauss.ini": K.GAUSS, "exp.ini": K.EXP, "poisson.ini": K.POISSON, "two_disks.ini": K.TWO_DISKS,
             "mitchell_schaefer.ini": K.MITCHELL_SCHAEFER, "two_disks_cell_data.ini": K.TWO_DISKS_CELL_DATA}
    allowed = {"rng", "Heaviside", "parser_type", "u_in_analytic", "u_out_analytic"}

    def norm(d):
        return {k: norm(v) for k, v in d.items()} if isinstance(d, dict) else " ".join(str(d).split())
    for name, text in pairs.items():
        ref = INI.parse_ini(open(f"{REF}/test/{name}").read())
        mine = INI.parse_ini(text)
        assert norm(INI.sub(ref, "compartments")) == norm(INI.sub(mine, "compartments")), name
        assert norm(INI.sub(INI.sub(ref, "model"), "scalar_field")) == norm(INI.sub(INI.sub(mine, "model"), "scalar_field")), name
        a, b = norm(INI.sub(ref, "parser_context")), norm(INI.sub(mine, "parser_context"))
        for key in set(a) | set(b):
            if key not in allowed:
                assert a.get(key) == b.get(key), (name, key)
        for key in ("is_linear",):
            assert str(INI.sub(ref, "model").get(key, "false")) == str(INI.sub(mine, "model").get(key, "false")), (name, key)
    # the Gray-Scott case is the documented example's model (bumps in 3-D added for the 3-D lattice)
    ref = INI.parse_ini(open(f"{REF}/doc/docusaurus/static/ini/next/grey_scott.ini").read())
    mine = INI.parse_ini(K.GRAY_SCOTT)
    fr, fm = norm(INI.sub(INI.sub(ref, "model"), "scalar_field")), norm(INI.sub(INI.sub(mine, "model"), "scalar_field"))
    for sp in ("U", "V"):
        for key in ("storage", "reaction", "cross_diffusion", "compartment"):
            assert fr[sp][key] == fm[sp][key], (sp, key)
    for key in ("F", "k", "D"):
        assert norm(INI.sub(ref, "parser_context"))[key] == norm(INI.sub(mine, "parser_context"))[key]
