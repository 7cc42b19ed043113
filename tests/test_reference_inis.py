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
# parser_context types outside the hot path (TIFF images): not built, must fail loudly
UNSUPPORTED = {"time_snap.ini": "tiff"}

pytestmark = pytest.mark.skipif(not FILES, reason="reference tree not mounted")


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_reference_ini(path):
    import dune_copasi_b200 as D
    name = os.path.basename(path)
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
