#!/usr/bin/env python
"""3-D Gray-Scott lattice, matrix based, BiCGSTAB + SSOR (the reference's default preconditioner): ms per step with the
self-scheduled sweeps and with one launch per level.   python tools/bench_ssor3d.py [cells] [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import dune_copasi_b200 as D  # noqa: E402
from dune_copasi_b200 import workloads as W  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
out = []
for sweep in (False, True):
    text = W.ini_text("grayscott", **{"model.time_step_operator.linear_solver.matrix_free": "false",
                                      "model.time_step_operator.linear_solver.preconditioner.type": "SSOR",
                                      "model.time_step_operator.linear_solver.b200.sor_sweep": "true" if sweep else "false"})
    cfg = D.Config(text)
    model = D.Model(cfg, 3)
    grid = D.Grid.structured(3, [cells] * 3)
    grid.bind(model)
    op = D.Operator(model, grid)
    st = D.Stepper(op, cfg)
    st.set_state(grid.interpolate(model, 0.0), 0.0)
    assert st.step(1.0)
    s0 = st.stats()
    t0 = time.perf_counter()
    for _ in range(steps):
        assert st.step(1.0)
    u, _ = st.get_state()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    s1 = st.stats()
    print(f"grayscott {cells}^3 ({op.ndofs} dofs) BiCGSTAB+SSOR sor_sweep={sweep}: {ms:.1f} ms/step, "
          f"{(s1['kernel_launches'] - s0['kernel_launches']) / steps:.0f} launches/step, "
          f"{(s1['linear_half_iterations'] - s0['linear_half_iterations']) / steps:.1f} half iterations/step", flush=True)
    out.append(u)
print("identical fields:", bool(np.array_equal(out[0], out[1])), "max abs difference", float(np.abs(out[0] - out[1]).max()))
