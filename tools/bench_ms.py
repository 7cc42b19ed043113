#!/usr/bin/env python
"""test/mitchell_schaefer.ini as written (RestartedGMRes + SSOR on the assembled matrix, the reference's default
preconditioner) on the 2^level x 2^level simplex lattice: ms per time step and launches per step, with the
self-scheduled SSOR sweeps (one launch per sweep) and with one launch per level.

    python tools/bench_ms.py [level] [steps]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import cases as K  # noqa: E402
import dune_copasi_b200 as D  # noqa: E402


def run(level, steps, sweep):
    n = 1 << level
    over = {"model.time_step_operator.linear_solver.type": "RestartedGMRes",
            "model.time_step_operator.linear_solver.preconditioner.type": "SSOR",
            "model.time_step_operator.linear_solver.matrix_free": "false",
            "model.time_step_operator.linear_solver.b200.sor_sweep": "true" if sweep else "false"}
    for kv in filter(None, os.environ.get("MS_SET", "").split(",")):
        k, v = kv.split("=")
        over[k] = v
    case = K.CASES["mitchell_schaefer"]
    cfg = D.Config(case.ini_with(**over))
    model = D.Model(cfg, 2)
    grid = D.Grid.structured(2, [n, n], [0.0, 0.0], [1.0, 1.0])
    grid.bind(model)
    op = D.Operator(model, grid)
    st = D.Stepper(op, cfg)
    st.set_state(grid.interpolate(model, 0.0), 0.0)
    dt = float(os.environ.get('MS_DT', '10.0'))
    for k in range(2):
        if not st.step(dt):
            print(f"level {level} sor_sweep={sweep}: warm-up step {k} failed", st.stats(), flush=True)
            return None
    s0 = st.stats()
    t0 = time.perf_counter()
    for k in range(steps):
        if not st.step(dt):
            print(f"level {level} sor_sweep={sweep}: step {k} failed", st.stats(), flush=True)
            return None
    u, t = st.get_state()       # synchronises
    ms = (time.perf_counter() - t0) * 1e3 / steps
    s1 = st.stats()
    d = {k: s1[k] - s0[k] for k in s1}
    print(f"mitchell_schaefer level {level} ({op.ndofs} dofs) GMRES+SSOR sor_sweep={sweep}: {ms:.2f} ms/step, "
          f"{d['kernel_launches'] / steps:.0f} launches/step, {d['linear_iterations'] / steps:.1f} Krylov iterations/step, "
          f"u in [{u[0::2].min():.4f}, {u[0::2].max():.4f}]", flush=True)
    return u


if __name__ == "__main__":
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    b = run(level, steps, False)
    a = run(level, steps, True)
    import numpy as np
    if a is not None and b is not None:
        print("identical fields:", bool(np.array_equal(a, b)), "max abs difference", float(np.abs(a - b).max()))
