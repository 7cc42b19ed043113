"""Level scheduling of the SOR-family sweeps (csrc/solver.cpp LinearSolver::build_levels): the level
construction restated in numpy, applied level by level with every row of a level updated from the same
snapshot (what concurrent threads may see), equals the sequential ascending / descending sweeps of
dune-istl's bsorf / bsorb -- also on a structurally non-symmetric pattern (mitchell_schaefer: u reads z,
z does not read u).  CPU only; the CUDA path is checked against the oracle's sequential sweeps in
tests/test_gpu_parity.py::test_sor_family_matches_oracle."""
import numpy as np
import pytest

import cases as K


def levels(rp, ci):
    n = rp.size - 1
    level = np.zeros(n, dtype=np.int64)
    for i in range(n):
        cols = ci[rp[i]:rp[i + 1]]
        lower = cols[cols < i]
        if lower.size:
            level[i] = max(level[i], level[lower].max() + 1)
        upper = cols[cols > i]
        level[upper] = np.maximum(level[upper], level[i] + 1)
    return level


def sweep_sequential(rp, ci, vals, d, v, w, backward):
    n = d.size
    for i in (range(n - 1, -1, -1) if backward else range(n)):
        sl = slice(rp[i], rp[i + 1])
        diag = vals[sl][ci[sl] == i][0]
        v[i] += w * (d[i] - vals[sl] @ v[ci[sl]]) / diag


def sweep_levels(rp, ci, vals, d, v, w, level, backward):
    order = np.unique(level)
    for l in (order[::-1] if backward else order):
        rows = np.nonzero(level == l)[0]
        snap = v.copy()                       # every row of the level reads the same state
        for i in rows:
            sl = slice(rp[i], rp[i + 1])
            diag = vals[sl][ci[sl] == i][0]
            v[i] = snap[i] + w * (d[i] - vals[sl] @ snap[ci[sl]]) / diag


@pytest.mark.parametrize("name", ["mitchell_schaefer", "grayscott3d", "two_disks"])
def test_level_sweeps_equal_sequential_sweeps(name):
    case = K.CASES[name]
    om = case.oracle()
    S = K.ORC.StepOperator(om)
    x = K.rand_state(om.ndofs, 1)
    vals = S._stage_jacobian(x, case.t0, 1.0, 0.5 * case.dt)
    rp, ci = S.rowptr, S.colidx
    lev = levels(rp, ci)
    # no two rows of a level are coupled, in either direction
    rows = np.repeat(np.arange(om.ndofs), np.diff(rp))
    off = rows != ci
    assert np.all(lev[rows[off]] != lev[ci[off]])
    assert np.all((lev[rows[off]] < lev[ci[off]]) == (rows[off] < ci[off]))
    d = K.rand_state(om.ndofs, 2, -1.0, 1.0)
    a, b = np.zeros(om.ndofs), np.zeros(om.ndofs)
    for backward in (False, True, False):
        sweep_sequential(rp, ci, vals, d, a, 0.9, backward)
        sweep_levels(rp, ci, vals, d, b, 0.9, lev, backward)
        assert np.array_equal(a, b)
