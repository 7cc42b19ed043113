import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import cases as K
import dune_copasi_b200 as D
from test_gpu_parity import make, rel
case, om, cfg, model, grid, op = make("grayscott2d")
x = K.rand_state(om.ndofs, 10)
t, wM, wA = 0.0, 1.0, 1.0
for tol in (1e-6, 1e-8, 1e-10, 1e-12):
    lcfg = D.Config("type = BiCGSTAB\npreconditioner.type = Jacobi\n")
    solver = D.Solver(op, lcfg)
    solver.linearize(t, wM, wA, x)
    b = K.rand_state(om.ndofs, 11, -1.0, 1.0)
    z, res = solver.solve(b, tol)
    S = K.ORC.StepOperator(om)
    vals = S._stage_jacobian(x, t, wM, wA)
    zo, ro = K.ORC.linear_solve(S.rowptr, S.colidx, vals, b, {"type": "BiCGSTAB", "preconditioner": {"type": "Jacobi"}}, tol)
    print(tol, 'gpu', res.converged, res.iterations, res.half_iterations, res.reduction, res.defect0, '| oracle', ro.converged, ro.iterations_x2, ro.reduction, ro.norm0, 'rel', rel(z, zo))
