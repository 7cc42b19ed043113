#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/mgpu_check.py [case] [nsteps] [matrix_free]

Every rank partitions the global mesh (dcb_grid_partition), steps its part with halo updates and
all-reduces through the library's NCCL communicator; the owned values are gathered on rank 0 and
compared with the serial oracle (<= 1e-10 relative L2)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cases as K  # noqa: E402
import dune_copasi_b200 as D  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "grayscott3d"
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    mf = (sys.argv[3] if len(sys.argv) > 3 else "1") == "1"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    case = K.ALL_CASES[name]
    over = {"model.time_step_operator.linear_solver.matrix_free": "true" if mf else "false"}
    for kv in filter(None, (sys.argv[4] if len(sys.argv) > 4 else "").split(",")):   # extra ini overrides
        k, v = kv.split("=")
        over[k] = v
    gmesh = case.mesh_fn()
    cfg = D.Config(case.ini_with(**over))
    model = D.Model(cfg, case.dim, gmesh.cell_keys)
    gglob = K.product_grid(case, gmesh)     # structured cases -> slab partition + structured kernels
    grid = gglob.partition(rank, world)
    grid.bind(model)
    op = D.Operator(model, grid)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.tensor(list(D.Comm.unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    comm = D.Comm(bytes(uid.cpu().tolist()), rank, world, op)
    st = D.Stepper(op, cfg, comm)
    st.set_state(grid.interpolate(model, case.t0), case.t0)
    for _ in range(nsteps):
        assert st.step(case.dt), "step failed"
    u, t = st.get_state()
    # [model.reduce] over the partition: every rank obtains the global values
    red = D.Reducer(op, cfg, comm)
    red_vals = red.apply_dev(st.time, st.state_dev(), raise_on_error=False) if red.keys else {}
    # gather (global vertex id, compartment-major dof values) of the owned dofs on rank 0
    om = case.oracle(**over) if rank == 0 else None
    gids = grid.global_vertex_ids()
    ranges = op.owned_ranges()
    ob, oe = grid.owned_vertex_range()
    payload = [gids, (ob, oe), [u[b:e] for b, e in ranges], grid.elem_compartment(), grid.elements()]
    gathered = [None] * world
    dist.gather_object(payload, gathered if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        S = K.ORC.StepOperator(om)
        uref = om.initial(case.t0)
        tt = case.t0
        for _ in range(nsteps):
            uref, good = S.apply(uref, tt, case.dt)
            assert good
            tt += case.dt
        got = np.full(om.ndofs, np.nan)
        mg = om.mesh
        for r, (gid_all, (ob, oe), vals, ecomp, elems) in enumerate(gathered):
            for c in range(om.ncomp):
                ns = om.comp_nspec[c]
                if ns == 0:
                    continue
                # local vertices of compartment c that are owned, ascending local id == ascending gid
                lv = np.unique(elems[ecomp == c])
                lv = lv[(lv >= ob) & (lv < oe)]
                gv = gid_all[lv]
                pos = np.searchsorted(mg.comp_vertices[c], gv)
                for s in range(ns):
                    got[mg.comp_offset[c] + pos * ns + s] = vals[c][s::ns]
        assert not np.isnan(got).any(), "some dofs were not owned by any rank"
        err = np.linalg.norm(got - uref) / np.linalg.norm(uref)
        print(f"mgpu_check {name} world={world} steps={nsteps} matrix_free={mf} peer_memory={comm.uses_peer_memory}: rel L2 err {err:.3e}")
        ok = err <= 1e-10
        if red_vals:
            ref_vals, _ = K.ORC.reduce(om, uref, tt)
            for key, v in ref_vals.items():
                dv = abs(red_vals[key] - v) / max(abs(v), 1e-300)
                print(f"  reduce {key}: {red_vals[key]:.12g} (serial oracle {v:.12g}, rel diff {dv:.2e})")
                ok = ok and dv <= 1e-9
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    if flag.item() != 1:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
