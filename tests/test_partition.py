"""Multi-GPU host logic on CPU: vertex-range partition, ghost layer, halo plan and owned ranges.

The N>1 data path is: every rank assembles on its local mesh (owned
vertices + one ghost layer), ghost values are refreshed owner->ghost before every operator
application, reductions run over owned dofs and are all-reduced.  Here the per-rank operator is
played by the oracle on the *local* mesh the product's partitioner produced, the exchange runs over
torch.distributed's gloo backend (world size 2 and 3, 127.0.0.1), and the result must match the
serial oracle on owned rows -- bit-exact maps, rounding-level values."""
import os
import socket

import numpy as np
import pytest

import cases as K


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _local_problem(case, rank, size, structured=False, method="auto"):
    import dune_copasi_b200 as D
    from oracle import core as ORC, ini as INI, mesh as OMESH
    gmesh = case.mesh_fn()
    cfg = D.Config(case.ini)
    model = D.Model(cfg, case.dim, gmesh.cell_keys)
    if structured:
        gglob = K.product_grid(case, gmesh)      # slab partition of the structured box
    else:
        gglob = D.Grid.from_arrays(case.dim, gmesh.coords, gmesh.elems, gmesh.cell_keys, gmesh.cell_data)
    gloc = gglob.partition(rank, size, method)
    gloc.bind(model)
    # the oracle on the same local arrays (cell data restricted through the element ids is not
    # needed for the cases used here)
    lmesh = OMESH.Mesh(dim=case.dim, coords=gloc.coords(), elems=gloc.elements())
    if case.element == "cube":      # a slab of a Q1 lattice is a Q1 lattice
        lmesh.etype = 1
        lmesh.lattice = tuple(int(np.unique(lmesh.coords[:, a]).size) for a in range(case.dim))
    if gmesh.cell_keys:
        lmesh.cell_keys = list(gmesh.cell_keys)
        lmesh.cell_data = np.ascontiguousarray(gmesh.cell_data[:, gloc.global_element_ids()])
    om_loc = ORC.Model(INI.parse_ini(case.ini), lmesh)
    return model, gloc, om_loc


def _global_dofs(om_glob, om_loc, gids):
    """local dof -> global dof through (compartment, global vertex, species)"""
    mg, ml = om_glob.mesh, om_loc.mesh
    out = -np.ones(om_loc.ndofs, dtype=np.int64)
    for c in range(om_glob.ncomp):
        ns = om_glob.comp_nspec[c]
        if ns == 0:
            continue
        gv = mg.comp_vertices[c]
        pos = -np.ones(mg.nv, dtype=np.int64)
        pos[gv] = np.arange(gv.size)
        lv = ml.comp_vertices[c]
        gl = pos[gids[lv]]
        assert (gl >= 0).all()
        for s in range(ns):
            out[ml.comp_offset[c] + np.arange(lv.size) * ns + s] = mg.comp_offset[c] + gl * ns + s
    assert (out >= 0).all()
    return out


def _worker(rank, size, port, name, structured, q, method="auto"):
    import torch.distributed as dist
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=size)
        import torch
        case = K.ALL_CASES[name]
        om = case.oracle()                      # serial reference
        model, gloc, oml = _local_problem(case, rank, size, structured, method)
        gids = gloc.global_vertex_ids()
        owner = gloc.vertex_owner()
        ob, oe = gloc.owned_vertex_range()
        assert oe - ob == gloc.n_owned
        is_owned = np.zeros(gloc.nv, dtype=bool)
        is_owned[ob:oe] = True
        assert (owner[is_owned] == rank).all() and (owner[~is_owned] != rank).all()
        # owned first then ghosts, each group ascending in global id; contiguous ranges for slab / range
        assert np.all(np.diff(gids[ob:oe]) >= 1) and len(np.unique(gids)) == gids.size
        assert np.all(np.diff(gids[:ob]) >= 1) and np.all(np.diff(gids[oe:]) >= 1)
        if structured or method == "range":
            assert np.all(np.diff(gids[ob:oe]) == 1)
        assert gloc.ndofs == oml.ndofs and np.array_equal(gloc.elem_dof(), oml.mesh.elem_dof)
        l2g = _global_dofs(om, oml, gids)
        # owned dof ranges: contiguous per compartment, and exactly the dofs on owned vertices
        owned_mask = np.zeros(oml.ndofs, dtype=bool)
        for c in range(oml.ncomp):
            ns = oml.comp_nspec[c]
            lv = oml.mesh.comp_vertices[c]
            for s in range(ns):
                owned_mask[oml.mesh.comp_offset[c] + np.arange(lv.size) * ns + s] = is_owned[lv]
        x_glob = K.rand_state(om.ndofs, 42)
        x = x_glob[l2g].copy()
        # ---- halo update over gloo following the product's plan
        x[~owned_mask] = np.nan
        plan = gloc.halo_plan(rank)
        reqs, recv_bufs = [], []
        for peer, send, recv in plan:
            sb = torch.from_numpy(x[send].copy())
            rb = torch.empty(recv.size, dtype=torch.float64)
            recv_bufs.append((recv, rb))
            reqs.append(dist.isend(sb, peer))
            reqs.append(dist.irecv(rb, peer))
        for r in reqs:
            r.wait()
        for recv, rb in recv_bufs:
            x[recv] = rb.numpy()
        assert not np.isnan(x).any(), "halo plan does not cover every ghost dof"
        assert np.array_equal(x, x_glob[l2g]), "ghost values differ from their owners'"
        # ---- owner-computes residual equals the serial residual on owned rows
        t = case.t0 + 0.1
        r_loc = np.zeros(oml.ndofs)
        oml.residual(1, t, 1.0, x, r_loc)
        oml.residual(0, t, 0.5, x, r_loc)
        r_ser = np.zeros(om.ndofs)
        om.residual(1, t, 1.0, x_glob, r_ser)
        om.residual(0, t, 0.5, x_glob, r_ser)
        err = np.abs(r_loc[owned_mask] - r_ser[l2g][owned_mask]).max() / np.abs(r_ser).max()
        assert err < 1e-13, err
        # ---- reductions over owned dofs + all-reduce equal the serial dot product
        part = torch.tensor([float(x[owned_mask] @ x[owned_mask]), float(owned_mask.sum())], dtype=torch.float64)
        dist.all_reduce(part)
        assert abs(part[0].item() - float(x_glob @ x_glob)) < 1e-9 * float(x_glob @ x_glob)
        assert int(part[1].item()) == om.ndofs            # every dof owned exactly once
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
        raise


@pytest.mark.parametrize("name,size,structured", [("grayscott3d", 2, False), ("cell3d", 2, False), ("grayscott2d", 3, False),
                                                  ("two_disks", 2, False), ("grayscott3d", 3, True), ("cell3d", 2, True),
                                                  ("grayscott3d_q1", 2, True), ("grayscott2d_q1", 3, True)])
def test_partition_halo_gloo(name, size, structured):
    _spawn(name, size, structured, "auto")


@pytest.mark.parametrize("name,size,method", [("cell3d", 3, "rcb"), ("two_disks", 4, "rcb"), ("grayscott3d", 2, "range"),
                                              ("cell3d", 3, "range"), ("two_disks_cell_data", 2, "rcb"), ("cell10_nested", 4, "rcb")])
def test_partition_methods_gloo(name, size, method):
    """Recursive coordinate bisection (any rank count) and contiguous vertex ranges on general meshes: the
    same owner-computes checks as above (bit-exact maps against the serial numbering, halo plan complete,
    owned rows equal the serial residual)."""
    _spawn(name, size, False, method)


def test_rcb_is_balanced_and_compact():
    """RCB: parts differ by at most one vertex per bisection level and every part is a box-shaped cloud (its
    bounding boxes overlap only at the cuts)."""
    import dune_copasi_b200 as D
    case = K.CASES["cell3d"]
    gmesh = case.mesh_fn()
    g = D.Grid.from_arrays(case.dim, gmesh.coords, gmesh.elems, gmesh.cell_keys, gmesh.cell_data)
    size = 5
    counts, boxes, seen = [], [], np.zeros(gmesh.coords.shape[0], dtype=int)
    for r in range(size):
        loc = g.partition(r, size, "rcb")
        ob, oe = loc.owned_vertex_range()
        gids = loc.global_vertex_ids()[ob:oe]
        seen[gids] += 1
        counts.append(gids.size)
        pts = gmesh.coords[gids]
        boxes.append((pts.min(0), pts.max(0)))
    assert (seen == 1).all()
    assert max(counts) - min(counts) <= 3
    vol = sum(np.prod(hi - lo) for lo, hi in boxes)
    allv = np.prod(gmesh.coords.max(0) - gmesh.coords.min(0))
    assert vol <= 1.0 * allv + 1e-12        # the boxes tile the domain (they do not overlap beyond the cut planes)


def _spawn(name, size, structured, method):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, size, port, name, structured, q, method)) for r in range(size)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"


@pytest.mark.parametrize("structured", [False, True])
def test_owned_ranges_are_contiguous_per_compartment(structured):
    case = K.CASES["cell3d"]
    total = 0
    for rank in range(3):
        model, gloc, oml = _local_problem(case, rank, 3, structured)
        ob, oe = gloc.owned_vertex_range()
        # Grid::owned_ranges: the owned dofs of a compartment are one contiguous range
        for c in range(oml.ncomp):
            lv = oml.mesh.comp_vertices[c]
            owned = (lv >= ob) & (lv < oe)
            idx = np.nonzero(owned)[0]
            if idx.size:
                assert idx[-1] - idx[0] + 1 == idx.size
            total += int(owned.sum()) * oml.comp_nspec[c]
    assert total == case.oracle().ndofs


@pytest.mark.parametrize("name", ["poisson_q1", "poisson"])
def test_dirichlet_constraints_on_slabs(name):
    """Slab partitions of a structured lattice: the cut planes are not boundary.  Every rank's constraint
    set equals the serial one restricted to its local vertices (owned and ghost), values included --
    for the Kuhn split (facet search) and for Q1 cells (lattice indices of the global box)."""
    import dune_copasi_b200 as D
    case = K.ALL_CASES[name]
    cfg = D.Config(case.ini)
    model = D.Model(cfg, case.dim, [])
    gglob = K.product_grid(case)
    gser = K.product_grid(case)
    gser.bind(model)
    d, v = gser.constraints(model)
    serial = dict(zip(d.tolist(), v.tolist()))           # single species: dof == global vertex id
    assert len(serial) > 0
    size = 3
    seen = set()
    for rank in range(size):
        gloc = gglob.partition(rank, size)
        gloc.bind(model)
        gids = gloc.global_vertex_ids()
        dl, vl = gloc.constraints(model)
        local = dict(zip(gids[dl].tolist(), vl.tolist()))
        expect = {g: serial[g] for g in gids.tolist() if g in serial}
        assert local == expect, (name, rank, len(local), len(expect))
        ob, oe = gloc.owned_vertex_range()
        seen.update(g for g in gids[ob:oe].tolist() if g in serial)
    assert seen == set(serial)
