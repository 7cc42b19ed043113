"""The synthetic tetrahedral mesh of BASELINE configs[4] (dune_copasi_b200/meshgen.py): conforming, three
nested closed compartments, well shaped, deterministic; bound identically by product and oracle; RCB
partitions of it keep the owner-computes invariants."""
import numpy as np
import pytest

import cases as K
from dune_copasi_b200 import meshgen as G


def _faces(elems):
    f = np.concatenate([elems[:, [1, 2, 3]], elems[:, [0, 2, 3]], elems[:, [0, 1, 3]], elems[:, [0, 1, 2]]], 0)
    owner = np.tile(np.arange(elems.shape[0]), 4)
    f = np.sort(f, 1)
    order = np.lexsort((f[:, 2], f[:, 1], f[:, 0]))
    return f[order], owner[order]


@pytest.mark.parametrize("order", ["morton", "random", "lattice"])
def test_mesh_is_conforming_and_nested(order):
    n = 10
    coords, elems, keys, data = G.nested_compartments(n, order=order)
    assert coords.shape == ((n + 1) ** 3, 3) and elems.shape == (6 * n ** 3, 4) and keys == ["gmsh_id"]
    st = G.mesh_stats(coords, elems, data)
    h = 2.0 / n
    assert st["min_volume"] > 5e-3 * h ** 3                    # no slivers
    assert set(st["volume_by_id"]) == {1, 2, 3}
    f, owner = _faces(elems)
    same = (f[1:] == f[:-1]).all(1)
    # every face belongs to one (boundary) or two (interior) tetrahedra -- never more: conforming
    triple = same[1:] & same[:-1]
    assert not triple.any()
    ids = data[0]
    a, b = owner[:-1][same], owner[1:][same]
    pairs = {tuple(sorted((int(ids[i]), int(ids[j])))) for i, j in zip(a, b) if ids[i] != ids[j]}
    assert pairs == {(1, 2), (2, 3)}                           # nucleus touches cytosol only, cytosol the shell
    single = np.ones(len(f), dtype=bool)
    single[1:] &= ~same
    single[:-1] &= ~same
    assert set(ids[owner[single]].astype(int)) == {3}          # the outer boundary belongs to the shell


def test_mesh_is_deterministic_and_orders_are_permutations():
    a = G.nested_compartments(8, seed=12345)
    b = G.nested_compartments(8, seed=12345)
    assert all(np.array_equal(x, y) for x, y in zip((a[0], a[1], a[3]), (b[0], b[1], b[3])))
    c = G.nested_compartments(8, seed=12345, order="random")
    assert np.allclose(np.sort(a[0].sum(1)), np.sort(c[0].sum(1)))
    va = np.sort(np.abs(np.linalg.det(a[0][a[1]][:, 1:] - a[0][a[1]][:, :1])))
    vc = np.sort(np.abs(np.linalg.det(c[0][c[1]][:, 1:] - c[0][c[1]][:, :1])))
    assert np.allclose(va, vc)


def test_product_and_oracle_bind_the_mesh_alike():
    import dune_copasi_b200 as D
    case = K.CASES["cell10_nested"]
    om = case.oracle()
    cfg, model, grid = K.product_objects(case)
    assert grid.ndofs == om.ndofs and np.array_equal(grid.elem_dof(), om.mesh.elem_dof)
    rp, ci = om.pattern()
    prp, pci = grid.pattern(model)
    assert np.array_equal(rp, prp) and np.array_equal(ci, pci)
    assert [n for n, _ in model.species()] == om.names and len(om.names) == 10
    assert D.lib() is not None
