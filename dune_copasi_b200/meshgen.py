"""Synthetic unstructured tetrahedral meshes for the multi-compartment workload (BASELINE.json configs[4],
SURVEY.md 8d "S-MC": three nested compartments -- nucleus inside cytosol inside an outer shell -- with
curved, conforming interfaces).

The reference reads such meshes from Gmsh files (`grid.path`, dune/copasi/grid/make_multi_domain_grid.hh:
40-75) and marks the compartments through the cell datum `gmsh_id`
(`compartments.<name>.expression = (gmsh_id == k)`); those files are git-LFS pointers in the reference
tree and gmsh is not in the image, so the mesh comes from this generator, in the same array form
(coordinates, connectivity, one cell datum) the C ABI takes (`dcb_grid_create`).

Construction: the cube [-1, 1]^3 as an n^3 lattice of Kuhn-split cells, vertices jittered by a seeded
random offset (tangentially only on the interfaces and on the outer boundary), then bent by the smooth map
p -> p * ((1 - a) + a g(p)), g_x = sqrt(1 - y^2/2 - z^2/2 + y^2 z^2/3) (cyclic; a = 1 would send the cube
onto the unit ball, a = 0.7 keeps the elements well shaped along the cube's edges): the nested cube shells
|p|_inf = const become nested closed curved surfaces, the compartment interfaces, and the mesh is conforming
across them.  (Projecting the shells onto exact spheres was tried first: every cell on a cube edge then
owns a simplex with all four vertices on one sphere, volume O(h^4) -- slivers that wreck the conditioning.)
Vertices and elements are then renumbered (Morton order of their positions by default, or a seeded random
permutation), so that nothing of the lattice survives in memory: the kernels see coordinates, connectivity
and cell data only.
"""
from __future__ import annotations

import itertools

import numpy as np


def _morton3(q):
    """interleave the bits of three 21-bit integer columns"""
    def spread(x):
        x = x.astype(np.uint64) & np.uint64(0x1FFFFF)
        x = (x | (x << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
        x = (x | (x << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
        x = (x | (x << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
        x = (x | (x << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
        x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
        return x
    return spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))


def nested_compartments(n: int, radii=(0.4, 0.8), seed: int = 12345, jitter: float = 0.2, order: str = "morton",
                        roundness: float = 0.7):
    """-> coords [nv, 3] float64, elems [ne, 4] int32, cell_keys ["gmsh_id"], cell_data [1, ne] float64.

    n      cells per axis of the underlying lattice (even); ne = 6 n^3, nv = (n + 1)^3
    radii  parametric size of the two interfaces (rounded to lattice shells): gmsh_id = 1 inside the first
           (nucleus), 2 between them (cytosol), 3 outside the second (shell / extracellular space)
    order  "morton" | "random" | "lattice": numbering of vertices and elements
    """
    if n < 4 or n % 2:
        raise ValueError("n must be even and >= 4")
    half = n // 2
    lev = [min(max(int(round(r * half)), 1), half - 1) for r in radii]
    if not lev[0] < lev[1]:
        raise ValueError("interfaces collapse on this lattice: increase n or separate the radii")
    rng = np.random.default_rng(seed)
    ax = np.arange(n + 1, dtype=np.int64)
    I, J, K = np.meshgrid(ax, ax, ax, indexing="ij")
    idx = np.stack([I.ravel(), J.ravel(), K.ravel()], 1)            # lattice index of every vertex, z fastest
    p = (idx - half) / half                                         # parametric position in [-1, 1]^3
    shell = np.abs(idx - half).max(1)
    h = 1.0 / half
    dp = rng.uniform(-jitter * h, jitter * h, p.shape)
    on_surface = (shell == lev[0]) | (shell == lev[1]) | (shell == half)
    at_max = np.abs(idx - half) == shell[:, None]
    dp[on_surface[:, None] & at_max] = 0.0                          # interfaces / boundary: tangential jitter only
    dp[shell == 0] = 0.0
    p = p + dp
    x2, y2, z2 = p[:, 0] ** 2, p[:, 1] ** 2, p[:, 2] ** 2
    g = np.stack([np.sqrt(1 - y2 / 2 - z2 / 2 + y2 * z2 / 3), np.sqrt(1 - z2 / 2 - x2 / 2 + z2 * x2 / 3),
                  np.sqrt(1 - x2 / 2 - y2 / 2 + x2 * y2 / 3)], 1)
    coords = p * ((1.0 - roundness) + roundness * g)
    # Kuhn split: from the lowest corner of a cell step the axes in the order of the permutation
    ci = np.arange(n, dtype=np.int64)
    CI, CJ, CK = np.meshgrid(ci, ci, ci, indexing="ij")
    cell = np.stack([CI.ravel(), CJ.ravel(), CK.ravel()], 1)
    s = np.array([(n + 1) * (n + 1), n + 1, 1], dtype=np.int64)
    base = cell @ s
    tets = []
    for perm in itertools.permutations(range(3)):
        v = [base]
        for a in perm:
            v.append(v[-1] + s[a])
        tets.append(np.stack(v, 1))
    elems = np.concatenate(tets, 0)
    cshell = np.floor(np.abs(cell + 0.5 - half).max(1)).astype(np.int64)   # lattice shell of the cell
    cid = np.where(cshell < lev[0], 1.0, np.where(cshell < lev[1], 2.0, 3.0))
    gmsh_id = np.tile(cid, 6)
    # orientation / validity
    X = coords[elems]
    vol = np.einsum("ij,ij->i", np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), X[:, 3] - X[:, 0]) / 6.0
    if not (np.abs(vol) > 1e-3 * h ** 3).all():      # (the Kuhn simplices alternate in orientation; |det| is what counts)
        raise RuntimeError("degenerate tetrahedra: reduce the jitter")
    # renumber
    nv, ne = coords.shape[0], elems.shape[0]
    if order == "lattice":
        vperm, eperm = np.arange(nv), np.arange(ne)
    elif order == "random":
        vperm, eperm = rng.permutation(nv), rng.permutation(ne)
    elif order == "morton":
        q = np.clip(((coords + 1.0) * 0.5 * 2097151.0), 0, 2097151).astype(np.uint64)
        vperm = np.argsort(_morton3(q), kind="stable")
        cen = X.mean(1)
        qe = np.clip(((cen + 1.0) * 0.5 * 2097151.0), 0, 2097151).astype(np.uint64)
        eperm = np.argsort(_morton3(qe), kind="stable")
    else:
        raise ValueError("order must be morton, random or lattice")
    new_id = np.empty(nv, dtype=np.int64)
    new_id[vperm] = np.arange(nv)
    coords = np.ascontiguousarray(coords[vperm])
    elems = np.ascontiguousarray(new_id[elems[eperm]].astype(np.int32))
    cell_data = np.ascontiguousarray(gmsh_id[eperm][None, :])
    return coords, elems, ["gmsh_id"], cell_data


def mesh_stats(coords, elems, cell_data):
    X = coords[elems]
    vol = np.abs(np.einsum("ij,ij->i", np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), X[:, 3] - X[:, 0])) / 6.0
    ids = cell_data[0]
    return {"vertices": int(coords.shape[0]), "tets": int(elems.shape[0]),
            "volume": float(vol.sum()), "min_volume": float(vol.min()),
            "volume_by_id": {int(k): float(vol[ids == k].sum()) for k in np.unique(ids)}}
