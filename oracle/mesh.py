"""ORACLE (test infrastructure) -- meshes, compartments, facets and the P1 DOF map in numpy.

Numbering rules (shared, by specification, with the product's C++ generator
dune_copasi_b200/csrc/mesh.cpp; tests compare the two bit-for-bit):

* structured simplex grid (reference: grid/make_multi_domain_grid.hh:76-100 builds the same
  *geometry* with StructuredGridFactory::createSimplexGrid + UG global refinement; UG's vertex
  order is not reproducible without UG -- SURVEY.md App. C.4 -- so the numbering is our own):
  vertices lexicographic, x fastest; cubes lexicographic, x fastest; each cube is split into the
  dim! Kuhn simplices obtained by walking from the cube's lowest corner along the axes in the
  order given by the lexicographically enumerated permutations of (0..dim-1).
* compartments: cell c belongs to compartment k iff compartments.<k>.expression evaluated at the
  cell centre (with the cell data) is != 0 (make_multi_domain_grid.hh:118-155); ids in ini order.
* DOF map (model_single_compartment_traits.hh:26-27,
  model_multi_compartment_traits.hh:22): compartments concatenated (Lexicographic); inside a
  compartment vertex-major / species-minor (EntityGrouping), vertices of the sub-domain in
  ascending global vertex id.
* facets: every element face once; interface facets (different compartments on both sides) and
  boundary facets, inside element = lower element index, sorted by (inside element, local face).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Mesh:
    dim: int
    coords: np.ndarray            # [nv, dim] float64
    elems: np.ndarray             # [ne, dim+1] int32
    cell_keys: list = field(default_factory=list)
    cell_data: np.ndarray | None = None   # [nkeys, ne]
    elem_comp: np.ndarray | None = None   # [ne] int32
    # facets
    f_in: np.ndarray | None = None
    f_out: np.ndarray | None = None
    f_lin: np.ndarray | None = None
    f_lout: np.ndarray | None = None
    # dof map
    comp_offset: np.ndarray | None = None     # [ncomp+1] int64
    comp_vertices: list | None = None         # per compartment sorted vertex ids
    elem_dof: np.ndarray | None = None        # [ne, dim+1] int64
    ndofs: int = 0
    etype: int = 0                            # 0 simplices, 1 axis-aligned Q1 cubes (non-reference extension)
    lattice: tuple | None = None              # vertices per axis of a structured grid

    @property
    def nv(self):
        return self.coords.shape[0]

    @property
    def ne(self):
        return self.elems.shape[0]

    def centers(self):
        return self.coords[self.elems].mean(axis=1)


def structured(dim: int, cells, origin=None, extent=None, element: str = "simplex") -> Mesh:
    """element = "simplex": Kuhn split (the reference's createSimplexGrid geometry);
    element = "cube": the lattice cells themselves as Q1 elements, corner m of a cell at the bit
    pattern of m (x = bit 0) -- BASELINE configs[3]'s "Q1", not a reference capability."""
    cells = [int(c) for c in cells][:dim]
    origin = np.zeros(dim) if origin is None else np.asarray(origin, float)[:dim]
    extent = np.ones(dim) if extent is None else np.asarray(extent, float)[:dim]
    nvs = [c + 1 for c in cells]
    # x fastest: vertex id = i + nvx*(j + nvy*k)
    g = np.meshgrid(*[np.arange(n) for n in reversed(nvs)], indexing="ij")
    vi = [a.ravel() for a in reversed(g)]                  # vi[0] = i (x index) ...
    coords = np.stack([origin[a] + extent[a] * (vi[a] / cells[a]) for a in range(dim)], axis=1)
    stride = [1]
    for a in range(1, dim):
        stride.append(stride[-1] * nvs[a - 1])
    g = np.meshgrid(*[np.arange(n) for n in reversed(cells)], indexing="ij")
    ci = [a.ravel().astype(np.int64) for a in reversed(g)]
    base = sum(ci[a] * stride[a] for a in range(dim))
    if element == "cube":
        el = np.stack([base + sum(((m >> a) & 1) * stride[a] for a in range(dim)) for m in range(1 << dim)], axis=1)
        return Mesh(dim=dim, coords=np.ascontiguousarray(coords), elems=np.ascontiguousarray(el.astype(np.int32)),
                    etype=1, lattice=tuple(nvs))
    if element != "simplex":
        raise ValueError("element must be 'simplex' or 'cube'")
    perms = list(itertools.permutations(range(dim)))
    el = np.empty((base.size, len(perms), dim + 1), dtype=np.int64)
    for p, perm in enumerate(perms):
        cur = base.copy()
        el[:, p, 0] = cur
        for k, ax in enumerate(perm):
            cur = cur + stride[ax]
            el[:, p, k + 1] = cur
    elems = el.reshape(-1, dim + 1).astype(np.int32)
    return Mesh(dim=dim, coords=np.ascontiguousarray(coords), elems=np.ascontiguousarray(elems), lattice=tuple(nvs))


def two_disks(nr_inner: int, nr_outer: int, ntheta: int) -> Mesh:
    """Conforming triangulation of the disk r<1 (gmsh_id 2) inside the annulus 1<r<2 (gmsh_id 1):
    stand-in for test/data/grids/two_disks.msh, a git-LFS pointer in the reference."""
    radii = np.concatenate([np.linspace(0, 1, nr_inner + 1)[1:], np.linspace(1, 2, nr_outer + 1)[1:]])
    th = 2 * np.pi * np.arange(ntheta) / ntheta
    pts = [np.zeros((1, 2))]
    for r in radii:
        pts.append(np.stack([r * np.cos(th), r * np.sin(th)], axis=1))
    coords = np.concatenate(pts)
    ring = lambda k: 1 + k * ntheta  # noqa: E731
    tris, ids = [], []
    for j in range(ntheta):
        jn = (j + 1) % ntheta
        tris.append((0, ring(0) + j, ring(0) + jn))
        ids.append(2)
    for k in range(len(radii) - 1):
        gid = 2 if k + 1 < nr_inner else 1
        for j in range(ntheta):
            jn = (j + 1) % ntheta
            a, b, c, d = ring(k) + j, ring(k) + jn, ring(k + 1) + j, ring(k + 1) + jn
            if (j + k) % 2 == 0:
                tris += [(a, c, d), (a, d, b)]
            else:
                tris += [(a, c, b), (b, c, d)]
            ids += [gid, gid]
    m = Mesh(dim=2, coords=coords, elems=np.asarray(tris, dtype=np.int32))
    m.cell_keys = ["gmsh_id"]
    m.cell_data = np.asarray(ids, dtype=np.float64)[None, :]
    return m


def build_facets(m: Mesh):
    """Interface + boundary facets (see module docstring)."""
    if m.etype == 1:
        # Q1 cubes: single-compartment lattices only, no facet terms; the boundary vertices (for
        # Dirichlet constraints) are the lattice's outer vertices
        if np.any(m.elem_comp < 0) or np.unique(m.elem_comp).size != 1:
            raise NotImplementedError("Q1 cube grids carry one compartment over the whole lattice")
        idx = np.unravel_index(np.arange(m.nv), tuple(reversed(m.lattice)))
        onb = np.zeros(m.nv, dtype=bool)
        for a, n in zip(idx, reversed(m.lattice)):
            onb |= (a == 0) | (a == n - 1)
        m.boundary_vertices = np.nonzero(onb)[0]
        m.f_in = np.zeros(0, dtype=np.int64); m.f_out = np.zeros(0, dtype=np.int64)
        m.f_lin = np.zeros(0, dtype=np.int32); m.f_lout = np.zeros(0, dtype=np.int32)
        return m
    dim, nd, ne = m.dim, m.dim + 1, m.ne
    # face opposite to local vertex a = all other vertices
    faces = np.empty((ne, nd, dim), dtype=np.int64)
    for a in range(nd):
        others = [b for b in range(nd) if b != a]
        faces[:, a, :] = np.sort(m.elems[:, others].astype(np.int64), axis=1)
    flat = faces.reshape(-1, dim)
    eidx = np.repeat(np.arange(ne, dtype=np.int64), nd)
    lidx = np.tile(np.arange(nd, dtype=np.int32), ne)
    order = np.lexsort(tuple(flat[:, k] for k in reversed(range(dim))))
    fs, es, ls = flat[order], eidx[order], lidx[order]
    same = np.all(fs[1:] == fs[:-1], axis=1)
    first = np.concatenate([[True], ~same])
    paired_with_next = np.concatenate([same, [False]])
    comp = m.elem_comp
    # interior pairs
    i0 = np.nonzero(first & paired_with_next)[0]
    ea, eb, la, lb = es[i0], es[i0 + 1], ls[i0], ls[i0 + 1]
    swap = ea > eb
    ea, eb = np.where(swap, eb, ea), np.where(swap, ea, eb)
    la, lb = np.where(swap, lb, la), np.where(swap, la, lb)
    keep = comp[ea] != comp[eb]
    f_in, f_out, f_lin, f_lout = ea[keep], eb[keep], la[keep], lb[keep]
    # boundary singles
    b0 = np.nonzero(first & ~paired_with_next)[0]
    bk = comp[es[b0]] >= 0
    f_in = np.concatenate([f_in, es[b0][bk]])
    f_out = np.concatenate([f_out, -np.ones(bk.sum(), dtype=np.int64)])
    f_lin = np.concatenate([f_lin, ls[b0][bk]])
    f_lout = np.concatenate([f_lout, -np.ones(bk.sum(), dtype=np.int32)])
    o = np.lexsort((f_lin, f_in))
    m.f_in, m.f_out = np.ascontiguousarray(f_in[o]), np.ascontiguousarray(f_out[o])
    m.f_lin = np.ascontiguousarray(f_lin[o].astype(np.int32))
    m.f_lout = np.ascontiguousarray(f_lout[o].astype(np.int32))
    # all boundary vertices of the mesh (for Dirichlet constraints: boundary_entity_mapper.hh:26-58)
    bv = np.unique(fs[b0].ravel())
    m.boundary_vertices = bv
    # first boundary facet (in facet order) touching each boundary vertex -> its normal is used
    return m


def build_dofmap(m: Mesh, comp_nspec):
    ncomp = len(comp_nspec)
    m.comp_vertices, offs = [], [0]
    m.elem_dof = -np.ones((m.ne, m.elems.shape[1]), dtype=np.int64)
    for c in range(ncomp):
        sel = m.elem_comp == c
        verts = np.unique(m.elems[sel])
        m.comp_vertices.append(verts)
        lv = -np.ones(m.nv, dtype=np.int64)
        lv[verts] = np.arange(verts.size)
        m.elem_dof[sel] = offs[-1] + lv[m.elems[sel]] * comp_nspec[c]
        offs.append(offs[-1] + verts.size * comp_nspec[c])
    m.comp_offset = np.asarray(offs, dtype=np.int64)
    m.ndofs = int(offs[-1])
    return m
