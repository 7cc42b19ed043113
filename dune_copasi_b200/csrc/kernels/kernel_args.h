// Argument blocks shared by the host (nvcc/g++) and the NVRTC-compiled assembly kernels.
// Plain types only; identical layout on both sides.
#ifndef DCB_KERNEL_ARGS_H
#define DCB_KERNEL_ARGS_H

struct DcVolArgs {
  const double* coords;        // [nv][DIM]
  const int* elems;            // [ne][DIM+1]
  const int* elem_ids;         // elements of this compartment (null: 0..n-1)
  const int* vdof;             // vertex -> dof of species 0 in this compartment (null: dof_offset + v*NS)
  const double* cell;          // [nkeys][ne_total]
  long long ne_total;
  long long n;                 // elements in this launch
  int dof_offset;
  double time, wM, wA;
  const double* x;             // linearisation point / coefficients
  const double* z;             // direction for Jacobian-apply
  double* r;                   // residual / apply output (accumulated)
  const long long* rowptr;     // CSR of the Jacobian
  const int* colidx;
  double* vals;
  double* bdiag;               // block diagonal, NS x NS per vertex: bdiag[dof0*NS + i*NS + j]
  const unsigned char* cmask;  // per-dof Dirichlet mask (null: none); masked z entries act as 0
  // gather form of the CSR fill: vertices of this compartment and the elements around them
  const int* verts;            // vertex ids (null: 0..n-1)
  const int* vptr;             // [n+1]
  const int* vel;              // (element << 2 | local vertex index)
  int gather_maxlen;           // longest row of the compartment (shared-memory slots per species row)
  // Packed connectivity of the compartment (3-D, model.assembly.b200.packed_conn): thread position t reads its four
  // vertex ids and its four dof bases with two 16-byte loads instead of walking elem_ids[t] -> elems[e] -> vdof[v]
  // -> x[dof] (four dependent loads; the kernels wait on `long_scoreboard` at 4 warps per scheduler)
  const int* pverts;           // [n][4] vertex ids (null: walk the mesh arrays)
  const int* pdofs;            // [n][4] dof of species 0 at those vertices
  // 16-byte gathers: the element kernels are bound by the L1 gather wavefronts, not by DRAM or the fp64 pipe
  const double* coords4;       // 3-D: [nv][4] coordinates padded to 32 bytes per vertex (null: use coords)
  int vec;                     // 1: x, z are 16-byte aligned and every dof block starts at an even offset (NS even)
};

struct DcPatchArgs {
  const double* coords;
  const int* patch_node_ptr;   // [npatch+1]
  const int* patch_nodes;      // vertex ids of each patch, ascending
  const int* patch_elem_ptr;   // [npatch+1]
  const unsigned short* lconn; // [ne][4] patch-local vertex indices (4th unused in 2-D)
  const unsigned short* adj;   // per patch node: (local element << 2 | local vertex) list
  const int* adj_ptr;          // [total patch nodes + 1]
  const int* vdof;
  const double* cell;          // [nkeys][ne] in patch element order
  long long ne_total;
  int npatch;
  int dof_offset;
  int max_nodes, max_elems;    // patch budgets (shared-memory layout)
  double time, wM, wA;
  const double* x;
  const double* z;
  double* r;
  double* bdiag;
  const unsigned char* cmask;
};

struct DcStructArgs {
  int n[3];                    // cells per axis of the (local) box
  double h[3], origin[3];
  double rh[3];                // 1 / h (the host divides once)
  double vol;                  // volume of one Kuhn simplex, prod(h) / dim! (an fp64 division: ~6 % of the apply kernel's
                               // issue slots when every thread redid it, profiles/r02_struct_apply_256_xpair_ncu.txt)
  long long cell_begin;        // this launch covers the cells [cell_begin, ncells)
  long long ncells;
  int march;                   // marching kernels: cells a thread walks along the last axis
  int dof_offset;
  double time, wM, wA;
  const double* x;
  const double* z;
  double* r;
  double* bdiag;
  const unsigned char* cmask;
  // Jacobian apply of the per-cell driver, scaled instantiation: direction = zscale .* z with zscale = relax * D^-1
  // (the Jacobi application of the Krylov solve formed while the corners are loaded, instead of a vector written
  // and re-read)
  const double* zscale;
  double zrelax;               // unused (the relaxation factor is folded into zscale by the host)
  const long long* rowptr;     // CSR of the Jacobian (Q1 fill only)
  const int* colidx;
  double* vals;
};

// Tile-marching drivers (kernels/assembly_tile.cuh): the structured kernels as an owner-computes
// sweep.  A CTA owns a tile of the lattice's first DIM-1 axes and walks `lz` cell layers up the last
// axis; vertex data of a plane is staged in shared memory, every result entry is written once with
// a plain store.  Vertices shared with a neighbouring tile / chunk ("cut" vertices) get their partial
// sums in `slots` and are finished by la::tile_fixup.
struct DcTileArgs {
  DcStructArgs s;              // lattice, weights; s.x = linearisation point, s.z = direction, s.r = result
  int lz;                      // cell layers per chunk along the marching (last) axis
  int ntx, nty;                // tiles along x and y (nty = 1 in 2-D)
  int own_lo, own_hi;          // vertex planes [own_lo, own_hi) of the marching axis enter the reductions
  int pro;                     // 0: direction = s.z;  1: p_out = r_in + beta (p_in - omega v_in), direction = relax dinv p_out
                               // 2: r_out = r_in - alpha v_in, direction = relax dinv r_out, partial 0 = |r_out|^2
  int epi;                     // 0: none;  1: partial 1 = <w, result>;  2: partial 1 = <result, r_out>, partial 2 = |result|^2
  int first;                   // pro 1: p_out = r_in (first BiCGSTAB iteration)
  int accumulate;              // result += instead of result =
  int identity;                // apply: rows of Dirichlet-constrained dofs are identity rows (result = s.z there)
  double relax;
  const double* r_in;
  const double* p_in;
  const double* v_in;
  const double* dinv;
  const double* w;
  double* r_out;
  double* p_out;
  const double* rho_new;       // device scalars of the BiCGSTAB recurrences (kernels/linalg.cu)
  const double* rho;
  const double* hptr;
  const double* trtt;
  double* slots;               // slots[(k - 1) * slot_stride + dof], k = 1..7: partial sums of cut vertices
  long long slot_stride;
  double* partials;            // [gridDim.x][4] reduction partials of this launch
};

struct DcFacetArgs {
  const double* coords;
  const int* elems;
  const long long* f_self;     // element on the side that owns the residual rows
  const long long* f_other;    // element on the other side (-1 on the boundary)
  const int* f_lself;          // local index of the vertex opposite to the facet
  const int* f_lother;
  const int* vdof_s;           // vertex -> dof maps of the two compartments (null: offset + v*NS)
  const int* vdof_t;
  const double* cell;
  long long ne_total;
  long long n;
  int dof_offset_s, dof_offset_t;
  int block_offset;            // first block of this list inside a fused launch over all pairs
  double time, wA;
  const double* x;
  const double* z;
  double* r;
  const long long* rowptr;
  const int* colidx;
  double* vals;
  double* bdiag;
  const unsigned char* cmask;
};

struct DcReduceArgs {
  const double* coords;
  const int* elems;
  const int* elem_ids;         // elements of this launch (one compartment, or the cells outside all)
  const int* vdof;             // vertex -> dof of species 0 of the compartment (null: offset + v*NS)
  const double* cell;
  long long ne_total;
  long long n;
  int dof_offset;
  double time;
  const double* x;
  const double* init;          // [DC_NRED] start value of every functional (its neutral element)
  double* partials;            // [gridDim.x][DC_NRED]
};

#endif
