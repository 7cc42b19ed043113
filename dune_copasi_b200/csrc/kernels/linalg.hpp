// Launchers of the statically compiled sm_100a kernels (linalg.cu): SpMV, fused BLAS-1 sweeps of
// BiCGSTAB / CG, Jacobi and block-Jacobi, Dirichlet masks.  All fp64, all HBM-bound.
// They replace the dune-istl vector/matrix arithmetic the reference drives through
// dune/copasi/solver/istl/factory/{iterative,preconditioner}.hh and block_jacobi.hh:46-128.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "peer.hpp"

namespace dcb {
namespace la {

// Deterministic grid reductions: every reducing kernel runs on a fixed grid and writes its result
// through a "last block sums the partials in order" epilogue.
struct ReduceWorkspace {
  double* partials = nullptr;   // [max_blocks * 4]
  unsigned* counter = nullptr;  // self-resetting ticket
  int max_blocks = 0;
};
constexpr int kMaxBlocks = 148 * 8;
// dofs that enter global reductions (multi-GPU: the owned dofs; one range per compartment)
struct Ranges {
  int n = 0;
  long long b[8], e[8];
  static Ranges all(long long len) { Ranges r; r.n = 1; r.b[0] = 0; r.e[0] = len; return r; }
};
// What a sweep does beyond its own arithmetic (all off by default):
//   peer       the collectives that follow it in the algorithm, fused into the kernel (peer.hpp);
//   host_out   reducing kernels: the final sums also go to this mapped pinned host address (no D2H copy);
//   zero_input the sweep clears the operator-result vector it has just consumed (v in bicg_p_prec, t in
//              bicg_final), so the next matrix-free application can accumulate without a fill launch.
struct Link {
  peer::Link peer;
  double* host_out = nullptr;
  int zero_input = 0;
};
void reduce_workspace_create(ReduceWorkspace* w, int max_blocks = kMaxBlocks);
void reduce_workspace_destroy(ReduceWorkspace* w);

// y = A x (CSR, sorted columns); rowptr32 != null selects 32-bit row pointers
void spmv_csr(int64_t nrows, const int64_t* rowptr, const int32_t* rowptr32, const int32_t* colidx,
              const double* vals, const double* x, double* y, int avg_nnz, cudaStream_t s);

// out[0] = <a,b>            (out is a device pointer)
void dot(const Ranges& own, const double* a, const double* b, double* out, const ReduceWorkspace& w, cudaStream_t s,
         const Link& L = Link());
// out[0] = <a,b>, out[1] = <c,d>
void dot2(const Ranges& own, const double* a, const double* b, const double* c, const double* d, double* out,
          const ReduceWorkspace& w, cudaStream_t s, const Link& L = Link());

// BiCGSTAB: p = r + beta (p - omega v)      (first = true: p = r)
void bicg_update_p(int64_t n, double* p, const double* r, const double* v, double beta, double omega,
                   bool first, cudaStream_t s);
// x += alpha y ; r -= alpha v ; out[0] = <r,r> ; out[1] = <rt,r>  (rt may be null -> out[1] = 0)
void axpy_pair_norm(int64_t n, const Ranges& own, double alpha, const double* y, double* x, const double* v, double* r,
                    const double* rt, double* out, const ReduceWorkspace& w, cudaStream_t s);
// fused BiCGSTAB sweeps; dinv == null skips the folded Jacobi application.
// Every step length is formed on the device from the (all-reduced) device-resident sums, so the
// host can enqueue the next sweep without waiting for a scalar:
//   rho = <rt,r> at the start of the iteration, hptr = <rt,v>, trtt = (<t,r>, <t,t>):
//   alpha = rho / h, omega = tr / tt, beta = (rho_new / rho) * (alpha / omega)  (dune-istl's order)
// p = r + beta (p - omega v) ; y = relax dinv p      (rho, hptr, trtt: those of the previous iteration)
void bicg_p_prec(int64_t n, double* p, const double* r, double* v, const double* rho_new, const double* rho,
                 const double* hptr, const double* trtt, bool first, const double* dinv, double relax, double* y,
                 const ReduceWorkspace& w, cudaStream_t s, const Link& L = Link());
// r -= alpha v ; out[0] = <r,r> ; y2 = relax dinv r
void bicg_r_prec(int64_t n, const Ranges& own, const double* rho, const double* hptr, const double* v, double* r,
                 const double* dinv, double relax, double* y2, double* out, const ReduceWorkspace& w, cudaStream_t s,
                 const Link& L = Link());
// xout = xin + alpha y1 + omega y2 ; r -= omega t ; out[0] = <r,r> ; out[1] = <rt,r>
void bicg_final(int64_t n, const Ranges& own, const double* rho, const double* hptr, const double* trtt, const double* y1,
                const double* y2, const double* xin, double* xout, double* t, double* r, const double* rt,
                double* out, const ReduceWorkspace& w, cudaStream_t s, const Link& L = Link());
// GMRES, one modified Gram-Schmidt step per pass: w -= (*coef) vprev (coef null: skipped), out[0] = <vnext, w>
// over `own` (vnext null: <w, w>); coef is the device-resident result of the previous step
void mgs_step(int64_t n, const Ranges& own, const double* coef, const double* vprev, double* w, const double* vnext,
              double* out, const ReduceWorkspace& ws, cudaStream_t s, const Link& L = Link());
// GMRES (modified Gram-Schmidt) with device-resident coefficients:
// y += sign * (*coef) * x
void axpy_dev(int64_t n, const double* coef, double sign, const double* x, double* y, cudaStream_t s);
// dst = src / sqrt(*norm2)
void normalize_dev(int64_t n, const double* src, const double* norm2, double* dst, cudaStream_t s);
// x *= a
void scale(int64_t n, double a, double* x, cudaStream_t s);
// y = a - b
void sub(int64_t n, const double* a, const double* b, double* y, cudaStream_t s);
// CG: p = q + beta p
void xpby(int64_t n, double* p, const double* q, double beta, cudaStream_t s);
// y += a x
void axpy(int64_t n, double a, const double* x, double* y, cudaStream_t s);
void copy(int64_t n, const double* x, double* y, cudaStream_t s);
void fill(int64_t n, double v, double* y, cudaStream_t s);

// preconditioners
void jacobi_apply(int64_t n, const double* dinv, double relax, const double* d, double* v, cudaStream_t s);
// One level of a level-scheduled SOR / Gauss-Seidel sweep (dune-istl bsorf / bsorb / dbgs on scalar
// entries): for every row i of the level  v_i += relax (d_i - sum_j a_ij v_j) / a_ii with the diagonal term
// in the sum (bsorf / bsorb), or, with skip_diag (dbgs), v_i = (d_i - sum_{j != i} a_ij v_j) / a_ii -- dbgs
// relaxes after the sweep: relax_blend.  Rows of one level are not coupled (structurally
// symmetric pattern), so the result is the sequential sweep's, whatever the order inside the level.
void sor_level(const int32_t* rows, int64_t count, const int64_t* rowptr, const int32_t* colidx, const double* vals,
               const double* d, double* v, double relax, bool skip_diag, cudaStream_t s);
// The same sweep as nlev calls of sor_level, self-scheduled in one launch: slots = the rows in level order, every
// level padded with -1 to a multiple of 32; dep_ptr / dep_idx = pattern of A + A^T; done[row] = epoch once a row is
// finished (epochs increase from sweep to sweep; `done` starts at 0, epochs at 1).  Bit-identical to the level loop.
void sor_sweep(const int32_t* slots, int64_t nslots, bool backward, const int64_t* rowptr, const int32_t* colidx,
               const double* vals, const int64_t* dep_ptr, const int32_t* dep_idx, const double* d, double* v, double relax,
               bool skip_diag, int* done, int epoch, cudaStream_t s);
// x = w x + (1 - w) xold   (the relaxation step that closes a dbgs sweep)
void relax_blend(int64_t n, double w, const double* xold, double* x, cudaStream_t s);
void csr_extract_diag_inv(int64_t n, const int64_t* rowptr, const int32_t* colidx, const double* vals,
                          double* dinv, cudaStream_t s);
// block diagonal of node blocks of size bs over dofs [dof0, dof0 + nblocks*bs): bdiag[dof*bs + j]
void csr_extract_block_diag(int64_t dof0, int64_t nblocks, int bs, const int64_t* rowptr,
                            const int32_t* colidx, const double* vals, double* bdiag, cudaStream_t s);
// in-place inversion of nblocks dense bs x bs blocks (Gauss-Jordan with partial pivoting,
// = FieldMatrix::invert used by DenseInverse, dense_inverse.hh:9-37), bs <= 19
void block_invert(int64_t nblocks, int bs, double* blocks, cudaStream_t s);
void block_jacobi_apply(int64_t dof0, int64_t nblocks, int bs, const double* binv, double relax,
                        const double* d, double* v, cudaStream_t s);
// scalar diagonal from the block diagonal: dinv[dof0 + b*bs + i] = 1 / bdiag[(dof0 + b*bs)*bs + i*bs + i]
void block_diag_to_dinv(int64_t dof0, int64_t nblocks, int bs, const double* bdiag, double* dinv, cudaStream_t s);

// d[i] = mask[i] ? 1 : 1/d[i]   (mask may be null)
void invert_diag(int64_t n, double* d, const unsigned char* mask, cudaStream_t s);

// Dirichlet handling
void set_values(int64_t n, const int32_t* idx, const double* vals, double* x, cudaStream_t s);   // x[idx] = vals
void zero_values(int64_t n, const int32_t* idx, double* x, cudaStream_t s);                      // x[idx] = 0
void copy_values(int64_t n, const int32_t* idx, const double* src, double* dst, cudaStream_t s); // dst[idx] = src[idx]
// rows and columns of constrained dofs -> identity (mask is per dof)
void csr_constrain(int64_t nrows, const int64_t* rowptr, const int32_t* colidx, double* vals,
                   const unsigned char* mask, cudaStream_t s);
// block diagonal rows/cols of constrained dofs -> identity
void bdiag_constrain(int64_t dof0, int64_t nblocks, int bs, double* bdiag, const unsigned char* mask, cudaStream_t s);

// Second pass of the tile-marching assembly drivers (kernels/assembly_tile.cuh): every vertex that
// lies on a tile edge or a chunk boundary plane ("cut") gets its partial sums added in slot order
// (slot 0 = y itself), its share of the reductions is added, and the last block forms the final sums:
// the partials of the main launch in block order, then those of this launch in block order.
struct TileFixup {
  int n[3];                 // cells per axis; 2-D lattices are passed as (nx, 0, ny)
  int tile[3];              // tile extents along x, y and the chunk length along the marching axis
  int ns;
  long long dof_offset;
  int own_lo, own_hi;       // vertex planes of the marching axis that enter the reductions
  int epi;                  // as DcTileArgs::epi
  int accumulate_unused;
  double* y;
  const double* slots;
  long long slot_stride;
  const double* w;          // epi 1: <w, y>
  const double* aux;        // epi 2: <y, aux>, |y|^2
  const unsigned char* cmask;   // identity rows: y = zraw there
  const double* zraw;
  const double* main_partials;  // [nmain][4]
  int nmain;
  double* out;              // out[q] = sum of partial q for every q with (out_mask >> q) & 1
  int out_mask;
};
void tile_fixup(const TileFixup& f, const ReduceWorkspace& w, cudaStream_t s);
// x += (rho / h) * relax * dinv * p     (BiCGSTAB stops after its first half step)
void bicg_x_half(int64_t n, const double* rho, const double* hptr, const double* dinv, double relax, const double* p,
                 double* x, cudaStream_t s);
// xout = (xin + alpha relax dinv p) + omega relax dinv r ; rout = r - omega t ; out[0] = <rout,rout> ; out[1] = <rt,rout>
void bicg_final_fold(int64_t n, const Ranges& own, const double* rho, const double* hptr, const double* trtt,
                     const double* dinv, double relax, const double* p, const double* r, const double* xin, double* xout,
                     double* t, double* rout, const double* rt, double* out, const ReduceWorkspace& w, cudaStream_t s,
                     const Link& L = Link());
// halo exchange helpers
void gather(int64_t n, const int32_t* idx, const double* x, double* buf, cudaStream_t s);   // buf[i] = x[idx[i]]
void scatter(int64_t n, const int32_t* idx, const double* buf, double* x, cudaStream_t s);  // x[idx[i]] = buf[i]

}  // namespace la
}  // namespace dcb
