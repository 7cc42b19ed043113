/* ORACLE -- test infrastructure, NOT product code.  Plain C (+OpenMP) restatement of the CPU
 * algorithm of DuneCopasi's CG-P1 hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.
 *
 * PARITY STATUS: the reference cannot be compiled here (needs ~10 DUNE modules incl. an unpinned
 * dune-pdelab branch, .ci/setup_dune:55-74), so this restatement is pinned only by the reference's
 * own coarse known-answer system tests (test/gauss.ini:53-55, test/exp.ini:33-35,
 * test/poisson.ini:29-32, test/two_disks.ini:13-19,57-59) and by element-level identities
 * (tests/test_oracle_kat.py).  Per-entry residual/Jacobian parity against the real reference is
 * "parity unpinned".
 *
 * What each function follows (paths relative to /root/reference):
 *   orc_eval                 byte-code walk ~ src/dune/copasi/parser/mu.cc:214-219 (muParser VM);
 *                            ExprTk/SymEngine evaluate the same grammar as ASTs.
 *   orc_residual_volume      dune/copasi/model/diffusion_reaction/local_operator.hh:417-491
 *   orc_jacobian_volume      local_operator.hh:541-707 (analytic), :713-765 (numerical, FD)
 *                            sink = CSR (matrix based) or PseudoJacobian :90-108 (matrix free apply)
 *   orc_residual_skeleton    local_operator.hh:777-971, :1398-1415
 *   orc_jacobian_skeleton    local_operator.hh:973-1199, :1354-1396, :1417-1452
 *   quadrature / basis       dune-geometry simplex rules of order 2, dune-localfunctions P1
 *                            (third party, values as listed in SURVEY.md App. A.3)
 *   etype 1 (Q1 cubes)       NOT a reference element (PkLocalFiniteElementMap is simplex-only,
 *                            model_single_compartment_traits.hh:23-24): the same loops with the
 *                            multilinear basis and the 2-point Gauss rule per axis, for BASELINE
 *                            configs[3]'s "Q1"; pinned by closed-form element matrices and the
 *                            gauss / poisson KATs (tests/test_q1_oracle.py), parity unpinned by
 *                            construction (nothing in the reference to compare with)
 *   orc_bicgstab / orc_cg    dune-istl BiCGSTABSolver / CGSolver operation order (third party,
 *                            SURVEY.md App. C.1), built by dune/copasi/solver/istl/factory/iterative.hh:36-73
 *   preconditioners          dune-istl SeqJac; dune/copasi/solver/istl/block_jacobi.hh:46-128
 *
 * Conscious deviation (documented in DESIGN.md): for a species that lives only on the *other*
 * side of an interface facet the reference pairs the other element's coefficients with this
 * element's shape functions by local index (local_operator.hh:903-916), which is only correct when
 * both elements number the shared vertices identically.  The oracle (and the product) evaluate the
 * geometrically intended value: the P1 trace of the other side's field on the facet.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int32_t dim;
  int64_t nv;
  const double* coords;      /* [nv*dim] */
  int64_t ne;
  const int32_t* elems;      /* [ne*(dim+1)] */
  const int32_t* elem_comp;  /* [ne] compartment id or -1 */
  int32_t nkeys;
  const double* cell_data;   /* [nkeys*ne] */
  const int64_t* elem_dof;   /* [ne*(dim+1)] dof of species 0 at local vertex a (own compartment) */
  /* interface + boundary facets */
  int64_t nf;
  const int64_t* f_in;       /* inside element */
  const int64_t* f_out;      /* outside element or -1 */
  const int32_t* f_lin;      /* local index of the vertex opposite to the facet in f_in */
  const int32_t* f_lout;
  int32_t etype;             /* 0: simplices (the reference's element type), 1: axis-aligned Q1 cubes
                                (extension without a reference counterpart) */
} OrcMesh;

/* term kinds; (i, j, k) per kind:
 *   K_DIFF (i, wrt j, -)            scalar D_ij            K_DIFF_JAC (i, wrt j, k)          dD_ij/du_k
 *   K_DIFF_T (i, wrt j, 3r+c)       tensor entry D_ij[r][c] K_DIFF_T_JAC (i, wrt j, 9k+3r+c)  its derivative
 *   K_VEL (i, axis, -)              velocity component      K_VEL_JAC (i, wrt k, axis)        d vel/du_k   */
enum { K_REACTION = 0, K_REACTION_JAC, K_STORAGE, K_STORAGE_JAC, K_DIFF, K_DIFF_JAC, K_OUTFLOW,
       K_OUTFLOW_JAC, K_VEL, K_VEL_JAC, K_DIFF_T, K_DIFF_T_JAC, K_NKIND };

typedef struct {
  int32_t ncomp, nspec;
  const int32_t* spec_comp;     /* [nspec] */
  const int32_t* spec_local;    /* [nspec] index inside its compartment */
  const int32_t* comp_ptr;      /* [ncomp+1] */
  const int32_t* comp_spec;     /* species ids grouped by compartment */
  int32_t nterms;
  const int32_t* terms;         /* [nterms*5] kind,i,j,k,prog ; sorted by (i,kind) */
  const int32_t* tptr;          /* [nspec*K_NKIND+1] */
  const int32_t* prog_ptr;      /* [nprog+1] in ints */
  const int32_t* code;
  const int32_t* const_ptr;     /* [nprog] */
  const double* consts;
  int32_t nslots, spec_base;
} OrcModel;

enum { SLOT_TIME = 0, SLOT_INTFAC, SLOT_ENTVOL, SLOT_INVOL, SLOT_INBND, SLOT_INSKEL, SLOT_POS = 6,
       SLOT_NORMAL = 9, SLOT_CELL = 12 };

/* ---------------------------------------------------------------- expression VM */
static double f1(int id, double a) {
  switch (id) {
    case 0: return sqrt(a); case 1: return exp(a); case 2: return log(a); case 3: return sin(a);
    case 4: return cos(a); case 5: return tan(a); case 6: return fabs(a); case 7: return floor(a);
    case 8: return ceil(a); case 9: return tanh(a); case 10: return sinh(a); case 11: return cosh(a);
    case 12: return asin(a); case 13: return acos(a); case 14: return atan(a); case 15: return log10(a);
    case 16: return log2(a); case 17: return (a > 0) - (a < 0); case 18: return exp2(a);
    case 19: return round(a);
  }
  return NAN;
}
static double f2(int id, double a, double b) {
  switch (id) {
    case 0: return a < b ? a : b; case 1: return a > b ? a : b; case 2: return atan2(a, b);
    case 3: return pow(a, b);
  }
  return NAN;
}

/* std::lerp as libstdc++ implements it (context.cc:93 calls std::lerp) */
static double lerp_std(double a, double b, double t) {
  if ((a <= 0 && b >= 0) || (a >= 0 && b <= 0)) return t * b + (1 - t) * a;
  if (t == 1) return b;
  const double x = a + t * (b - a);
  return (t > 1) == (b > a) ? (b < x ? x : b) : (x < b ? x : b);
}
/* tabulated context function; rec = kind, samples, clamp, domain..., range... (expr.py Table.record)
   kind 0: context.cc:83-95 (lower_bound + lerp); kind 1: context.cc:256-277, literally (returns sample k) */
static double tab_eval(const double* rec, double x) {
  const int kind = (int)rec[0], n = (int)rec[1], clamp = (int)rec[2];
  if (kind == 0) {
    const double *dom = rec + 3, *rng = rec + 3 + n;
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (dom[mid] < x) lo = mid + 1; else hi = mid; }
    if (lo == 0) return rng[0];
    if (lo == n) return rng[n - 1];
    return lerp_std(rng[lo - 1], rng[lo], (x - dom[lo - 1]) / (dom[lo] - dom[lo - 1]));
  }
  const double d0 = rec[3], d1 = rec[4], *rng = rec + 5;
  const int m = n - 1;   /* intervals */
  if (clamp) x = d0 < x ? x : d0;
  else if (x < d0 || x > d1 || x != x) return NAN;   /* the reference throws */
  double whole;
  modf((x - d0) * ((double)m / (d1 - d0)), &whole);
  const int k = whole <= 0.0 ? 0 : (whole >= (double)m ? m : (int)whole);
  return rng[k];
}

/* image function (type = tiff); rec = 2, rows, cols, x_res, y_res, x_off, y_off, values[rows][cols] (oracle/tiff.py).
   TIFFGrayscale::operator(), src/dune/copasi/common/tiff_grayscale.cc:91-105: float arithmetic, uint32 truncation
   (the row index wraps as in the reference), clamped to the image; negative offsets (undefined in the reference) -> 0 */
static uint32_t tiff_pixel(float v) { return v <= 0.0f ? 0u : (v >= 4294967040.0f ? 4294967295u : (uint32_t)v); }
static double tab2_eval(const double* rec, double x, double y) {
  const uint32_t rows = (uint32_t)rec[1], cols = (uint32_t)rec[2];
  const float x_res = (float)rec[3], y_res = (float)rec[4], x_off = (float)rec[5], y_off = (float)rec[6];
  uint32_t px = tiff_pixel(x_res * ((float)x - x_off));
  uint32_t line = rows - tiff_pixel(y_res * ((float)y - y_off)) - 1u;
  if (px > cols - 1) px = cols - 1;
  if (line > rows - 1) line = rows - 1;
  return rec[7 + (size_t)line * cols + px];
}

double orc_eval(const int32_t* code, int n, const double* consts, const double* ctx) {
  double st[64];
  int sp = 0;
  for (int pc = 0; pc < n; pc += 2) {
    int op = code[pc], arg = code[pc + 1];
    switch (op) {
      case 0: st[sp++] = consts[arg]; break;
      case 1: st[sp++] = ctx[arg]; break;
      case 2: sp--; st[sp - 1] += st[sp]; break;
      case 3: sp--; st[sp - 1] -= st[sp]; break;
      case 4: sp--; st[sp - 1] *= st[sp]; break;
      case 5: sp--; st[sp - 1] /= st[sp]; break;
      case 6: sp--; st[sp - 1] = pow(st[sp - 1], st[sp]); break;
      case 7: st[sp - 1] = -st[sp - 1]; break;
      case 8: sp--; st[sp - 1] = st[sp - 1] < st[sp]; break;
      case 9: sp--; st[sp - 1] = st[sp - 1] > st[sp]; break;
      case 10: sp--; st[sp - 1] = st[sp - 1] <= st[sp]; break;
      case 11: sp--; st[sp - 1] = st[sp - 1] >= st[sp]; break;
      case 12: sp--; st[sp - 1] = st[sp - 1] == st[sp]; break;
      case 13: sp--; st[sp - 1] = st[sp - 1] != st[sp]; break;
      case 14: sp--; st[sp - 1] = (st[sp - 1] != 0.0) && (st[sp] != 0.0); break;
      case 15: sp--; st[sp - 1] = (st[sp - 1] != 0.0) || (st[sp] != 0.0); break;
      case 16: st[sp - 1] = (st[sp - 1] == 0.0); break;
      case 17: sp -= 2; st[sp - 1] = (st[sp - 1] != 0.0) ? st[sp] : st[sp + 1]; break;
      case 18: st[sp - 1] = f1(arg, st[sp - 1]); break;
      case 19: sp--; st[sp - 1] = f2(arg, st[sp - 1], st[sp]); break;
      case 20: sp--; st[sp - 1] = fmod(st[sp - 1], st[sp]); break;
      case 21: st[sp - 1] = tab_eval(consts + arg, st[sp - 1]); break;
      case 22: sp--; st[sp - 1] = tab2_eval(consts + arg, st[sp - 1], st[sp]); break;
    }
  }
  return st[0];
}

static inline double run(const OrcModel* P, int prog, const double* ctx) {
  return orc_eval(P->code + P->prog_ptr[prog], P->prog_ptr[prog + 1] - P->prog_ptr[prog],
                  P->consts + P->const_ptr[prog], ctx);
}

/* one stand-alone program on many contexts (rows of `nslots` doubles) */
void orc_eval_rows(const int32_t* code, int n, const double* consts, const double* ctxs, int64_t nrows,
                   int64_t nslots, double* out) {
  for (int64_t i = 0; i < nrows; ++i) out[i] = orc_eval(code, n, consts, ctxs + i * nslots);
}

void orc_eval_many(const OrcModel* P, int prog, int64_t n, const double* ctxs, double* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = run(P, prog, ctxs + i * P->nslots);
}

/* ---------------------------------------------------------------- quadrature (order 2) */
#define QA 0.585410196624968515
#define QB 0.138196601125010504
static const double Q2[3][2] = {{4. / 6., 1. / 6.}, {1. / 6., 4. / 6.}, {1. / 6., 1. / 6.}};
static const double Q3[4][3] = {{QA, QB, QB}, {QB, QA, QB}, {QB, QB, QA}, {QB, QB, QB}};
static const double Q1[2] = {0.5 - 0.28867513459481288225, 0.5 + 0.28867513459481288225};

static int quad(int dim, double pts[][3], double* wts) {
  if (dim == 1) { for (int q = 0; q < 2; ++q) { pts[q][0] = Q1[q]; wts[q] = 0.5; } return 2; }
  if (dim == 2) { for (int q = 0; q < 3; ++q) { pts[q][0] = Q2[q][0]; pts[q][1] = Q2[q][1]; wts[q] = 1. / 6.; } return 3; }
  for (int q = 0; q < 4; ++q) { for (int k = 0; k < 3; ++k) pts[q][k] = Q3[q][k]; wts[q] = 1. / 24.; }
  return 4;
}

/* P1 reference basis at xi: phi0 = 1 - sum xi, phi_i = xi_{i-1} */
static void p1(int dim, const double* xi, double* phi) {
  double s = 0;
  for (int k = 0; k < dim; ++k) { phi[k + 1] = xi[k]; s += xi[k]; }
  phi[0] = 1.0 - s;
}

/* geometry of a simplex: corners X[nd][dim]; returns det, fills grads[nd][dim] (global gradients) */
static double simplex_geo(int dim, double (*X)[3], double (*G)[3]) {
  double B[3][3], Bi[3][3], det;
  for (int k = 0; k < dim; ++k) for (int c = 0; c < dim; ++c) B[c][k] = X[k + 1][c] - X[0][c];
  /* x = x0 + B xi ;  grad phi_a = B^{-T} grad_ref phi_a */
  if (dim == 2) {
    det = B[0][0] * B[1][1] - B[0][1] * B[1][0];
    Bi[0][0] = B[1][1] / det; Bi[0][1] = -B[0][1] / det;
    Bi[1][0] = -B[1][0] / det; Bi[1][1] = B[0][0] / det;
  } else {
    double c00 = B[1][1] * B[2][2] - B[1][2] * B[2][1];
    double c01 = B[1][2] * B[2][0] - B[1][0] * B[2][2];
    double c02 = B[1][0] * B[2][1] - B[1][1] * B[2][0];
    det = B[0][0] * c00 + B[0][1] * c01 + B[0][2] * c02;
    Bi[0][0] = c00 / det; Bi[1][0] = c01 / det; Bi[2][0] = c02 / det;
    Bi[0][1] = (B[0][2] * B[2][1] - B[0][1] * B[2][2]) / det;
    Bi[1][1] = (B[0][0] * B[2][2] - B[0][2] * B[2][0]) / det;
    Bi[2][1] = (B[0][1] * B[2][0] - B[0][0] * B[2][1]) / det;
    Bi[0][2] = (B[0][1] * B[1][2] - B[0][2] * B[1][1]) / det;
    Bi[1][2] = (B[0][2] * B[1][0] - B[0][0] * B[1][2]) / det;
    Bi[2][2] = (B[0][0] * B[1][1] - B[0][1] * B[1][0]) / det;
  }
  /* grad phi_{k+1} = row k of B^{-1}; grad phi_0 = -sum */
  for (int c = 0; c < dim; ++c) G[0][c] = 0;
  for (int k = 0; k < dim; ++k)
    for (int c = 0; c < dim; ++c) { G[k + 1][c] = Bi[k][c]; G[0][c] -= Bi[k][c]; }
  return det;
}

/* ---- element type dispatch.  etype 1 = multilinear Q1 basis on an axis-aligned box whose corner m
 * sits at the bit pattern of m (x = bit 0), integrated with the 2-point Gauss rule per axis (what
 * dune-geometry returns for a cube at order 2).  Not a reference capability (PkLocalFiniteElementMap
 * is simplex-only, model_single_compartment_traits.hh:23-24): same weak form, different element. */
#define MAXND 8
static inline int mesh_nd(const OrcMesh* M) { return M->etype ? 1 << M->dim : M->dim + 1; }
static inline double elem_fact(const OrcMesh* M) { return M->etype ? 1.0 : (M->dim == 2 ? 2.0 : 6.0); }

static int elem_quad(const OrcMesh* M, double pts[][3], double* wts) {
  if (!M->etype) return quad(M->dim, pts, wts);
  const int nq = 1 << M->dim;
  for (int q = 0; q < nq; ++q) {
    for (int k = 0; k < 3; ++k) pts[q][k] = k < M->dim ? Q1[(q >> k) & 1] : 0.0;
    wts[q] = 1.0 / nq;
  }
  return nq;
}

static void elem_basis(const OrcMesh* M, const double* xi, double* phi) {
  if (!M->etype) { p1(M->dim, xi, phi); return; }
  for (int m = 0; m < (1 << M->dim); ++m) {
    double v = 1.0;
    for (int k = 0; k < M->dim; ++k) v *= ((m >> k) & 1) ? xi[k] : 1.0 - xi[k];
    phi[m] = v;
  }
}

/* global gradients of the basis at reference point xi (constant on a simplex); returns det */
static double elem_geo(const OrcMesh* M, double (*X)[3], const double* xi, double (*G)[3]) {
  if (!M->etype) return simplex_geo(M->dim, X, G);
  double det = 1.0, h[3] = {1, 1, 1};
  for (int k = 0; k < M->dim; ++k) { h[k] = X[1 << k][k] - X[0][k]; det *= h[k]; }
  for (int m = 0; m < (1 << M->dim); ++m)
    for (int k = 0; k < M->dim; ++k) {
      double v = (((m >> k) & 1) ? 1.0 : -1.0) / h[k];
      for (int l = 0; l < M->dim; ++l)
        if (l != k) v *= ((m >> l) & 1) ? xi[l] : 1.0 - xi[l];
      G[m][k] = v;
    }
  return det;
}

static inline void atomic_add(double* p, double v, int par) {
  if (par) {
#pragma omp atomic
    *p += v;
  } else
    *p += v;
}

/* ---------------------------------------------------------------- sinks for Jacobian entries */
typedef struct {
  int mode;                 /* 0: CSR add, 1: y += v * z[col]  (PseudoJacobian) */
  const int64_t* rowptr; const int32_t* colidx; double* vals;
  const double* z; double* y;
  int par;
} Sink;

static inline void sink_add(const Sink* S, int64_t row, int64_t col, double v) {
  if (S->mode == 1) { atomic_add(&S->y[row], v * S->z[col], S->par); return; }
  int64_t lo = S->rowptr[row], hi = S->rowptr[row + 1] - 1;
  while (lo <= hi) {
    int64_t mid = (lo + hi) >> 1;
    int32_t c = S->colidx[mid];
    if (c == col) { atomic_add(&S->vals[mid], v, S->par); return; }
    if (c < col) lo = mid + 1; else hi = mid - 1;
  }
  abort(); /* entry outside the pattern: pattern builder and Jacobian disagree */
}

#define TERMS(P, g, kind, t0, t1) \
  int t0 = (P)->tptr[(g) * K_NKIND + (kind)], t1 = (P)->tptr[(g) * K_NKIND + (kind) + 1]

static void load_element(const OrcMesh* M, int64_t e, double (*X)[3]) {
  int nd = mesh_nd(M);
  for (int a = 0; a < nd; ++a) {
    int64_t v = M->elems[e * nd + a];
    for (int c = 0; c < 3; ++c) X[a][c] = c < M->dim ? M->coords[v * M->dim + c] : 0.0;
  }
}

static void set_cell(const OrcMesh* M, int64_t e, double* ctx) {
  for (int k = 0; k < M->nkeys; ++k) ctx[SLOT_CELL + k] = M->cell_data[(int64_t)k * M->ne + e];
}

/* values + gradients of all species of compartment c at a point with basis values phi */
static void eval_fields(const OrcMesh* M, const OrcModel* P, int c, int64_t e, const double* x,
                        const double* phi, double (*G)[3], double* ctx) {
  int nd = mesh_nd(M);
  for (int t = P->comp_ptr[c]; t < P->comp_ptr[c + 1]; ++t) {
    int g = P->comp_spec[t], s = P->spec_local[g];
    double val = 0, gr[3] = {0, 0, 0};
    for (int a = 0; a < nd; ++a) {
      double xa = x[M->elem_dof[e * nd + a] + s];
      val += xa * phi[a];
      for (int k = 0; k < M->dim; ++k) gr[k] += xa * G[a][k];
    }
    double* slot = ctx + P->spec_base + 4 * g;
    slot[0] = val; slot[1] = gr[0]; slot[2] = gr[1]; slot[3] = gr[2];
  }
}

/* ---------------------------------------------------------------- volume residual
 * form 0 = stiffness (reaction, diffusion), 1 = mass (storage).  r += w * R_form(x)          */
void orc_residual_volume(const OrcMesh* M, const OrcModel* P, int form, double time, double w,
                         const double* x, double* r, int par) {
  const int dim = M->dim, nd = mesh_nd(M);
  double qp[MAXND][3], qw[MAXND];
  const int nq = elem_quad(M, qp, qw);
  double fact = elem_fact(M);
#pragma omp parallel if (par)
  {
    double* ctx = (double*)calloc(P->nslots, sizeof(double));
#pragma omp for schedule(static)
    for (int64_t e = 0; e < M->ne; ++e) {
      int c = M->elem_comp[e];
      if (c < 0) continue;
      double X[MAXND][3], G[MAXND][3], loc[16][MAXND];
      load_element(M, e, X);
      double det = elem_geo(M, X, qp[0], G);
      ctx[SLOT_TIME] = time; ctx[SLOT_ENTVOL] = fabs(det) / fact; ctx[SLOT_INVOL] = 1;
      set_cell(M, e, ctx);
      int ns = P->comp_ptr[c + 1] - P->comp_ptr[c];
      for (int s = 0; s < ns; ++s) for (int a = 0; a < nd; ++a) loc[s][a] = 0;
      for (int q = 0; q < nq; ++q) {
        double phi[MAXND];
        elem_basis(M, qp[q], phi);
        if (M->etype) elem_geo(M, X, qp[q], G);   /* gradients vary inside a cube */
        for (int k = 0; k < 3; ++k) {
          double p = 0;
          for (int a = 0; a < nd; ++a) p += phi[a] * X[a][k];
          ctx[SLOT_POS + k] = p;
        }
        double factor = qw[q] * fabs(det);
        ctx[SLOT_INTFAC] = factor;
        eval_fields(M, P, c, e, x, phi, G, ctx);
        for (int t = P->comp_ptr[c]; t < P->comp_ptr[c + 1]; ++t) {
          int g = P->comp_spec[t], s = P->spec_local[g];
          double scalar = 0, flux[3] = {0, 0, 0};
          if (form == 0) {
            TERMS(P, g, K_REACTION, r0, r1);
            if (r1 > r0) scalar = -run(P, P->terms[r0 * 5 + 4], ctx);
            TERMS(P, g, K_DIFF, d0, d1);
            for (int d = d0; d < d1; ++d) {
              int j = P->terms[d * 5 + 2];
              double D = run(P, P->terms[d * 5 + 4], ctx);
              const double* gj = ctx + P->spec_base + 4 * j + 1;
              for (int k = 0; k < dim; ++k) flux[k] -= D * gj[k];
            }
            /* tensor diffusion: out[r] = sum_c D[r][c] in[c]  (functor_factory_parser.impl.hh:79-107) */
            TERMS(P, g, K_DIFF_T, t0, t1);
            for (int d = t0; d < t1; ++d) {
              int j = P->terms[d * 5 + 2], rr = P->terms[d * 5 + 3] / 3, cc = P->terms[d * 5 + 3] % 3;
              double D = run(P, P->terms[d * 5 + 4], ctx);
              flux[rr] -= D * ctx[P->spec_base + 4 * j + 1 + cc];
            }
            /* advection: flux += velocity * u_i  (local_operator.hh:481) */
            TERMS(P, g, K_VEL, v0, v1);
            for (int d = v0; d < v1; ++d)
              flux[P->terms[d * 5 + 2]] += run(P, P->terms[d * 5 + 4], ctx) * ctx[P->spec_base + 4 * g];
          } else {
            TERMS(P, g, K_STORAGE, s0, s1);
            if (s1 > s0) scalar += ctx[P->spec_base + 4 * g] * run(P, P->terms[s0 * 5 + 4], ctx);
          }
          for (int a = 0; a < nd; ++a) {
            double fl = 0;
            for (int k = 0; k < dim; ++k) fl += flux[k] * G[a][k];
            loc[s][a] += (scalar * phi[a] - fl) * factor;
          }
        }
      }
      for (int s = 0; s < ns; ++s)
        for (int a = 0; a < nd; ++a) atomic_add(&r[M->elem_dof[e * nd + a] + s], w * loc[s][a], par);
      ctx[SLOT_INVOL] = 0;
    }
    free(ctx);
  }
}

/* ---------------------------------------------------------------- volume Jacobian (analytic)   */
static void jacobian_volume(const OrcMesh* M, const OrcModel* P, int form, double time, double w,
                            const double* x, const Sink* S) {
  const int dim = M->dim, nd = mesh_nd(M);
  double qp[MAXND][3], qw[MAXND];
  const int nq = elem_quad(M, qp, qw);
  double fact = elem_fact(M);
#pragma omp parallel if (S->par)
  {
    double* ctx = (double*)calloc(P->nslots, sizeof(double));
#pragma omp for schedule(static)
    for (int64_t e = 0; e < M->ne; ++e) {
      int c = M->elem_comp[e];
      if (c < 0) continue;
      double X[MAXND][3], G[MAXND][3];
      load_element(M, e, X);
      double det = elem_geo(M, X, qp[0], G);
      ctx[SLOT_TIME] = time; ctx[SLOT_ENTVOL] = fabs(det) / fact; ctx[SLOT_INVOL] = 1;
      set_cell(M, e, ctx);
      const int64_t* ed = M->elem_dof + e * nd;
      for (int q = 0; q < nq; ++q) {
        double phi[MAXND];
        elem_basis(M, qp[q], phi);
        if (M->etype) elem_geo(M, X, qp[q], G);
        for (int k = 0; k < 3; ++k) {
          double p = 0;
          for (int a = 0; a < nd; ++a) p += phi[a] * X[a][k];
          ctx[SLOT_POS + k] = p;
        }
        double factor = qw[q] * fabs(det);
        ctx[SLOT_INTFAC] = factor;
        eval_fields(M, P, c, e, x, phi, G, ctx);
        for (int t = P->comp_ptr[c]; t < P->comp_ptr[c + 1]; ++t) {
          int g = P->comp_spec[t], si = P->spec_local[g];
          if (form == 0) {
            TERMS(P, g, K_REACTION, r0, r1);
            if (r1 > r0) {
              TERMS(P, g, K_REACTION_JAC, j0, j1);
              for (int jt = j0; jt < j1; ++jt) {
                int wrt = P->terms[jt * 5 + 2], sj = P->spec_local[wrt];
                double jac = run(P, P->terms[jt * 5 + 4], ctx);
                for (int a = 0; a < nd; ++a)
                  for (int b = 0; b < nd; ++b)
                    sink_add(S, ed[a] + si, ed[b] + sj, w * (-jac * phi[a] * phi[b] * factor));
              }
            }
            TERMS(P, g, K_DIFF, d0, d1);
            for (int d = d0; d < d1; ++d) {
              int wrt = P->terms[d * 5 + 2], sj = P->spec_local[wrt];
              double D = run(P, P->terms[d * 5 + 4], ctx);
              for (int a = 0; a < nd; ++a) {
                for (int b = 0; b < nd; ++b) {
                  double v = 0;
                  for (int k = 0; k < dim; ++k) v += D * G[a][k] * G[b][k];
                  sink_add(S, ed[a] + si, ed[b] + sj, w * v * factor);
                }
              }
            }
            /* non-linear diffusion d D_ij / d u_k -- literal restatement incl. the index
               transposition and the use of grad u_k (local_operator.hh:688-700) */
            TERMS(P, g, K_DIFF_JAC, e0, e1);
            for (int et = e0; et < e1; ++et) {
              int kk = P->terms[et * 5 + 3], sk = P->spec_local[kk];
              double dD = run(P, P->terms[et * 5 + 4], ctx);
              const double* gk = ctx + P->spec_base + 4 * kk + 1;
              for (int a = 0; a < nd; ++a)
                for (int b = 0; b < nd; ++b) {
                  double v = 0;
                  for (int k = 0; k < dim; ++k) v += dD * gk[k] * G[b][k];
                  sink_add(S, ed[a] + si, ed[b] + sk, w * phi[a] * v * factor);
                }
            }
            /* tensor diffusion, entry by entry: (D grad phi_a) . grad phi_b -- the tensor acts on the
               TEST gradient as written at local_operator.hh:679-684 */
            TERMS(P, g, K_DIFF_T, t0, t1);
            for (int d = t0; d < t1; ++d) {
              int wrt = P->terms[d * 5 + 2], sj = P->spec_local[wrt];
              int rr = P->terms[d * 5 + 3] / 3, cc = P->terms[d * 5 + 3] % 3;
              double D = run(P, P->terms[d * 5 + 4], ctx);
              for (int a = 0; a < nd; ++a)
                for (int b = 0; b < nd; ++b)
                  sink_add(S, ed[a] + si, ed[b] + sj, w * (D * G[a][cc] * G[b][rr]) * factor);
            }
            TERMS(P, g, K_DIFF_T_JAC, u0, u1);
            for (int d = u0; d < u1; ++d) {
              int code3 = P->terms[d * 5 + 3], kk = code3 / 9, rr = (code3 % 9) / 3, cc = code3 % 3;
              int sk = P->spec_local[kk];
              double dD = run(P, P->terms[d * 5 + 4], ctx);
              double gkc = ctx[P->spec_base + 4 * kk + 1 + cc];
              for (int a = 0; a < nd; ++a)
                for (int b = 0; b < nd; ++b)
                  sink_add(S, ed[a] + si, ed[b] + sk, w * phi[a] * (dD * gkc * G[b][rr]) * factor);
            }
            /* advection (local_operator.hh:643-671), literal incl. the roles of the indices:
               entry (test a, trial b) = -phi_a (vel . grad phi_b) */
            TERMS(P, g, K_VEL, v0, v1);
            for (int d = v0; d < v1; ++d) {
              int ax = P->terms[d * 5 + 2];
              double vel = run(P, P->terms[d * 5 + 4], ctx);
              for (int a = 0; a < nd; ++a)
                for (int b = 0; b < nd; ++b)
                  sink_add(S, ed[a] + si, ed[b] + si, w * (-(vel * phi[a]) * G[b][ax]) * factor);
            }
            TERMS(P, g, K_VEL_JAC, w0, w1);
            for (int d = w0; d < w1; ++d) {
              int kk = P->terms[d * 5 + 2], ax = P->terms[d * 5 + 3], sk = P->spec_local[kk];
              double adv = run(P, P->terms[d * 5 + 4], ctx) * ctx[P->spec_base + 4 * g];
              for (int a = 0; a < nd; ++a)
                for (int b = 0; b < nd; ++b)
                  sink_add(S, ed[a] + si, ed[b] + sk, w * (-phi[a] * (adv * G[b][ax])) * factor);
            }
          } else {
            TERMS(P, g, K_STORAGE, s0, s1);
            if (s1 > s0) {
              double stg = run(P, P->terms[s0 * 5 + 4], ctx);
              for (int a = 0; a < nd; ++a)
                for (int b = 0; b < nd; ++b)
                  sink_add(S, ed[a] + si, ed[b] + si, w * stg * phi[a] * phi[b] * factor);
              TERMS(P, g, K_STORAGE_JAC, j0, j1);
              for (int jt = j0; jt < j1; ++jt) {
                int wrt = P->terms[jt * 5 + 2], sj = P->spec_local[wrt];
                double jac = run(P, P->terms[jt * 5 + 4], ctx);
                double val = ctx[P->spec_base + 4 * g];
                for (int a = 0; a < nd; ++a)
                  for (int b = 0; b < nd; ++b)
                    sink_add(S, ed[a] + si, ed[b] + sj, w * jac * val * phi[a] * phi[b] * factor);
              }
            }
          }
        }
      }
      ctx[SLOT_INVOL] = 0;
    }
    free(ctx);
  }
}

void orc_jacobian_volume(const OrcMesh* M, const OrcModel* P, int form, double time, double w,
                         const double* x, const int64_t* rowptr, const int32_t* colidx,
                         double* vals, int par) {
  Sink S = {0, rowptr, colidx, vals, 0, 0, par};
  jacobian_volume(M, P, form, time, w, x, &S);
}

/* matrix-free y += w * J_form(x) z  (local_operator.hh:510-524 via PseudoJacobian) */
void orc_jacobian_apply_volume(const OrcMesh* M, const OrcModel* P, int form, double time,
                               double w, const double* x, const double* z, double* y, int par) {
  Sink S = {1, 0, 0, 0, z, y, par};
  jacobian_volume(M, P, form, time, w, x, &S);
}

/* numerical (finite difference) volume Jacobian, local_operator.hh:713-765:
 * delta = eps (1 + |x_j|), one-sided, column by column on the element's local vector.  */
void orc_jacobian_volume_numerical(const OrcMesh* M, const OrcModel* P, int form, double time,
                                   double w, double eps, const double* x, const int64_t* rowptr,
                                   const int32_t* colidx, double* vals) {
  const int nd = mesh_nd(M);
  Sink S = {0, rowptr, colidx, vals, 0, 0, 0};
  /* element-local evaluation through a one-element sub-mesh view */
  for (int64_t e = 0; e < M->ne; ++e) {
    int c = M->elem_comp[e];
    if (c < 0) continue;
    int ns = P->comp_ptr[c + 1] - P->comp_ptr[c];
    OrcMesh one = *M;
    int64_t ldof[MAXND];
    double xl[128], down[128], up[128];
    for (int a = 0; a < nd; ++a) {
      ldof[a] = (int64_t)a * ns;
      for (int s = 0; s < ns; ++s) xl[a * ns + s] = x[M->elem_dof[e * nd + a] + s];
    }
    one.ne = 1; one.elems = M->elems + e * nd; one.elem_comp = M->elem_comp + e;
    one.elem_dof = ldof; one.nf = 0;
    /* cell data of element e must be seen at local index 0 */
    double cd[32];
    for (int k = 0; k < M->nkeys; ++k) cd[k] = M->cell_data[(int64_t)k * M->ne + e];
    one.cell_data = cd;
    memset(down, 0, sizeof(down));
    orc_residual_volume(&one, P, form, time, 1.0, xl, down, 0);
    for (int b = 0; b < nd; ++b)
      for (int sj = 0; sj < ns; ++sj) {
        int col = b * ns + sj;
        double keep = xl[col], delta = eps * (1.0 + fabs(keep));
        xl[col] += delta;
        memset(up, 0, sizeof(up));
        orc_residual_volume(&one, P, form, time, 1.0, xl, up, 0);
        for (int a = 0; a < nd; ++a)
          for (int si = 0; si < ns; ++si) {
            int row = a * ns + si;
            double v = (up[row] - down[row]) / delta;
            if (v != 0.0 || 1) {
              /* only entries inside the pattern are stored (implicit build mode drops none in the
                 reference because the FD Jacobian is dense per element; we mirror by skipping
                 exact zeros outside the pattern) */
              int64_t R = M->elem_dof[e * nd + a] + si, C = M->elem_dof[e * nd + b] + sj;
              int64_t lo = rowptr[R], hi = rowptr[R + 1] - 1, hit = -1;
              while (lo <= hi) {
                int64_t mid = (lo + hi) >> 1;
                if (colidx[mid] == C) { hit = mid; break; }
                if (colidx[mid] < C) lo = mid + 1; else hi = mid - 1;
              }
              if (hit >= 0) vals[hit] += w * v;
              else if (v != 0.0) abort();
            }
          }
        xl[col] = keep;
      }
  }
  (void)S;
}

/* ---------------------------------------------------------------- skeleton / boundary
 * Facet f between e_i (compartment ci) and e_o (co != ci), or boundary facet (e_o = -1).
 * Disjoint compartments: a species of ci fires outflow.<co> on an interface facet and
 * outflow.<ci> on a boundary facet (local_operator.hh:839-852).                               */
typedef struct {
  double Xi[4][3], Gi[4][3], Xo[4][3], Go[4][3];
  double area, normal[3];
  int fv_i[3], fv_o[3]; /* local indices of the facet vertices in e_i / matching ones in e_o */
} FacetGeo;

static void facet_geo(const OrcMesh* M, int64_t f, FacetGeo* F) {
  const int dim = M->dim, nd = dim + 1;
  int64_t ei = M->f_in[f], eo = M->f_out[f];
  int mi = M->f_lin[f];
  load_element(M, ei, F->Xi);
  simplex_geo(dim, F->Xi, F->Gi);
  int n = 0;
  for (int a = 0; a < nd; ++a) if (a != mi) F->fv_i[n++] = a;
  if (eo >= 0) {
    load_element(M, eo, F->Xo);
    simplex_geo(dim, F->Xo, F->Go);
    for (int k = 0; k < dim; ++k) {
      int32_t gv = M->elems[ei * nd + F->fv_i[k]];
      F->fv_o[k] = -1;
      for (int a = 0; a < nd; ++a) if (M->elems[eo * nd + a] == gv) F->fv_o[k] = a;
    }
  }
  /* outer unit normal of the inside element: -grad(phi_m)/|grad(phi_m)| */
  double nn = 0;
  for (int k = 0; k < dim; ++k) nn += F->Gi[mi][k] * F->Gi[mi][k];
  nn = sqrt(nn);
  for (int k = 0; k < 3; ++k) F->normal[k] = k < dim ? -F->Gi[mi][k] / nn : 0.0;
  if (dim == 2) {
    double dx = F->Xi[F->fv_i[1]][0] - F->Xi[F->fv_i[0]][0], dy = F->Xi[F->fv_i[1]][1] - F->Xi[F->fv_i[0]][1];
    F->area = sqrt(dx * dx + dy * dy);
  } else {
    double u[3], v[3];
    for (int k = 0; k < 3; ++k) {
      u[k] = F->Xi[F->fv_i[1]][k] - F->Xi[F->fv_i[0]][k];
      v[k] = F->Xi[F->fv_i[2]][k] - F->Xi[F->fv_i[0]][k];
    }
    double cx = u[1] * v[2] - u[2] * v[1], cy = u[2] * v[0] - u[0] * v[2], cz = u[0] * v[1] - u[1] * v[0];
    F->area = 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
  }
}

/* model.b200.reference_compat: cross-side values as local_operator.hh:903-916 / :939-941 compute them -- the
 * coefficients of the element across the facet are paired, local index by local index, with the shape
 * functions (and their gradients) of the element whose residual rows are being assembled (bound through
 * quad_proj_i / quad_proj_o).  0: the P1 trace at the physical point (what the formula means when both
 * elements number the shared vertices alike).  Set per call by oracle/core.py. */
static int g_ref_compat = 1;
void orc_set_reference_compat(int on) { g_ref_compat = on; }

/* fill ctx with every species of both sides at the facet point with facet barycentrics lam, as seen from
 * `self_side` (the side whose rows are assembled) */
static void facet_fields(const OrcMesh* M, const OrcModel* P, int64_t f, const FacetGeo* F,
                         const double* lam, const double* x, double* ctx, int self_side) {
  const int dim = M->dim, nd = dim + 1;
  for (int side = 0; side < 2; ++side) {
    int64_t e = side == 0 ? M->f_in[f] : M->f_out[f];
    if (e < 0) continue;
    int c = M->elem_comp[e];
    if (c < 0) continue;
    /* whose shape functions: its own, or (compat, other side) those of the assembling element */
    const int basis = (g_ref_compat && M->f_out[f] >= 0) ? self_side : side;
    double phi[4] = {0, 0, 0, 0};
    const int* fv = basis == 0 ? F->fv_i : F->fv_o;
    for (int k = 0; k < dim; ++k) phi[fv[k]] = lam[k];
    double(*G)[3] = basis == 0 ? (double(*)[3])F->Gi : (double(*)[3])F->Go;
    eval_fields(M, P, c, e, x, phi, G, ctx);
    (void)nd;
  }
}

static int facet_quad(int dim, double lam[][3], double* wts) {
  /* rule of order 2 on the (dim-1)-simplex, expressed in facet barycentric coordinates */
  if (dim == 2) {
    for (int q = 0; q < 2; ++q) { lam[q][0] = 1.0 - Q1[q]; lam[q][1] = Q1[q]; wts[q] = 0.5; }
    return 2;
  }
  for (int q = 0; q < 3; ++q) {
    lam[q][0] = 1.0 - Q2[q][0] - Q2[q][1]; lam[q][1] = Q2[q][0]; lam[q][2] = Q2[q][1];
    wts[q] = 1. / 6.;
  }
  return 3;
}

/* mode 0: residual r += w*T ; mode 1: jacobian into sink */
static void skeleton_range(const OrcMesh* M, const OrcModel* P, double time, double w, const double* x,
                           double* r, const Sink* S, int mode, int64_t f_begin, int64_t f_end) {
  const int dim = M->dim, nd = dim + 1;
  double lam[3][3], qw[3];
  const int nq = facet_quad(dim, lam, qw);
  double* ctx = (double*)calloc(P->nslots, sizeof(double));
  for (int64_t f = f_begin; f < f_end; ++f) {
    int64_t ei = M->f_in[f], eo = M->f_out[f];
    int ci = M->elem_comp[ei], co = eo >= 0 ? M->elem_comp[eo] : -1;
    if (eo >= 0 && ci == co) continue;
    FacetGeo F;
    facet_geo(M, f, &F);
    double ie = dim == 2 ? F.area : 2.0 * F.area; /* integrationElement of the facet map */
    for (int side = 0; side < 2; ++side) {
      int64_t e = side == 0 ? ei : eo;
      int cs = side == 0 ? ci : co;            /* own compartment */
      int ct = eo >= 0 ? (side == 0 ? co : ci) : ci; /* outflow target compartment */
      if (e < 0 || cs < 0 || ct < 0) continue;
      const int* fv = side == 0 ? F.fv_i : F.fv_o;
      for (int q = 0; q < nq; ++q) {
        memset(ctx, 0, sizeof(double) * P->nslots);
        ctx[SLOT_TIME] = time; ctx[SLOT_ENTVOL] = F.area;
        ctx[SLOT_INBND] = eo < 0; ctx[SLOT_INSKEL] = eo >= 0;
        for (int k = 0; k < M->nkeys; ++k) ctx[SLOT_CELL + k] = M->cell_data[(int64_t)k * M->ne + e];
        for (int k = 0; k < 3; ++k) {
          double p = 0;
          for (int a = 0; a < dim; ++a) p += lam[q][a] * F.Xi[F.fv_i[a]][k];
          ctx[SLOT_POS + k] = p;
          ctx[SLOT_NORMAL + k] = side == 0 ? F.normal[k] : -F.normal[k];
        }
        double factor = qw[q] * ie;
        ctx[SLOT_INTFAC] = factor;
        facet_fields(M, P, f, &F, lam[q], x, ctx, side);
        for (int t = P->comp_ptr[cs]; t < P->comp_ptr[cs + 1]; ++t) {
          int g = P->comp_spec[t], si = P->spec_local[g];
          TERMS(P, g, K_OUTFLOW, o0, o1);
          for (int ot = o0; ot < o1; ++ot) {
            if (P->terms[ot * 5 + 2] != ct) continue;
            if (mode == 0) {
              double T = run(P, P->terms[ot * 5 + 4], ctx);
              for (int k = 0; k < dim; ++k)
                r[M->elem_dof[e * nd + fv[k]] + si] += w * T * lam[q][k] * factor;
            } else {
              TERMS(P, g, K_OUTFLOW_JAC, j0, j1);
              for (int jt = j0; jt < j1; ++jt) {
                if (P->terms[jt * 5 + 2] != ct) continue;
                int wrt = P->terms[jt * 5 + 3], sj = P->spec_local[wrt], cw = P->spec_comp[wrt];
                double jac = run(P, P->terms[jt * 5 + 4], ctx);
                /* block choice: the side where the wrt species lives (local_operator.hh:1129-1131) */
                int64_t ew; const int* fw;
                if (cw == cs) { ew = e; fw = fv; }
                else if (eo >= 0 && cw == (side == 0 ? co : ci)) {
                  ew = side == 0 ? eo : ei;
                  /* compat: column dof_j of the other element's basis carries phi_j of this element (:1133-1143) */
                  fw = g_ref_compat ? fv : (side == 0 ? F.fv_o : F.fv_i);
                }
                else continue;
                for (int a = 0; a < dim; ++a)
                  for (int b = 0; b < dim; ++b)
                    sink_add(S, M->elem_dof[e * nd + fv[a]] + si, M->elem_dof[ew * nd + fw[b]] + sj,
                             w * jac * lam[q][a] * lam[q][b] * factor);
              }
            }
          }
        }
      }
    }
  }
  free(ctx);
}

static void skeleton(const OrcMesh* M, const OrcModel* P, double time, double w, const double* x,
                     double* r, const Sink* S, int mode) {
  skeleton_range(M, P, time, w, x, r, S, mode, 0, M->nf);
}

void orc_residual_skeleton(const OrcMesh* M, const OrcModel* P, double time, double w,
                           const double* x, double* r) {
  skeleton(M, P, time, w, x, r, 0, 0);
}
void orc_jacobian_skeleton(const OrcMesh* M, const OrcModel* P, double time, double w,
                           const double* x, const int64_t* rowptr, const int32_t* colidx, double* vals) {
  Sink S = {0, rowptr, colidx, vals, 0, 0, 0};
  skeleton(M, P, time, w, x, 0, &S, 1);
}
void orc_jacobian_apply_skeleton(const OrcMesh* M, const OrcModel* P, double time, double w,
                                 const double* x, const double* z, double* y) {
  Sink S = {1, 0, 0, 0, z, y, 0};
  skeleton(M, P, time, w, x, 0, &S, 1);
}

/* numerical skeleton / boundary Jacobian, local_operator.hh:1205-1343: one-sided differences of the
 * facet residual, column by column over the coefficients of both sides, delta = eps (1 + |x_col|).
 * Columns are the dofs at the facet's vertices (the links of the skeleton pattern); entries
 * outside the pattern are dropped.  For out-side columns the reference sizes delta with `coeff_in` read at
 * the out-side node (:1298): reproduced under reference_compat (see `partner` below), otherwise the out-side
 * coefficient itself is used.  `n` = number of dofs. */
void orc_jacobian_skeleton_numerical(const OrcMesh* M, const OrcModel* P, double time, double w, double eps,
                                     int64_t n, const double* x, const int64_t* rowptr,
                                     const int32_t* colidx, double* vals) {
  const int dim = M->dim, nd = dim + 1;
  double* xw = (double*)malloc(8 * n);
  double* down = (double*)calloc(n, 8);
  double* up = (double*)calloc(n, 8);
  memcpy(xw, x, 8 * n);
  for (int64_t f = 0; f < M->nf; ++f) {
    int64_t ei = M->f_in[f], eo = M->f_out[f];
    int ci = M->elem_comp[ei], co = eo >= 0 ? M->elem_comp[eo] : -1;
    if (eo >= 0 && ci == co) continue;
    int64_t dofs[2 * 4 * 32], partner[2 * 4 * 32];
    int nd_f = 0;
    for (int side = 0; side < 2; ++side) {
      int64_t e = side == 0 ? ei : eo;
      int c = side == 0 ? ci : co;
      if (e < 0 || c < 0) continue;
      int m = side == 0 ? M->f_lin[f] : M->f_lout[f];
      int ns = P->comp_ptr[c + 1] - P->comp_ptr[c];
      for (int a = 0; a < nd; ++a)
        if (a != m || (g_ref_compat && eo >= 0))   /* compat: the local index pairing reaches every vertex */
          for (int s = 0; s < ns; ++s) {
            dofs[nd_f] = M->elem_dof[e * nd + a] + s;
            /* compat (:1298): delta of an out-side column is sized with coeff_in read at the out-side node --
             * the inside container at the same (species slot, local vertex), 0 where it has no such entry */
            partner[nd_f] = -1;
            if (side == 1 && g_ref_compat) {
              int nsi = P->comp_ptr[ci + 1] - P->comp_ptr[ci];
              partner[nd_f] = s < nsi ? M->elem_dof[ei * nd + a] + s : -2;
            }
            ++nd_f;
          }
    }
    for (int k = 0; k < nd_f; ++k) down[dofs[k]] = 0.0;
    skeleton_range(M, P, time, 1.0, xw, down, 0, 0, f, f + 1);
    for (int cidx = 0; cidx < nd_f; ++cidx) {
      int64_t col = dofs[cidx];
      double keep = xw[col];
      double sized = partner[cidx] == -1 ? keep : partner[cidx] == -2 ? 0.0 : xw[partner[cidx]];
      double delta = eps * (1.0 + fabs(sized));
      xw[col] = keep + delta;
      for (int k = 0; k < nd_f; ++k) up[dofs[k]] = 0.0;
      skeleton_range(M, P, time, 1.0, xw, up, 0, 0, f, f + 1);
      for (int k = 0; k < nd_f; ++k) {
        int64_t row = dofs[k];
        double v = (up[row] - down[row]) / delta;
        if (v == 0.0) continue;
        int64_t lo = rowptr[row], hi = rowptr[row + 1] - 1;
        while (lo <= hi) {
          int64_t mid = (lo + hi) >> 1;
          if (colidx[mid] == col) { vals[mid] += w * v; break; }
          if (colidx[mid] < col) lo = mid + 1; else hi = mid - 1;
        }
      }
      xw[col] = keep;
    }
  }
  free(xw); free(down); free(up);
}

/* ---------------------------------------------------------------- linear algebra (dune-istl order) */
void orc_spmv(int64_t n, const int64_t* rowptr, const int32_t* colidx, const double* vals,
              const double* x, double* y, int par) {
#pragma omp parallel for schedule(static) if (par)
  for (int64_t i = 0; i < n; ++i) {
    double s = 0;
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) s += vals[k] * x[colidx[k]];
    y[i] = s;
  }
}

/* elementwise vector sweeps: identical results for any thread count (dune-istl's are sequential; the
 * bench's CPU arm turns the threads on so that the baseline is not held back by them) */
#define PAR_PRAGMA(x) _Pragma(#x)
#define PAR_FOR(par) PAR_PRAGMA(omp parallel for schedule(static) if (par))
static int g_par_blas = 0;

static double dot(int64_t n, const double* a, const double* b, int par) {
  double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static) if (par)
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

/* preconditioner: kind 0 none (Richardson w=1), 1 Jacobi (SeqJac, 1 sweep from v=0: v = w D^-1 d),
 * 2 BlockJacobi with node blocks of size bs (block_jacobi.hh:46-128 with iterations=1: v = w Dblk^-1 d),
 * 3 SSOR, 4 SOR, 5 GaussSeidel = dune-istl SeqSSOR / SeqSOR / SeqGS on scalar entries in the dof order
 * (registry: solver/istl/factory/preconditioner.hh:101-104; SSOR is DUNE_COPASI_DEFAULT_PRECONDITIONER,
 * :17).  Third party, restated from dune-istl's gsetc.hh as published:
 *   bsorf: for rows i ascending   x_i += w (d_i - sum_j a_ij x_j) / a_ii   (current x, diagonal included)
 *   bsorb: the same for rows descending
 *   dbgs:  xold = x; for rows i ascending   x_i = (d_i - sum_{j != i} a_ij x_j) / a_ii;  then x = w x + (1 - w) xold
 *   SeqSSOR::apply = n x (bsorf; bsorb), SeqSOR::apply = n x bsorf, SeqGS::apply = n x dbgs, v = 0 on entry */
typedef struct {
  int kind, bs; double relax; double* dinv; int64_t n;
  int iters; const int64_t* rowptr; const int32_t* colidx; const double* vals;   /* sweeps > 1 */
} Prec;

/* `preconditioner.iterations` of the next solver calls (kept out of the solver signatures) */
static int g_prec_iters = 1;
void orc_set_preconditioner_iterations(int iters) { g_prec_iters = iters < 1 ? 1 : iters; }

static void prec_setup(Prec* Pc, int64_t n, const int64_t* rowptr, const int32_t* colidx,
                       const double* vals) {
  Pc->n = n;
  Pc->iters = g_prec_iters; Pc->rowptr = rowptr; Pc->colidx = colidx; Pc->vals = vals;
  if (Pc->kind == 1) {
    Pc->dinv = (double*)malloc(sizeof(double) * n);
    for (int64_t i = 0; i < n; ++i) {
      double d = 0;
      for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) if (colidx[k] == i) d = vals[k];
      Pc->dinv[i] = 1.0 / d;
    }
  } else if (Pc->kind == 2) {
    int bs = Pc->bs;
    int64_t nb = n / bs;
    Pc->dinv = (double*)malloc(sizeof(double) * nb * bs * bs);
    for (int64_t I = 0; I < nb; ++I) {
      double A[19 * 19], B[19 * 19];
      for (int a = 0; a < bs; ++a)
        for (int b = 0; b < bs; ++b) {
          double d = 0;
          int64_t row = I * bs + a, col = I * bs + b;
          for (int64_t k = rowptr[row]; k < rowptr[row + 1]; ++k) if (colidx[k] == col) d = vals[k];
          A[a * bs + b] = d; B[a * bs + b] = a == b;
        }
      /* Gauss-Jordan with partial pivoting (FieldMatrix::invert, dense_inverse.hh:9-37) */
      for (int p = 0; p < bs; ++p) {
        int piv = p;
        for (int i2 = p + 1; i2 < bs; ++i2) if (fabs(A[i2 * bs + p]) > fabs(A[piv * bs + p])) piv = i2;
        if (piv != p)
          for (int j = 0; j < bs; ++j) {
            double t = A[p * bs + j]; A[p * bs + j] = A[piv * bs + j]; A[piv * bs + j] = t;
            t = B[p * bs + j]; B[p * bs + j] = B[piv * bs + j]; B[piv * bs + j] = t;
          }
        double ip = 1.0 / A[p * bs + p];
        for (int j = 0; j < bs; ++j) { A[p * bs + j] *= ip; B[p * bs + j] *= ip; }
        for (int i2 = 0; i2 < bs; ++i2) if (i2 != p) {
          double fct = A[i2 * bs + p];
          for (int j = 0; j < bs; ++j) { A[i2 * bs + j] -= fct * A[p * bs + j]; B[i2 * bs + j] -= fct * B[p * bs + j]; }
        }
      }
      memcpy(Pc->dinv + I * bs * bs, B, sizeof(double) * bs * bs);
    }
  } else
    Pc->dinv = 0;
}

static void prec_sweep(const Prec* Pc, const double* d, double* v);

/* v = 0, then `iters` sweeps.
 * Jacobi = dune-istl SeqJac: v += w D^-1 (d - A v) with the old iterate in every row.
 * BlockJacobi = block_jacobi.hh:102-127 as written: the right-hand side copy is modified
 * cumulatively, b_k = b_{k-1} - A v_{k-1}, which is the true defect only for the first two sweeps. */
static void sor_sweep(const Prec* Pc, const double* d, double* v, int backward, int skip_diag) {
  const int64_t n = Pc->n;
  for (int64_t q = 0; q < n; ++q) {
    const int64_t i = backward ? n - 1 - q : q;
    double rhs = d[i], diag = 1.0;
    for (int64_t k = Pc->rowptr[i]; k < Pc->rowptr[i + 1]; ++k) {
      const int64_t j = Pc->colidx[k];
      if (j == i) { diag = Pc->vals[k]; if (skip_diag) continue; }
      rhs -= Pc->vals[k] * v[j];
    }
    if (skip_diag) v[i] = rhs / diag;   /* dbgs: the unrelaxed value inside the sweep ... */
    else v[i] += Pc->relax * (rhs / diag);
  }
}

static void prec_apply(const Prec* Pc, const double* d, double* v) {
  if (Pc->kind >= 3) {
    memset(v, 0, sizeof(double) * Pc->n);
    double* xold = Pc->kind == 5 ? malloc(sizeof(double) * Pc->n) : 0;
    for (int it = 0; it < Pc->iters; ++it) {
      if (xold) memcpy(xold, v, sizeof(double) * Pc->n);
      sor_sweep(Pc, d, v, 0, Pc->kind == 5);
      /* ... and x = w x + (1 - w) xold once the sweep is through (dune-istl gsetc.hh, dbgs) */
      if (xold) for (int64_t i = 0; i < Pc->n; ++i) v[i] = Pc->relax * v[i] + (1.0 - Pc->relax) * xold[i];
      if (Pc->kind == 3) sor_sweep(Pc, d, v, 1, 0);
    }
    free(xold);
    return;
  }
  prec_sweep(Pc, d, v);
  if (Pc->iters <= 1 || Pc->kind == 0) return;
  int64_t n = Pc->n;
  double *b = malloc(8 * n), *t = malloc(8 * n), *c = malloc(8 * n);
  memcpy(b, d, 8 * n);
  for (int it = 1; it < Pc->iters; ++it) {
    orc_spmv(n, Pc->rowptr, Pc->colidx, Pc->vals, v, t, 0);
    if (Pc->kind == 1) for (int64_t i = 0; i < n; ++i) b[i] = d[i] - t[i];
    else for (int64_t i = 0; i < n; ++i) b[i] -= t[i];
    prec_sweep(Pc, b, c);
    for (int64_t i = 0; i < n; ++i) v[i] += c[i];
  }
  free(b); free(t); free(c);
}

static void prec_sweep(const Prec* Pc, const double* d, double* v) {
  int64_t n = Pc->n;
  if (Pc->kind == 0) { for (int64_t i = 0; i < n; ++i) v[i] = d[i]; }
  else if (Pc->kind == 1) { PAR_FOR(g_par_blas) for (int64_t i = 0; i < n; ++i) v[i] = Pc->relax * Pc->dinv[i] * d[i]; }
  else {
    int bs = Pc->bs;
    for (int64_t I = 0; I < n / bs; ++I)
      for (int a = 0; a < bs; ++a) {
        double s = 0;
        for (int b = 0; b < bs; ++b) s += Pc->dinv[I * bs * bs + a * bs + b] * d[I * bs + b];
        v[I * bs + a] = Pc->relax * s;
      }
  }
}

typedef struct { int32_t iterations_x2; int32_t converged; double reduction; double norm0; } OrcResult;

/* dune-istl BiCGSTABSolver::apply.  On exit x holds the solution; b is consumed. */
void orc_bicgstab(int64_t n, const int64_t* rowptr, const int32_t* colidx, const double* vals,
                  double* x, double* b, double reduction, int maxit, int prec_kind, int bs,
                  double relax, int par, OrcResult* res) {
  Prec Pc = {prec_kind, bs, relax, 0, 0, 1, 0, 0, 0};
  g_par_blas = par;   /* elementwise sweeps over all host threads when the caller asks for it (bench CPU arm) */
  prec_setup(&Pc, n, rowptr, colidx, vals);
  double *r = b, *rt = malloc(8 * n), *p = calloc(n, 8), *v = calloc(n, 8), *t = malloc(8 * n),
         *y = malloc(8 * n);
  orc_spmv(n, rowptr, colidx, vals, x, t, par);
  PAR_FOR(par) for (int64_t i = 0; i < n; ++i) { r[i] -= t[i]; rt[i] = r[i]; }
  double norm0 = sqrt(dot(n, r, r, par)), norm = norm0;
  double rho = 1, alpha = 1, omega = 1, rho_new, h;
  double it = 0;
  res->converged = 0; res->norm0 = norm0;
  if (norm0 < 1e-30) { res->converged = 1; res->iterations_x2 = 0; res->reduction = 0; goto done; }
  for (it = 0.5; it < maxit; it += 0.5) {
    rho_new = dot(n, rt, r, par);
    if (fabs(rho) <= 1e-80 || fabs(omega) <= 1e-80) break;
    if (it < 1) { PAR_FOR(par) for (int64_t i = 0; i < n; ++i) p[i] = r[i]; }
    else {
      double beta = (rho_new / rho) * (alpha / omega);
      PAR_FOR(par) for (int64_t i = 0; i < n; ++i) p[i] = r[i] + beta * (p[i] - omega * v[i]);
    }
    prec_apply(&Pc, p, y);
    orc_spmv(n, rowptr, colidx, vals, y, v, par);
    h = dot(n, rt, v, par);
    if (fabs(h) < 1e-80) break;
    alpha = rho_new / h;
    PAR_FOR(par) for (int64_t i = 0; i < n; ++i) { x[i] += alpha * y[i]; r[i] -= alpha * v[i]; }
    norm = sqrt(dot(n, r, r, par));
    if (norm < reduction * norm0 || norm < 1e-30) { res->converged = 1; break; }
    it += 0.5;
    prec_apply(&Pc, r, y);
    orc_spmv(n, rowptr, colidx, vals, y, t, par);
    omega = dot(n, t, r, par) / dot(n, t, t, par);
    PAR_FOR(par) for (int64_t i = 0; i < n; ++i) { x[i] += omega * y[i]; r[i] -= omega * t[i]; }
    rho = rho_new;
    norm = sqrt(dot(n, r, r, par));
    if (norm < reduction * norm0 || norm < 1e-30) { res->converged = 1; break; }
  }
  res->iterations_x2 = (int32_t)(2 * it + 0.5);
  res->reduction = norm / norm0;
done:
  free(rt); free(p); free(v); free(t); free(y); free(Pc.dinv);
}

/* dune-istl CGSolver::apply: preconditioned CG, convergence on ||r||_2 */
void orc_cg(int64_t n, const int64_t* rowptr, const int32_t* colidx, const double* vals, double* x,
            double* b, double reduction, int maxit, int prec_kind, int bs, double relax, int par,
            OrcResult* res) {
  Prec Pc = {prec_kind, bs, relax, 0, 0, 1, 0, 0, 0};
  prec_setup(&Pc, n, rowptr, colidx, vals);
  double *r = b, *p = malloc(8 * n), *q = malloc(8 * n);
  orc_spmv(n, rowptr, colidx, vals, x, q, par);
  for (int64_t i = 0; i < n; ++i) r[i] -= q[i];
  double norm0 = sqrt(dot(n, r, r, par)), norm = norm0;
  res->converged = 0; res->norm0 = norm0; res->iterations_x2 = 0; res->reduction = 1;
  if (norm0 < 1e-30) { res->converged = 1; res->reduction = 0; goto done; }
  prec_apply(&Pc, r, p);
  double rholast = dot(n, p, r, par);
  int i = 1;
  for (; i <= maxit; ++i) {
    orc_spmv(n, rowptr, colidx, vals, p, q, par);
    double alpha = dot(n, p, q, par);
    double lambda = rholast / alpha;
    for (int64_t k = 0; k < n; ++k) { x[k] += lambda * p[k]; r[k] -= lambda * q[k]; }
    norm = sqrt(dot(n, r, r, par));
    if (norm < reduction * norm0 || norm < 1e-30) { res->converged = 1; break; }
    prec_apply(&Pc, r, q);
    double rho = dot(n, q, r, par);
    double beta = rho / rholast;
    for (int64_t k = 0; k < n; ++k) p[k] = q[k] + beta * p[k];
    rholast = rho;
  }
  res->iterations_x2 = 2 * (i <= maxit ? i : maxit);
  res->reduction = norm / norm0;
done:
  free(p); free(q); free(Pc.dinv);
}

/* dune-istl RestartedGMResSolver::apply (third party; used by the reference's own inis through
 * solver/istl/factory/iterative.hh:64, restart default 40): left preconditioned GMRES(m), modified
 * Gram-Schmidt, Givens rotations, convergence on the preconditioned defect ||W^-1 (b - A x)||. */
static void givens_gen(double dx, double dy, double* cs, double* sn) {
  double ndx = fabs(dx), ndy = fabs(dy);
  if (ndy < 1e-300) { *cs = 1.0; *sn = 0.0; }
  else if (ndx < 1e-300) { *cs = 0.0; *sn = 1.0; }
  else { double nrm = sqrt(dx * dx + dy * dy); *cs = dx / nrm; *sn = dy / nrm; }
}
static void givens_apply(double* dx, double* dy, double cs, double sn) {
  double t = cs * (*dx) + sn * (*dy);
  *dy = -sn * (*dx) + cs * (*dy);
  *dx = t;
}

void orc_gmres(int64_t n, const int64_t* rowptr, const int32_t* colidx, const double* vals, double* x,
               double* b, double reduction, int maxit, int restart, int prec_kind, int bs, double relax,
               int par, OrcResult* res) {
  Prec Pc = {prec_kind, bs, relax, 0, 0, 1, 0, 0, 0};
  prec_setup(&Pc, n, rowptr, colidx, vals);
  const int m = restart;
  double* V = malloc(sizeof(double) * (size_t)(m + 1) * n);
  double *w = malloc(8 * n), *b2 = malloc(8 * n), *tmp = malloc(8 * n);
  double* H = calloc((size_t)(m + 1) * m, 8);
  double *s = calloc(m + 1, 8), *cs = calloc(m, 8), *sn = calloc(m, 8), *yv = calloc(m, 8);
#define Hh(r, c) H[(size_t)(r) * m + (c)]
  memcpy(b2, b, 8 * n);
  orc_spmv(n, rowptr, colidx, vals, x, tmp, par);
  for (int64_t i = 0; i < n; ++i) b[i] -= tmp[i];
  prec_apply(&Pc, b, V);
  double norm = sqrt(dot(n, V, V, par)), norm0 = norm;
  res->norm0 = norm0; res->converged = 0; res->iterations_x2 = 0; res->reduction = 1;
  if (norm0 < 1e-30) { res->converged = 1; res->reduction = 0; goto done; }
  int j = 1;
  while (j <= maxit && !res->converged) {
    int i = 0;
    for (int64_t k = 0; k < n; ++k) V[k] *= 1.0 / norm;
    s[0] = norm;
    for (i = 1; i < m + 1; ++i) s[i] = 0.0;
    for (i = 0; i < m && j <= maxit && !res->converged; ++i, ++j) {
      double* vi = V + (size_t)i * n;
      double* vn = V + (size_t)(i + 1) * n;
      orc_spmv(n, rowptr, colidx, vals, vi, vn, par);
      prec_apply(&Pc, vn, w);
      for (int k = 0; k < i + 1; ++k) {
        double* vk = V + (size_t)k * n;
        double h = dot(n, vk, w, par);
        Hh(k, i) = h;
        for (int64_t q = 0; q < n; ++q) w[q] -= h * vk[q];
      }
      double hn = sqrt(dot(n, w, w, par));
      Hh(i + 1, i) = hn;
      if (fabs(hn) < 1e-80) { j = maxit + 1; break; }   /* breakdown: dune-istl throws SolverAbort */
      for (int64_t q = 0; q < n; ++q) vn[q] = w[q] * (1.0 / hn);
      for (int k = 0; k < i; ++k) givens_apply(&Hh(k, i), &Hh(k + 1, i), cs[k], sn[k]);
      givens_gen(Hh(i, i), Hh(i + 1, i), &cs[i], &sn[i]);
      givens_apply(&Hh(i, i), &Hh(i + 1, i), cs[i], sn[i]);
      givens_apply(&s[i], &s[i + 1], cs[i], sn[i]);
      norm = fabs(s[i + 1]);
      if (norm < reduction * norm0 || norm < 1e-30) res->converged = 1;
    }
    /* update: solve the triangular system, x += sum y_k v_k */
    for (int a = i - 1; a >= 0; --a) {
      double acc = s[a];
      for (int c2 = a + 1; c2 < i; ++c2) acc -= Hh(a, c2) * yv[c2];
      yv[a] = acc / Hh(a, a);
    }
    for (int a = 0; a < i; ++a) {
      const double* va = V + (size_t)a * n;
      for (int64_t q = 0; q < n; ++q) x[q] += yv[a] * va[q];
    }
    if (!res->converged && j <= maxit) {   /* dune-istl: the outer loop's condition */
      memcpy(b, b2, 8 * n);
      orc_spmv(n, rowptr, colidx, vals, x, tmp, par);
      for (int64_t q = 0; q < n; ++q) b[q] -= tmp[q];
      prec_apply(&Pc, b, V);
      norm = sqrt(dot(n, V, V, par));
    }
  }
  res->iterations_x2 = 2 * (j - 1);
  res->reduction = norm / norm0;
done:
#undef Hh
  free(V); free(w); free(b2); free(tmp); free(H); free(s); free(cs); free(sn); free(yv); free(Pc.dinv);
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
