// =================================================================================================
// Q1 (multilinear) elements on the cells of a structured lattice -- BASELINE configs[3]'s element.
//
// NOT a reference capability: DuneCopasi's basis is PkLocalFiniteElementMap on simplices
// (model_single_compartment_traits.hh:23-24, grid/make_multi_domain_grid.hh:84-90).
// The weak form is the reference's (local_operator.hh:417-491 residual, :541-707 Jacobian,
// :510-524 apply) with the element swapped: corner m of a cell sits at the bit pattern of m
// (x = bit 0), the rule is dune-geometry's order-2 rule of a cube, i.e. 2 Gauss points per axis.
// Checked against the oracle's own Q1 element (oracle.c, etype 1).
//
// One thread integrates one cell.  fp64 is the bound (as for the simplex kernels), so everything
// is done by sum factorisation with the symmetric 1-D matrix [[A, B], [B, A]] (A, B = the two
// linear shape functions at the near Gauss point): corner values -> point values, and point
// integrands -> corner sums, cost d 2^d multiply-adds per field each instead of 4^d.  Diffusion
// coefficients are point independent here (checked by the host), so the stiffness action is the
// exact tensor form  sum_k |cell|/h_k^2  K1_k (x) M1 (x) M1  applied to the corner values.
// MODE 0: residual, 1: Jacobian apply, 2: block diagonal, 3: scalar diagonal, 4: CSR values.
#define DQ_NC (1 << DC_DIM)   // == DC_NCORN of assembly_structured.cuh, whose drivers these kernels share
#define DQ_A 0.78867513459481288225
#define DQ_B 0.21132486540518711775

// v[m] <- sum_m' prod_k (bit_k(m) == bit_k(m') ? pa : pb) v[m'], in place
__device__ __forceinline__ void dq_tensor(double* v, const double pa, const double pb) {
#pragma unroll
  for (int k = 0; k < DC_DIM; ++k)
#pragma unroll
    for (int m = 0; m < DQ_NC; ++m) {
      if ((m >> k) & 1) continue;
      const double lo = v[m], hi = v[m | (1 << k)];
      v[m] = pa * lo + pb * hi;
      v[m | (1 << k)] = pb * lo + pa * hi;
    }
}

// the same map for (pa, pb) = (A, B): A + B = 1, so  lo' = lo + B (hi - lo),  hi' = hi - B (hi - lo)
// -- three fp64 operations per pair instead of four
__device__ __forceinline__ void dq_interp(double* v) {
#pragma unroll
  for (int k = 0; k < DC_DIM; ++k)
#pragma unroll
    for (int m = 0; m < DQ_NC; ++m) {
      if ((m >> k) & 1) continue;
      const double lo = v[m], hi = v[m | (1 << k)], d = hi - lo;
      v[m] = lo + DQ_B * d;
      v[m | (1 << k)] = hi - DQ_B * d;
    }
}

// d/dx_k of the multilinear interpolant of the corner values v at the 2^d Gauss points
__device__ __forceinline__ void dq_gradient(const double* v, int k, double rh, double* g) {
#pragma unroll
  for (int m = 0; m < DQ_NC; ++m) {
    if ((m >> k) & 1) continue;
    g[m] = g[m | (1 << k)] = (v[m | (1 << k)] - v[m]) * rh;
  }
#pragma unroll
  for (int l = 0; l < DC_DIM; ++l) {
    if (l == k) continue;
#pragma unroll
    for (int m = 0; m < DQ_NC; ++m) {
      if ((m >> l) & 1) continue;
      const double lo = g[m], hi = g[m | (1 << l)], d = hi - lo;
      g[m] = lo + DQ_B * d;
      g[m | (1 << l)] = hi - DQ_B * d;
    }
  }
}

// y += sum_k wk[k] (K1 along k) (x) (M1 along the other axes) v ; K1 = [[1,-1],[-1,1]], M1 = [[1/3,1/6],[1/6,1/3]]
__device__ __forceinline__ void dq_stiffness(const double* v, const double* wk, double scale, double* y) {
#pragma unroll
  for (int k = 0; k < DC_DIM; ++k) {
    double t[DQ_NC];
#pragma unroll
    for (int m = 0; m < DQ_NC; ++m)
      if (!((m >> k) & 1)) t[m] = v[m | (1 << k)] - v[m];
#pragma unroll
    for (int l = 0; l < DC_DIM; ++l) {
      if (l == k) continue;
#pragma unroll
      for (int m = 0; m < DQ_NC; ++m) {
        if (((m >> k) & 1) || ((m >> l) & 1)) continue;
        // 6 M1 = [[2, 1], [1, 2]]: additions only, the factors 1/6 go into the weight
        const double lo = t[m], hi = t[m | (1 << l)], sum = lo + hi;
        t[m] = lo + sum;
        t[m | (1 << l)] = hi + sum;
      }
    }
    const double w = wk[k] * scale * (DC_DIM == 3 ? 1.0 / 36.0 : 1.0 / 6.0);
#pragma unroll
    for (int m = 0; m < DQ_NC; ++m)
      if (!((m >> k) & 1)) {
        y[m] -= w * t[m];
        y[m | (1 << k)] += w * t[m];
      }
  }
}

// shape function of corner b at Gauss point q
__device__ __forceinline__ constexpr double dq_phi(int q, int b) {
  double v = 1.0;
  for (int k = 0; k < DC_DIM; ++k) v *= (((q ^ b) >> k) & 1) ? DQ_B : DQ_A;
  return v;
}
// entry (a, b) of the axis-k factor K1_k (x) M1 (x) M1 of the cell stiffness matrix
__device__ __forceinline__ constexpr double dq_kfactor(int k, int a, int b) {
  double v = (((a ^ b) >> k) & 1) ? -1.0 : 1.0;
  for (int l = 0; l < DC_DIM; ++l)
    if (l != k) v *= (((a ^ b) >> l) & 1) ? (1.0 / 6.0) : (1.0 / 3.0);
  return v;
}

// per-cell constants and the evaluation context
struct DqGeo {
  double adet, f, rh[DC_DIM], wk[DC_DIM], x0[DC_DIM];
};
__device__ __forceinline__ void dq_setup(const DcStructArgs& a, const int* idx, DqGeo& g, DcCtx& c) {
  g.adet = 1.0;
#pragma unroll
  for (int k = 0; k < DC_DIM; ++k) { g.adet *= a.h[k]; g.rh[k] = a.rh[k]; }
#pragma unroll
  for (int k = 0; k < DC_DIM; ++k) g.wk[k] = g.adet * g.rh[k] * g.rh[k];
  g.f = g.adet / DQ_NC;   // weight 2^-d per point times |det| = |cell|
  c.time = a.time; c.entity_volume = g.adet; c.integration_factor = g.f;
  c.in_volume = 1.0; c.in_boundary = 0.0; c.in_skeleton = 0.0;
  c.nrm[0] = c.nrm[1] = c.nrm[2] = 0.0; c.pos[2] = 0.0;
#pragma unroll
  for (int k = 0; k < DC_DIM; ++k) g.x0[k] = a.origin[k] + idx[k] * a.h[k];
}
__device__ __forceinline__ void dq_set_pos(const DcStructArgs& a, const DqGeo& g, DcCtx& c, int q) {
#pragma unroll
  for (int k = 0; k < DC_DIM; ++k) c.pos[k] = g.x0[k] + (((q >> k) & 1) ? DQ_A : DQ_B) * a.h[k];
}
// point-independent diffusion coefficients, jd[i][j] = wA * D_ij
template <int C>
__device__ __forceinline__ void dq_diffusion(const DcStructArgs& a, const DqGeo& g, DcCtx& c,
                                             double (*jd)[DcComp<C>::NS]) {
  typedef DcComp<C> M;
  constexpr int NS = M::NS;
  double u0[NS], g0[NS][DC_DIM];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    u0[s] = 0.0;
#pragma unroll
    for (int k = 0; k < DC_DIM; ++k) g0[s][k] = 0.0;
  }
#pragma unroll
  for (int k = 0; k < DC_DIM; ++k) c.pos[k] = g.x0[k] + 0.5 * a.h[k];
  M::jac_diff(c, u0, g0, a.wA, jd);
}

// residual (MODE 0) / Jacobian apply (MODE 1) of one cell: Uc/Zc = corner values [corner][species],
// acc = corner sums [corner][species] (overwritten)
template <int C, int MODE>
__device__ __forceinline__ void dc_q1_cell(const DcStructArgs& a, const int* idx, const double (*Uc)[DcComp<C>::NS],
                                           const double (*Zc)[DcComp<C>::NS], double (*out)[DcComp<C>::NS]) {
  typedef DcComp<C> M;
  constexpr int NS = M::NS;
  DqGeo g;
  DcCtx c;
  dq_setup(a, idx, g, c);
  // species-major copies so that the tensor maps run over contiguous registers
  double U[NS][DQ_NC], Z[MODE == 1 ? NS : 1][DQ_NC];
#pragma unroll
  for (int m = 0; m < DQ_NC; ++m)
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      U[s][m] = Uc[m][s];
      if (MODE == 1) Z[s][m] = Zc[m][s];
    }
  double jd[NS][NS];
  if (M::HAS_DIFF) dq_diffusion<C>(a, g, c, jd);
  // gradients at the Gauss points (only models whose coefficients read grad_* keep this alive)
  double G[DQ_NC][NS][DC_DIM];
#pragma unroll
  for (int s = 0; s < NS; ++s)
#pragma unroll
    for (int k = 0; k < DC_DIM; ++k) {
      double gr[DQ_NC];
      dq_gradient(U[s], k, g.rh[k], gr);
#pragma unroll
      for (int q = 0; q < DQ_NC; ++q) G[q][s][k] = gr[q];
    }
  // ---- diffusion first (needs the corner values), then the corner arrays turn into point arrays
  double acc[NS][DQ_NC];
#pragma unroll
  for (int i = 0; i < NS; ++i) {
#pragma unroll
    for (int m = 0; m < DQ_NC; ++m) acc[i][m] = 0.0;
    if (M::HAS_DIFF) {
#pragma unroll
      for (int j = 0; j < NS; ++j)
        if (M::dpair(i, j)) dq_stiffness(MODE == 0 ? U[j] : Z[j], g.wk, jd[i][j], acc[i]);
    }
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    dq_interp(U[s]);
    if (MODE == 1) dq_interp(Z[s]);
  }
#pragma unroll
  for (int q = 0; q < DQ_NC; ++q) {
    double u[NS];
    dq_set_pos(a, g, c, q);
#pragma unroll
    for (int s = 0; s < NS; ++s) u[s] = U[s][q];
    if (MODE == 0) {
      double sc[NS];
      M::scalar(c, u, G[q], a.wM, a.wA, sc);
#pragma unroll
      for (int s = 0; s < NS; ++s) U[s][q] = sc[s];
    } else {
      double jm[NS][NS], w[NS];
      M::jac_mass(c, u, G[q], a.wM, a.wA, jm);
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        w[i] = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j)
          if (M::pair(i, j)) w[i] += jm[i][j] * Z[j][q];
      }
#pragma unroll
      for (int i = 0; i < NS; ++i) Z[i][q] = w[i];
    }
  }
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    double* S = MODE == 0 ? U[s] : Z[s];
    dq_interp(S);
#pragma unroll
    for (int m = 0; m < DQ_NC; ++m) out[m][s] = acc[s][m] + g.f * S[m];
  }
}

// MODE 2: block diagonal, 3: scalar diagonal, 4: CSR values -- one thread per cell
template <int C, int MODE>
__device__ __forceinline__ void dc_q1_matrix_kernel(const DcStructArgs& a) {
  typedef DcComp<C> M;
  constexpr int NS = M::NS;
  const long long cell = a.cell_begin + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (cell >= a.ncells) return;
  int idx[3];
  dc_cell_index(a, cell, idx);
  long long stride[3] = {1, a.n[0] + 1, (long long)(a.n[0] + 1) * (a.n[1] + 1)};
  long long base = 0;
#pragma unroll
  for (int k = 0; k < DC_DIM; ++k) base += idx[k] * stride[k];
  double U[NS][DQ_NC];
  long long dof[DQ_NC];
#pragma unroll
  for (int m = 0; m < DQ_NC; ++m) {
    long long v = base;
#pragma unroll
    for (int k = 0; k < DC_DIM; ++k) v += ((m >> k) & 1) * stride[k];
    dof[m] = a.dof_offset + v * NS;
#pragma unroll
    for (int s = 0; s < NS; ++s) U[s][m] = a.x[dof[m] + s];
  }
  DqGeo g;
  DcCtx c;
  dq_setup(a, idx, g, c);
  const double f = g.f;
  const double* wk = g.wk;
  double jd[NS][NS];
  if (M::HAS_DIFF) dq_diffusion<C>(a, g, c, jd);
  double G[DQ_NC][NS][DC_DIM];
#pragma unroll
  for (int s = 0; s < NS; ++s)
#pragma unroll
    for (int k = 0; k < DC_DIM; ++k) {
      double gr[DQ_NC];
      dq_gradient(U[s], k, g.rh[k], gr);
#pragma unroll
      for (int q = 0; q < DQ_NC; ++q) G[q][s][k] = gr[q];
    }
  // ---- Jacobian coefficients at the Gauss points
  double J[NS][NS][DQ_NC];
#pragma unroll
  for (int s = 0; s < NS; ++s) dq_interp(U[s]);
#pragma unroll
  for (int q = 0; q < DQ_NC; ++q) {
    double u[NS], jm[NS][NS];
    dq_set_pos(a, g, c, q);
#pragma unroll
    for (int s = 0; s < NS; ++s) u[s] = U[s][q];
    M::jac_mass(c, u, G[q], a.wM, a.wA, jm);
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
      for (int j = 0; j < NS; ++j) J[i][j][q] = jm[i][j];
  }
  if (MODE == 2 || MODE == 3) {
    // diagonal entries: sum_q J(q) phi_m(q)^2, plus the (corner independent) stiffness diagonal
    double kd = 0.0;
#pragma unroll
    for (int k = 0; k < DC_DIM; ++k) kd += wk[k] * dq_kfactor(k, 0, 0);
#pragma unroll
    for (int i = 0; i < NS; ++i)
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        if (!M::pair(i, j) || (MODE == 3 && i != j)) continue;
        dq_tensor(J[i][j], DQ_A * DQ_A, DQ_B * DQ_B);
        const double d = M::HAS_DIFF ? jd[i][j] * kd : 0.0;
#pragma unroll
        for (int m = 0; m < DQ_NC; ++m) {
          const double v = f * J[i][j][m] + d;
          if (MODE == 2) dc_atomic_add(&a.bdiag[dof[m] * NS + i * NS + j], v);
          else dc_atomic_add(&a.bdiag[dof[m] + i], v);
        }
      }
    return;
  }
  // ---- MODE 4: CSR values, entry (i, a; j, b) = f sum_q J_ij(q) phi_a(q) phi_b(q) + jd_ij K_ab
#pragma unroll
  for (int i = 0; i < NS; ++i)
#pragma unroll
    for (int b = 0; b < DQ_NC; ++b) {
      double T[NS][DQ_NC];
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        if (!M::pair(i, j)) continue;
#pragma unroll
        for (int q = 0; q < DQ_NC; ++q) T[j][q] = J[i][j][q] * dq_phi(q, b);
        dq_interp(T[j]);
      }
#pragma unroll
      for (int m = 0; m < DQ_NC; ++m) {
        double kab = 0.0;
#pragma unroll
        for (int k = 0; k < DC_DIM; ++k) kab += wk[k] * dq_kfactor(k, m, b);
        // the coupled species of corner b sit next to each other in row (m, i): one search per block row
        long long p = -1;
        const int row = (int)dof[m] + i;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          if (!M::pair(i, j)) continue;
          const double v = f * T[j][m] + (M::HAS_DIFF ? jd[i][j] * kab : 0.0);
          const int col = (int)dof[b] + j;
          if (p >= 0 && p + 1 < a.rowptr[row + 1] && a.colidx[p + 1] == col) ++p;
          else p = dc_csr_find(a.rowptr, a.colidx, row, col);
          if (p >= 0) dc_atomic_add(&a.vals[p], v);
        }
      }
    }
}

template <int C, int MODE, bool SCALED = false, bool NOMASK = false>
__device__ __forceinline__ void dc_q1_kernel(const DcStructArgs& a) {
  constexpr int NS = DcComp<C>::NS;
  if constexpr (MODE >= 2) {
    dc_q1_matrix_kernel<C, MODE>(a);
  } else {
    auto cell = [&](const int* idx, const double (*U)[NS], const double (*Z)[NS], double (*acc)[NS]) {
      dc_q1_cell<C, MODE>(a, idx, U, Z, acc);
    };
    dc_struct_per_cell<C, MODE, decltype(cell), SCALED, NOMASK>(a, cell);
  }
}
template <int C, int MODE>
__device__ __forceinline__ void dc_q1_march_kernel(const DcStructArgs& a) {
  constexpr int NS = DcComp<C>::NS;
  dc_struct_march<C, MODE>(a, [&](const int* idx, const double (*U)[NS], const double (*Z)[NS], double (*acc)[NS]) {
    dc_q1_cell<C, MODE>(a, idx, U, Z, acc);
  });
}
