// =================================================================================================
// Structured simplex grids (the reference's default grid: StructuredGridFactory::createSimplexGrid,
// grid/make_multi_domain_grid.hh:76-100): connectivity and geometry are implicit.
//
// One thread integrates one lattice cell = dim! Kuhn simplices.  The 2^dim corner values are
// loaded once (coalesced along x) instead of (dim+1) gathers per simplex, no coordinates and no
// connectivity are read, and the cell's contributions are pre-summed per corner before one fp64
// reduction per corner value.  fp64 is the bound of this kernel, so the arithmetic is folded:
//  * mass/reaction part per simplex with the symmetric-rule identities of DcElem, accumulated
//    straight into the corner sums;
//  * diffusion part per *cell*: along a Kuhn path the gradient components are edge differences
//    du/dx_k = (u(m|k) - u(m))/h_k, and every lattice edge (m -> m|k) is used by
//    |m|! (d-1-|m|)! of the d! paths, so the cell's stiffness action is a sum over its
//    d 2^(d-1) edges instead of d! simplices x d path edges.
// Same weak form, quadrature points and weights as DcElem (requires point-independent diffusion
// coefficients, which the host checks).  MODE 0: residual, 1: Jacobian apply, 2: block diagonal
// (ns x ns per vertex, block-Jacobi), 3: scalar diagonal into a dof-indexed vector (Jacobi).
#if DC_DIM == 2
#define DC_NPERM 2
#define DC_NCORN 4
__device__ __forceinline__ constexpr int dc_perm(int p, int t) { return p == 0 ? t : 1 - t; }
#else
#define DC_NPERM 6
#define DC_NCORN 8
__device__ __forceinline__ constexpr int dc_perm(int p, int t) {
  // lexicographically enumerated permutations of (0,1,2)
  return p == 0 ? t : p == 1 ? (t == 0 ? 0 : 3 - t) : p == 2 ? (t == 0 ? 1 : t == 1 ? 0 : 2)
       : p == 3 ? (t == 0 ? 1 : t == 1 ? 2 : 0) : p == 4 ? (t == 0 ? 2 : t == 1 ? 0 : 1) : 2 - t;
}
#endif
// corner mask of local vertex k of simplex p: axes p_0..p_{k-1} stepped
__device__ __forceinline__ constexpr int dc_corner(int p, int k) {
  int m = 0;
  for (int t = 0; t < k; ++t) m |= 1 << dc_perm(p, t);
  return m;
}
// number of Kuhn paths through the lattice edge that leaves corner m: |m|! (d-1-|m|)!
__device__ __forceinline__ constexpr int dc_edge_paths(int m) {
  int bits = 0;
  for (int k = 0; k < DC_DIM; ++k) bits += (m >> k) & 1;
  int a = 1, b = 1;
  for (int i = 2; i <= bits; ++i) a *= i;
  for (int i = 2; i <= DC_DIM - 1 - bits; ++i) b *= i;
  return a * b;
}

template <int C, int MODE>
__device__ __forceinline__ void dc_structured_kernel(const DcStructArgs& a) {
  typedef DcComp<C> M;
  constexpr int NS = M::NS;
  constexpr int NV = MODE == 2 ? NS * NS : NS;
  const long long cell = a.cell_begin + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (cell >= a.ncells) return;
  int idx[3];
  {
    long long rem = cell;
    idx[0] = (int)(rem % a.n[0]); rem /= a.n[0];
#if DC_DIM == 3
    idx[1] = (int)(rem % a.n[1]); idx[2] = (int)(rem / a.n[1]);
#else
    idx[1] = (int)rem; idx[2] = 0;
#endif
  }
  long long stride[3] = {1, a.n[0] + 1, (long long)(a.n[0] + 1) * (a.n[1] + 1)};
  long long base = 0;
#pragma unroll
  for (int k = 0; k < DC_DIM; ++k) base += idx[k] * stride[k];
  // ---- corner data
  double U[DC_NCORN][NS], Z[MODE == 1 ? DC_NCORN : 1][NS], acc[DC_NCORN][NV];
  long long dof[DC_NCORN];
#pragma unroll
  for (int m = 0; m < DC_NCORN; ++m) {
    long long v = base;
#pragma unroll
    for (int k = 0; k < DC_DIM; ++k) v += ((m >> k) & 1) * stride[k];
    dof[m] = a.dof_offset + v * NS;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      U[m][s] = a.x[dof[m] + s];
      if (MODE == 1) Z[m][s] = (a.cmask && a.cmask[dof[m] + s]) ? 0.0 : a.z[dof[m] + s];
    }
#pragma unroll
    for (int s = 0; s < NV; ++s) acc[m][s] = 0.0;
  }
  double adet = 1.0, rh[DC_DIM];
#pragma unroll
  for (int k = 0; k < DC_DIM; ++k) { adet *= a.h[k]; rh[k] = 1.0 / a.h[k]; }
  const double f = DC_QW * adet, vol = adet / DC_FACT;
  const double ABf = DC_PAB * f, Bf = DC_PB * f;
  DcCtx c;
  c.time = a.time; c.entity_volume = vol; c.integration_factor = f;
  c.in_volume = 1.0; c.in_boundary = 0.0; c.in_skeleton = 0.0;
  c.nrm[0] = c.nrm[1] = c.nrm[2] = 0.0; c.pos[2] = 0.0;
  double x0[DC_DIM];
#pragma unroll
  for (int k = 0; k < DC_DIM; ++k) x0[k] = a.origin[k] + idx[k] * a.h[k];

  // ---- mass / reaction part, simplex by simplex
#pragma unroll
  for (int p = 0; p < DC_NPERM; ++p) {
    double xl[NS][DC_ND], BS[NS], gu[NS][DC_DIM];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < DC_ND; ++k) { xl[s][k] = U[dc_corner(p, k)][s]; t += xl[s][k]; }
      BS[s] = DC_PB * t;
#pragma unroll
      for (int k = 0; k < DC_DIM; ++k) gu[s][dc_perm(p, k)] = (xl[s][k + 1] - xl[s][k]) * rh[dc_perm(p, k)];
    }
    // sum of the vertex coordinates: axis p_t is stepped by the vertices t+1..d
    double XS[DC_DIM];
#pragma unroll
    for (int t = 0; t < DC_DIM; ++t) XS[dc_perm(p, t)] = DC_ND * x0[dc_perm(p, t)] + (DC_DIM - t) * a.h[dc_perm(p, t)];
    auto set_pos = [&](int q) {
#pragma unroll
      for (int k = 0; k < DC_DIM; ++k) {
        const int v = DC_VQ(q);
        bool stepped = false;
#pragma unroll
        for (int t = 0; t < DC_DIM; ++t) stepped |= (t < v && dc_perm(p, t) == k);
        c.pos[k] = DC_PB * XS[k] + DC_PAB * (x0[k] + (stepped ? a.h[k] : 0.0));
      }
    };
    if (MODE == 0) {
      double T[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) T[s] = 0.0;
#pragma unroll
      for (int q = 0; q < DC_NQ; ++q) {
        double u[NS], sc[NS];
        set_pos(q);
#pragma unroll
        for (int s = 0; s < NS; ++s) u[s] = BS[s] + DC_PAB * xl[s][DC_VQ(q)];
        M::scalar(c, u, gu, a.wM, a.wA, sc);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          T[s] += sc[s];
          acc[dc_corner(p, DC_VQ(q))][s] += ABf * sc[s];
        }
      }
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const double bt = Bf * T[s];
#pragma unroll
        for (int k = 0; k < DC_ND; ++k) acc[dc_corner(p, k)][s] += bt;
      }
    } else if (MODE == 1) {
      double zl[NS][DC_ND], BZ[NS], T[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < DC_ND; ++k) { zl[s][k] = Z[dc_corner(p, k)][s]; t += zl[s][k]; }
        BZ[s] = DC_PB * t;
        T[s] = 0.0;
      }
#pragma unroll
      for (int q = 0; q < DC_NQ; ++q) {
        double u[NS], zq[NS], jm[NS][NS];
        set_pos(q);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          u[s] = BS[s] + DC_PAB * xl[s][DC_VQ(q)];
          zq[s] = BZ[s] + DC_PAB * zl[s][DC_VQ(q)];
        }
        M::jac_mass(c, u, gu, a.wM, a.wA, jm);
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          double w = 0.0;
#pragma unroll
          for (int j = 0; j < NS; ++j)
            if (M::pair(i, j)) w += jm[i][j] * zq[j];
          T[i] += w;
          acc[dc_corner(p, DC_VQ(q))][i] += ABf * w;
        }
      }
#pragma unroll
      for (int i = 0; i < NS; ++i) {
        const double bt = Bf * T[i];
#pragma unroll
        for (int k = 0; k < DC_ND; ++k) acc[dc_corner(p, k)][i] += bt;
      }
    } else {
      double JS[NS][NS], JV[DC_ND][NS][NS];
#pragma unroll
      for (int i = 0; i < NS; ++i)
#pragma unroll
        for (int j = 0; j < NS; ++j) JS[i][j] = 0.0;
#pragma unroll
      for (int q = 0; q < DC_NQ; ++q) {
        double u[NS];
        set_pos(q);
#pragma unroll
        for (int s = 0; s < NS; ++s) u[s] = BS[s] + DC_PAB * xl[s][DC_VQ(q)];
        M::jac_mass(c, u, gu, a.wM, a.wA, JV[DC_VQ(q)]);
#pragma unroll
        for (int i = 0; i < NS; ++i)
#pragma unroll
          for (int j = 0; j < NS; ++j) JS[i][j] += JV[DC_VQ(q)][i][j];
      }
#pragma unroll
      for (int k = 0; k < DC_ND; ++k)
#pragma unroll
        for (int i = 0; i < NS; ++i)
#pragma unroll
          for (int j = 0; j < NS; ++j)
            if (M::pair(i, j) && (MODE == 2 || i == j))
              acc[dc_corner(p, k)][MODE == 2 ? i * NS + j : i] +=
                  (DC_PB * DC_PB * JS[i][j] + (DC_PA * DC_PA - DC_PB * DC_PB) * JV[k][i][j]) * f;
    }
  }

  // ---- diffusion part, lattice edge by lattice edge
  if (M::HAS_DIFF) {
    double u0[NS], g0[NS][DC_DIM], jd[NS][NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      u0[s] = 0.0;
#pragma unroll
      for (int k = 0; k < DC_DIM; ++k) g0[s][k] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < DC_DIM; ++k) c.pos[k] = x0[k] + 0.5 * a.h[k];
    M::jac_diff(c, u0, g0, a.wA, jd);   // jd[i][j] = wA * D_ij (point independent)
#pragma unroll
    for (int m = 0; m < DC_NCORN; ++m)
#pragma unroll
      for (int k = 0; k < DC_DIM; ++k) {
        if ((m >> k) & 1) continue;
        const int m2 = m | (1 << k);
        const double wgt = dc_edge_paths(m) * vol * rh[k] * rh[k];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < NS; ++j)
              if (M::pair(i, j)) {
                acc[m][i * NS + j] += wgt * jd[i][j];
                acc[m2][i * NS + j] += wgt * jd[i][j];
              }
          } else if (MODE == 3) {
            if (M::pair(i, i)) {
              acc[m][i] += wgt * jd[i][i];
              acc[m2][i] += wgt * jd[i][i];
            }
          } else {
            double d = 0.0;
#pragma unroll
            for (int j = 0; j < NS; ++j)
              if (M::pair(i, j)) d += jd[i][j] * (MODE == 0 ? (U[m2][j] - U[m][j]) : (Z[m2][j] - Z[m][j]));
            d *= wgt;
            acc[m2][i] += d;
            acc[m][i] -= d;
          }
        }
      }
  }

  // ---- one reduction per corner value
#pragma unroll
  for (int m = 0; m < DC_NCORN; ++m) {
    if (MODE == 2) {
#pragma unroll
      for (int s = 0; s < NV; ++s) dc_atomic_add(&a.bdiag[dof[m] * NS + s], acc[m][s]);
    } else if (MODE == 3) {   // scalar diagonal, vector layout
#pragma unroll
      for (int s = 0; s < NS; ++s) dc_atomic_add(&a.bdiag[dof[m] + s], acc[m][s]);
    } else {
#pragma unroll
      for (int s = 0; s < NS; ++s) dc_atomic_add(&a.r[dof[m] + s], acc[m][s]);
    }
  }
}
