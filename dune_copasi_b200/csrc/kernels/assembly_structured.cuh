// =================================================================================================
// Structured simplex grids (the reference's default grid: StructuredGridFactory::createSimplexGrid,
// grid/make_multi_domain_grid.hh:76-100): connectivity and geometry are implicit.
//
// One thread integrates one lattice cell = dim! Kuhn simplices.  The 2^dim corner values are
// loaded once (coalesced along x) instead of (dim+1) gathers per simplex, no coordinates and no
// connectivity are read, the gradients along a Kuhn path are plain edge differences:
//     du/dx_{p_t} = (u_{t+1} - u_t) / h_{p_t},   |det| = prod h,
// and the cell's contributions are pre-summed per corner before one fp64 reduction per corner
// value.  Same weak form, same quadrature points and weights as DcElem.
// MODE 0: residual, 1: Jacobian apply, 2: block diagonal.
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

template <int C, int MODE>
__device__ __forceinline__ void dc_structured_kernel(const DcStructArgs& a) {
  typedef DcComp<C> M;
  constexpr int NS = M::NS;
  constexpr int NV = MODE == 2 ? NS * NS : NS;
  const long long cell = blockIdx.x * (long long)blockDim.x + threadIdx.x;
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
  DcCtx c;
  c.time = a.time; c.entity_volume = vol; c.integration_factor = f;
  c.in_volume = 1.0; c.in_boundary = 0.0; c.in_skeleton = 0.0;
  c.nrm[0] = c.nrm[1] = c.nrm[2] = 0.0; c.pos[2] = 0.0;
  double x0[DC_DIM];
#pragma unroll
  for (int k = 0; k < DC_DIM; ++k) x0[k] = a.origin[k] + idx[k] * a.h[k];

#pragma unroll
  for (int p = 0; p < DC_NPERM; ++p) {
    // vertex k of this simplex sits at corner dc_corner(p,k); its coordinates are x0 + h on the
    // axes stepped so far
    double xl[NS][DC_ND], S[NS], gu[NS][DC_DIM];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < DC_ND; ++k) { xl[s][k] = U[dc_corner(p, k)][s]; t += xl[s][k]; }
      S[s] = t;
#pragma unroll
      for (int k = 0; k < DC_DIM; ++k) gu[s][dc_perm(p, k)] = (xl[s][k + 1] - xl[s][k]) * rh[dc_perm(p, k)];
    }
    // sum of the vertex coordinates: axis p_t is stepped by the vertices t+1..d
    double XS[DC_DIM];
#pragma unroll
    for (int t = 0; t < DC_DIM; ++t) XS[dc_perm(p, t)] = DC_ND * x0[dc_perm(p, t)] + (DC_DIM - t) * a.h[dc_perm(p, t)];
    auto set_pos = [&](int q) {
      // X of vertex v: x0 + h on axes p_0..p_{v-1}
#pragma unroll
      for (int k = 0; k < DC_DIM; ++k) {
        const int v = DC_VQ(q);
        bool stepped = false;
#pragma unroll
        for (int t = 0; t < DC_DIM; ++t) stepped |= (t < v && dc_perm(p, t) == k);
        c.pos[k] = DC_PB * XS[k] + DC_PAB * (x0[k] + (stepped ? a.h[k] : 0.0));
      }
    };
    double loc[NV == NS ? NS : 1][DC_ND];
    if (MODE == 0) {
      double T[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) T[s] = 0.0;
#pragma unroll
      for (int q = 0; q < DC_NQ; ++q) {
        double u[NS], sc[NS];
        set_pos(q);
#pragma unroll
        for (int s = 0; s < NS; ++s) u[s] = DC_PB * S[s] + DC_PAB * xl[s][DC_VQ(q)];
        M::scalar(c, u, gu, a.wM, a.wA, sc);
#pragma unroll
        for (int s = 0; s < NS; ++s) { T[s] += sc[s]; loc[s][DC_VQ(q)] = DC_PAB * sc[s]; }
      }
#pragma unroll
      for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int k = 0; k < DC_ND; ++k) loc[s][k] = (loc[s][k] + DC_PB * T[s]) * f;
      if (M::HAS_DIFF) {
        double u0[NS], fl[NS][DC_DIM];
#pragma unroll
        for (int s = 0; s < NS; ++s) u0[s] = 0.0;
#pragma unroll
        for (int k = 0; k < DC_DIM; ++k) c.pos[k] = XS[k] / DC_ND;
        M::flux(c, u0, gu, a.wA, fl);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          // fl . grad(phi_k) |T| with grad(phi_0) = -e_{p0}/h, grad(phi_t) = e_{p(t-1)}/h - e_{pt}/h
          double w[DC_DIM];
#pragma unroll
          for (int t = 0; t < DC_DIM; ++t) w[t] = fl[s][dc_perm(p, t)] * rh[dc_perm(p, t)] * vol;
          loc[s][0] += w[0];
#pragma unroll
          for (int t = 1; t < DC_DIM; ++t) loc[s][t] -= w[t - 1] - w[t];
          loc[s][DC_DIM] -= w[DC_DIM - 1];
        }
      }
    } else if (MODE == 1) {
      double zl[NS][DC_ND], ZS[NS], gz[NS][DC_DIM], T[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < DC_ND; ++k) { zl[s][k] = Z[dc_corner(p, k)][s]; t += zl[s][k]; }
        ZS[s] = t;
        T[s] = 0.0;
#pragma unroll
        for (int k = 0; k < DC_DIM; ++k) gz[s][dc_perm(p, k)] = (zl[s][k + 1] - zl[s][k]) * rh[dc_perm(p, k)];
      }
#pragma unroll
      for (int q = 0; q < DC_NQ; ++q) {
        double u[NS], zq[NS], jm[NS][NS];
        set_pos(q);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          u[s] = DC_PB * S[s] + DC_PAB * xl[s][DC_VQ(q)];
          zq[s] = DC_PB * ZS[s] + DC_PAB * zl[s][DC_VQ(q)];
        }
        M::jac_mass(c, u, gu, a.wM, a.wA, jm);
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          double w = 0.0;
#pragma unroll
          for (int j = 0; j < NS; ++j)
            if (M::pair(i, j)) w += jm[i][j] * zq[j];
          T[i] += w;
          loc[i][DC_VQ(q)] = DC_PAB * w;
        }
      }
#pragma unroll
      for (int i = 0; i < NS; ++i)
#pragma unroll
        for (int k = 0; k < DC_ND; ++k) loc[i][k] = (loc[i][k] + DC_PB * T[i]) * f;
      if (M::HAS_DIFF) {
        double u0[NS], jd[NS][NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) u0[s] = 0.0;
#pragma unroll
        for (int k = 0; k < DC_DIM; ++k) c.pos[k] = XS[k] / DC_ND;
        M::jac_diff(c, u0, gu, a.wA, jd);
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          double w[DC_DIM];
#pragma unroll
          for (int t = 0; t < DC_DIM; ++t) {
            double fl = 0.0;
#pragma unroll
            for (int j = 0; j < NS; ++j)
              if (M::pair(i, j)) fl += jd[i][j] * gz[j][dc_perm(p, t)];
            w[t] = fl * rh[dc_perm(p, t)] * vol;
          }
          loc[i][0] -= w[0];
#pragma unroll
          for (int t = 1; t < DC_DIM; ++t) loc[i][t] += w[t - 1] - w[t];
          loc[i][DC_DIM] += w[DC_DIM - 1];
        }
      }
    }
    if (MODE != 2) {
#pragma unroll
      for (int k = 0; k < DC_ND; ++k)
#pragma unroll
        for (int s = 0; s < NS; ++s) acc[dc_corner(p, k)][s] += loc[s][k];
    } else {
      double JS[NS][NS], JV[DC_ND][NS][NS], DD[NS][NS];
#pragma unroll
      for (int i = 0; i < NS; ++i)
#pragma unroll
        for (int j = 0; j < NS; ++j) { JS[i][j] = 0.0; DD[i][j] = 0.0; }
#pragma unroll
      for (int q = 0; q < DC_NQ; ++q) {
        double u[NS];
        set_pos(q);
#pragma unroll
        for (int s = 0; s < NS; ++s) u[s] = DC_PB * S[s] + DC_PAB * xl[s][DC_VQ(q)];
        M::jac_mass(c, u, gu, a.wM, a.wA, JV[DC_VQ(q)]);
#pragma unroll
        for (int i = 0; i < NS; ++i)
#pragma unroll
          for (int j = 0; j < NS; ++j) JS[i][j] += JV[DC_VQ(q)][i][j];
      }
      if (M::HAS_DIFF) {
        double u0[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) u0[s] = 0.0;
#pragma unroll
        for (int k = 0; k < DC_DIM; ++k) c.pos[k] = XS[k] / DC_ND;
        M::jac_diff(c, u0, gu, a.wA, DD);
      }
#pragma unroll
      for (int k = 0; k < DC_ND; ++k) {
        // |grad phi_k|^2
        double gg = 0.0;
        if (k > 0) gg += rh[dc_perm(p, k - 1)] * rh[dc_perm(p, k - 1)];
        if (k < DC_DIM) gg += rh[dc_perm(p, k)] * rh[dc_perm(p, k)];
#pragma unroll
        for (int i = 0; i < NS; ++i)
#pragma unroll
          for (int j = 0; j < NS; ++j)
            if (M::pair(i, j))
              acc[dc_corner(p, k)][i * NS + j] +=
                  (DC_PB * DC_PB * JS[i][j] + (DC_PA * DC_PA - DC_PB * DC_PB) * JV[k][i][j]) * f + DD[i][j] * gg * vol;
      }
    }
  }
  // ---- one reduction per corner value
#pragma unroll
  for (int m = 0; m < DC_NCORN; ++m) {
    if (MODE == 2) {
#pragma unroll
      for (int s = 0; s < NV; ++s) dc_atomic_add(&a.bdiag[dof[m] * NS + s], acc[m][s]);
    } else {
#pragma unroll
      for (int s = 0; s < NS; ++s) dc_atomic_add(&a.r[dof[m] + s], acc[m][s]);
    }
  }
}
