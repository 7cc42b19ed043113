// ---- facets: transmission conditions / outflow boundary ------------------------------------------
// One thread per (facet, side) entry of the list built for the directional pair P (CS -> CT).
// Everything the quadrature loop needs is gathered once in load() with compile-time indexing only
// (the facet's vertices are selected by predicated copies, not by runtime subscripts), so that the
// per-facet state stays in registers in the unrolled residual / apply kernels.
template <int P>
struct DcFacet {
  typedef DcOutflow<P> O;
  static constexpr int NSS = O::NSS, NST = O::NST;
  double Xf[DC_DIM][DC_DIM];                  // coordinates of the facet's vertices
  double Gsf[DC_DIM][DC_DIM], Gtf[DC_DIM][DC_DIM];   // shape-function gradients of those vertices on both sides
  int dofs[DC_DIM], doft[DC_DIM];             // dof of species 0 at the facet vertices
  double xs[NSS][DC_DIM], xt[NST][DC_DIM];    // coefficients at the facet vertices
  double gs[NSS][DC_DIM], gt[NST][DC_DIM];    // gradients (element-wise constants)
  double area, ie;
  bool self_inside;                           // this side is the facet's inside element (the lower element id)
  DcCtx c;

  __device__ __forceinline__ bool load(const DcFacetArgs& a, long long* fout) {
    const long long f = (blockIdx.x - a.block_offset) * (long long)blockDim.x + threadIdx.x;
    if (f >= a.n) return false;
    *fout = f;
    const long long es = a.f_self[f], et = a.f_other[f];
    const int ms = a.f_lself[f];
    self_inside = et < 0 || es < et;
    int vs[DC_ND], ds[DC_ND], gvf[DC_DIM];
    double Xs[DC_ND][DC_DIM], Gs[DC_ND][DC_DIM], xall[NSS][DC_ND];
#pragma unroll
    for (int k = 0; k < DC_ND; ++k) {
      vs[k] = a.elems[es * DC_ND + k];
#pragma unroll
      for (int cc = 0; cc < DC_DIM; ++cc) Xs[k][cc] = a.coords[(long long)vs[k] * DC_DIM + cc];
      ds[k] = a.vdof_s ? a.vdof_s[vs[k]] : a.dof_offset_s + vs[k] * NSS;
#pragma unroll
      for (int s = 0; s < NSS; ++s) xall[s][k] = a.x[ds[k] + s];
    }
    dc_geometry(Xs, Gs);
#pragma unroll
    for (int s = 0; s < NSS; ++s)
#pragma unroll
      for (int k = 0; k < DC_DIM; ++k) {
        double acc = 0.0;
#pragma unroll
        for (int b = 0; b < DC_ND; ++b) acc += xall[s][b] * Gs[b][k];
        gs[s][k] = acc;
      }
    // facet vertices = the element's vertices except the opposite one, in ascending local order:
    // slot n holds vertex n (n < ms) or n + 1 -- written as selects between two compile-time
    // subscripts so that none of the arrays is addressed dynamically
    double gm[DC_DIM];   // gradient of the opposite vertex's shape function
#pragma unroll
    for (int cc = 0; cc < DC_DIM; ++cc) {
      double g = Gs[0][cc];
#pragma unroll
      for (int k = 1; k < DC_ND; ++k) g = ms == k ? Gs[k][cc] : g;
      gm[cc] = g;
    }
#pragma unroll
    for (int n = 0; n < DC_DIM; ++n) {
      const bool lo = n < ms;
      gvf[n] = lo ? vs[n] : vs[n + 1];
      dofs[n] = lo ? ds[n] : ds[n + 1];
#pragma unroll
      for (int cc = 0; cc < DC_DIM; ++cc) {
        Xf[n][cc] = lo ? Xs[n][cc] : Xs[n + 1][cc];
        Gsf[n][cc] = lo ? Gs[n][cc] : Gs[n + 1][cc];
      }
#pragma unroll
      for (int s = 0; s < NSS; ++s) xs[s][n] = lo ? xall[s][n] : xall[s][n + 1];
    }
    // other side
#pragma unroll
    for (int s = 0; s < NST; ++s)
#pragma unroll
      for (int k = 0; k < DC_DIM; ++k) { xt[s][k] = 0.0; gt[s][k] = 0.0; }
#pragma unroll
    for (int n = 0; n < DC_DIM; ++n) {
      doft[n] = 0;
#pragma unroll
      for (int cc = 0; cc < DC_DIM; ++cc) Gtf[n][cc] = 0.0;
    }
    if (!O::BOUNDARY && et >= 0) {
      double Xt[DC_ND][DC_DIM], Gt[DC_ND][DC_DIM], xtall[NST][DC_ND];
      int vt[DC_ND], dt[DC_ND];
#pragma unroll
      for (int k = 0; k < DC_ND; ++k) {
        vt[k] = a.elems[et * DC_ND + k];
#pragma unroll
        for (int cc = 0; cc < DC_DIM; ++cc) Xt[k][cc] = a.coords[(long long)vt[k] * DC_DIM + cc];
        dt[k] = a.vdof_t ? a.vdof_t[vt[k]] : a.dof_offset_t + vt[k] * O::NST_REAL;
#pragma unroll
        for (int s = 0; s < O::NST_REAL; ++s) xtall[s][k] = a.x[dt[k] + s];
      }
#if DC_REF_COMPAT
      // local_operator.hh:903-916 / :939-941 as written: the coefficients of the element across the facet are
      // paired, local index by local index, with this element's shape functions and gradients
      (void)Xt; (void)Gt;
#pragma unroll
      for (int s = 0; s < O::NST_REAL; ++s)
#pragma unroll
        for (int k = 0; k < DC_DIM; ++k) {
          double acc = 0.0;
#pragma unroll
          for (int b = 0; b < DC_ND; ++b) acc += xtall[s][b] * Gs[b][k];
          gt[s][k] = acc;
        }
#pragma unroll
      for (int n = 0; n < DC_DIM; ++n) {
        const bool lo = n < ms;
        doft[n] = lo ? dt[n] : dt[n + 1];
#pragma unroll
        for (int s = 0; s < O::NST_REAL; ++s) xt[s][n] = lo ? xtall[s][n] : xtall[s][n + 1];
#pragma unroll
        for (int cc = 0; cc < DC_DIM; ++cc) Gtf[n][cc] = Gsf[n][cc];
      }
#else
      dc_geometry(Xt, Gt);
#pragma unroll
      for (int s = 0; s < O::NST_REAL; ++s)
#pragma unroll
        for (int k = 0; k < DC_DIM; ++k) {
          double acc = 0.0;
#pragma unroll
          for (int b = 0; b < DC_ND; ++b) acc += xtall[s][b] * Gt[b][k];
          gt[s][k] = acc;
        }
      // match the facet vertices by global id
#pragma unroll
      for (int n = 0; n < DC_DIM; ++n) {
        doft[n] = a.vdof_t ? a.vdof_t[gvf[n]] : a.dof_offset_t + gvf[n] * O::NST_REAL;
#pragma unroll
        for (int k = 0; k < DC_ND; ++k)
          if (vt[k] == gvf[n]) {
#pragma unroll
            for (int s = 0; s < O::NST_REAL; ++s) xt[s][n] = xtall[s][k];
#pragma unroll
            for (int cc = 0; cc < DC_DIM; ++cc) Gtf[n][cc] = Gt[k][cc];
          }
      }
#endif
    }
    // facet measure and unit outer normal of the own side: -grad(phi_m)/|grad(phi_m)|
    double nn = 0.0;
#pragma unroll
    for (int k = 0; k < DC_DIM; ++k) nn += gm[k] * gm[k];
    nn = sqrt(nn);
#pragma unroll
    for (int k = 0; k < 3; ++k) c.nrm[k] = k < DC_DIM ? -gm[k < DC_DIM ? k : 0] / nn : 0.0;
#if DC_DIM == 2
    {
      const double dx = Xf[1][0] - Xf[0][0], dy = Xf[1][1] - Xf[0][1];
      area = sqrt(dx * dx + dy * dy);
      ie = area;
    }
#else
    {
      double u[3], v[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) { u[k] = Xf[1][k] - Xf[0][k]; v[k] = Xf[2][k] - Xf[0][k]; }
      const double cx = u[1] * v[2] - u[2] * v[1], cy = u[2] * v[0] - u[0] * v[2], cz = u[0] * v[1] - u[1] * v[0];
      area = 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
      ie = 2.0 * area;
    }
#endif
    c.time = a.time;
    c.entity_volume = area;
    c.in_volume = 0.0;
    c.in_boundary = O::BOUNDARY ? 1.0 : 0.0;
    c.in_skeleton = O::BOUNDARY ? 0.0 : 1.0;
    c.pos[2] = 0.0;
#pragma unroll
    for (int k = 0; k < DC_NKEYS; ++k) c.cell[k] = a.cell[(long long)k * a.ne_total + es];
    return true;
  }

  // facet quadrature (order 2) in facet barycentric coordinates
  __device__ __forceinline__ static double lam(int q, int m) {
#if DC_DIM == 2
    const double g = 0.28867513459481288225;
    return (m == 0) == (q == 0) ? 0.5 + g : 0.5 - g;
#else
    return (q == 0 && m == 1) || (q == 1 && m == 2) || (q == 2 && m == 0) ? (4.0 / 6.0) : (1.0 / 6.0);
#endif
  }
  __device__ __forceinline__ static double qw() { return DC_DIM == 2 ? 0.5 : 1.0 / 6.0; }

  __device__ __forceinline__ double point(int q, double* us, double* ut) {
#pragma unroll
    for (int k = 0; k < DC_DIM; ++k) {
      double p = 0.0;
#pragma unroll
      for (int m = 0; m < DC_DIM; ++m) p += lam(q, m) * Xf[m][k];
      c.pos[k] = p;
    }
#pragma unroll
    for (int s = 0; s < NSS; ++s) {
      double v = 0.0;
#pragma unroll
      for (int m = 0; m < DC_DIM; ++m) v += xs[s][m] * lam(q, m);
      us[s] = v;
    }
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      double v = 0.0;
#pragma unroll
      for (int m = 0; m < DC_DIM; ++m) v += xt[s][m] * lam(q, m);
      ut[s] = v;
    }
    const double factor = qw() * ie;
    c.integration_factor = factor;
    return factor;
  }

  // own-side facet residual loc[s][m] = wA sum_q T_s lam_m factor   (local_operator.hh:920-927)
  __device__ __forceinline__ void residual(double wA, double (*loc)[DC_DIM]) {
#pragma unroll
    for (int s = 0; s < NSS; ++s)
#pragma unroll
      for (int m = 0; m < DC_DIM; ++m) loc[s][m] = 0.0;
#pragma unroll
    for (int q = 0; q < DC_DIM; ++q) {
      double us[NSS], ut[NST], T[NSS];
      const double factor = point(q, us, ut);
      O::flux(c, us, gs, ut, gt, T);
#pragma unroll
      for (int s = 0; s < NSS; ++s)
#pragma unroll
        for (int m = 0; m < DC_DIM; ++m) loc[s][m] += wA * T[s] * lam(q, m) * factor;
    }
  }
};

template <int P>
__device__ __forceinline__ void dc_skeleton_residual(const DcFacetArgs& a) {
  typedef DcOutflow<P> O;
  DcFacet<P> F;
  long long f;
  if (!F.load(a, &f)) return;
  double loc[O::NSS][DC_DIM];
  F.residual(a.wA, loc);
#pragma unroll
  for (int s = 0; s < O::NSS; ++s)
#pragma unroll
    for (int m = 0; m < DC_DIM; ++m) dc_atomic_add(&a.r[F.dofs[m] + s], loc[s][m]);
}

// Numerical skeleton / boundary Jacobian (local_operator.hh:1205-1343): one-sided differences of
// the own-side facet residual with respect to the coefficients of both sides at the facet's
// vertices, delta = eps (1 + |x|).  `sink(i, ma, other, j, mb, value)`: other = 0 own-side column.
// Only the species pairs of the skeleton pattern are kept.
template <int P, class Sink>
__device__ __noinline__ void dc_fd_skeleton(DcFacet<P>& F, double wA, Sink sink) {
  typedef DcOutflow<P> O;
  double down[O::NSS][DC_DIM], up[O::NSS][DC_DIM];
  F.residual(wA, down);
#pragma unroll 1
  for (int side = 0; side < (O::BOUNDARY ? 1 : 2); ++side) {
    const int nsp = side == 0 ? O::NSS : O::NST_REAL;
#pragma unroll 1
    for (int j = 0; j < nsp; ++j)
#pragma unroll 1
      for (int mb = 0; mb < DC_DIM; ++mb) {
        double* coef = side == 0 ? &F.xs[j][mb] : &F.xt[j][mb];
        double* grad = side == 0 ? F.gs[j] : F.gt[j];
        const double* shape = side == 0 ? F.Gsf[mb] : F.Gtf[mb];
        const double keep = *coef;
        double gkeep[DC_DIM];
#if DC_REF_COMPAT
        // :1298 sizes delta of an out-side column with coeff_in read at the out-side node: the own-side
        // container at the same (species slot, local vertex), 0 where it has no such entry
        // (inside = this side when self_inside; slots pair the two elements by local index here)
        const bool out_column = (side == 0) != F.self_inside;
        const double partner = side == 0 ? (j < O::NST_REAL ? F.xt[j < O::NST_REAL ? j : 0][mb] : 0.0)
                                         : (j < O::NSS ? F.xs[j < O::NSS ? j : 0][mb] : 0.0);
        const double sized = out_column ? partner : keep;
#else
        const double sized = keep;
#endif
        const double delta = DC_FD_EPS * (1.0 + fabs(sized));
        *coef = keep + delta;
#pragma unroll
        for (int k = 0; k < DC_DIM; ++k) { gkeep[k] = grad[k]; grad[k] += delta * shape[k]; }
        F.residual(wA, up);
#pragma unroll 1
        for (int i = 0; i < O::NSS; ++i) {
          if (side == 0 ? !O::pair_s(i, j) : !O::pair_t(i, j)) continue;
#pragma unroll 1
          for (int ma = 0; ma < DC_DIM; ++ma) sink(i, ma, side, j, mb, (up[i][ma] - down[i][ma]) / delta);
        }
        *coef = keep;
#pragma unroll
        for (int k = 0; k < DC_DIM; ++k) grad[k] = gkeep[k];
      }
  }
}

// Matrix-free apply (the kernel of every Krylov iteration, local_operator.hh:1354-1396): the direction
// is evaluated at the quadrature point first, y[i,ma] += wA sum_q lam_ma (js z_s(q) + jt z_t(q))_i factor
template <int P>
__device__ __forceinline__ void dc_skeleton_apply(const DcFacetArgs& a, DcFacet<P>& F) {
  typedef DcOutflow<P> O;
  double zs[O::NSS][DC_DIM], zt[O::NST][DC_DIM], loc[O::NSS][DC_DIM];
#pragma unroll
  for (int m = 0; m < DC_DIM; ++m) {
#pragma unroll
    for (int j = 0; j < O::NSS; ++j) {
      const int col = F.dofs[m] + j;
      zs[j][m] = (a.cmask && a.cmask[col]) ? 0.0 : a.z[col];
      loc[j][m] = 0.0;
    }
#pragma unroll
    for (int j = 0; j < O::NST; ++j) {
      const int col = F.doft[m] + j;
      zt[j][m] = (O::BOUNDARY || j >= O::NST_REAL) ? 0.0 : ((a.cmask && a.cmask[col]) ? 0.0 : a.z[col]);
    }
  }
#pragma unroll
  for (int q = 0; q < DC_DIM; ++q) {
    double us[O::NSS], ut[O::NST], js[O::NSS][O::NSS], jt[O::NSS][O::NST];
    const double factor = F.point(q, us, ut);
    O::jacobian(F.c, us, F.gs, ut, F.gt, js, jt);
#pragma unroll
    for (int i = 0; i < O::NSS; ++i) {
      double w = 0.0;
#pragma unroll
      for (int j = 0; j < O::NSS; ++j) {
        if (!O::pair_s(i, j)) continue;
        double zq = 0.0;
#pragma unroll
        for (int m = 0; m < DC_DIM; ++m) zq += zs[j][m] * F.lam(q, m);
        w += js[i][j] * zq;
      }
      if (!O::BOUNDARY) {
#pragma unroll
        for (int j = 0; j < O::NST_REAL; ++j) {
          if (!O::pair_t(i, j)) continue;
          double zq = 0.0;
#pragma unroll
          for (int m = 0; m < DC_DIM; ++m) zq += zt[j][m] * F.lam(q, m);
          w += jt[i][j] * zq;
        }
      }
#pragma unroll
      for (int m = 0; m < DC_DIM; ++m) loc[i][m] += a.wA * w * F.lam(q, m) * factor;
    }
  }
#pragma unroll
  for (int i = 0; i < O::NSS; ++i)
#pragma unroll
    for (int m = 0; m < DC_DIM; ++m) dc_atomic_add(&a.r[F.dofs[m] + i], loc[i][m]);
}

// MODE 0: CSR values; 1: y += J z; 2: block diagonal
template <int P, int MODE>
__device__ __forceinline__ void dc_skeleton_jacobian(const DcFacetArgs& a) {
  typedef DcOutflow<P> O;
  DcFacet<P> F;
  long long f;
  if (!F.load(a, &f)) return;
  if (DC_NUMJAC) {
    dc_fd_skeleton<P>(F, a.wA, [&](int i, int ma, int other, int j, int mb, double v) {
      const int row = F.dofs[ma] + i;
      const int col = (other ? F.doft[mb] : F.dofs[mb]) + j;
      if (MODE == 0) {
        const long long p = dc_csr_find(a.rowptr, a.colidx, row, col);
        if (p >= 0) dc_atomic_add(&a.vals[p], v);
      } else if (MODE == 1) {
        dc_atomic_add(&a.r[row], v * ((a.cmask && a.cmask[col]) ? 0.0 : a.z[col]));
      } else if (!other && ma == mb) {
        dc_atomic_add(&a.bdiag[(long long)F.dofs[ma] * O::NSS + i * O::NSS + j], v);
      }
    });
    return;
  }
  if (MODE == 1) {
    dc_skeleton_apply<P>(a, F);
    return;
  }
  // assembled forms (once per linearisation): rolled loops keep the code small
#pragma unroll 1
  for (int q = 0; q < DC_DIM; ++q) {
    double us[O::NSS], ut[O::NST], js[O::NSS][O::NSS], jt[O::NSS][O::NST];
    const double factor = F.point(q, us, ut);
    O::jacobian(F.c, us, F.gs, ut, F.gt, js, jt);
#pragma unroll 1
    for (int i = 0; i < O::NSS; ++i)
#pragma unroll 1
      for (int ma = 0; ma < DC_DIM; ++ma) {
        const int row = F.dofs[ma] + i;
#pragma unroll 1
        for (int mb = 0; mb < DC_DIM; ++mb) {
          const double w = a.wA * F.lam(q, ma) * F.lam(q, mb) * factor;
#pragma unroll 1
          for (int j = 0; j < O::NSS; ++j) {
            if (!O::pair_s(i, j)) continue;
            const int col = F.dofs[mb] + j;
            const double v = js[i][j] * w;
            if (MODE == 0) {
              const long long p = dc_csr_find(a.rowptr, a.colidx, row, col);
              if (p >= 0) dc_atomic_add(&a.vals[p], v);
            } else if (ma == mb) {
              dc_atomic_add(&a.bdiag[(long long)F.dofs[ma] * O::NSS + i * O::NSS + j], v);
            }
          }
          if (!O::BOUNDARY && MODE == 0) {
#pragma unroll 1
            for (int j = 0; j < O::NST_REAL; ++j) {
              if (!O::pair_t(i, j)) continue;
              const int col = F.doft[mb] + j;
              const long long p = dc_csr_find(a.rowptr, a.colidx, row, col);
              if (p >= 0) dc_atomic_add(&a.vals[p], jt[i][j] * w);
            }
          }
        }
      }
  }
}
