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
__device__ __forceinline__ constexpr int dc_corner_ref(int p, int k) {
  int m = 0;
  for (int t = 0; t < k; ++t) m |= 1 << dc_perm(p, t);
  return m;
}
// the same as a packed table (DC_DIM bits per vertex): the kernels call this several hundred times in
// fully unrolled loops, and folding the loop form above cost NVRTC minutes per model
__device__ __forceinline__ constexpr int dc_corner(int p, int k) {
#if DC_DIM == 2
  return ((p == 0 ? 0x34 : 0x38) >> (2 * k)) & 3;
#else
  return ((p == 0 ? 0xEC8 : p == 1 ? 0xF48 : p == 2 ? 0xED0 : p == 3 ? 0xF90 : p == 4 ? 0xF60 : 0xFA0) >> (3 * k)) & 7;
#endif
}
constexpr bool dc_corner_table_ok() {
  for (int p = 0; p < DC_NPERM; ++p)
    for (int k = 0; k <= DC_DIM; ++k)
      if (dc_corner(p, k) != dc_corner_ref(p, k)) return false;
  return true;
}
static_assert(dc_corner_table_ok(), "packed corner table disagrees with the permutation walk");
// simplices p that contain corner m, as a bit mask
__device__ __forceinline__ constexpr int dc_simplices_of(int m) {
  int mask = 0;
  for (int p = 0; p < DC_NPERM; ++p)
    for (int k = 0; k <= DC_DIM; ++k)
      if (dc_corner(p, k) == m) mask |= 1 << p;
  return mask;
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

// dc_simplices_of as literals (checked at compile time)
__device__ __forceinline__ constexpr int dc_simplices_table(int m) {
#if DC_DIM == 2
  return m == 0 || m == 3 ? 3 : m == 1 ? 1 : 2;
#else
  return m == 0 || m == 7 ? 63 : m == 1 ? 3 : m == 2 ? 12 : m == 4 ? 48 : m == 3 ? 5 : m == 5 ? 18 : 40;
#endif
}
constexpr bool dc_simplices_table_ok() {
  for (int m = 0; m < DC_NCORN; ++m)
    if (dc_simplices_table(m) != dc_simplices_of(m)) return false;
  return true;
}
static_assert(dc_simplices_table_ok(), "simplex-of-corner table disagrees with the corner table");

// contributions of one lattice cell: U/Z = corner values [corner][species], acc = corner sums (overwritten)
template <int C, int MODE>
__device__ __forceinline__ void dc_struct_cell(const DcStructArgs& a, const int* idx,
                                               const double (*U)[DcComp<C>::NS], const double (*Z)[DcComp<C>::NS],
                                               double (*acc)[MODE == 2 ? DcComp<C>::NS * DcComp<C>::NS : DcComp<C>::NS]) {
  typedef DcComp<C> M;
  constexpr int NS = M::NS;
  constexpr int NV = MODE == 2 ? NS * NS : NS;
#pragma unroll
  for (int m = 0; m < DC_NCORN; ++m)
#pragma unroll
    for (int s = 0; s < NV; ++s) acc[m][s] = 0.0;
  double adet = 1.0, rh[DC_DIM];
#pragma unroll
  for (int k = 0; k < DC_DIM; ++k) { adet *= a.h[k]; rh[k] = a.rh[k]; }
#if DC_HOST_VOL
  const double f = DC_QW * adet, vol = a.vol;   // = adet / DC_FACT, divided once on the host
#else
  const double f = DC_QW * adet, vol = adet / DC_FACT;
#endif
  const double ABf = DC_PAB * f, Bf = DC_PB * f;
  DcCtx c;
  c.time = a.time; c.entity_volume = vol; c.integration_factor = f;
  c.in_volume = 1.0; c.in_boundary = 0.0; c.in_skeleton = 0.0;
  c.nrm[0] = c.nrm[1] = c.nrm[2] = 0.0; c.pos[2] = 0.0;
  double x0[DC_DIM];
#pragma unroll
  for (int k = 0; k < DC_DIM; ++k) x0[k] = a.origin[k] + idx[k] * a.h[k];

  // ---- mass / reaction part, simplex by simplex
  double TT[MODE <= 1 ? DC_NPERM : 1][NS];   // sum over the points of simplex p (residual / apply)
#pragma unroll
  for (int p = 0; p < DC_NPERM; ++p) {
    double xl[NS][DC_ND], BS[NS], gu[NS][DC_DIM];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
#pragma unroll
      for (int k = 0; k < DC_ND; ++k) xl[s][k] = U[dc_corner(p, k)][s];
      // every Kuhn simplex runs from corner 0 to corner 2^d - 1: that pair is summed once per cell
      double t = U[0][s] + U[DC_NCORN - 1][s];
#if DC_DIM == 3
      t += xl[s][1] + xl[s][2];
#else
      t += xl[s][1];
#endif
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
      for (int s = 0; s < NS; ++s) TT[p][s] = T[s];
    } else if (MODE == 1) {
      double zl[NS][DC_ND], BZ[NS], T[NS];
#pragma unroll
      for (int s = 0; s < NS; ++s) {
#pragma unroll
        for (int k = 0; k < DC_ND; ++k) zl[s][k] = Z[dc_corner(p, k)][s];
        double t = Z[0][s] + Z[DC_NCORN - 1][s];
#if DC_DIM == 3
        t += zl[s][1] + zl[s][2];
#else
        t += zl[s][1];
#endif
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
      for (int i = 0; i < NS; ++i) TT[p][i] = T[i];
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

  // the B * (sum over the points) part of every simplex goes to its d+1 corners: summed per corner
  // first (corner 0 and the last corner see all simplices -- the same sum --, the others d!/ (d+1 choose ..) of them)
  if (MODE <= 1) {
#pragma unroll
    for (int m = 0; m < DC_NCORN; ++m)
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        double t = 0.0;
#pragma unroll
        for (int p = 0; p < DC_NPERM; ++p)
          if ((dc_simplices_table(m) >> p) & 1) t += TT[p][s];
        acc[m][s] += Bf * t;
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
    double cw[DC_DIM], jdw[NS][NS][DC_DIM];   // |simplex| / h_k^2 and the coefficients times it
#pragma unroll
    for (int k = 0; k < DC_DIM; ++k) {
      cw[k] = vol * rh[k] * rh[k];
#pragma unroll
      for (int i = 0; i < NS; ++i)
#pragma unroll
        for (int j = 0; j < NS; ++j) jdw[i][j][k] = jd[i][j] * cw[k];
    }
#pragma unroll
    for (int m = 0; m < DC_NCORN; ++m)
#pragma unroll
      for (int k = 0; k < DC_DIM; ++k) {
        if ((m >> k) & 1) continue;
        const int m2 = m | (1 << k);
        const double wgt = dc_edge_paths(m) * cw[k];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < NS; ++j)
              if (M::dpair(i, j)) {
                acc[m][i * NS + j] += wgt * jd[i][j];
                acc[m2][i * NS + j] += wgt * jd[i][j];
              }
          } else if (MODE == 3) {
            if (M::dpair(i, i)) {
              acc[m][i] += wgt * jd[i][i];
              acc[m2][i] += wgt * jd[i][i];
            }
          } else {
            // only the couplings the model writes (dpair); the edge weight is folded into the coefficient
            double d = 0.0;
            bool any = false;
#pragma unroll
            for (int j = 0; j < NS; ++j)
              if (M::dpair(i, j)) {
                d += (dc_edge_paths(m) * jdw[i][j][k]) * (MODE == 0 ? (U[m2][j] - U[m][j]) : (Z[m2][j] - Z[m][j]));
                any = true;
              }
            if (any) {
              acc[m2][i] += d;
              acc[m][i] -= d;
            }
          }
        }
      }
  }

}

// decode a linear cell index
__device__ __forceinline__ void dc_cell_index(const DcStructArgs& a, long long cell, int* idx) {
  long long rem = cell;
  idx[0] = (int)(rem % a.n[0]); rem /= a.n[0];
#if DC_DIM == 3
  idx[1] = (int)(rem % a.n[1]); idx[2] = (int)(rem / a.n[1]);
#else
  idx[1] = (int)rem; idx[2] = 0;
#endif
}

// ---- driver 1: one thread per cell (all modes).  The lanes of a warp hold consecutive cells of a lattice row, so
// the x = 1 corners of lane l are the x = 0 corners of lane l + 1: those halves are summed through a warp shuffle
// first, and only one reduction per *vertex* value leaves the warp (plus the last lane's x = 1 half).  ncu on the
// 256^3 apply had the L2's atomic unit at 90 % of its peak (`lts__d_atomic_input_cycles_active`,
// profiles/r02_struct_apply_256_ncu.txt) next to the fp64 pipe at 77 %: this halves the `RED` sectors.
// SCALED (apply only): the direction is (zrelax * zscale) .* z, formed while the corners are loaded -- a separate
// instantiation, because a run-time switch in the load phase (16 conditional loads) cost the plain apply 10 %
// (same box: 0.838 -> 0.922 ms; the loads no longer issued as one batch)
// NOMASK (apply only): the operator has no Dirichlet rows (the host checks), the mask is not looked at
template <int C, int MODE, class CellFn, bool SCALED = false, bool NOMASK = false>
__device__ __forceinline__ void dc_struct_per_cell(const DcStructArgs& a, CellFn cell_fn) {
  typedef DcComp<C> M;
  constexpr int NS = M::NS;
  constexpr int NV = MODE == 2 ? NS * NS : NS;
  const long long cell = a.cell_begin + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool active = cell < a.ncells;     // idle lanes stay for the shuffles
  int idx[3] = {0, 0, 0};
  double acc[DC_NCORN][NV];
  // dof of corner m = d0 + off(m): one 64-bit base and two 32-bit plane strides instead of 2^d 64-bit indices
  long long d0 = 0;
  const int o1 = (a.n[0] + 1) * NS, o2 = DC_DIM == 3 ? o1 * (a.n[1] + 1) : 0;
  auto off = [&](int m) { return ((m & 1) ? NS : 0) + ((m & 2) ? o1 : 0) + ((m & 4) ? o2 : 0); };
  if (active) {
    dc_cell_index(a, cell, idx);
    d0 = a.dof_offset + (idx[0] * (long long)NS + idx[1] * (long long)o1 + idx[2] * (long long)o2);
    double U[DC_NCORN][NS], Z[MODE == 1 ? DC_NCORN : 1][NS];
    const double* xb = a.x + d0;
    const double* zb = MODE == 1 ? a.z + d0 : nullptr;
    const double* sb = (MODE == 1 && SCALED) ? a.zscale + d0 : nullptr;
    const unsigned char* mb = a.cmask ? a.cmask + d0 : nullptr;
#if DC_DIM == 3
    // The only vertex row a warp is the first to touch is (y + 1, z + 1); everything else was loaded by the row or
    // the plane before and hits L1 / L2.  ncu had 16 % of the apply kernel's samples on the first use of the loaded
    // corners (`long_scoreboard`, DRAM latency at 3 warps per scheduler): the same row one plane ahead -- the first
    // touch of the cell n0 * n1 positions later -- is pulled into L2 now.  No extra DRAM traffic: it is demanded later.
    if (DC_STRUCT_PREFETCH && (MODE == 0 || MODE == 1) && idx[2] + 2 <= a.n[2]) {
      const long long ahead = (long long)o1 + 2ll * o2;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(xb + ahead));
      if (MODE == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(zb + ahead));
      if (MODE == 1 && SCALED) asm volatile("prefetch.global.L2 [%0];" ::"l"(sb + ahead));
    }
#endif
#pragma unroll
    for (int m = 0; m < DC_NCORN; ++m) {
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        U[m][s] = xb[off(m) + s];
        // (the unscaled form is kept exactly as it was: written as "load, then select" the same statement compiled to
        // a kernel 11 % slower -- 131 predicated instructions instead of 22 short branches, same-box A/B)
        if (MODE == 1 && !SCALED && !NOMASK) Z[m][s] = (mb && mb[off(m) + s]) ? 0.0 : zb[off(m) + s];
        if (MODE == 1 && !SCALED && NOMASK) Z[m][s] = zb[off(m) + s];
        // zscale = relax * dinv, rounded once: the product k_bicg_p_prec / k_bicg_r_prec would have stored
        // (only operators without Dirichlet rows take this path: apply_scale_ready)
        if (MODE == 1 && SCALED) Z[m][s] = sb[off(m) + s] * zb[off(m) + s];
      }
    }
    cell_fn(idx, U, Z, acc);
  } else {
#pragma unroll
    for (int m = 0; m < DC_NCORN; ++m)
#pragma unroll
      for (int s = 0; s < NV; ++s) acc[m][s] = 0.0;
  }
  // lane l - 1 holds the cell to the left iff this cell is not the first of its row; it keeps its x = 1 half iff
  // it is the last lane, the last cell of its row or the last cell of the launch
  const int lane = threadIdx.x & 31;
  const bool take = active && lane > 0 && idx[0] > 0;
  const bool keep = lane == 31 || idx[0] + 1 == a.n[0] || cell + 1 >= a.ncells;
#pragma unroll
  for (int m = 1; m < DC_NCORN; m += 2)
#pragma unroll
    for (int s = 0; s < NV; ++s) {
      const double t = __shfl_up_sync(0xffffffffu, acc[m][s], 1);
      if (take) acc[m - 1][s] += t;
    }
  if (!active) return;
#pragma unroll
  for (int m = 0; m < DC_NCORN; ++m) {
    if ((m & 1) && !keep) continue;
    if (MODE == 2) {
      double* out = a.bdiag + (d0 + off(m)) * NS;
#pragma unroll
      for (int s = 0; s < NV; ++s) dc_atomic_add(out + s, acc[m][s]);
    } else {   // residual / apply into r; scalar diagonal into a vector laid out like r
      double* out = (MODE == 3 ? a.bdiag : a.r) + d0 + off(m);
#pragma unroll
      for (int s = 0; s < NS; ++s) dc_atomic_add(out + s, acc[m][s]);
    }
  }
}

// ---- driver 2 (residual and apply): register marching along the last axis, lanes along x.
// A thread walks a.march cells up the last axis: the top face of one cell is the bottom face of the
// next, so per cell only the new face is loaded and only the finished bottom face is reduced; along
// x the lanes of a warp hold neighbouring cells, so the x = 1 half of every face is exchanged by
// warp shuffles instead of being loaded / reduced twice.  Per cell and species: 2 loads and 2
// reductions (3-D) instead of 8 and 8 -- the per-cell driver kept the L2 at 75-90 % of its peak
// throughput (profiles/r01_struct_apply_256_ncu.txt, r01_q1_apply_v1_256_ncu.txt).
// The launch covers whole layers: cells [cell_begin, ncells) must be multiples of the layer size.
#define DC_NFACE (DC_NCORN / 2)
template <int C, int MODE, class CellFn>
__device__ __forceinline__ void dc_struct_march(const DcStructArgs& a, CellFn cell_fn) {
  typedef DcComp<C> M;
  constexpr int NS = M::NS;
  constexpr int L = DC_DIM - 1;
  constexpr unsigned FULL = 0xffffffffu;
  const long long layer = (long long)a.n[0] * (DC_DIM == 3 ? a.n[1] : 1);
  const int k0 = (int)(a.cell_begin / layer), k1 = (int)(a.ncells / layer);
  const int nchunk = (k1 - k0 + a.march - 1) / a.march;
  const long long nthreads = layer * nchunk;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool valid = t < nthreads;
  const long long tt = valid ? t : nthreads - 1;
  int idx[3] = {0, 0, 0};
  idx[0] = (int)(tt % a.n[0]);
#if DC_DIM == 3
  idx[1] = (int)((tt / a.n[0]) % a.n[1]);
#endif
  const int chunk = (int)(tt / layer);
  const int kb = k0 + chunk * a.march, ke = min(kb + a.march, k1);
  const unsigned lane = threadIdx.x & 31u;
  // lane + 1 holds the cell at x + 1 of the same row and chunk (hence the same layer range)
  const bool has_right = valid && lane < 31u && idx[0] + 1 < a.n[0];
  const bool has_left = valid && lane > 0u && idx[0] > 0;
  const long long stride[3] = {1, a.n[0] + 1, (long long)(a.n[0] + 1) * (a.n[1] + 1)};
  long long vcol = idx[0];
#if DC_DIM == 3
  vcol += idx[1] * stride[1];
#endif
  long long foff[DC_NFACE];
#pragma unroll
  for (int f = 0; f < DC_NFACE; ++f) {
    foff[f] = 0;
#pragma unroll
    for (int k = 0; k < L; ++k) foff[f] += ((f >> k) & 1) * stride[k];
  }
  auto load_face = [&](int plane, bool on, double (*Uf)[NS], double (*Zf)[NS]) {
    const long long vb = vcol + plane * stride[L];
#pragma unroll
    for (int f = 0; f < DC_NFACE; ++f) {
      if (f & 1) continue;
      const long long d = a.dof_offset + (vb + foff[f]) * NS;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        Uf[f][s] = on ? a.x[d + s] : 0.0;
        if (MODE == 1) Zf[f][s] = on ? ((a.cmask && a.cmask[d + s]) ? 0.0 : a.z[d + s]) : 0.0;
      }
    }
#pragma unroll
    for (int f = 1; f < DC_NFACE; f += 2) {
      const long long d = a.dof_offset + (vb + foff[f]) * NS;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const double ru = __shfl_down_sync(FULL, Uf[f - 1][s], 1);
        Uf[f][s] = has_right ? ru : (on ? a.x[d + s] : 0.0);
        if (MODE == 1) {
          const double rz = __shfl_down_sync(FULL, Zf[f - 1][s], 1);
          Zf[f][s] = has_right ? rz : (on ? ((a.cmask && a.cmask[d + s]) ? 0.0 : a.z[d + s]) : 0.0);
        }
      }
    }
  };
  auto flush_face = [&](int plane, bool on, double (*af)[NS]) {
    const long long vb = vcol + plane * stride[L];
#pragma unroll
    for (int f = 1; f < DC_NFACE; f += 2)
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const double recv = __shfl_up_sync(FULL, af[f][s], 1);
        if (has_left && on) af[f - 1][s] += recv;   // (idle iterations of a short last chunk still shuffle)
      }
    if (!on) return;
#pragma unroll
    for (int f = 0; f < DC_NFACE; ++f) {
      if ((f & 1) && has_right) continue;   // the right neighbour reduces that corner
      const long long d = a.dof_offset + (vb + foff[f]) * NS;
#pragma unroll
      for (int s = 0; s < NS; ++s) dc_atomic_add(&a.r[d + s], af[f][s]);
    }
  };
  double Ub[DC_NFACE][NS], Zb[MODE == 1 ? DC_NFACE : 1][NS], accb[DC_NFACE][NS];
  load_face(kb, valid, Ub, Zb);
#pragma unroll
  for (int f = 0; f < DC_NFACE; ++f)
#pragma unroll
    for (int s = 0; s < NS; ++s) accb[f][s] = 0.0;
  for (int kk = 0; kk < a.march; ++kk) {
    const int k = kb + kk;
    const bool act = valid && k < ke;
    double Ut[DC_NFACE][NS], Zt[MODE == 1 ? DC_NFACE : 1][NS];
    load_face(k + 1, act, Ut, Zt);
    double acct[DC_NFACE][NS];
    if (act) {
      double U[DC_NCORN][NS], Z[MODE == 1 ? DC_NCORN : 1][NS], acc[DC_NCORN][NS];
#pragma unroll
      for (int f = 0; f < DC_NFACE; ++f)
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          U[f][s] = Ub[f][s];
          U[f + DC_NFACE][s] = Ut[f][s];
          if (MODE == 1) { Z[f][s] = Zb[f][s]; Z[f + DC_NFACE][s] = Zt[f][s]; }
        }
      idx[L] = k;
      cell_fn(idx, U, Z, acc);
#pragma unroll
      for (int f = 0; f < DC_NFACE; ++f)
#pragma unroll
        for (int s = 0; s < NS; ++s) { accb[f][s] += acc[f][s]; acct[f][s] = acc[f + DC_NFACE][s]; }
    }
    flush_face(k, act, accb);
    if (act) {
#pragma unroll
      for (int f = 0; f < DC_NFACE; ++f)
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          Ub[f][s] = Ut[f][s];
          if (MODE == 1) Zb[f][s] = Zt[f][s];
          accb[f][s] = acct[f][s];
        }
    }
  }
  flush_face(ke, valid && ke > kb, accb);
}

template <int C, int MODE, bool SCALED = false, bool NOMASK = false>
__device__ __forceinline__ void dc_structured_kernel(const DcStructArgs& a) {
  constexpr int NS = DcComp<C>::NS;
  constexpr int NV = MODE == 2 ? NS * NS : NS;
  auto cell = [&](const int* idx, const double (*U)[NS], const double (*Z)[NS], double (*acc)[NV]) {
    dc_struct_cell<C, MODE>(a, idx, U, Z, acc);
  };
  dc_struct_per_cell<C, MODE, decltype(cell), SCALED, NOMASK>(a, cell);
}
template <int C, int MODE>
__device__ __forceinline__ void dc_structured_march_kernel(const DcStructArgs& a) {
  constexpr int NS = DcComp<C>::NS;
  dc_struct_march<C, MODE>(a, [&](const int* idx, const double (*U)[NS], const double (*Z)[NS], double (*acc)[NS]) {
    dc_struct_cell<C, MODE>(a, idx, U, Z, acc);
  });
}
