// =================================================================================================
// Tile-marching driver of the structured kernels: owner computes, every result entry is written
// once with a plain store (no fp64 atomics, no zero-fill of the result, deterministic sums), the
// vector update that produces the direction and the dot products that consume the result are fused
// into the same sweep.  Serves the matrix-free Jacobian apply of local_operator.hh:510-524 behind
// MatrixFreeAdapter::apply (model/make_step_operator.hh:62-95) and the residual of :417-491; the
// cell integrals are the same functions the per-cell / marching drivers call (dc_struct_cell,
// dc_q1_cell), so the arithmetic per cell is unchanged.
//
// A CTA owns a tile of 32 x (DC_TILE_W * DC_TILE_R) cells of the first DIM-1 axes (2-D: 32 cells) and
// walks A.lz cell layers up the last axis.  Lanes are x neighbours, warp w integrates the cell rows
// [w R, (w+1) R) of a layer one after the other:
//   stage    vertex planes travel global -> shared memory with cp.async, two layers ahead of the
//            cells that read them, one copy per vertex and tile.  The prologue runs when a plane has
//            landed: the BiCGSTAB update that produces the direction (p = r + beta (p - omega v) or
//            r -= alpha v, then the folded Jacobi relax * dinv * .) is evaluated from the staged
//            operands; the CTA that owns the vertex writes the updated vector back.
//   compute  the cell's 2^DIM corner values come from the two staged planes.
//   combine  the corner sums of a cell go to two accumulator planes in shared memory (the plane
//            under and the plane above the layer; the upper one becomes the lower one of the next
//            layer): the x = 1 corners travel to the neighbouring lane by shuffle, rows inside a
//            warp's block are touched by that warp only, and the row a block shares with the next
//            warp goes to a spill row of its own -- no two threads ever write one address between
//            two barriers, the first write to an address is a store (nothing is zero-filled).
//   epilogue the finished plane is written: complete vertices with a plain store and straight into
//            the reductions (<w, y>, <y, r>, |y|^2) while the value is in a register; vertices
//            that a neighbouring tile / chunk also contributes to ("cut": on a tile edge or a chunk
//            boundary plane) go to the slot of this contributor -- slot 0 is the result vector
//            itself -- and la::tile_fixup (linalg.cu) adds the slots of a cut vertex in slot order,
//            finishes the reductions and forms the final sums in block order.
// Slot of a contributor: bit a set = the contributor lies on the upper side of the cut of axis a.
// Two barriers per layer.  Shared memory per vertex of a staged plane (33 x (W R + 1) vertices, NS
// doubles each, vertex-major so that two species move as one 16-byte access): 3 planes of u, 3 of the
// direction, 2 of what the epilogue reads (w, the updated r or the old result), 4 raw operands of the
// prologue, 2 accumulator planes: 14 (+ spill rows) -- 70.7 KB for two species at 32 x 8.

#if DC_DIM == 3
#define DC_TW DC_TILE_W
#define DC_TR DC_TILE_R
#define DC_TPY (DC_TILE_W * DC_TILE_R + 1)
#else
#define DC_TW 1
#define DC_TR 1
#define DC_TPY 1
#endif
#define DC_TPX 33
#define DC_TPL (DC_TPX * DC_TPY)
#define DC_TILE_THREADS (32 * DC_TW)

// NS doubles of one vertex; even NS: 16-byte accesses (the host guarantees the alignment)
template <int NS>
__device__ __forceinline__ void dc_vld(const double* p, double* out) {
  if constexpr (NS % 2 == 0) {
#pragma unroll
    for (int g = 0; g < NS / 2; ++g) {
      const double2 t = reinterpret_cast<const double2*>(p)[g];
      out[2 * g] = t.x; out[2 * g + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int s = 0; s < NS; ++s) out[s] = p[s];
  }
}
template <int NS>
__device__ __forceinline__ void dc_vst(double* p, const double* v) {
  if constexpr (NS % 2 == 0) {
#pragma unroll
    for (int g = 0; g < NS / 2; ++g) reinterpret_cast<double2*>(p)[g] = make_double2(v[2 * g], v[2 * g + 1]);
  } else {
#pragma unroll
    for (int s = 0; s < NS; ++s) p[s] = v[s];
  }
}
template <int NS>
__device__ __forceinline__ void dc_vcp_async(double* smem, const double* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  if constexpr (NS % 2 == 0) {
#pragma unroll
    for (int g = 0; g < NS / 2; ++g)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa + 16 * g), "l"(gmem + 2 * g) : "memory");
  } else {
#pragma unroll
    for (int s = 0; s < NS; ++s)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa + 8 * s), "l"(gmem + s) : "memory");
  }
}
__device__ __forceinline__ void dc_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void dc_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// The cell integrals of dc_struct_cell (MODE 0 residual / 1 Jacobian apply) with the corner values read
// where they are used: `ldu(m, out)` / `ldz(m, out)` fetch the NS values of corner m (shared memory in
// the tile driver).  Same quadrature, same sums per corner; what changes is the life time of the
// operands -- four corners per simplex instead of the 2^DIM corners of both fields for the whole
// cell -- and the B * (sum over the points) part, which goes straight to the simplex's corners.
// ~100 registers instead of ~166, i.e. 16 resident warps per SM instead of 12.
template <int C, int MODE, class LdU, class LdZ>
__device__ __forceinline__ void dc_struct_cell_stream(const DcStructArgs& a, const int* idx, LdU ldu, LdZ ldz,
                                                      double (*acc)[DcComp<C>::NS]) {
  typedef DcComp<C> M;
  constexpr int NS = M::NS;
#pragma unroll
  for (int m = 0; m < DC_NCORN; ++m)
#pragma unroll
    for (int s = 0; s < NS; ++s) acc[m][s] = 0.0;
  double adet = 1.0, rh[DC_DIM];
#pragma unroll
  for (int k = 0; k < DC_DIM; ++k) { adet *= a.h[k]; rh[k] = a.rh[k]; }
  const double f = DC_QW * adet, vol = adet / DC_FACT;
  const double ABf = DC_PAB * f, Bf = DC_PB * f;
  DcCtx c;
  c.time = a.time; c.entity_volume = vol; c.integration_factor = f;
  c.in_volume = 1.0; c.in_boundary = 0.0; c.in_skeleton = 0.0;
  c.nrm[0] = c.nrm[1] = c.nrm[2] = 0.0; c.pos[2] = 0.0;
  double x0[DC_DIM];
#pragma unroll
  for (int k = 0; k < DC_DIM; ++k) x0[k] = a.origin[k] + idx[k] * a.h[k];
  // every Kuhn simplex runs from corner 0 to corner 2^d - 1: that pair is summed once per cell
  double CU[NS], CZ[MODE == 1 ? NS : 1];
  {
    double lo[NS], hi[NS];
    ldu(0, lo); ldu(DC_NCORN - 1, hi);
#pragma unroll
    for (int s = 0; s < NS; ++s) CU[s] = lo[s] + hi[s];
    if (MODE == 1) {
      ldz(0, lo); ldz(DC_NCORN - 1, hi);
#pragma unroll
      for (int s = 0; s < NS; ++s) CZ[s] = lo[s] + hi[s];
    }
  }
#pragma unroll
  for (int p = 0; p < DC_NPERM; ++p) {
    // (keeps the compiler from hoisting the corner loads of all simplices to the top: their life time is the point)
    asm volatile("" ::: "memory");
    double xl[DC_ND][NS], BS[NS], gu[NS][DC_DIM];
#pragma unroll
    for (int k = 0; k < DC_ND; ++k) ldu(dc_corner(p, k), xl[k]);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
#if DC_DIM == 3
      BS[s] = DC_PB * (CU[s] + (xl[1][s] + xl[2][s]));
#else
      BS[s] = DC_PB * (CU[s] + xl[1][s]);
#endif
#pragma unroll
      for (int k = 0; k < DC_DIM; ++k) gu[s][dc_perm(p, k)] = (xl[k + 1][s] - xl[k][s]) * rh[dc_perm(p, k)];
    }
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
    double T[NS];
    if (MODE == 0) {
#pragma unroll
      for (int q = 0; q < DC_NQ; ++q) {
        double u[NS], sc[NS];
        set_pos(q);
#pragma unroll
        for (int s = 0; s < NS; ++s) u[s] = BS[s] + DC_PAB * xl[DC_VQ(q)][s];
        M::scalar(c, u, gu, a.wM, a.wA, sc);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          T[s] = q == 0 ? sc[s] : T[s] + sc[s];
          acc[dc_corner(p, DC_VQ(q))][s] += ABf * sc[s];
        }
      }
    } else {
      double zl[DC_ND][NS], BZ[NS];
#pragma unroll
      for (int k = 0; k < DC_ND; ++k) ldz(dc_corner(p, k), zl[k]);
#pragma unroll
      for (int s = 0; s < NS; ++s) {
#if DC_DIM == 3
        BZ[s] = DC_PB * (CZ[s] + (zl[1][s] + zl[2][s]));
#else
        BZ[s] = DC_PB * (CZ[s] + zl[1][s]);
#endif
      }
#pragma unroll
      for (int q = 0; q < DC_NQ; ++q) {
        double u[NS], zq[NS], jm[NS][NS];
        set_pos(q);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          u[s] = BS[s] + DC_PAB * xl[DC_VQ(q)][s];
          zq[s] = BZ[s] + DC_PAB * zl[DC_VQ(q)][s];
        }
        M::jac_mass(c, u, gu, a.wM, a.wA, jm);
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          double w = 0.0;
          bool any = false;
#pragma unroll
          for (int j = 0; j < NS; ++j)
            if (M::pair(i, j)) { w = any ? w + jm[i][j] * zq[j] : jm[i][j] * zq[j]; any = true; }
          T[i] = q == 0 ? w : T[i] + w;
          acc[dc_corner(p, DC_VQ(q))][i] += ABf * w;
        }
      }
    }
    // the B * (sum over the points) part: to the d+1 corners of this simplex
#pragma unroll
    for (int k = 0; k < DC_ND; ++k)
#pragma unroll
      for (int s = 0; s < NS; ++s) acc[dc_corner(p, k)][s] += Bf * T[s];
  }

  // ---- diffusion part, lattice edge by lattice edge (as in dc_struct_cell)
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
    asm volatile("" ::: "memory");
    double V[DC_NCORN][NS];
#pragma unroll
    for (int m = 0; m < DC_NCORN; ++m) {
      if (MODE == 0) ldu(m, V[m]);
      else ldz(m, V[m]);
    }
#pragma unroll
    for (int k = 0; k < DC_DIM; ++k) {
      const double cw = vol * rh[k] * rh[k];
#pragma unroll
      for (int m = 0; m < DC_NCORN; ++m) {
        if ((m >> k) & 1) continue;
        const int m2 = m | (1 << k);
#pragma unroll
        for (int i = 0; i < NS; ++i) {
          double d = 0.0;
          bool any = false;
#pragma unroll
          for (int j = 0; j < NS; ++j)
            if (M::dpair(i, j)) {
              const double t = (dc_edge_paths(m) * (jd[i][j] * cw)) * (V[m2][j] - V[m][j]);
              d = any ? d + t : t;
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

template <int C, int MODE, class CellFn>
__device__ __forceinline__ void dc_tile_march(const DcTileArgs& A, CellFn cell_fn) {
  typedef DcComp<C> M;
  constexpr int NS = M::NS;
  constexpr int L = DC_DIM - 1;
  constexpr int NYC = DC_DIM == 3 ? 2 : 1;             // corners of a cell along y
  constexpr int W = DC_TW, R = DC_TR, TPX = DC_TPX, TPL = DC_TPL, NT = DC_TILE_THREADS;
  constexpr int PLANE = TPL * NS;                      // doubles of one staged plane
  constexpr int NTRIP = (TPL + NT - 1) / NT;           // vertices of a plane per thread
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) double dc_tile_smem[];
  double* const ringU = dc_tile_smem;                  // [3][TPL][NS] linearisation point
  double* const ringZ = ringU + 3 * PLANE;             // [3][TPL][NS] direction
  double* const ringA = ringZ + 3 * PLANE;             // [2][TPL][NS] what the epilogue reads
  double* const raw = ringA + 2 * PLANE;               // [4][TPL][NS] r, p, v, dinv of the plane being staged
  double* const Oacc = raw + 4 * PLANE;                // [2][TPL][NS] accumulator planes
  double* const Bacc = Oacc + 2 * PLANE;               // [2][W][TPX][NS] spill rows
  const DcStructArgs& a = A.s;

  // ---- work item: (tile, chunk)
  int b = blockIdx.x;
  const int tx = b % A.ntx;
  b /= A.ntx;
#if DC_DIM == 3
  const int ty = b % A.nty;
  b /= A.nty;
#else
  const int ty = 0;
#endif
  const int x0 = tx * 32, y0 = ty * (W * R);
  const int TXa = min(32, a.n[0] - x0);
  const int TYa = DC_DIM == 3 ? min(W * R, a.n[1] - y0) : 0;   // 2-D: the only vertex "row" is 0
  const int nL = a.n[L];
  const int kb = b * A.lz, ke = min(kb + A.lz, nL);
  const int lx = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool xlo_cut = x0 > 0, xhi_cut = x0 + TXa < a.n[0];
  const bool ylo_cut = DC_DIM == 3 && y0 > 0, yhi_cut = DC_DIM == 3 && y0 + TYa < a.n[1];
  const int vs1 = a.n[0] + 1;
  const int vsL = DC_DIM == 3 ? vs1 * (a.n[1] + 1) : vs1;
  const long long dof0 = a.dof_offset + (long long)(x0 + (DC_DIM == 3 ? y0 * vs1 : 0)) * NS;
  auto plane_base = [&](int kp) { return dof0 + (long long)kp * vsL * NS; };

  // ---- this thread's vertices of a staged plane: dof offset inside the plane, ownership / cut flags
  int voff[NTRIP];
  unsigned vflag[NTRIP];   // 1: owned by this tile, 2 / 4: on the lower / upper x cut, 8 / 16: y cuts
#pragma unroll
  for (int j = 0; j < NTRIP; ++j) {
    const int pv = threadIdx.x + j * NT;
    const int vx = pv % TPX, vy = pv / TPX;
    const bool valid = pv < TPL && vx <= TXa && vy <= TYa;
    voff[j] = valid ? (vx + vy * vs1) * NS : -1;
    vflag[j] = ((vx < TXa || !xhi_cut) && (vy < TYa || !yhi_cut) ? 1u : 0u) | (vx == 0 && xlo_cut ? 2u : 0u) |
               (vx == TXa && xhi_cut ? 4u : 0u) | (DC_DIM == 3 && vy == 0 && ylo_cut ? 8u : 0u) |
               (DC_DIM == 3 && vy == TYa && yhi_cut ? 16u : 0u);
  }

  // ---- step lengths of the fused BiCGSTAB updates, formed as kernels/linalg.cu forms them
  double alpha = 0.0, beta = 0.0, omega = 0.0;
  if (A.pro == 1 && !A.first) {
    const double rho = *A.rho;
    alpha = rho / *A.hptr;
    omega = A.trtt[0] / A.trtt[1];
    beta = (*A.rho_new / rho) * (alpha / omega);
  } else if (A.pro == 2) {
    alpha = *A.rho / *A.hptr;
  }
  double red[3] = {0.0, 0.0, 0.0};
  const bool aux_copy = (MODE == 1 && A.epi == 1) || A.accumulate;   // the epilogue reads w / the old result

  // ---- vertex plane kp: start the copies global -> shared memory (this thread's vertices)
  auto issue = [&](int kp) {
    const long long pb = plane_base(kp);
    double* U = ringU + (kp % 3) * PLANE;
    double* Z = ringZ + (kp % 3) * PLANE;
#pragma unroll
    for (int j = 0; j < NTRIP; ++j) {
      if (voff[j] < 0) continue;
      const int e = (threadIdx.x + j * NT) * NS;
      const long long d = pb + voff[j];
      dc_vcp_async<NS>(U + e, a.x + d);
      if (MODE == 1) {
        if (A.pro == 0) {
          dc_vcp_async<NS>(Z + e, a.z + d);
        } else {
          dc_vcp_async<NS>(raw + e, A.r_in + d);
          if (A.pro == 1 && !A.first) dc_vcp_async<NS>(raw + PLANE + e, A.p_in + d);
          if (A.pro == 2 || !A.first) dc_vcp_async<NS>(raw + 2 * PLANE + e, A.v_in + d);
          dc_vcp_async<NS>(raw + 3 * PLANE + e, A.dinv + d);
        }
      }
    }
  };
  // what the epilogue of plane kq reads, one layer ahead of it
  auto issue_aux = [&](int kq) {
    if (!aux_copy) return;
    const long long pb = plane_base(kq);
    const double* src = A.accumulate ? a.r : A.w;
    double* Q = ringA + (kq & 1) * PLANE;
#pragma unroll
    for (int j = 0; j < NTRIP; ++j)
      if (voff[j] >= 0) dc_vcp_async<NS>(Q + (threadIdx.x + j * NT) * NS, src + pb + voff[j]);
  };

  // ---- vertex plane kp has landed: the prologue on this thread's vertices
  auto process = [&](int kp) {
    if (MODE != 1 || (A.pro == 0 && !a.cmask)) return;
    const long long pb = plane_base(kp);
    double* Z = ringZ + (kp % 3) * PLANE;
    double* Q = ringA + (kp & 1) * PLANE;
    const bool plane_owned = kp < ke || ke == nL;         // a chunk owns its planes [kb, ke), the last one also nL
    const bool plane_counts = kp >= A.own_lo && kp < A.own_hi;
#pragma unroll
    for (int j = 0; j < NTRIP; ++j) {
      if (voff[j] < 0) continue;
      const int e = (threadIdx.x + j * NT) * NS;
      const long long d = pb + voff[j];
      const bool owner = plane_owned && (vflag[j] & 1u);
      double z[NS];
      if (A.pro == 0) {
        dc_vld<NS>(Z + e, z);
#pragma unroll
        for (int s = 0; s < NS; ++s)
          if (a.cmask[d + s]) z[s] = 0.0;
      } else if (A.pro == 1) {
        double r[NS], p[NS], v[NS], di[NS];
        dc_vld<NS>(raw + e, r);
        dc_vld<NS>(raw + 3 * PLANE + e, di);
        if (!A.first) { dc_vld<NS>(raw + PLANE + e, p); dc_vld<NS>(raw + 2 * PLANE + e, v); }
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          p[s] = A.first ? r[s] : r[s] + beta * (p[s] - omega * v[s]);
          z[s] = A.relax * di[s] * p[s];
        }
        if (owner) dc_vst<NS>(A.p_out + d, p);
      } else {
        double r[NS], v[NS], di[NS];
        dc_vld<NS>(raw + e, r);
        dc_vld<NS>(raw + 2 * PLANE + e, v);
        dc_vld<NS>(raw + 3 * PLANE + e, di);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          r[s] = r[s] - alpha * v[s];
          z[s] = A.relax * di[s] * r[s];
          if (owner && plane_counts) red[0] += r[s] * r[s];
        }
        if (owner) dc_vst<NS>(A.r_out + d, r);
        dc_vst<NS>(Q + e, r);
      }
      dc_vst<NS>(Z + e, z);
    }
  };

  // ---- finished plane k: its accumulator plane holds this CTA's sums (epilogue)
  auto finish = [&](int k) {
    const long long pb = plane_base(k);
    const double* O = Oacc + (k & 1) * PLANE;
    const double* B = Bacc + (k & 1) * (W * TPX * NS);
    const double* Q = ringA + (k & 1) * PLANE;
    const bool zlo = k == kb && kb > 0, zhi = k == ke && ke < nL;
    const bool plane_counts = k >= A.own_lo && k < A.own_hi;
#pragma unroll
    for (int j = 0; j < NTRIP; ++j) {
      if (voff[j] < 0) continue;
      const int pv = threadIdx.x + j * NT, e = pv * NS;
      const long long d = pb + voff[j];
      const bool cut = (vflag[j] & 30u) || zlo || zhi;
      const int slot = ((vflag[j] & 2u) ? 1 : 0) | ((vflag[j] & 8u) ? 2 : 0) | (zlo ? 4 : 0);
      // row vy of the plane: cells of row vy wrote O (their lower y corners) and so did row vy - 1 unless it is
      // the last row of a warp's block, whose upper y corners went to that warp's spill row
      double val[NS];
      const int vy = pv / TPX, vx = pv % TPX;
      const bool in_o = DC_DIM == 2 || vy < TYa || vy % R != 0, in_b = DC_DIM == 3 && vy > 0 && vy % R == 0;
      if (in_o) {
        dc_vld<NS>(O + e, val);
      } else {
#pragma unroll
        for (int s = 0; s < NS; ++s) val[s] = 0.0;
      }
      if (in_b) {
        double sp[NS];
        dc_vld<NS>(B + ((vy / R - 1) * TPX + vx) * NS, sp);
#pragma unroll
        for (int s = 0; s < NS; ++s) val[s] += sp[s];
      }
      if (cut && slot != 0) {
        dc_vst<NS>(A.slots + (long long)(slot - 1) * A.slot_stride + d, val);
        continue;
      }
      double q[NS];
      if (aux_copy || A.epi == 2) dc_vld<NS>(Q + e, q);
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        if (A.accumulate) val[s] += q[s];
        if (!cut) {
          if (MODE == 1 && A.identity && a.cmask[d + s]) val[s] = a.z[d + s];   // identity row
          if (plane_counts) {
            if (A.epi == 1 || A.epi == 2) red[1] += val[s] * q[s];
            if (A.epi == 2) red[2] += val[s] * val[s];
          }
        }
      }
      dc_vst<NS>(a.r + d, val);
    }
  };

  // ---- march
  int idx[3] = {x0 + lx, 0, 0};
  issue(kb);
  issue_aux(kb);
  dc_cp_async_commit();
  dc_cp_async_wait_all();
  process(kb);
  issue(kb + 1);
  dc_cp_async_commit();
  for (int k = kb; k < ke; ++k) {
    dc_cp_async_wait_all();        // plane k+1 (and the epilogue operand of plane k) have landed
    process(k + 1);
    if (k + 2 <= ke) issue(k + 2);
    issue_aux(k + 1);
    dc_cp_async_commit();
    __syncthreads();               // planes k, k+1 staged and processed; the previous epilogue is through
    const double* Ub = ringU + (k % 3) * PLANE;
    const double* Ut = ringU + ((k + 1) % 3) * PLANE;
    const double* Zb = ringZ + (k % 3) * PLANE;
    const double* Zt = ringZ + ((k + 1) % 3) * PLANE;
    idx[L] = k;
#pragma unroll 1
    for (int c = 0; c < R; ++c) {
      const int row = w * R + c;
      const bool active = lx < TXa && (DC_DIM == 2 || row < TYa);
      double acc[DC_NCORN][NS];
      if (active) {
        const int e0 = (row * TPX + lx) * NS;
        auto corner = [&](int m) { return e0 + (((DC_DIM == 3 ? (m >> 1) & 1 : 0)) * TPX + (m & 1)) * NS; };
        auto ldu = [&](int m, double* out) { dc_vld<NS>(((m >> L) ? Ut : Ub) + corner(m), out); };
        auto ldz = [&](int m, double* out) { dc_vld<NS>(((m >> L) ? Zt : Zb) + corner(m), out); };
        if (DC_DIM == 3) idx[1] = y0 + row;
        cell_fn(idx, ldu, ldz, acc);
      } else {
#pragma unroll
        for (int m = 0; m < DC_NCORN; ++m)
#pragma unroll
          for (int s = 0; s < NS; ++s) acc[m][s] = 0.0;
      }
      // corner sums -> accumulator planes (zt = 0: the plane under the layer, 1: the plane above)
#pragma unroll
      for (int zt = 0; zt < 2; ++zt) {
        double* O = Oacc + ((k + zt) & 1) * PLANE;
        double* B = Bacc + ((k + zt) & 1) * (W * TPX * NS);
        const bool fresh = zt == 1 || k == kb;     // nothing has been written to this plane by an earlier layer
#pragma unroll
        for (int fy = 0; fy < NYC; ++fy) {
          const int m0 = (zt << L) | (fy << 1);
          // rows inside the block: this thread wrote them with the previous cell; the row shared with the
          // next block goes to this warp's spill row
          const bool spill = DC_DIM == 3 && fy == 1 && c == R - 1;
          const bool store = fresh && (fy == 1 || c == 0);
          double* dst = spill ? B + (w * TPX + lx) * NS : O + ((row + fy) * TPX + lx) * NS;
          double v[NS], rgt[NS];
#pragma unroll
          for (int s = 0; s < NS; ++s) {
            rgt[s] = acc[m0 | 1][s];
            const double left = __shfl_up_sync(FULL, rgt[s], 1);
            v[s] = lx > 0 ? acc[m0][s] + left : acc[m0][s];
          }
          if (active) {
            if (!store) {
              double old[NS];
              dc_vld<NS>(dst, old);
#pragma unroll
              for (int s = 0; s < NS; ++s) v[s] += old[s];
            }
            dc_vst<NS>(dst, v);
            if (lx + 1 == TXa) {   // the last vertex column of the tile
              if (!store) {
                double old[NS];
                dc_vld<NS>(dst + NS, old);
#pragma unroll
                for (int s = 0; s < NS; ++s) rgt[s] += old[s];
              }
              dc_vst<NS>(dst + NS, rgt);
            }
          }
        }
      }
    }
    __syncthreads();               // the plane under the layer is complete
    finish(k);
  }
  dc_cp_async_wait_all();
  finish(ke);

  // ---- reduction partials of this CTA (summed in block order by la::tile_fixup)
  __syncthreads();
  double* sm = dc_tile_smem;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    double v = red[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    if (lane == 0) sm[q * 32 + w] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = 0.0;
    for (int ww = 0; ww < W; ++ww) v += sm[threadIdx.x * 32 + ww];
    A.partials[(size_t)blockIdx.x * 4 + threadIdx.x] = v;
  }
}

// cell functions behind the driver (functors with a templated call: the corner loaders are lambdas of the driver)
template <int C, int MODE>
struct DcTileCellP1 {
  const DcStructArgs& a;
  template <class LdU, class LdZ>
  __device__ __forceinline__ void operator()(const int* idx, LdU ldu, LdZ ldz, double (*acc)[DcComp<C>::NS]) const {
    dc_struct_cell_stream<C, MODE>(a, idx, ldu, ldz, acc);
  }
};
template <int C, int MODE>
__device__ __forceinline__ void dc_tile_kernel(const DcTileArgs& A) {
  dc_tile_march<C, MODE>(A, DcTileCellP1<C, MODE>{A.s});
}
#ifdef DC_TILE_Q1
template <int C, int MODE>
struct DcTileCellQ1 {
  const DcStructArgs& a;
  template <class LdU, class LdZ>
  __device__ __forceinline__ void operator()(const int* idx, LdU ldu, LdZ ldz, double (*acc)[DcComp<C>::NS]) const {
    constexpr int NS = DcComp<C>::NS;
    double U[DC_NCORN][NS], Z[MODE == 1 ? DC_NCORN : 1][NS];
#pragma unroll
    for (int m = 0; m < DC_NCORN; ++m) {
      ldu(m, U[m]);
      if (MODE == 1) ldz(m, Z[m]);
    }
    dc_q1_cell<C, MODE>(a, idx, U, Z, acc);
  }
};
template <int C, int MODE>
__device__ __forceinline__ void dc_tile_q1_kernel(const DcTileArgs& A) {
  dc_tile_march<C, MODE>(A, DcTileCellQ1<C, MODE>{A.s});
}
#endif
