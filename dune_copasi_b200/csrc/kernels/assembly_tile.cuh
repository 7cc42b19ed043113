// =================================================================================================
// Tile-marching driver of the structured kernels: owner computes, every result entry is written
// once with a plain store (no fp64 atomics, no zero-fill of the result, deterministic sums), the
// vector update that produces the direction and the dot products that consume the result are fused
// into the same sweep.  Serves the matrix-free Jacobian apply of local_operator.hh:510-524 behind
// MatrixFreeAdapter::apply (model/make_step_operator.hh:62-95) and the residual of :417-491; the
// cell integrals are the same functions the per-cell / marching drivers call (dc_struct_cell,
// dc_q1_cell), so the arithmetic per cell is unchanged.
//
// A CTA owns a tile of DC_TILE_X (x DC_TILE_Y) cells of the first DIM-1 axes and walks A.lz cell
// layers up the last axis, one thread per cell column:
//   stage    vertex planes travel global -> shared memory with cp.async, two layers ahead of the
//            cells that read them (the loads of plane k+2 are in flight while layer k is integrated),
//            one copy per vertex and tile.  The prologue runs when a plane has landed: the BiCGSTAB
//            update that produces the direction (p = r + beta (p - omega v) or r -= alpha v, then
//            the folded Jacobi relax * dinv * .) is evaluated from the staged operands; the CTA
//            that owns the vertex writes the updated vector back.
//   compute  the cell's 2^DIM corner values come from the two staged planes.
//   combine  contributions to the finished (lower) plane: a thread carries its own corners in
//            registers from layer to layer; the 2^(DIM-1) threads around a vertex add up in shared
//            memory in barrier-separated phases (every phase writes distinct addresses); in 3-D the
//            x neighbours are lanes of one warp and exchange by shuffle, which leaves two phases.
//   epilogue the finished plane is written: complete vertices with a plain store and straight into
//            the reductions (<w, y>, <y, r>, |y|^2) while the value is in a register; vertices
//            that a neighbouring tile / chunk also contributes to ("cut": on a tile edge or a chunk
//            boundary plane) go to the slot of this contributor -- slot 0 is the result vector
//            itself -- and la::tile_fixup (linalg.cu) adds the slots of a cut vertex in slot order,
//            finishes the reductions and forms the final sums in block order.
// Slot of a contributor: bit a set = the contributor lies on the upper side of the cut of axis a.
//
// Shared memory: 3 staged planes x 3 fields (u, direction, auxiliary = what the epilogue reads: w,
// the updated r or the old result), one result plane and 2 x 4 raw operand planes of the prologue,
// NS * (DC_TILE_X + 1) * (DC_TILE_Y + 1) doubles each (18 planes: 47.5 KB for two species at 32 x 4).

#if DC_DIM == 3
#define DC_TILE_TY DC_TILE_Y
#define DC_TILE_PY (DC_TILE_Y + 1)
#else
#define DC_TILE_TY 1
#define DC_TILE_PY 1
#endif
#define DC_TILE_PX (DC_TILE_X + 1)
#define DC_TILE_PLANE (DC_TILE_PX * DC_TILE_PY)
#define DC_TILE_THREADS (DC_TILE_X * DC_TILE_TY)

__device__ __forceinline__ void dc_cp_async8(double* smem, const double* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void dc_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void dc_cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <int C, int MODE, class CellFn>
__device__ __forceinline__ void dc_tile_march(const DcTileArgs& A, CellFn cell_fn) {
  typedef DcComp<C> M;
  constexpr int NS = M::NS;
  constexpr int L = DC_DIM - 1;
  constexpr int NFACE = DC_NCORN / 2;
  constexpr int TPX = DC_TILE_PX, TPL = DC_TILE_PLANE, NT = DC_TILE_THREADS;
  constexpr int FIELD = NS * TPL, SLOT = 3 * FIELD, RAW = 4 * FIELD;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ double dc_tile_smem[];
  double* const ring = dc_tile_smem;                 // [3 planes][u, direction, auxiliary][NS][TPL]
  double* const O = dc_tile_smem + 3 * SLOT;         // [NS][TPL] sums of the finished plane
  double* const raw = dc_tile_smem + 3 * SLOT + FIELD;   // [2 planes][r, p, v, dinv][NS][TPL]
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
  const int x0 = tx * DC_TILE_X, y0 = ty * DC_TILE_TY;
  const int TXa = min(DC_TILE_X, a.n[0] - x0);
  const int TYa = DC_DIM == 3 ? min(DC_TILE_TY, a.n[1] - y0) : 0;   // 2-D: the only "row" is 0
  const int nL = a.n[L];
  const int kb = b * A.lz, ke = min(kb + A.lz, nL);
  const int lx = threadIdx.x % DC_TILE_X, ly = threadIdx.x / DC_TILE_X;
  const bool active = lx < TXa && (DC_DIM == 2 || ly < TYa);
  const bool xlo_cut = x0 > 0, xhi_cut = x0 + TXa < a.n[0];
  const bool ylo_cut = DC_DIM == 3 && y0 > 0, yhi_cut = DC_DIM == 3 && y0 + TYa < a.n[1];
  const int vs1 = a.n[0] + 1;
  const int vsL = DC_DIM == 3 ? vs1 * (a.n[1] + 1) : vs1;
  const int vbase = x0 + (DC_DIM == 3 ? y0 * vs1 : 0);

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

  // ---- vertex plane kp: start the copies global -> shared memory (this thread's vertices)
  auto issue = [&](int kp) {
    double* R = ring + (kp % 3) * SLOT;
    double* W = raw + (kp & 1) * RAW;
    for (int pv = threadIdx.x; pv < TPL; pv += NT) {
      const int vx = pv % TPX, vy = pv / TPX;
      if (vx > TXa || vy > TYa) continue;
      const long long d = a.dof_offset + (long long)(vbase + vx + vy * vs1 + kp * vsL) * NS;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const int e = s * TPL + pv;
        dc_cp_async8(R + e, a.x + d + s);
        if (MODE == 1) {
          if (A.pro == 0) {
            dc_cp_async8(R + FIELD + e, a.z + d + s);
          } else {
            dc_cp_async8(W + e, A.r_in + d + s);
            if (A.pro == 1 && !A.first) dc_cp_async8(W + FIELD + e, A.p_in + d + s);
            if (A.pro == 2 || !A.first) dc_cp_async8(W + 2 * FIELD + e, A.v_in + d + s);
            dc_cp_async8(W + 3 * FIELD + e, A.dinv + d + s);
          }
          if (A.epi == 1) dc_cp_async8(R + 2 * FIELD + e, A.w + d + s);
        }
        if (A.accumulate) dc_cp_async8(R + 2 * FIELD + e, a.r + d + s);
      }
    }
    dc_cp_async_commit();
  };

  // ---- vertex plane kp has landed: the prologue on this thread's vertices
  auto process = [&](int kp) {
    if (MODE != 1 || (A.pro == 0 && !a.cmask)) return;
    double* R = ring + (kp % 3) * SLOT;
    const double* W = raw + (kp & 1) * RAW;
    const bool plane_owned = kp < ke || ke == nL;         // a chunk owns its planes [kb, ke), the last one also nL
    const bool plane_counts = kp >= A.own_lo && kp < A.own_hi;
    for (int pv = threadIdx.x; pv < TPL; pv += NT) {
      const int vx = pv % TPX, vy = pv / TPX;
      if (vx > TXa || vy > TYa) continue;
      const long long d = a.dof_offset + (long long)(vbase + vx + vy * vs1 + kp * vsL) * NS;
      const bool owner = plane_owned && (vx < TXa || !xhi_cut) && (vy < TYa || !yhi_cut);
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const int e = s * TPL + pv;
        if (A.pro == 0) {
          if (a.cmask[d + s]) R[FIELD + e] = 0.0;
        } else if (A.pro == 1) {
          const double ri = W[e];
          const double pi = A.first ? ri : ri + beta * (W[FIELD + e] - omega * W[2 * FIELD + e]);
          if (owner) A.p_out[d + s] = pi;
          R[FIELD + e] = A.relax * W[3 * FIELD + e] * pi;
        } else {
          const double ri = W[e] - alpha * W[2 * FIELD + e];
          if (owner) {
            A.r_out[d + s] = ri;
            if (plane_counts) red[0] += ri * ri;
          }
          R[FIELD + e] = A.relax * W[3 * FIELD + e] * ri;
          R[2 * FIELD + e] = ri;
        }
      }
    }
  };

  // ---- finished plane k: O holds this CTA's sums (epilogue)
  auto finish = [&](int k) {
    const double* R = ring + (k % 3) * SLOT;
    const bool zlo = k == kb && kb > 0, zhi = k == ke && ke < nL;
    const bool plane_counts = k >= A.own_lo && k < A.own_hi;
    for (int pv = threadIdx.x; pv < TPL; pv += NT) {
      const int vx = pv % TPX, vy = pv / TPX;
      if (vx > TXa || vy > TYa) continue;
      const bool xlo = vx == 0 && xlo_cut, xhi = vx == TXa && xhi_cut;
      const bool ylo = DC_DIM == 3 && vy == 0 && ylo_cut, yhi = DC_DIM == 3 && vy == TYa && yhi_cut;
      const bool cut = xlo || xhi || ylo || yhi || zlo || zhi;
      const int slot = (xlo ? 1 : 0) | (ylo ? 2 : 0) | (zlo ? 4 : 0);
      const long long d = a.dof_offset + (long long)(vbase + vx + vy * vs1 + k * vsL) * NS;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const int e = s * TPL + pv;
        double val = O[e];
        if (cut && slot != 0) {
          A.slots[(slot - 1) * A.slot_stride + d + s] = val;
          continue;
        }
        if (A.accumulate) val += R[2 * FIELD + e];
        if (!cut) {
          if (MODE == 1 && A.identity && a.cmask[d + s]) val = a.z[d + s];   // identity row
          if (plane_counts) {
            if (A.epi == 1) red[1] += R[2 * FIELD + e] * val;
            if (A.epi == 2) { red[1] += val * R[2 * FIELD + e]; red[2] += val * val; }
          }
        }
        a.r[d + s] = val;
      }
    }
  };

  // ---- combine: the cells around a vertex of the finished plane
  double accb[NFACE][NS];
  auto combine = [&]() {
#if DC_DIM == 3
    // lanes are x neighbours: the x = 1 corners travel by shuffle, one phase per y row
#pragma unroll
    for (int fy = 0; fy < 2; ++fy) {
      const bool first = fy == 0 || ly + 1 == TYa;   // the first phase that reaches a vertex stores, the other adds
      const int pv = (ly + fy) * TPX + lx;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const double mine = active ? accb[2 * fy][s] : 0.0, right = active ? accb[2 * fy + 1][s] : 0.0;
        const double left = __shfl_up_sync(FULL, right, 1);
        if (active) {
          const double v = lx > 0 ? mine + left : mine;
          O[s * TPL + pv] = first ? v : O[s * TPL + pv] + v;
          if (lx + 1 == TXa) O[s * TPL + pv + 1] = first ? right : O[s * TPL + pv + 1] + right;
        }
      }
      __syncthreads();
    }
#else
#pragma unroll
    for (int f = 0; f < NFACE; ++f) {
      if (active) {
        const int pv = lx + f;
        const bool first = f == 0 || lx + 1 == TXa;
#pragma unroll
        for (int s = 0; s < NS; ++s) O[s * TPL + pv] = first ? accb[f][s] : O[s * TPL + pv] + accb[f][s];
      }
      __syncthreads();
    }
#endif
  };

#pragma unroll
  for (int f = 0; f < NFACE; ++f)
#pragma unroll
    for (int s = 0; s < NS; ++s) accb[f][s] = 0.0;
  int idx[3] = {x0 + lx, DC_DIM == 3 ? y0 + ly : 0, 0};
  issue(kb);
  issue(kb + 1);
  dc_cp_async_wait<1>();
  process(kb);
  for (int k = kb; k < ke; ++k) {
    if (k + 2 <= ke) issue(k + 2);
    else dc_cp_async_commit();
    dc_cp_async_wait<1>();
    process(k + 1);
    __syncthreads();   // planes k, k+1 staged and processed; the previous epilogue has read O
    double acct[NFACE][NS];
    if (active) {
      const double* Rb = ring + (k % 3) * SLOT;
      const double* Rt = ring + ((k + 1) % 3) * SLOT;
      double U[DC_NCORN][NS], Z[MODE == 1 ? DC_NCORN : 1][NS], acc[DC_NCORN][NS];
#pragma unroll
      for (int m = 0; m < DC_NCORN; ++m) {
        const double* R = (m >> L) ? Rt : Rb;
        const int pv = (ly + (DC_DIM == 3 ? (m >> 1) & 1 : 0)) * TPX + lx + (m & 1);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          U[m][s] = R[s * TPL + pv];
          if (MODE == 1) Z[m][s] = R[FIELD + s * TPL + pv];
        }
      }
      idx[L] = k;
      cell_fn(idx, U, Z, acc);
#pragma unroll
      for (int f = 0; f < NFACE; ++f)
#pragma unroll
        for (int s = 0; s < NS; ++s) { accb[f][s] += acc[f][s]; acct[f][s] = acc[f + NFACE][s]; }
    }
    combine();
    finish(k);
    if (active) {
#pragma unroll
      for (int f = 0; f < NFACE; ++f)
#pragma unroll
        for (int s = 0; s < NS; ++s) accb[f][s] = acct[f][s];
    }
  }
  dc_cp_async_wait<0>();
  __syncthreads();   // the last epilogue has read O
  combine();
  finish(ke);

  // ---- reduction partials of this CTA (summed in block order by la::tile_fixup)
  __syncthreads();
  double* sm = dc_tile_smem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    double v = red[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    if (lane == 0) sm[q * 32 + warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = 0.0;
    for (int w = 0; w < (NT + 31) / 32; ++w) v += sm[threadIdx.x * 32 + w];
    A.partials[(size_t)blockIdx.x * 4 + threadIdx.x] = v;
  }
}

template <int C, int MODE>
__device__ __forceinline__ void dc_tile_kernel(const DcTileArgs& A) {
  constexpr int NS = DcComp<C>::NS;
  dc_tile_march<C, MODE>(A, [&](const int* idx, const double (*U)[NS], const double (*Z)[NS], double (*acc)[NS]) {
    dc_struct_cell<C, MODE>(A.s, idx, U, Z, acc);
  });
}
#ifdef DC_TILE_Q1
template <int C, int MODE>
__device__ __forceinline__ void dc_tile_q1_kernel(const DcTileArgs& A) {
  constexpr int NS = DcComp<C>::NS;
  dc_tile_march<C, MODE>(A, [&](const int* idx, const double (*U)[NS], const double (*Z)[NS], double (*acc)[NS]) {
    dc_q1_cell<C, MODE>(A.s, idx, U, Z, acc);
  });
}
#endif
