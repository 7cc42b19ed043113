// =================================================================================================
// Patch kernels: one CTA per patch of elements (a compact, Morton-ordered run of the mesh).
//   1. the patch's vertex data (coordinates, coefficients) and its vertex->element adjacency are
//      staged once into shared memory with near-coalesced loads (patch vertex lists are ascending
//      runs of global ids);
//   2. phase A: each thread integrates elements (same DcElem code as the element kernels) reading
//      shared memory through 16-bit patch-local connectivity and writes the element results to a
//      shared buffer (conflict-free, element-major);
//   3. phase B: each patch vertex is owned by exactly one thread, which gathers the results of its
//      incident elements through the adjacency list in a fixed order -- no atomics and a
//      deterministic summation order inside the patch -- and issues one fp64 reduction per value
//      to global memory (vertices shared between patches are the only ones that meet another
//      writer).
// MODE 0: residual, 1: Jacobian apply (matrix free), 2: block diagonal of the Jacobian.
#ifndef DC_PATCH_THREADS
#define DC_PATCH_THREADS 256
#endif
#ifndef DC_PATCH_MINB
#define DC_PATCH_MINB 3
#endif

template <int C, int MODE>
__device__ __forceinline__ void dc_patch_kernel(const DcPatchArgs& a) {
  typedef DcComp<C> M;
  constexpr int NS = M::NS;
  constexpr int NV = MODE == 2 ? NS * NS : NS;  // values per (element, local vertex)
  constexpr int T = DC_PATCH_THREADS;
  extern __shared__ double dc_smem[];
  const int PN = a.max_nodes, PE = a.max_elems;
  double* sX = dc_smem;                                   // [PN][DIM]
  double* sU = sX + PN * DC_DIM;                          // [PN][NS]
  double* sZ = sU + PN * NS;                              // [PN][NS]   (MODE 1)
  double* sC = sZ + (MODE == 1 ? PN * NS : 0);            // [ND*NV][PE]
  unsigned short* sAdj = reinterpret_cast<unsigned short*>(sC + DC_ND * NV * PE);   // [PE*ND]
  unsigned short* sPtr = sAdj + PE * DC_ND;               // [PN+1]
  const int tid = threadIdx.x;
  for (int p = blockIdx.x; p < a.npatch; p += gridDim.x) {
    const int n0 = a.patch_node_ptr[p], np = a.patch_node_ptr[p + 1] - n0;
    const int e0 = a.patch_elem_ptr[p], ne = a.patch_elem_ptr[p + 1] - e0;
    const int adj0 = a.adj_ptr[n0];
    for (int ln = tid; ln < np; ln += T) {
      const int v = a.patch_nodes[n0 + ln];
#pragma unroll
      for (int k = 0; k < DC_DIM; ++k) sX[ln * DC_DIM + k] = a.coords[(long long)v * DC_DIM + k];
      const int dof = a.vdof ? a.vdof[v] : a.dof_offset + v * NS;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        sU[ln * NS + s] = a.x[dof + s];
        if (MODE == 1) sZ[ln * NS + s] = (a.cmask && a.cmask[dof + s]) ? 0.0 : a.z[dof + s];
      }
      sPtr[ln] = (unsigned short)(a.adj_ptr[n0 + ln] - adj0);
    }
    if (tid == 0) sPtr[np] = (unsigned short)(ne * DC_ND);
    for (int k = tid; k < ne * DC_ND; k += T) sAdj[k] = a.adj[adj0 + k];
    __syncthreads();
    // ---- phase A: integrate elements
    for (int el = tid; el < ne; el += T) {
      const ushort4 lc = reinterpret_cast<const ushort4*>(a.lconn)[e0 + el];
      const int lv[4] = {lc.x, lc.y, lc.z, lc.w};
      DcElem<C> E;
#pragma unroll
      for (int k = 0; k < DC_ND; ++k) {
#pragma unroll
        for (int c = 0; c < DC_DIM; ++c) E.X[k][c] = sX[lv[k] * DC_DIM + c];
#pragma unroll
        for (int s = 0; s < NS; ++s) E.xl[s][k] = sU[lv[k] * NS + s];
      }
      E.init_ctx(a.time, a.cell, a.ne_total, e0 + el);
      E.finish_geometry();
      if (MODE == 2) {
        double JS[NS][NS], JV[DC_ND][NS][NS], DD[NS][NS];
        E.jacobian_coefficients(a.wM, a.wA, JS, JV, DD);
#pragma unroll
        for (int k = 0; k < DC_ND; ++k)
#pragma unroll
          for (int i = 0; i < NS; ++i)
#pragma unroll
            for (int j = 0; j < NS; ++j)
              sC[(k * NV + i * NS + j) * PE + el] = M::pair(i, j) ? E.jacobian_entry(JS, JV, DD, i, k, j, k) : 0.0;
      } else {
        double loc[NS][DC_ND];
        if (MODE == 0) {
          E.residual(a.wM, a.wA, loc);
        } else {
          double zl[NS][DC_ND];
#pragma unroll
          for (int k = 0; k < DC_ND; ++k)
#pragma unroll
            for (int s = 0; s < NS; ++s) zl[s][k] = sZ[lv[k] * NS + s];
          E.jacobian_apply(a.wM, a.wA, zl, loc);
        }
#pragma unroll
        for (int k = 0; k < DC_ND; ++k)
#pragma unroll
          for (int s = 0; s < NS; ++s) sC[(k * NV + s) * PE + el] = loc[s][k];
      }
    }
    __syncthreads();
    // ---- phase B: every vertex gathers its incident element results (fixed order) and adds its
    //      total to global memory
    for (int ln = tid; ln < np; ln += T) {
      double acc[NV];
#pragma unroll
      for (int s = 0; s < NV; ++s) acc[s] = 0.0;
      const int kb = sPtr[ln], ke = sPtr[ln + 1];
      for (int k = kb; k < ke; ++k) {
        const unsigned ent = sAdj[k];
        const int el = (int)(ent >> 2), lk = (int)(ent & 3u);
#pragma unroll
        for (int s = 0; s < NV; ++s) acc[s] += sC[(lk * NV + s) * PE + el];
      }
      const int v = a.patch_nodes[n0 + ln];
      const int dof = a.vdof ? a.vdof[v] : a.dof_offset + v * NS;
      if (MODE == 2) {
#pragma unroll
        for (int s = 0; s < NV; ++s) dc_atomic_add(&a.bdiag[(long long)dof * NS + s], acc[s]);
      } else {
#pragma unroll
        for (int s = 0; s < NS; ++s) dc_atomic_add(&a.r[dof + s], acc[s]);
      }
    }
    __syncthreads();
  }
}
