// ---- [model.reduce]: evaluation / reduction functionals over the grid ------------------------------
// Device side of dune/copasi/model/diffusion_reaction/reduce.hh:38-203: for every cell and every
// point of a quadrature rule of order 4, value_k = reduction_k(evaluation_k(), value_k).  The
// generated part (reduce.cpp) provides DC_NRED, DcReduce<C>::eval (all evaluation expressions with
// the fields of compartment C bound, other species read 0 as after LocalEquations::clear) and
// dc_reduce_op (the reduction expressions; default a + b).
//
// One thread walks a strided set of cells sequentially, the block combines its threads with the
// same operation in a fixed tree, and the host folds the per-block partials in block order, so the
// result is deterministic.  Like the reference's multi-threaded path, every partial starts from
// `initial.value`, which therefore has to be the neutral element of the reduction.
//
// Quadrature (dune-geometry's tables are not in the reference tree, parity unpinned): triangle =
// the 6-point rule of degree 4 (Dunavant), tetrahedron = the 15-point rule of degree 5 (Stroud T3:5-1).

#define DC_RED_THREADS 128
#if DC_DIM == 2
#define DC_RED_NQ 6
__device__ __forceinline__ double dc_red_point(int q, double* lam) {
  const double a = q < 3 ? 0.445948490915965 : 0.091576213509771;
  const double w = q < 3 ? 0.223381589678011 : 0.109951743655322;
  const int odd = q % 3;   // the vertex that carries 1 - 2a
#pragma unroll
  for (int k = 0; k < 3; ++k) lam[k] = k == odd ? 1.0 - 2.0 * a : a;
  return 0.5 * w;
}
#else
#define DC_RED_NQ 15
__device__ __forceinline__ double dc_red_point(int q, double* lam) {
  const double s15 = 3.872983346207417;
  if (q == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) lam[k] = 0.25;
    return (16.0 / 135.0) / 6.0;
  }
  if (q < 9) {
    const bool first = q < 5;
    const double a = first ? (7.0 - s15) / 34.0 : (7.0 + s15) / 34.0;
    const double w = first ? (2665.0 + 14.0 * s15) / 37800.0 : (2665.0 - 14.0 * s15) / 37800.0;
    const int odd = (q - 1) & 3;
#pragma unroll
    for (int k = 0; k < 4; ++k) lam[k] = k == odd ? 1.0 - 3.0 * a : a;
    return w / 6.0;
  }
  // six points (b, b, 1/2 - b, 1/2 - b): the pair of vertices that carries b
  const double b = (10.0 - 2.0 * s15) / 40.0;
  const int p = q - 9;
  const int i0 = p < 3 ? 0 : (p < 5 ? 1 : 2);
  const int i1 = p < 3 ? p + 1 : (p < 5 ? p - 1 : 3);
#pragma unroll
  for (int k = 0; k < 4; ++k) lam[k] = (k == i0 || k == i1) ? b : 0.5 - b;
  return (10.0 / 189.0) / 6.0;
}
#endif

template <int C>
__device__ __forceinline__ void dc_reduce_kernel(const DcReduceArgs& a) {
  typedef DcReduce<C> R;
  constexpr int NS = R::NS;
  double acc[DC_NRED];
#pragma unroll
  for (int k = 0; k < DC_NRED; ++k) acc[k] = a.init[k];
  DcCtx c;
  c.time = a.time;
  c.in_volume = 1.0; c.in_boundary = 0.0; c.in_skeleton = 0.0;
  c.nrm[0] = c.nrm[1] = c.nrm[2] = 0.0;
  c.pos[2] = 0.0;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < a.n; t += (long long)gridDim.x * blockDim.x) {
    const long long e = a.elem_ids ? (long long)a.elem_ids[t] : t;
    double X[DC_ND][DC_DIM], G[DC_ND][DC_DIM], xl[NS][DC_ND], gu[NS][DC_DIM];
#pragma unroll
    for (int k = 0; k < DC_ND; ++k) {
      const int v = a.elems[e * DC_ND + k];
#pragma unroll
      for (int d = 0; d < DC_DIM; ++d) X[k][d] = a.coords[(long long)v * DC_DIM + d];
      if (R::NS_REAL > 0) {
        const int dof = a.vdof ? a.vdof[v] : a.dof_offset + v * NS;
#pragma unroll
        for (int s = 0; s < NS; ++s) xl[s][k] = a.x[dof + s];
      } else {
        xl[0][k] = 0.0;
      }
    }
    const double adet = dc_geometry(X, G);
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
      for (int d = 0; d < DC_DIM; ++d) {
        double g = 0.0;
#pragma unroll
        for (int k = 0; k < DC_ND; ++k) g += xl[s][k] * G[k][d];
        gu[s][d] = g;
      }
    c.entity_volume = adet / DC_FACT;
#pragma unroll
    for (int k = 0; k < DC_NKEYS; ++k) c.cell[k] = a.cell[(long long)k * a.ne_total + e];
#pragma unroll 1
    for (int q = 0; q < DC_RED_NQ; ++q) {
      double lam[DC_ND], u[NS], val[DC_NRED];
      const double w = dc_red_point(q, lam);
#pragma unroll
      for (int d = 0; d < DC_DIM; ++d) {
        double p = 0.0;
#pragma unroll
        for (int k = 0; k < DC_ND; ++k) p += lam[k] * X[k][d];
        c.pos[d] = p;
      }
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < DC_ND; ++k) v += lam[k] * xl[s][k];
        u[s] = v;
      }
      c.integration_factor = w * adet;
      R::eval(c, u, gu, val);
#pragma unroll
      for (int k = 0; k < DC_NRED; ++k) acc[k] = dc_reduce_op(k, val[k], acc[k]);
    }
  }
  __shared__ double sh[DC_RED_THREADS * DC_NRED];
#pragma unroll
  for (int k = 0; k < DC_NRED; ++k) sh[threadIdx.x * DC_NRED + k] = acc[k];
  __syncthreads();
  for (int stride = DC_RED_THREADS / 2; stride > 0; stride >>= 1) {
    if ((int)threadIdx.x < stride) {
#pragma unroll
      for (int k = 0; k < DC_NRED; ++k)
        sh[threadIdx.x * DC_NRED + k] = dc_reduce_op(k, sh[(threadIdx.x + stride) * DC_NRED + k], sh[threadIdx.x * DC_NRED + k]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < DC_NRED; ++k) a.partials[(long long)blockIdx.x * DC_NRED + k] = sh[k];
  }
}

// ---- Q1 cells (lattice cubes as multilinear elements, kernels/assembly_q1.cuh): the order-4 rule of a cube is the
// 3-point Gauss rule per axis (dune-geometry's cube rules are Gauss-Legendre tensor products); corner a sits at the
// bit pattern of a (x = bit 0), points run x fastest.  The gradient of a field varies inside the cell and is formed
// per point.  Restated in oracle/core.py (reduce, etype 1).
#define DC_Q1_ND (1 << DC_DIM)
#if DC_DIM == 2
#define DC_Q1_NQ 9
#else
#define DC_Q1_NQ 27
#endif
template <int C>
__device__ __forceinline__ void dc_reduce_q1_kernel(const DcReduceArgs& a) {
  typedef DcReduce<C> R;
  constexpr int NS = R::NS;
  double acc[DC_NRED];
#pragma unroll
  for (int k = 0; k < DC_NRED; ++k) acc[k] = a.init[k];
  DcCtx c;
  c.time = a.time;
  c.in_volume = 1.0; c.in_boundary = 0.0; c.in_skeleton = 0.0;
  c.nrm[0] = c.nrm[1] = c.nrm[2] = 0.0;
  c.pos[2] = 0.0;
  const double g1[3] = {0.5 - 0.3872983346207417, 0.5, 0.5 + 0.3872983346207417};   // 1/2 -+ sqrt(0.15)
  const double w1[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < a.n; t += (long long)gridDim.x * blockDim.x) {
    const long long e = a.elem_ids ? (long long)a.elem_ids[t] : t;
    double X[DC_Q1_ND][DC_DIM], xl[NS][DC_Q1_ND], h[DC_DIM];
#pragma unroll
    for (int k = 0; k < DC_Q1_ND; ++k) {
      const int v = a.elems[e * DC_Q1_ND + k];
#pragma unroll
      for (int d = 0; d < DC_DIM; ++d) X[k][d] = a.coords[(long long)v * DC_DIM + d];
      if (R::NS_REAL > 0) {
        const int dof = a.vdof ? a.vdof[v] : a.dof_offset + v * NS;
#pragma unroll
        for (int s = 0; s < NS; ++s) xl[s][k] = a.x[dof + s];
      } else {
        xl[0][k] = 0.0;
      }
    }
    double det = 1.0;
#pragma unroll
    for (int d = 0; d < DC_DIM; ++d) { h[d] = X[1 << d][d] - X[0][d]; det *= h[d]; }
    c.entity_volume = det;
#pragma unroll
    for (int k = 0; k < DC_NKEYS; ++k) c.cell[k] = a.cell[(long long)k * a.ne_total + e];
#pragma unroll 1
    for (int q = 0; q < DC_Q1_NQ; ++q) {
      int qi[3] = {q % 3, (q / 3) % 3, q / 9};
      double w = 1.0, p[DC_DIM];
#pragma unroll
      for (int d = 0; d < DC_DIM; ++d) { p[d] = g1[qi[d]]; w *= w1[qi[d]]; }
      double lam[DC_Q1_ND], u[NS], gu[NS][DC_DIM], val[DC_NRED];
#pragma unroll
      for (int m = 0; m < DC_Q1_ND; ++m) {
        double f = 1.0;
#pragma unroll
        for (int d = 0; d < DC_DIM; ++d) f *= ((m >> d) & 1) ? p[d] : 1.0 - p[d];
        lam[m] = f;
      }
#pragma unroll
      for (int d = 0; d < DC_DIM; ++d) {
        double x = 0.0;
#pragma unroll
        for (int m = 0; m < DC_Q1_ND; ++m) x += lam[m] * X[m][d];
        c.pos[d] = x;
      }
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        double v = 0.0;
#pragma unroll
        for (int m = 0; m < DC_Q1_ND; ++m) v += lam[m] * xl[s][m];
        u[s] = v;
#pragma unroll
        for (int r = 0; r < DC_DIM; ++r) {
          double g = 0.0;
#pragma unroll
          for (int m = 0; m < DC_Q1_ND; ++m) {
            double f = ((m >> r) & 1) ? 1.0 : -1.0;
#pragma unroll
            for (int d = 0; d < DC_DIM; ++d)
              if (d != r) f *= ((m >> d) & 1) ? p[d] : 1.0 - p[d];
            g += xl[s][m] * f;
          }
          gu[s][r] = g / h[r];
        }
      }
      c.integration_factor = w * det;
      R::eval(c, u, gu, val);
#pragma unroll
      for (int k = 0; k < DC_NRED; ++k) acc[k] = dc_reduce_op(k, val[k], acc[k]);
    }
  }
  __shared__ double sh[DC_RED_THREADS * DC_NRED];
#pragma unroll
  for (int k = 0; k < DC_NRED; ++k) sh[threadIdx.x * DC_NRED + k] = acc[k];
  __syncthreads();
  for (int stride = DC_RED_THREADS / 2; stride > 0; stride >>= 1) {
    if ((int)threadIdx.x < stride) {
#pragma unroll
      for (int k = 0; k < DC_NRED; ++k)
        sh[threadIdx.x * DC_NRED + k] = dc_reduce_op(k, sh[(threadIdx.x + stride) * DC_NRED + k], sh[threadIdx.x * DC_NRED + k]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < DC_NRED; ++k) a.partials[(long long)blockIdx.x * DC_NRED + k] = sh[k];
  }
}
