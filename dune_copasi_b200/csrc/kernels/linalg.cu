// Statically compiled sm_100a kernels of the Krylov solve: SpMV, fused BLAS-1 sweeps, Jacobi and
// block-Jacobi, Dirichlet masks, halo pack/unpack.  Everything here is HBM-bound fp64 streaming
// work: one double per thread and grid stride (coalesced 256-byte requests per warp; measured at the
// HBM copy bandwidth, profiles/r02_timeline_n1.json), grids sized as multiples of the 148 SMs,
// reductions deterministic (fixed grid, ordered final sum), collectives fused into the last block.
#include "linalg.hpp"

#include "peer_device.cuh"

#include "../util.hpp"

namespace dcb {
namespace la {

namespace {

constexpr int kThreads = 256;
constexpr int kNumSms = 148;

inline int grid_for(int64_t n, int per_thread = 1) {
  int64_t blocks = (n + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
  if (blocks < 1) blocks = 1;
  if (blocks > kMaxBlocks) blocks = kMaxBlocks;  // grid-stride beyond 8 CTAs per SM
  return (int)blocks;
}

__device__ __forceinline__ bool in_ranges(const Ranges& r, long long i) {
  bool hit = false;
#pragma unroll 1
  for (int k = 0; k < r.n; ++k) hit |= (i >= r.b[k] && i < r.e[k]);
  return hit;
}

// block reduce NOUT values, then "last block sums the partials in order".  With a link the last block goes on:
// it publishes the halo planes the blocks have pushed while streaming, all-reduces the sums over the peer
// mailboxes (peer_device.cuh) and mirrors them into mapped host memory -- no k_halo / k_allreduce launch and no
// device-to-host copy between two sweeps.
template <int NOUT>
__device__ void grid_reduce(double (&v)[NOUT], double* partials, unsigned* counter, double* out, const Link& L = Link()) {
  __shared__ double sm[NOUT][kThreads / 32];
  __shared__ double fin[peer::kMaxWords];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NOUT; ++k) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sm[k][warp] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
      double x = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) x += sm[k][w];
      partials[(size_t)blockIdx.x * NOUT + k] = x;
    }
    __threadfence();
    const unsigned ticket = atomicInc(counter, gridDim.x - 1);  // wraps back to 0: self resetting
    last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (L.peer.push && threadIdx.x == 0) peer::publish_halo(L.peer.m, L.peer.h, L.peer.halo_seq);
#pragma unroll
  for (int k = 0; k < NOUT; ++k) {
    double x = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += kThreads) x += partials[(size_t)b * NOUT + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sm[k][warp] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
      double x = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) x += sm[k][w];
      fin[k] = x;
    }
  }
  if (L.peer.reduce) peer::block_allreduce(L.peer.m, fin, NOUT, L.peer.ar_seq, L.peer.h.error);
  else __syncthreads();
  if (threadIdx.x < NOUT) {
    out[threadIdx.x] = fin[threadIdx.x];
    if (L.host_out) L.host_out[threadIdx.x] = fin[threadIdx.x];
  }
}

// sweeps without a reduction: count the blocks, the last one publishes the pushed planes
__device__ __forceinline__ void finish_push(unsigned* counter, const Link& L) {
  if (!L.peer.push) return;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicInc(counter, gridDim.x - 1);
    if (ticket == gridDim.x - 1) peer::publish_halo(L.peer.m, L.peer.h, L.peer.halo_seq);
  }
}

// ---------------------------------------------------------------- SpMV
// T lanes cooperate on one row; consecutive lanes read consecutive nonzeros (coalesced).
template <int T, class RP>
__global__ void __launch_bounds__(kThreads) k_spmv(int64_t nrows, const RP* __restrict__ rowptr,
                                                   const int32_t* __restrict__ colidx,
                                                   const double* __restrict__ vals,
                                                   const double* __restrict__ x, double* __restrict__ y) {
  const int sub = threadIdx.x % T;
  const int64_t rows_per_block = kThreads / T;
  for (int64_t row = blockIdx.x * rows_per_block + threadIdx.x / T; row < nrows;
       row += (int64_t)gridDim.x * rows_per_block) {
    const int64_t b = rowptr[row], e = rowptr[row + 1];
    double acc = 0.0;
    for (int64_t k = b + sub; k < e; k += T) acc += vals[k] * __ldg(&x[colidx[k]]);
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, T);
    if (sub == 0) y[row] = acc;
  }
}

template <class RP>
void spmv_launch(int64_t nrows, const RP* rp, const int32_t* ci, const double* va, const double* x,
                 double* y, int avg, cudaStream_t s) {
  // rows per block = 256/T; grid capped at a multiple of the SM count with grid-stride rows
  auto grid = [&](int T) {
    int64_t b = (nrows * T + kThreads - 1) / kThreads;
    int64_t cap = (int64_t)kNumSms * 64;
    return (int)std::max<int64_t>(1, std::min(b, cap));
  };
  if (avg <= 3) k_spmv<2, RP><<<grid(2), kThreads, 0, s>>>(nrows, rp, ci, va, x, y);
  else if (avg <= 10) k_spmv<4, RP><<<grid(4), kThreads, 0, s>>>(nrows, rp, ci, va, x, y);
  else if (avg <= 48) k_spmv<8, RP><<<grid(8), kThreads, 0, s>>>(nrows, rp, ci, va, x, y);
  else if (avg <= 128) k_spmv<16, RP><<<grid(16), kThreads, 0, s>>>(nrows, rp, ci, va, x, y);
  else k_spmv<32, RP><<<grid(32), kThreads, 0, s>>>(nrows, rp, ci, va, x, y);
}

// ---------------------------------------------------------------- BLAS-1
__global__ void __launch_bounds__(kThreads) k_dot(Ranges own, const double* __restrict__ a,
                                                  const double* __restrict__ b, double* partials,
                                                  unsigned* counter, double* out, Link L) {
  double v[1] = {0.0};
  for (int k = 0; k < own.n; ++k)
    for (long long i = own.b[k] + blockIdx.x * (long long)kThreads + threadIdx.x; i < own.e[k];
         i += (long long)gridDim.x * kThreads)
      v[0] += a[i] * b[i];
  grid_reduce<1>(v, partials, counter, out, L);
}

__global__ void __launch_bounds__(kThreads) k_dot2(Ranges own, const double* __restrict__ a,
                                                   const double* __restrict__ b,
                                                   const double* __restrict__ c,
                                                   const double* __restrict__ d, double* partials,
                                                   unsigned* counter, double* out, Link L) {
  double v[2] = {0.0, 0.0};
  for (int k = 0; k < own.n; ++k)
    for (long long i = own.b[k] + blockIdx.x * (long long)kThreads + threadIdx.x; i < own.e[k];
         i += (long long)gridDim.x * kThreads) {
      v[0] += a[i] * b[i];
      v[1] += c[i] * d[i];
    }
  grid_reduce<2>(v, partials, counter, out, L);
}

__global__ void __launch_bounds__(kThreads) k_bicg_p(int64_t n, double* __restrict__ p,
                                                     const double* __restrict__ r,
                                                     const double* __restrict__ v, double beta,
                                                     double omega, bool first) {
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    p[i] = first ? r[i] : r[i] + beta * (p[i] - omega * v[i]);
}

__global__ void __launch_bounds__(kThreads) k_axpy_pair_norm(int64_t n, Ranges own, double alpha,
                                                             const double* __restrict__ y,
                                                             double* __restrict__ x,
                                                             const double* __restrict__ v,
                                                             double* __restrict__ r,
                                                             const double* __restrict__ rt,
                                                             double* partials, unsigned* counter,
                                                             double* out) {
  double acc[2] = {0.0, 0.0};
  const bool single = own.n == 1 && own.b[0] == 0 && own.e[0] == n;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    x[i] += alpha * y[i];
    const double ri = r[i] - alpha * v[i];
    r[i] = ri;
    if (single || in_ranges(own, i)) {
      acc[0] += ri * ri;
      if (rt) acc[1] += rt[i] * ri;
    }
  }
  grid_reduce<2>(acc, partials, counter, out);
}

// ---- fused BiCGSTAB sweeps (Jacobi folded into the producing kernel when dinv != null) ----------
// Every step length is formed on the device from the (all-reduced) device-resident sums, in the
// operation order of the host formulas of dune-istl, so that the host never has to wait for a
// scalar before it can enqueue the next sweep:
//   rho_new = <rt,r> of this iteration, rho = the one before, h = <rt,v>, trtt = (<t,r>, <t,t>)
//   alpha = rho / h (of the same iteration), omega = tr / tt, beta = (rho_new / rho) * (alpha / omega)
// p = r + beta (p - omega v) ; y = relax * dinv * p
__global__ void __launch_bounds__(kThreads) k_bicg_p_prec(int64_t n, double* __restrict__ p,
                                                          const double* __restrict__ r,
                                                          double* v,
                                                          const double* __restrict__ rho_new_p,
                                                          const double* __restrict__ rho_p,
                                                          const double* __restrict__ hptr,
                                                          const double* __restrict__ trtt, bool first,
                                                          const double* __restrict__ dinv, double relax,
                                                          double* __restrict__ y, unsigned* counter, Link L) {
  double beta = 0.0, omega = 0.0;
  if (!first) {
    const double rho = *rho_p, alpha = rho / *hptr;
    omega = trtt[0] / trtt[1];
    beta = (*rho_new_p / rho) * (alpha / omega);
  }
  const int par = (int)(L.peer.halo_seq & 1ull);
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const double pi = first ? r[i] : r[i] + beta * (p[i] - omega * v[i]);
    p[i] = pi;
    if (L.zero_input) v[i] = 0.0;   // v is done: the operator application that follows accumulates into it
    double out = pi;                // what the operator application reads: y, or p itself (it applies D^-1 on the fly)
    if (dinv) {
      out = relax * dinv[i] * pi;
      y[i] = out;
    }
    if (L.peer.push) peer::push_entry(L.peer.m, L.peer.h, par, i, out);
  }
  finish_push(counter, L);
}
// r -= alpha v ; out[0] = <r,r> ; y2 = relax * dinv * r
__global__ void __launch_bounds__(kThreads) k_bicg_r_prec(int64_t n, Ranges own,
                                                          const double* __restrict__ rho_p,
                                                          const double* __restrict__ hptr,
                                                          const double* __restrict__ v,
                                                          double* __restrict__ r,
                                                          const double* __restrict__ dinv, double relax,
                                                          double* __restrict__ y2, double* partials,
                                                          unsigned* counter, double* out, Link L) {
  double acc[1] = {0.0};
  const bool single = own.n == 1 && own.b[0] == 0 && own.e[0] == n;
  const double alpha = *rho_p / *hptr;   // alpha = rho'/<rt,v> from the device-resident reductions
  const int par = (int)(L.peer.halo_seq & 1ull);
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const double ri = r[i] - alpha * v[i];
    r[i] = ri;
    double out = ri;
    if (dinv) {
      out = relax * dinv[i] * ri;
      y2[i] = out;
    }
    if (L.peer.push) peer::push_entry(L.peer.m, L.peer.h, par, i, out);
    if (single || in_ranges(own, i)) acc[0] += ri * ri;
  }
  if (L.peer.push) __threadfence_system();
  grid_reduce<1>(acc, partials, counter, out, L);
}
// xout = xin + alpha y1 + omega y2 ; r -= omega t ; out[0] = <r,r> ; out[1] = <rt,r>
__global__ void __launch_bounds__(kThreads) k_bicg_final(int64_t n, Ranges own,
                                                         const double* __restrict__ rho_p,
                                                         const double* __restrict__ hptr,
                                                         const double* __restrict__ trtt,
                                                         const double* __restrict__ y1,
                                                         const double* __restrict__ y2,
                                                         const double* __restrict__ xin,
                                                         double* __restrict__ xout,
                                                         double* t,
                                                         double* __restrict__ r,
                                                         const double* __restrict__ rt, double* partials,
                                                         unsigned* counter, double* out, Link L) {
  double acc[2] = {0.0, 0.0};
  const bool single = own.n == 1 && own.b[0] == 0 && own.e[0] == n;
  const double alpha = *rho_p / *hptr, omega = trtt[0] / trtt[1];
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    xout[i] = (xin[i] + alpha * y1[i]) + omega * y2[i];
    const double ri = r[i] - omega * t[i];
    r[i] = ri;
    if (L.zero_input) t[i] = 0.0;   // t is done: the next operator application accumulates into it
    if (single || in_ranges(own, i)) {
      acc[0] += ri * ri;
      acc[1] += rt[i] * ri;
    }
  }
  grid_reduce<2>(acc, partials, counter, out, L);
}

// One step of modified Gram-Schmidt in a single pass: w -= (*coef) vprev (skipped when coef is null), then
// out[0] = <vnext, w> over the owned entries (vnext null: <w, w>).  The unfused pair (k_axpy_dev, k_dot) reads w twice.
__global__ void __launch_bounds__(kThreads) k_mgs_step(int64_t n, Ranges own, const double* __restrict__ coef,
                                                       const double* __restrict__ vprev, double* __restrict__ w,
                                                       const double* __restrict__ vnext, double* partials,
                                                       unsigned* counter, double* out, Link L) {
  double acc[1] = {0.0};
  const bool single = own.n == 1 && own.b[0] == 0 && own.e[0] == n;
  const double a = coef ? *coef : 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    double wi = w[i];
    if (coef) { wi -= a * vprev[i]; w[i] = wi; }
    if (single || in_ranges(own, i)) acc[0] += (vnext ? vnext[i] : wi) * wi;
  }
  grid_reduce<1>(acc, partials, counter, out, L);
}

__global__ void __launch_bounds__(kThreads) k_xpby(int64_t n, double* __restrict__ p,
                                                   const double* __restrict__ q, double beta) {
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    p[i] = q[i] + beta * p[i];
}
__global__ void __launch_bounds__(kThreads) k_axpy(int64_t n, double a, const double* __restrict__ x,
                                                   double* __restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    y[i] += a * x[i];
}
__global__ void __launch_bounds__(kThreads) k_fill(int64_t n, double v, double* __restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    y[i] = v;
}

// ---------------------------------------------------------------- preconditioners
__global__ void __launch_bounds__(kThreads) k_jacobi(int64_t n, const double* __restrict__ dinv,
                                                     double relax, const double* __restrict__ d,
                                                     double* __restrict__ v) {
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    v[i] = relax * dinv[i] * d[i];
}

__global__ void __launch_bounds__(kThreads) k_diag_inv(int64_t n, const int64_t* __restrict__ rowptr,
                                                       const int32_t* __restrict__ colidx,
                                                       const double* __restrict__ vals,
                                                       double* __restrict__ dinv) {
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    double d = 0.0;
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
      if (colidx[k] == i) d = vals[k];
    dinv[i] = 1.0 / d;
  }
}

__global__ void __launch_bounds__(kThreads) k_block_diag(int64_t dof0, int64_t nrows, int bs,
                                                         const int64_t* __restrict__ rowptr,
                                                         const int32_t* __restrict__ colidx,
                                                         const double* __restrict__ vals,
                                                         double* __restrict__ bdiag) {
  for (int64_t t = blockIdx.x * (int64_t)kThreads + threadIdx.x; t < nrows; t += (int64_t)gridDim.x * kThreads) {
    const int64_t row = dof0 + t, blk0 = dof0 + (t / bs) * bs;
    for (int j = 0; j < bs; ++j) bdiag[row * bs + j] = 0.0;
    for (int64_t k = rowptr[row]; k < rowptr[row + 1]; ++k) {
      const int64_t c = colidx[k];
      if (c >= blk0 && c < blk0 + bs) bdiag[row * bs + (c - blk0)] = vals[k];
    }
  }
}

__global__ void __launch_bounds__(128) k_block_invert(int64_t nblocks, int bs, double* __restrict__ blocks) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  double A[19 * 19], B[19 * 19];
  double* g = blocks + b * bs * bs;
  for (int i = 0; i < bs; ++i)
    for (int j = 0; j < bs; ++j) { A[i * bs + j] = g[i * bs + j]; B[i * bs + j] = i == j ? 1.0 : 0.0; }
  for (int p = 0; p < bs; ++p) {
    int piv = p;
    for (int i = p + 1; i < bs; ++i)
      if (fabs(A[i * bs + p]) > fabs(A[piv * bs + p])) piv = i;
    if (piv != p)
      for (int j = 0; j < bs; ++j) {
        double t = A[p * bs + j]; A[p * bs + j] = A[piv * bs + j]; A[piv * bs + j] = t;
        t = B[p * bs + j]; B[p * bs + j] = B[piv * bs + j]; B[piv * bs + j] = t;
      }
    const double ip = 1.0 / A[p * bs + p];
    for (int j = 0; j < bs; ++j) { A[p * bs + j] *= ip; B[p * bs + j] *= ip; }
    for (int i = 0; i < bs; ++i)
      if (i != p) {
        const double f = A[i * bs + p];
        for (int j = 0; j < bs; ++j) { A[i * bs + j] -= f * A[p * bs + j]; B[i * bs + j] -= f * B[p * bs + j]; }
      }
  }
  for (int i = 0; i < bs * bs; ++i) g[i] = B[i];
}

__global__ void __launch_bounds__(kThreads) k_block_jacobi(int64_t dof0, int64_t nrows, int bs,
                                                           const double* __restrict__ binv, double relax,
                                                           const double* __restrict__ d,
                                                           double* __restrict__ v) {
  for (int64_t t = blockIdx.x * (int64_t)kThreads + threadIdx.x; t < nrows; t += (int64_t)gridDim.x * kThreads) {
    const int64_t row = dof0 + t, blk0 = dof0 + (t / bs) * bs;
    double acc = 0.0;
    for (int j = 0; j < bs; ++j) acc += binv[row * bs + j] * d[blk0 + j];
    v[row] = relax * acc;
  }
}

__global__ void __launch_bounds__(kThreads) k_bdiag_dinv(int64_t dof0, int64_t nrows, int bs,
                                                         const double* __restrict__ bdiag,
                                                         double* __restrict__ dinv) {
  for (int64_t t = blockIdx.x * (int64_t)kThreads + threadIdx.x; t < nrows; t += (int64_t)gridDim.x * kThreads) {
    const int64_t row = dof0 + t;
    dinv[row] = 1.0 / bdiag[row * bs + (t % bs)];
  }
}

__global__ void __launch_bounds__(kThreads) k_axpy_dev(int64_t n, const double* __restrict__ coef, double sign,
                                                       const double* __restrict__ x, double* __restrict__ y) {
  const double a = sign * *coef;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    y[i] += a * x[i];
}
__global__ void __launch_bounds__(kThreads) k_normalize_dev(int64_t n, const double* __restrict__ src,
                                                            const double* __restrict__ norm2,
                                                            double* __restrict__ dst) {
  const double a = 1.0 / sqrt(*norm2);
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    dst[i] = src[i] * a;
}
__global__ void __launch_bounds__(kThreads) k_scale(int64_t n, double a, double* __restrict__ x) {
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    x[i] *= a;
}
__global__ void __launch_bounds__(kThreads) k_sub(int64_t n, const double* __restrict__ a,
                                                  const double* __restrict__ b, double* __restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    y[i] = a[i] - b[i];
}

__global__ void __launch_bounds__(kThreads) k_invert_diag(int64_t n, double* __restrict__ d,
                                                          const unsigned char* __restrict__ mask) {
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    d[i] = (mask && mask[i]) ? 1.0 : 1.0 / d[i];
}

// ---------------------------------------------------------------- Dirichlet / halo
__global__ void k_set_values(int64_t n, const int32_t* idx, const double* vals, double* x) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) x[idx[i]] = vals ? vals[i] : 0.0;
}
__global__ void k_copy_values(int64_t n, const int32_t* idx, const double* src, double* dst) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = src[idx[i]];
}
__global__ void __launch_bounds__(kThreads) k_csr_constrain(int64_t nrows, const int64_t* __restrict__ rowptr,
                                                            const int32_t* __restrict__ colidx,
                                                            double* __restrict__ vals,
                                                            const unsigned char* __restrict__ mask) {
  for (int64_t row = blockIdx.x * (int64_t)kThreads + threadIdx.x; row < nrows; row += (int64_t)gridDim.x * kThreads) {
    const bool rc = mask[row];
    for (int64_t k = rowptr[row]; k < rowptr[row + 1]; ++k) {
      const int32_t c = colidx[k];
      if (rc || mask[c]) vals[k] = (rc && c == row) ? 1.0 : 0.0;
    }
  }
}
__global__ void __launch_bounds__(kThreads) k_bdiag_constrain(int64_t dof0, int64_t nrows, int bs,
                                                              double* __restrict__ bdiag,
                                                              const unsigned char* __restrict__ mask) {
  for (int64_t t = blockIdx.x * (int64_t)kThreads + threadIdx.x; t < nrows; t += (int64_t)gridDim.x * kThreads) {
    const int64_t row = dof0 + t, blk0 = dof0 + (t / bs) * bs;
    const bool rc = mask[row];
    for (int j = 0; j < bs; ++j)
      if (rc || mask[blk0 + j]) bdiag[row * bs + j] = (rc && blk0 + j == row) ? 1.0 : 0.0;
  }
}
__global__ void k_gather(int64_t n, const int32_t* idx, const double* x, double* buf) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) buf[i] = x[idx[i]];
}
__global__ void k_scatter(int64_t n, const int32_t* idx, const double* buf, double* x) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) x[idx[i]] = buf[i];
}

// ---------------------------------------------------------------- tile fix-up
// One thread per cut vertex, three families: x-cut planes, y-cut planes (vertices that are not
// x-cut), chunk boundary planes (vertices that are neither).
__global__ void __launch_bounds__(kThreads) k_tile_fixup(TileFixup f, double* partials, unsigned* counter) {
  const long long P0 = f.n[0] + 1, P1 = f.n[1] + 1, P2 = f.n[2] + 1;
  const long long ncx = f.n[0] > 0 ? (f.n[0] - 1) / f.tile[0] : 0;
  const long long ncy = f.n[1] > 0 ? (f.n[1] - 1) / f.tile[1] : 0;
  const long long ncz = f.n[2] > 0 ? (f.n[2] - 1) / f.tile[2] : 0;
  const long long famx = ncx * P1 * P2, famy = ncy * P0 * P2, famz = ncz * P0 * P1;
  const long long total = famx + famy + famz;
  double red[3] = {0.0, 0.0, 0.0};
  for (long long t = blockIdx.x * (long long)kThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kThreads) {
    long long vx, vy, vz;
    if (t < famx) {
      vx = (t % ncx + 1) * f.tile[0];
      vy = (t / ncx) % P1;
      vz = t / (ncx * P1);
    } else if (t < famx + famy) {
      const long long u = t - famx;
      vx = u % P0;
      vy = ((u / P0) % ncy + 1) * f.tile[1];
      vz = u / (P0 * ncy);
    } else {
      const long long u = t - famx - famy;
      vx = u % P0;
      vy = (u / P0) % P1;
      vz = (u / (P0 * P1) + 1) * f.tile[2];
    }
    const bool xc = vx > 0 && vx < f.n[0] && vx % f.tile[0] == 0;
    const bool yc = vy > 0 && vy < f.n[1] && vy % f.tile[1] == 0;
    const bool zc = vz > 0 && vz < f.n[2] && vz % f.tile[2] == 0;
    if (t >= famx && xc) continue;               // counted in the x family
    if (t >= famx + famy && yc) continue;        // counted in the y family
    const int m = (xc ? 1 : 0) | (yc ? 2 : 0) | (zc ? 4 : 0);
    const long long d = f.dof_offset + (vx + vy * P0 + vz * P0 * P1) * f.ns;
    const bool counts = vz >= f.own_lo && vz < f.own_hi;
    for (int s = 0; s < f.ns; ++s) {
      double val = f.y[d + s];
      for (int k = 1; k < 8; ++k)
        if ((k & ~m) == 0) val += f.slots[(k - 1) * f.slot_stride + d + s];
      if (f.cmask && f.cmask[d + s]) val = f.zraw[d + s];
      f.y[d + s] = val;
      if (counts) {
        if (f.epi == 1) red[1] += f.w[d + s] * val;
        if (f.epi == 2) { red[1] += val * f.aux[d + s]; red[2] += val * val; }
      }
    }
  }
  // block sums, then the last block adds everything up in a fixed order
  __shared__ double sm[3][kThreads / 32];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    double x = red[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sm[q][warp] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      double x = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) x += sm[q][w];
      partials[(size_t)blockIdx.x * 4 + q] = x;
    }
    __threadfence();
    const unsigned ticket = atomicInc(counter, gridDim.x - 1);
    last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    if (!((f.out_mask >> q) & 1)) continue;
    // fixed order: strided partial sums per thread, lanes, warps -- the same for every run
    double x = 0.0;
    for (int b = threadIdx.x; b < f.nmain; b += kThreads) x += f.main_partials[(size_t)b * 4 + q];
    for (unsigned b = threadIdx.x; b < gridDim.x; b += kThreads) x += partials[(size_t)b * 4 + q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sm[q][warp] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      if (!((f.out_mask >> q) & 1)) continue;
      double x = 0.0;
      for (int w = 0; w < kThreads / 32; ++w) x += sm[q][w];
      f.out[q] = x;
    }
  }
}

__global__ void __launch_bounds__(kThreads) k_bicg_x_half(int64_t n, const double* __restrict__ rho_p,
                                                          const double* __restrict__ hptr,
                                                          const double* __restrict__ dinv, double relax,
                                                          const double* __restrict__ p, double* __restrict__ x) {
  const double alpha = *rho_p / *hptr;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    x[i] += alpha * (relax * dinv[i] * p[i]);
}

// the closing sweep of a BiCGSTAB iteration with the Jacobi applications recomputed from their
// arguments (y1 = relax dinv p, y2 = relax dinv r: the same products k_bicg_p_prec / k_bicg_r_prec store)
__global__ void __launch_bounds__(kThreads) k_bicg_final_fold(int64_t n, Ranges own, const double* __restrict__ rho_p,
                                                              const double* __restrict__ hptr,
                                                              const double* __restrict__ trtt,
                                                              const double* __restrict__ dinv, double relax,
                                                              const double* __restrict__ p,
                                                              const double* r,
                                                              const double* __restrict__ xin, double* __restrict__ xout,
                                                              double* t, double* rout,
                                                              const double* __restrict__ rt, double* partials,
                                                              unsigned* counter, double* out, Link L) {
  double acc[2] = {0.0, 0.0};
  const bool single = own.n == 1 && own.b[0] == 0 && own.e[0] == n;
  const double alpha = *rho_p / *hptr, omega = trtt[0] / trtt[1];
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const double di = relax * dinv[i], rh = r[i];
    const double y1 = di * p[i], y2 = di * rh;
    xout[i] = (xin[i] + alpha * y1) + omega * y2;
    const double ri = rh - omega * t[i];
    rout[i] = ri;          // rout may be r itself
    if (L.zero_input) t[i] = 0.0;
    if (single || in_ranges(own, i)) {
      acc[0] += ri * ri;
      acc[1] += rt[i] * ri;
    }
  }
  grid_reduce<2>(acc, partials, counter, out, L);
}

inline void check_launch() { DCB_CUDA(cudaGetLastError()); }

}  // namespace

void reduce_workspace_create(ReduceWorkspace* w, int max_blocks) {
  w->max_blocks = max_blocks;
  DCB_CUDA(cudaMalloc(&w->partials, sizeof(double) * max_blocks * 4));
  DCB_CUDA(cudaMalloc(&w->counter, sizeof(unsigned)));
  DCB_CUDA(cudaMemset(w->counter, 0, sizeof(unsigned)));
}
void reduce_workspace_destroy(ReduceWorkspace* w) {
  if (w->partials) cudaFree(w->partials);
  if (w->counter) cudaFree(w->counter);
  w->partials = nullptr; w->counter = nullptr;
}

void spmv_csr(int64_t nrows, const int64_t* rowptr, const int32_t* rowptr32, const int32_t* colidx,
              const double* vals, const double* x, double* y, int avg_nnz, cudaStream_t s) {
  if (nrows == 0) return;
  if (rowptr32) spmv_launch<int32_t>(nrows, rowptr32, colidx, vals, x, y, avg_nnz, s);
  else spmv_launch<int64_t>(nrows, rowptr, colidx, vals, x, y, avg_nnz, s);
  check_launch();
}

static int64_t ranges_len(const Ranges& r) {
  int64_t n = 0;
  for (int k = 0; k < r.n; ++k) n = std::max<int64_t>(n, r.e[k] - r.b[k]);
  return n;
}

void dot(const Ranges& own, const double* a, const double* b, double* out, const ReduceWorkspace& w, cudaStream_t s,
         const Link& L) {
  k_dot<<<grid_for(ranges_len(own), 4), kThreads, 0, s>>>(own, a, b, w.partials, w.counter, out, L);
  check_launch();
}
void dot2(const Ranges& own, const double* a, const double* b, const double* c, const double* d, double* out,
          const ReduceWorkspace& w, cudaStream_t s, const Link& L) {
  k_dot2<<<grid_for(ranges_len(own), 4), kThreads, 0, s>>>(own, a, b, c, d, w.partials, w.counter, out, L);
  check_launch();
}
void bicg_update_p(int64_t n, double* p, const double* r, const double* v, double beta, double omega,
                   bool first, cudaStream_t s) {
  k_bicg_p<<<grid_for(n, 2), kThreads, 0, s>>>(n, p, r, v, beta, omega, first);
  check_launch();
}
void axpy_pair_norm(int64_t n, const Ranges& own, double alpha, const double* y, double* x, const double* v,
                    double* r, const double* rt, double* out, const ReduceWorkspace& w, cudaStream_t s) {
  k_axpy_pair_norm<<<grid_for(n, 2), kThreads, 0, s>>>(n, own, alpha, y, x, v, r, rt, w.partials, w.counter, out);
  check_launch();
}
void bicg_p_prec(int64_t n, double* p, const double* r, double* v, const double* rho_new, const double* rho,
                 const double* hptr, const double* trtt, bool first, const double* dinv, double relax, double* y,
                 const ReduceWorkspace& w, cudaStream_t s, const Link& L) {
  k_bicg_p_prec<<<grid_for(n, 2), kThreads, 0, s>>>(n, p, r, v, rho_new, rho, hptr, trtt, first, dinv, relax, y, w.counter, L);
  check_launch();
}
void bicg_r_prec(int64_t n, const Ranges& own, const double* rho, const double* hptr, const double* v, double* r,
                 const double* dinv, double relax, double* y2, double* out, const ReduceWorkspace& w, cudaStream_t s,
                 const Link& L) {
  k_bicg_r_prec<<<grid_for(n, 2), kThreads, 0, s>>>(n, own, rho, hptr, v, r, dinv, relax, y2, w.partials, w.counter, out, L);
  check_launch();
}
void bicg_final(int64_t n, const Ranges& own, const double* rho, const double* hptr, const double* trtt, const double* y1,
                const double* y2, const double* xin, double* xout, double* t, double* r, const double* rt,
                double* out, const ReduceWorkspace& w, cudaStream_t s, const Link& L) {
  k_bicg_final<<<grid_for(n, 2), kThreads, 0, s>>>(n, own, rho, hptr, trtt, y1, y2, xin, xout, t, r, rt, w.partials, w.counter, out, L);
  check_launch();
}
void mgs_step(int64_t n, const Ranges& own, const double* coef, const double* vprev, double* w, const double* vnext,
              double* out, const ReduceWorkspace& ws, cudaStream_t s, const Link& L) {
  k_mgs_step<<<grid_for(n, 2), kThreads, 0, s>>>(n, own, coef, vprev, w, vnext, ws.partials, ws.counter, out, L);
  check_launch();
}
void xpby(int64_t n, double* p, const double* q, double beta, cudaStream_t s) {
  k_xpby<<<grid_for(n, 2), kThreads, 0, s>>>(n, p, q, beta);
  check_launch();
}
void axpy(int64_t n, double a, const double* x, double* y, cudaStream_t s) {
  k_axpy<<<grid_for(n, 2), kThreads, 0, s>>>(n, a, x, y);
  check_launch();
}
void copy(int64_t n, const double* x, double* y, cudaStream_t s) {
  if (n) DCB_CUDA(cudaMemcpyAsync(y, x, n * sizeof(double), cudaMemcpyDeviceToDevice, s));
}
void fill(int64_t n, double v, double* y, cudaStream_t s) {
  if (v == 0.0) { if (n) DCB_CUDA(cudaMemsetAsync(y, 0, n * sizeof(double), s)); return; }
  k_fill<<<grid_for(n, 2), kThreads, 0, s>>>(n, v, y);
  check_launch();
}
namespace {
__global__ void __launch_bounds__(128) k_sor_level(const int32_t* __restrict__ rows, int64_t count,
                                                   const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                   const double* __restrict__ vals, const double* __restrict__ d,
                                                   double* v, double relax, bool skip_diag) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= count) return;
  const int i = rows[t];
  double rhs = d[i], diag = 1.0;
  for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
    const int j = colidx[k];
    const double a = vals[k];
    if (j == i) {
      diag = a;
      if (skip_diag) continue;
    }
    rhs -= a * v[j];
  }
  if (skip_diag) v[i] = rhs / diag;   // dbgs assigns the unrelaxed value; the caller blends with the old iterate
  else v[i] += relax * (rhs / diag);
}
// Self-scheduled sweep: the whole forward (or backward) sweep in ONE launch.  Slot p of `slots` holds a row (or -1:
// padding, so that the 32 rows of a warp belong to one level and never wait for each other); slots are in level order.
// A row spins until every coupled row that the sequential sweep visits earlier carries this sweep's epoch, then
// computes exactly what k_sor_level computes and publishes itself.  Every wait is for a row in an earlier slot and
// blocks start in index order, so the sweep cannot deadlock; coupled rows are symmetrised on the host (dep_ptr /
// dep_idx = pattern of A + A^T), which also keeps a row from overwriting a value a smaller row still has to read.
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__global__ void __launch_bounds__(128) k_sor_sweep(const int32_t* __restrict__ slots, int64_t nslots, bool backward,
                                                   const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                   const double* __restrict__ vals, const int64_t* __restrict__ dep_ptr,
                                                   const int32_t* __restrict__ dep_idx, const double* __restrict__ d,
                                                   double* v, double relax, bool skip_diag, int* done, int epoch) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nslots) return;
  const int i = slots[backward ? nslots - 1 - t : t];
  if (i < 0) return;
  for (int64_t k = dep_ptr[i]; k < dep_ptr[i + 1]; ++k) {
    const int j = dep_idx[k];
    if (backward ? j > i : j < i)
      while (ld_acquire(done + j) != epoch) __nanosleep(40);   // back off: spinning warps would crowd the L2 the chain runs through
  }
  double rhs = d[i], diag = 1.0;
  for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
    const int j = colidx[k];
    const double a = vals[k];
    if (j == i) {
      diag = a;
      if (skip_diag) continue;
    }
    rhs -= a * __ldcg(v + j);   // L2: other SMs write v during the sweep
  }
  if (skip_diag) v[i] = rhs / diag;
  else v[i] = __ldcg(v + i) + relax * (rhs / diag);
  st_release(done + i, epoch);
}
__global__ void __launch_bounds__(kThreads) k_relax_blend(int64_t n, double w, const double* __restrict__ xold,
                                                          double* __restrict__ x) {
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    x[i] = w * x[i] + (1.0 - w) * xold[i];
}
}  // namespace
void relax_blend(int64_t n, double w, const double* xold, double* x, cudaStream_t s) {
  k_relax_blend<<<grid_for(n, 2), kThreads, 0, s>>>(n, w, xold, x);
  check_launch();
}
void sor_level(const int32_t* rows, int64_t count, const int64_t* rowptr, const int32_t* colidx, const double* vals,
               const double* d, double* v, double relax, bool skip_diag, cudaStream_t s) {
  if (count <= 0) return;
  k_sor_level<<<(unsigned)((count + 127) / 128), 128, 0, s>>>(rows, count, rowptr, colidx, vals, d, v, relax, skip_diag);
  check_launch();
}
void sor_sweep(const int32_t* slots, int64_t nslots, bool backward, const int64_t* rowptr, const int32_t* colidx,
               const double* vals, const int64_t* dep_ptr, const int32_t* dep_idx, const double* d, double* v, double relax,
               bool skip_diag, int* done, int epoch, cudaStream_t s) {
  if (nslots <= 0) return;
  k_sor_sweep<<<(unsigned)((nslots + 127) / 128), 128, 0, s>>>(slots, nslots, backward, rowptr, colidx, vals, dep_ptr, dep_idx, d,
                                                               v, relax, skip_diag, done, epoch);
  check_launch();
}
void jacobi_apply(int64_t n, const double* dinv, double relax, const double* d, double* v, cudaStream_t s) {
  k_jacobi<<<grid_for(n, 2), kThreads, 0, s>>>(n, dinv, relax, d, v);
  check_launch();
}
void csr_extract_diag_inv(int64_t n, const int64_t* rowptr, const int32_t* colidx, const double* vals,
                          double* dinv, cudaStream_t s) {
  k_diag_inv<<<grid_for(n), kThreads, 0, s>>>(n, rowptr, colidx, vals, dinv);
  check_launch();
}
void csr_extract_block_diag(int64_t dof0, int64_t nblocks, int bs, const int64_t* rowptr,
                            const int32_t* colidx, const double* vals, double* bdiag, cudaStream_t s) {
  if (nblocks == 0) return;
  k_block_diag<<<grid_for(nblocks * bs), kThreads, 0, s>>>(dof0, nblocks * bs, bs, rowptr, colidx, vals, bdiag);
  check_launch();
}
void block_invert(int64_t nblocks, int bs, double* blocks, cudaStream_t s) {
  if (nblocks == 0) return;
  if (bs > 19) fail("BlockJacobi: block size ", bs, " > 19 (DenseInverse limit, direct.hh:36-74)");
  k_block_invert<<<(unsigned)((nblocks + 127) / 128), 128, 0, s>>>(nblocks, bs, blocks);
  check_launch();
}
void block_jacobi_apply(int64_t dof0, int64_t nblocks, int bs, const double* binv, double relax,
                        const double* d, double* v, cudaStream_t s) {
  if (nblocks == 0) return;
  k_block_jacobi<<<grid_for(nblocks * bs), kThreads, 0, s>>>(dof0, nblocks * bs, bs, binv, relax, d, v);
  check_launch();
}
void block_diag_to_dinv(int64_t dof0, int64_t nblocks, int bs, const double* bdiag, double* dinv, cudaStream_t s) {
  if (nblocks == 0) return;
  k_bdiag_dinv<<<grid_for(nblocks * bs), kThreads, 0, s>>>(dof0, nblocks * bs, bs, bdiag, dinv);
  check_launch();
}
void axpy_dev(int64_t n, const double* coef, double sign, const double* x, double* y, cudaStream_t s) {
  k_axpy_dev<<<grid_for(n, 2), kThreads, 0, s>>>(n, coef, sign, x, y);
  check_launch();
}
void normalize_dev(int64_t n, const double* src, const double* norm2, double* dst, cudaStream_t s) {
  k_normalize_dev<<<grid_for(n, 2), kThreads, 0, s>>>(n, src, norm2, dst);
  check_launch();
}
void scale(int64_t n, double a, double* x, cudaStream_t s) {
  k_scale<<<grid_for(n, 2), kThreads, 0, s>>>(n, a, x);
  check_launch();
}
void sub(int64_t n, const double* a, const double* b, double* y, cudaStream_t s) {
  k_sub<<<grid_for(n, 2), kThreads, 0, s>>>(n, a, b, y);
  check_launch();
}
void invert_diag(int64_t n, double* d, const unsigned char* mask, cudaStream_t s) {
  k_invert_diag<<<grid_for(n, 2), kThreads, 0, s>>>(n, d, mask);
  check_launch();
}
void set_values(int64_t n, const int32_t* idx, const double* vals, double* x, cudaStream_t s) {
  if (n == 0) return;
  k_set_values<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, idx, vals, x);
  check_launch();
}
void zero_values(int64_t n, const int32_t* idx, double* x, cudaStream_t s) { set_values(n, idx, nullptr, x, s); }
void copy_values(int64_t n, const int32_t* idx, const double* src, double* dst, cudaStream_t s) {
  if (n == 0) return;
  k_copy_values<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, idx, src, dst);
  check_launch();
}
void csr_constrain(int64_t nrows, const int64_t* rowptr, const int32_t* colidx, double* vals,
                   const unsigned char* mask, cudaStream_t s) {
  k_csr_constrain<<<grid_for(nrows), kThreads, 0, s>>>(nrows, rowptr, colidx, vals, mask);
  check_launch();
}
void bdiag_constrain(int64_t dof0, int64_t nblocks, int bs, double* bdiag, const unsigned char* mask, cudaStream_t s) {
  if (nblocks == 0) return;
  k_bdiag_constrain<<<grid_for(nblocks * bs), kThreads, 0, s>>>(dof0, nblocks * bs, bs, bdiag, mask);
  check_launch();
}
void tile_fixup(const TileFixup& f, const ReduceWorkspace& w, cudaStream_t s) {
  const long long P0 = f.n[0] + 1, P1 = f.n[1] + 1, P2 = f.n[2] + 1;
  const long long ncx = f.n[0] > 0 ? (f.n[0] - 1) / f.tile[0] : 0, ncy = f.n[1] > 0 ? (f.n[1] - 1) / f.tile[1] : 0,
                  ncz = f.n[2] > 0 ? (f.n[2] - 1) / f.tile[2] : 0;
  const long long total = ncx * P1 * P2 + ncy * P0 * P2 + ncz * P0 * P1;
  // latency bound gathers: one vertex per thread as long as the partials buffer allows
  const long long blocks = std::max<long long>(1, std::min<long long>((total + kThreads - 1) / kThreads, w.max_blocks));
  k_tile_fixup<<<(unsigned)blocks, kThreads, 0, s>>>(f, w.partials, w.counter);
  check_launch();
}
void bicg_x_half(int64_t n, const double* rho, const double* hptr, const double* dinv, double relax, const double* p,
                 double* x, cudaStream_t s) {
  k_bicg_x_half<<<grid_for(n, 2), kThreads, 0, s>>>(n, rho, hptr, dinv, relax, p, x);
  check_launch();
}
void bicg_final_fold(int64_t n, const Ranges& own, const double* rho, const double* hptr, const double* trtt,
                     const double* dinv, double relax, const double* p, const double* r, const double* xin, double* xout,
                     double* t, double* rout, const double* rt, double* out, const ReduceWorkspace& w, cudaStream_t s,
                     const Link& L) {
  k_bicg_final_fold<<<grid_for(n, 2), kThreads, 0, s>>>(n, own, rho, hptr, trtt, dinv, relax, p, r, xin, xout, t, rout, rt,
                                                        w.partials, w.counter, out, L);
  check_launch();
}
void gather(int64_t n, const int32_t* idx, const double* x, double* buf, cudaStream_t s) {
  if (n == 0) return;
  k_gather<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, idx, x, buf);
  check_launch();
}
void scatter(int64_t n, const int32_t* idx, const double* buf, double* x, cudaStream_t s) {
  if (n == 0) return;
  k_scatter<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, idx, buf, x);
  check_launch();
}

}  // namespace la
}  // namespace dcb
