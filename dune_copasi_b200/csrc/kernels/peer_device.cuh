// Device side of the peer-memory collectives: used by the stand-alone kernels of peer.cu and, fused, by the last
// block / the streaming loops of the Krylov sweeps in linalg.cu.  Protocol: see peer.cu.
#pragma once
#include "peer.hpp"

namespace dcb {
namespace peer {

__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_data(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
// wait until *f == seq; gives up after kSpinLimit polls and raises *error (the host checks it after the solve)
__device__ __forceinline__ void wait_flag(const unsigned long long* f, unsigned long long seq, int* error) {
  for (long long spins = 0; ld_flag(f) != seq; ++spins)
    if (spins > kSpinLimit) {
      if (error) *error = 1;
      break;
    }
}

// All-reduce of vals[0..n) (shared memory, n <= kMaxWords) over all ranks; called by every thread of one block
// (>= max(size, n) threads).  Sums are formed in rank order on every rank: deterministic, identical everywhere.
__device__ __forceinline__ void block_allreduce(const Mailboxes& m, double* vals, int n, unsigned long long seq, int* error) {
  const int p = threadIdx.x, par = (int)(seq & 1ull);
  __syncthreads();
  if (p < m.size) {
    char* box = m.box[p];
    double* dst = (double*)(box + ar_val_offset(m.size, par, m.rank));
    for (int i = 0; i < n; ++i) dst[i] = vals[i];
    __threadfence_system();
    st_flag((unsigned long long*)(box + ar_flag_offset(m.size, par, m.rank)), seq);
    // wait for peer p's contribution in the local mailbox
    wait_flag((const unsigned long long*)(m.box[m.rank] + ar_flag_offset(m.size, par, p)), seq, error);
  }
  __threadfence_system();
  __syncthreads();
  if (p < n) {
    double sum = 0.0;
    for (int q = 0; q < m.size; ++q)
      sum += ld_data((const double*)(m.box[m.rank] + ar_val_offset(m.size, par, q)) + p);
    vals[p] = sum;
  }
  __syncthreads();
}

// entry i of a vector that is being written: if it lies in a send range, it also goes to that neighbour's slot
__device__ __forceinline__ void push_entry(const Mailboxes& m, const HaloArgs& h, int par, long long i, double v) {
#pragma unroll
  for (int k = 0; k < 2; ++k)
    if (k < h.npeers && i >= h.send_off[k] && i < h.send_off[k] + h.send_n[k])
      ((double*)(m.box[h.peer[k]] + halo_data_offset(m.size, m.cap, h.remote_slot[k], par)))[i - h.send_off[k]] = v;
}
// after every block's pushes have been fenced (system scope) and counted: one thread publishes the exchange
__device__ __forceinline__ void publish_halo(const Mailboxes& m, const HaloArgs& h, unsigned long long seq) {
  const int par = (int)(seq & 1ull);
  __threadfence_system();
  for (int k = 0; k < h.npeers; ++k)
    st_flag((unsigned long long*)(m.box[h.peer[k]] + halo_flag_offset(m.size, h.remote_slot[k], par)), seq);
}

}  // namespace peer
}  // namespace dcb
