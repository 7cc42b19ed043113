// Collectives over NVLink peer memory (sm_100a), written for the two exchanges that sit on the
// critical path of every Krylov half step on a slab-partitioned lattice:
//   * all-reduce of 1..8 doubles (dot products, norms) across all ranks,
//   * halo update of contiguous vertex planes with the two slab neighbours.
// NCCL needs ~20-30 us for each of them at these sizes; a half step of the 256^3 problem on 8 GPUs
// is ~0.25 ms of kernels, so three NCCL calls were a third of the step.  Here every rank owns a
// "mailbox" in device memory that its peers have mapped through CUDA IPC: a rank *stores* its
// contribution straight into the peers' mailboxes (P2P writes over NVSwitch), publishes it with a
// sequence flag after a system-scope fence, and spins on its own flags for the peers' data.
// Two parities of every slot make the protocol safe without a barrier: a rank can be at most one
// exchange ahead of a peer, because finishing exchange s needs the peer's flag s, which the peer
// writes only after it has finished reading exchange s-1.
// Sums are formed in rank order on every rank: deterministic and identical everywhere.
#include "peer.hpp"
#include "peer_device.cuh"

namespace dcb {
namespace peer {

namespace {

// one block, one thread per peer
__global__ void __launch_bounds__(64) k_allreduce(Mailboxes m, double* data, int n, unsigned long long seq, int* error) {
  __shared__ double mine[kMaxWords];
  if (threadIdx.x < n) mine[threadIdx.x] = data[threadIdx.x];
  block_allreduce(m, mine, n, seq, error);
  if (threadIdx.x < n) data[threadIdx.x] = mine[threadIdx.x];
}

// grid of G blocks: push the send ranges into the neighbours' mailboxes, publish, wait, pull
__global__ void __launch_bounds__(256) k_halo(Mailboxes m, HaloArgs h, double* x, unsigned long long seq) {
  const int par = (int)(seq & 1ull);
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  for (int k = 0; k < h.npeers; ++k) {
    // slot of the receiver that is reserved for data arriving from this side
    double* dst = (double*)(m.box[h.peer[k]] + halo_data_offset(m.size, m.cap, h.remote_slot[k], par));
    const double* src = x + h.send_off[k];
    for (long long i = tid; i < h.send_n[k]; i += nth) dst[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicAdd(h.counter, 1u);
    last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    *h.counter = 0;
    __threadfence_system();
    for (int k = 0; k < h.npeers; ++k)
      st_flag((unsigned long long*)(m.box[h.peer[k]] + halo_flag_offset(m.size, h.remote_slot[k], par)), seq);
  }
  for (int k = 0; k < h.npeers; ++k) {
    if (threadIdx.x == 0)
      wait_flag((const unsigned long long*)(m.box[m.rank] + halo_flag_offset(m.size, h.local_slot[k], par)), seq, h.error);
    __syncthreads();
    __threadfence_system();
    const double* src = (const double*)(m.box[m.rank] + halo_data_offset(m.size, m.cap, h.local_slot[k], par));
    double* dst = x + h.recv_off[k];
    for (long long i = tid; i < h.recv_n[k]; i += nth) dst[i] = ld_data(src + i);
  }
}

// second half of k_halo on its own: the planes of exchange `seq` were pushed by the neighbours' sweeps (push_entry)
__global__ void __launch_bounds__(256) k_halo_pull(Mailboxes m, HaloArgs h, double* x, unsigned long long seq) {
  const int par = (int)(seq & 1ull);
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  for (int k = 0; k < h.npeers; ++k) {
    if (threadIdx.x == 0)
      wait_flag((const unsigned long long*)(m.box[m.rank] + halo_flag_offset(m.size, h.local_slot[k], par)), seq, h.error);
    __syncthreads();
    __threadfence_system();
    const double* src = (const double*)(m.box[m.rank] + halo_data_offset(m.size, m.cap, h.local_slot[k], par));
    double* dst = x + h.recv_off[k];
    for (long long i = tid; i < h.recv_n[k]; i += nth) dst[i] = ld_data(src + i);
  }
}

// general halo plan in one launch: gather the send lists straight into the neighbours' slots, publish, wait for the
// neighbours' flags, scatter the slots into the ghost entries (replaces gather kernels + ncclSend/Recv + scatter
// kernels: 15 launches per exchange with 7 neighbours)
__global__ void __launch_bounds__(256) k_halo_general(Mailboxes m, GeneralHaloArgs h, double* x, unsigned long long seq) {
  const int par = (int)(seq & 1ull);
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  for (int k = 0; k < h.npeers; ++k) {
    double* dst = (double*)(m.box[h.peer[k]] + halo_data_offset(m.size, m.cap, m.rank, par));
    const long long b = h.send_ptr[k], n = h.send_ptr[k + 1] - b;
    for (long long i = tid; i < n; i += nth) dst[i] = x[h.send_idx[b + i]];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicAdd(h.counter, 1u);
    last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    *h.counter = 0;
    __threadfence_system();
    for (int k = 0; k < h.npeers; ++k)
      st_flag((unsigned long long*)(m.box[h.peer[k]] + halo_flag_offset(m.size, m.rank, par)), seq);
  }
  for (int k = 0; k < h.npeers; ++k) {
    if (threadIdx.x == 0)
      wait_flag((const unsigned long long*)(m.box[m.rank] + halo_flag_offset(m.size, h.peer[k], par)), seq, h.error);
    __syncthreads();
    __threadfence_system();
    const double* src = (const double*)(m.box[m.rank] + halo_data_offset(m.size, m.cap, h.peer[k], par));
    const long long b = h.recv_ptr[k], n = h.recv_ptr[k + 1] - b;
    for (long long i = tid; i < n; i += nth) x[h.recv_idx[b + i]] = ld_data(src + i);
  }
}

}  // namespace

void halo_general(const Mailboxes& m, const GeneralHaloArgs& h, double* x, unsigned long long seq, cudaStream_t s) {
  k_halo_general<<<kHaloBlocks, 256, 0, s>>>(m, h, x, seq);
}

void allreduce(const Mailboxes& m, double* data, int n, unsigned long long seq, int* error, cudaStream_t s) {
  k_allreduce<<<1, 64, 0, s>>>(m, data, n, seq, error);
}
void halo_pull(const Mailboxes& m, const HaloArgs& h, double* x, unsigned long long seq, cudaStream_t s) {
  k_halo_pull<<<kHaloBlocks, 256, 0, s>>>(m, h, x, seq);
}

void halo(const Mailboxes& m, const HaloArgs& h, double* x, unsigned long long seq, cudaStream_t s) {
  k_halo<<<kHaloBlocks, 256, 0, s>>>(m, h, x, seq);
}

}  // namespace peer
}  // namespace dcb
