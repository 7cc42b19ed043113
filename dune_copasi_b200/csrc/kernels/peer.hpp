// Peer-memory collectives (peer.cu): mailbox layout shared by host and device, launchers.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace dcb {
namespace peer {

constexpr int kMaxRanks = 16;     // one NVSwitch domain
constexpr int kMaxWords = 8;      // doubles per all-reduce
constexpr int kHaloBlocks = 64;

struct Mailboxes {
  char* box[kMaxRanks];           // box[rank] is local, the others are IPC mappings of the peers'
  int rank, size;
  long long cap;                  // doubles per halo slot (the largest receive list of any rank)
};

// byte offsets inside a mailbox
__host__ __device__ inline size_t ar_val_offset(int size, int par, int src) {
  return ((size_t)(par * size + src) * kMaxWords) * sizeof(double);
}
__host__ __device__ inline size_t ar_flag_offset(int size, int par, int src) {
  return (size_t)2 * size * kMaxWords * sizeof(double) + (size_t)(par * size + src) * 8;
}
// Halo slots.  Slab partitions use two (slot 0: data from the lower neighbour, slot 1: from the higher one);
// general partitions (RCB over unstructured meshes: several neighbours, index lists) use one slot per source rank.
__host__ __device__ inline size_t halo_flag_offset(int size, int slot, int par) {
  return (size_t)2 * size * kMaxWords * sizeof(double) + (size_t)2 * size * 8 + (size_t)(slot * 2 + par) * 8;
}
__host__ __device__ inline size_t halo_data_offset(int size, long long cap, int slot, int par) {
  size_t head = (size_t)2 * size * kMaxWords * sizeof(double) + (size_t)2 * size * 8 + (size_t)kMaxRanks * 2 * 8;
  head = (head + 255) / 256 * 256;
  return head + (size_t)(slot * 2 + par) * (size_t)cap * sizeof(double);
}
inline size_t mailbox_bytes(int size, long long cap, int nslots = 2) { return halo_data_offset(size, cap, nslots, 0); }

struct HaloArgs {
  int npeers;                     // <= 2
  int peer[2];
  int local_slot[2], remote_slot[2];
  long long send_off[2], send_n[2], recv_off[2], recv_n[2];
  unsigned* counter;              // self-resetting ticket of the push phase
  // spins are bounded: a peer that never publishes (died, diverged) raises *error instead of hanging the device
  int* error;
};
constexpr long long kSpinLimit = 1ll << 29;   // polls of a flag before giving up: minutes of wall time (a rank may sit in NVRTC for a while)

// general partitions: peer k sends the entries send_idx[send_ptr[k] .. send_ptr[k+1]) of the vector (both sides
// list them in the same canonical order) and fills recv_idx[recv_ptr[k] .. recv_ptr[k+1]); slot = source rank
struct GeneralHaloArgs {
  int npeers;
  int peer[kMaxRanks];
  long long send_ptr[kMaxRanks + 1], recv_ptr[kMaxRanks + 1];
  const int* send_idx;
  const int* recv_idx;
  unsigned* counter;
  int* error;
};

void allreduce(const Mailboxes& m, double* data, int n, unsigned long long seq, int* error, cudaStream_t s);
// one launch per exchange whatever the number of neighbours: pack + push, publish, wait, pull + unpack
void halo_general(const Mailboxes& m, const GeneralHaloArgs& h, double* x, unsigned long long seq, cudaStream_t s);
void halo(const Mailboxes& m, const HaloArgs& h, double* x, unsigned long long seq, cudaStream_t s);

// ---- exchanges fused into the kernels of the Krylov sweeps (kernels/linalg.cu, peer_device.cuh) ------------
// Link of one kernel launch to the collectives that follow it in the algorithm.  The producing kernel does the
// communication itself instead of handing over to k_allreduce / k_halo:
//   * reduce: the last block of a reducing kernel all-reduces the sums it has just formed over the mailboxes
//     (exchange number ar_seq) before it writes them;
//   * push:   a sweep that writes a vector also stores the entries of the send ranges into the neighbours' halo
//     slots (exchange number halo_seq) while it streams; its last block publishes the flags.  The receiver
//     copies the planes into its ghost range with halo_pull right before the operator application.
// A link with reduce = push = 0 leaves the kernel local (single GPU, NCCL fallback).
struct Link {
  Mailboxes m;
  int reduce = 0, push = 0;
  unsigned long long ar_seq = 0, halo_seq = 0;
  HaloArgs h;
};
void halo_pull(const Mailboxes& m, const HaloArgs& h, double* x, unsigned long long seq, cudaStream_t s);

}  // namespace peer
}  // namespace dcb
