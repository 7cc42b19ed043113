// Multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.  The reference has no
// distributed path (solver/istl/factory/inverse.hh:34-35 throws "Parallel solvers have not been
// implemented!"), so this is new design: a vertex partition with one layer of ghost elements,
// owner -> ghost halo updates before every operator application and fp64 all-reduces of the
// Krylov scalars.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "kernels/peer.hpp"
#include "util.hpp"

namespace dcb {

struct HaloPlan {
  // per peer rank: local dof indices to send (owned here, ghost there) and to receive
  std::vector<int> peers;
  std::vector<std::vector<int32_t>> send_idx, recv_idx;
};

struct Communicator {
  int rank = 0, size = 1;
  virtual ~Communicator() = default;
  virtual void allreduce_sum(double* dev, int n, cudaStream_t s) = 0;   // in place
  virtual void halo_update(double* x, cudaStream_t s) = 0;              // owner -> ghost copies
  // Collectives fused into the Krylov sweeps (kernels/linalg.cu): when *_links_ready(), the solver asks for the
  // next all-reduce / halo exchange as a *link* and hands it to the kernel that produces the data; that kernel
  // then does the exchange itself (last block / streaming loop).  Every rank must ask in the same order.
  virtual bool reduce_links_ready() const { return false; }   // all-reduces over the peer mailboxes
  virtual bool push_links_ready() const { return false; }     // halo of contiguous planes (slab partitions)
  virtual void link_reduce(peer::Link* l) { (void)l; }   // l->reduce = 1, next all-reduce sequence number
  virtual void link_push(peer::Link* l) { (void)l; }     // l->push = 1, next halo sequence number
  // ghost entries of x from the planes the neighbours pushed in the exchange of the last link_push
  virtual void halo_pull(double* x, cudaStream_t s) { halo_update(x, s); }
  // a bounded flag spin gave up (a peer died or diverged): checked by the solver after every solve
  virtual bool peer_error() { return false; }
  long long launches = 0;
  // small all-reduces (and, on slab partitions, halo updates) run over NVLink peer memory
  // (kernels/peer.cu) instead of NCCL; DCB_PEER_COLLECTIVES=0 turns that off
  bool peer_active = false;
};

// NCCL-backed communicator; libnccl is resolved at run time (dlopen) so the library loads on
// hosts without NCCL/driver.  `unique_id` is the 128-byte ncclUniqueId created by rank 0
// (nccl_unique_id) and distributed by the caller (e.g. torch.distributed broadcast).
void nccl_unique_id(char out[128]);
Communicator* nccl_communicator_create(const char unique_id[128], int rank, int size, const HaloPlan& plan);

}  // namespace dcb
