#include "comm.hpp"

#include <dlfcn.h>

#include <cstring>

#include "kernels/linalg.hpp"

namespace dcb {

namespace {

// minimal NCCL surface, resolved with dlopen so that the library itself has no link-time
// dependency on libnccl (it must load on the CPU-only build box)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct Nccl {
  void* h = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

Nccl& nccl() {
  static Nccl n;
  if (n.h) return n;
  // if torch is loaded its bundled libnccl.so.2 is already mapped and is the one we get
  n.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!n.h) n.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!n.h) fail("cannot load libnccl.so.2: ", dlerror());
  auto sym = [&](const char* name) {
    void* p = dlsym(n.h, name);
    if (!p) fail("libnccl: missing symbol ", name);
    return p;
  };
  n.GetUniqueId = (decltype(n.GetUniqueId))sym("ncclGetUniqueId");
  n.CommInitRank = (decltype(n.CommInitRank))sym("ncclCommInitRank");
  n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
  n.AllReduce = (decltype(n.AllReduce))sym("ncclAllReduce");
  n.Send = (decltype(n.Send))sym("ncclSend");
  n.Recv = (decltype(n.Recv))sym("ncclRecv");
  n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
  n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
  n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
  return n;
}

#define DCB_NCCL(call)                                                                 \
  do {                                                                                 \
    int r_ = (call);                                                                   \
    if (r_ != ncclSuccess) fail("NCCL error: ", nccl().GetErrorString(r_), " in " #call); \
  } while (0)

struct NcclCommunicator : Communicator {
  ncclComm_t comm = nullptr;
  HaloPlan plan;
  std::vector<DeviceBuffer<int32_t>> send_idx, recv_idx;
  std::vector<DeviceBuffer<double>> send_buf, recv_buf;

  ~NcclCommunicator() override {
    if (comm) nccl().CommDestroy(comm);
  }
  void allreduce_sum(double* dev, int n, cudaStream_t s) override {
    DCB_NCCL(nccl().AllReduce(dev, dev, (size_t)n, ncclFloat64, ncclSum, comm, s));
    launches++;
  }
  // contiguous index lists (slab partitions of structured grids: whole vertex planes) are sent
  // and received in place, without pack / unpack kernels
  std::vector<long long> send_off, recv_off;   // >= 0: contiguous range starting there

  void halo_update(double* x, cudaStream_t s) override {
    const size_t np = plan.peers.size();
    if (np == 0) return;
    for (size_t k = 0; k < np; ++k)
      if (send_off[k] < 0 && send_idx[k].n) {
        la::gather((int64_t)send_idx[k].n, send_idx[k].p, x, send_buf[k].p, s);
        launches++;
      }
    DCB_NCCL(nccl().GroupStart());
    for (size_t k = 0; k < np; ++k) {
      const size_t ns = plan.send_idx[k].size(), nr = plan.recv_idx[k].size();
      if (ns) DCB_NCCL(nccl().Send(send_off[k] >= 0 ? x + send_off[k] : send_buf[k].p, ns, ncclFloat64, plan.peers[k], comm, s));
      if (nr) DCB_NCCL(nccl().Recv(recv_off[k] >= 0 ? x + recv_off[k] : recv_buf[k].p, nr, ncclFloat64, plan.peers[k], comm, s));
    }
    DCB_NCCL(nccl().GroupEnd());
    launches++;
    for (size_t k = 0; k < np; ++k)
      if (recv_off[k] < 0 && recv_idx[k].n) {
        la::scatter((int64_t)recv_idx[k].n, recv_idx[k].p, recv_buf[k].p, x, s);
        launches++;
      }
  }
};

}  // namespace

void nccl_unique_id(char out[128]) {
  ncclUniqueId id;
  DCB_NCCL(nccl().GetUniqueId(&id));
  std::memcpy(out, id.internal, 128);
}

Communicator* nccl_communicator_create(const char unique_id[128], int rank, int size, const HaloPlan& plan) {
  require_device();
  auto* c = new NcclCommunicator();
  c->rank = rank;
  c->size = size;
  c->plan = plan;
  ncclUniqueId id;
  std::memcpy(id.internal, unique_id, 128);
  DCB_NCCL(nccl().CommInitRank(&c->comm, size, id, rank));
  const size_t np = plan.peers.size();
  c->send_idx.resize(np); c->recv_idx.resize(np); c->send_buf.resize(np); c->recv_buf.resize(np);
  auto contiguous = [](const std::vector<int32_t>& idx) -> long long {
    for (size_t i = 1; i < idx.size(); ++i)
      if (idx[i] != idx[0] + (int32_t)i) return -1;
    return idx.empty() ? -1 : idx[0];
  };
  c->send_off.resize(np); c->recv_off.resize(np);
  for (size_t k = 0; k < np; ++k) {
    c->send_off[k] = contiguous(plan.send_idx[k]);
    c->recv_off[k] = contiguous(plan.recv_idx[k]);
    if (c->send_off[k] < 0) {
      c->send_idx[k].upload(plan.send_idx[k]);
      c->send_buf[k].alloc(plan.send_idx[k].size());
    }
    if (c->recv_off[k] < 0) {
      c->recv_idx[k].upload(plan.recv_idx[k]);
      c->recv_buf[k].alloc(plan.recv_idx[k].size());
    }
  }
  DCB_CUDA(cudaDeviceSynchronize());
  return c;
}

}  // namespace dcb
