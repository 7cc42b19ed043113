#include "comm.hpp"

#include <dlfcn.h>

#include <algorithm>
#include <cstring>

#include <cstdlib>

#include "kernels/linalg.hpp"
#include "kernels/peer.hpp"

namespace dcb {

namespace {

// minimal NCCL surface, resolved with dlopen so that the library itself has no link-time
// dependency on libnccl (it must load on the CPU-only build box)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMin = 3 };
enum { ncclInt8 = 0 };

struct Nccl {
  void* h = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
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
  n.AllGather = (decltype(n.AllGather))sym("ncclAllGather");
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

  // ---- peer-memory path (kernels/peer.cu); NCCL stays the path for everything it does not cover
  peer::Mailboxes boxes{};
  peer::HaloArgs hargs{};
  DeviceBuffer<unsigned> ticket;
  bool peer_halo = false;
  unsigned long long ar_seq = 0, halo_seq = 0;

  ~NcclCommunicator() override {
    if (peer_active) {
      cudaDeviceSynchronize();
      for (int r = 0; r < size; ++r)
        if (r != rank && boxes.box[r]) cudaIpcCloseMemHandle(boxes.box[r]);
      if (boxes.box[rank]) cudaFree(boxes.box[rank]);
    }
    if (comm) nccl().CommDestroy(comm);
  }
  DeviceBuffer<int> error_flag;   // raised by a bounded spin that gave up

  bool reduce_links_ready() const override { return peer_active; }
  bool push_links_ready() const override { return peer_active && peer_halo; }
  void link_reduce(peer::Link* l) override {
    l->m = boxes; l->h = hargs; l->reduce = 1; l->ar_seq = ++ar_seq;
  }
  void link_push(peer::Link* l) override {
    l->m = boxes; l->h = hargs; l->push = 1; l->halo_seq = ++halo_seq;
  }
  void halo_pull(double* x, cudaStream_t s) override {
    peer::halo_pull(boxes, hargs, x, halo_seq, s);
    DCB_CUDA(cudaGetLastError());
    launches++;
  }
  bool peer_error() override {
    if (!peer_active) return false;
    int e = 0;
    DCB_CUDA(cudaMemcpy(&e, error_flag.p, sizeof e, cudaMemcpyDeviceToHost));
    return e != 0;
  }
  void allreduce_sum(double* dev, int n, cudaStream_t s) override {
    if (peer_active && n <= peer::kMaxWords) {
      peer::allreduce(boxes, dev, n, ++ar_seq, error_flag.p, s);
      DCB_CUDA(cudaGetLastError());
    } else {
      DCB_NCCL(nccl().AllReduce(dev, dev, (size_t)n, ncclFloat64, ncclSum, comm, s));
    }
    launches++;
  }
  // Map every rank's mailbox (CUDA IPC).  Collective; all ranks end up with the same answer.
  void setup_peer_memory() {
    const char* env = std::getenv("DCB_PEER_COLLECTIVES");
    bool want = !(env && env[0] == '0') && size > 1 && size <= peer::kMaxRanks;
    // slab halo: at most one lower and one higher neighbour, both with contiguous ranges; anything else goes
    // through the general (index list) exchange with one slot per source rank
    bool halo_ok = plan.peers.size() <= 2;
    long long cap = 1;
    for (size_t k = 0; k < plan.peers.size(); ++k) {
      halo_ok = halo_ok && send_off[k] >= 0 && recv_off[k] >= 0;
      cap = std::max<long long>(cap, (long long)plan.recv_idx[k].size());
    }
    if (plan.peers.size() == 2) halo_ok = halo_ok && ((plan.peers[0] < rank) != (plan.peers[1] < rank));
    const char* genv = std::getenv("DCB_PEER_GENERAL_HALO");
    const bool want_general = !(genv && genv[0] == '0');
    // agree on: everybody wants it, everybody's halo fits the slots, the largest slot
    DeviceBuffer<double> d(3);
    double h[3] = {want ? 1.0 : 0.0, halo_ok ? 1.0 : 0.0, -(double)cap};
    DCB_CUDA(cudaMemcpy(d.p, h, sizeof h, cudaMemcpyHostToDevice));
    DCB_NCCL(nccl().AllReduce(d.p, d.p, 3, ncclFloat64, ncclMin, comm, 0));
    DCB_CUDA(cudaMemcpy(h, d.p, sizeof h, cudaMemcpyDeviceToHost));
    if (h[0] < 0.5) return;
    const bool halo_all = h[1] > 0.5;
    const bool general = !halo_all && want_general;   // the same on every rank: h[1] is the global minimum
    cap = (halo_all || general) ? (long long)(-h[2]) : 1;
    // own mailbox, zeroed, exported
    const size_t bytes = peer::mailbox_bytes(size, cap, general ? size : 2);
    char* mine = nullptr;
    cudaIpcMemHandle_t handle;
    std::memset(&handle, 0, sizeof handle);
    bool ok = cudaMalloc(&mine, bytes) == cudaSuccess && cudaMemset(mine, 0, bytes) == cudaSuccess &&
              cudaDeviceSynchronize() == cudaSuccess && cudaIpcGetMemHandle(&handle, mine) == cudaSuccess;
    DeviceBuffer<char> hsend(sizeof handle), hall(sizeof handle * (size_t)size);
    DCB_CUDA(cudaMemcpy(hsend.p, &handle, sizeof handle, cudaMemcpyHostToDevice));
    DCB_NCCL(nccl().AllGather(hsend.p, hall.p, sizeof handle, ncclInt8, comm, 0));
    std::vector<cudaIpcMemHandle_t> handles(size);
    DCB_CUDA(cudaMemcpy(handles.data(), hall.p, sizeof handle * (size_t)size, cudaMemcpyDeviceToHost));
    for (int r = 0; r < peer::kMaxRanks; ++r) boxes.box[r] = nullptr;
    for (int r = 0; r < size && ok; ++r) {
      if (r == rank) { boxes.box[r] = mine; continue; }
      void* ptr = nullptr;
      ok = cudaIpcOpenMemHandle(&ptr, handles[r], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
      boxes.box[r] = (char*)ptr;
    }
    cudaGetLastError();   // a failed mapping is not an error of the run: NCCL carries on
    h[0] = ok ? 1.0 : 0.0;
    DCB_CUDA(cudaMemcpy(d.p, h, sizeof(double), cudaMemcpyHostToDevice));
    DCB_NCCL(nccl().AllReduce(d.p, d.p, 1, ncclFloat64, ncclMin, comm, 0));
    DCB_CUDA(cudaMemcpy(h, d.p, sizeof(double), cudaMemcpyDeviceToHost));
    if (h[0] < 0.5) {
      for (int r = 0; r < size; ++r)
        if (r != rank && boxes.box[r]) cudaIpcCloseMemHandle(boxes.box[r]);
      if (mine) cudaFree(mine);
      for (int r = 0; r < peer::kMaxRanks; ++r) boxes.box[r] = nullptr;
      return;
    }
    boxes.rank = rank; boxes.size = size; boxes.cap = cap;
    peer_active = true;
    error_flag.alloc(1);
    error_flag.zero();
    hargs.error = error_flag.p;
    peer_halo = halo_all;
    if (peer_halo) {
      ticket.alloc(1);
      ticket.zero();
      hargs.npeers = (int)plan.peers.size();
      hargs.counter = ticket.p;
      for (size_t k = 0; k < plan.peers.size(); ++k) {
        hargs.peer[k] = plan.peers[k];
        hargs.local_slot[k] = plan.peers[k] < rank ? 0 : 1;    // where that peer's data lands here
        hargs.remote_slot[k] = rank < plan.peers[k] ? 0 : 1;   // where ours lands there
        hargs.send_off[k] = send_off[k]; hargs.send_n[k] = (long long)plan.send_idx[k].size();
        hargs.recv_off[k] = recv_off[k]; hargs.recv_n[k] = (long long)plan.recv_idx[k].size();
      }
    }
    if (general) {
      peer_general = true;
      ticket.alloc(1);
      ticket.zero();
      gargs.npeers = (int)plan.peers.size();
      gargs.counter = ticket.p;
      gargs.error = error_flag.p;
      std::vector<int32_t> sidx, ridx;
      gargs.send_ptr[0] = gargs.recv_ptr[0] = 0;
      for (size_t k = 0; k < plan.peers.size(); ++k) {
        gargs.peer[k] = plan.peers[k];
        sidx.insert(sidx.end(), plan.send_idx[k].begin(), plan.send_idx[k].end());
        ridx.insert(ridx.end(), plan.recv_idx[k].begin(), plan.recv_idx[k].end());
        gargs.send_ptr[k + 1] = (long long)sidx.size();
        gargs.recv_ptr[k + 1] = (long long)ridx.size();
      }
      if (!sidx.empty()) gsend.upload(sidx);
      if (!ridx.empty()) grecv.upload(ridx);
      gargs.send_idx = gsend.p;
      gargs.recv_idx = grecv.p;
    }
    DCB_CUDA(cudaDeviceSynchronize());
  }
  peer::GeneralHaloArgs gargs{};
  DeviceBuffer<int32_t> gsend, grecv;
  bool peer_general = false;
  // contiguous index lists (slab partitions of structured grids: whole vertex planes) are sent
  // and received in place, without pack / unpack kernels
  std::vector<long long> send_off, recv_off;   // >= 0: contiguous range starting there

  void halo_update(double* x, cudaStream_t s) override {
    const size_t np = plan.peers.size();
    if (peer_halo) {   // every rank takes this branch or none does (setup_peer_memory agreed on it)
      peer::halo(boxes, hargs, x, ++halo_seq, s);
      DCB_CUDA(cudaGetLastError());
      launches++;
      return;
    }
    if (peer_general) {
      peer::halo_general(boxes, gargs, x, ++halo_seq, s);
      DCB_CUDA(cudaGetLastError());
      launches++;
      return;
    }
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
  c->setup_peer_memory();
  return c;
}

}  // namespace dcb
