// Common helpers: error transport across the C ABI, CUDA checks, device buffers.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace dcb {

// Exceptions never cross the C ABI: capi.cpp catches and stores the message (dc_last_error()).
// Mirrors the reference's convention of configuration errors as exceptions caught in main
// (src/dune_copasi.cc:446-461) and operator failures as error conditions.
struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

template <class... A>
[[noreturn]] inline void fail(A&&... a) {
  std::ostringstream os;
  (os << ... << a);
  throw Error(os.str());
}

#define DCB_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      ::dcb::fail("CUDA error ", cudaGetErrorName(e_), " (", cudaGetErrorString(e_), ") at ", \
                  __FILE__, ":", __LINE__, " in " #call);                                \
  } while (0)

// The product path has no CPU fallback: anything that computes needs a device.
inline void require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    fail("no CUDA device available: dune_copasi_b200 has no CPU fallback (", cudaGetErrorString(e), ")");
}

template <class T>
struct DeviceBuffer {
  T* p = nullptr;
  size_t n = 0;
  DeviceBuffer() = default;
  explicit DeviceBuffer(size_t n_) { alloc(n_); }
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  DeviceBuffer(DeviceBuffer&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DeviceBuffer& operator=(DeviceBuffer&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DeviceBuffer() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr; n = 0;
  }
  void alloc(size_t n_) {
    release();
    n = n_;
    if (n) DCB_CUDA(cudaMalloc(&p, n * sizeof(T)));
  }
  void upload(const T* h, size_t n_, cudaStream_t s = 0) {
    if (n_ != n) alloc(n_);
    if (n) DCB_CUDA(cudaMemcpyAsync(p, h, n * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const std::vector<T>& h, cudaStream_t s = 0) { upload(h.data(), h.size(), s); }
  void zero(cudaStream_t s = 0) {
    if (n) DCB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
  }
  void download(T* h, cudaStream_t s = 0) const {
    if (n) DCB_CUDA(cudaMemcpyAsync(h, p, n * sizeof(T), cudaMemcpyDeviceToHost, s));
    DCB_CUDA(cudaStreamSynchronize(s));
  }
};

template <class T>
struct PinnedBuffer {
  T* p = nullptr;
  size_t n = 0;
  PinnedBuffer() = default;
  PinnedBuffer(const PinnedBuffer&) = delete;
  PinnedBuffer& operator=(const PinnedBuffer&) = delete;
  ~PinnedBuffer() { if (p) cudaFreeHost(p); }
  void alloc(size_t n_) {
    if (p) cudaFreeHost(p);
    p = nullptr; n = n_;
    if (n) DCB_CUDA(cudaMallocHost(&p, n * sizeof(T)));
  }
};

}  // namespace dcb
