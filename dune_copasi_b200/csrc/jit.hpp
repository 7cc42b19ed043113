// NVRTC lowering of one model: generated model source + assembly kernel templates -> sm_100a cubin,
// loaded through the CUDA runtime's library API.  This replaces the reference's run-time expression
// interpreters (src/dune/copasi/parser/{mu,exprtk,symengine}.cc) on the hot path: expressions are
// compiled once per model and fused into the element kernels.
#pragma once
#include <cuda_runtime.h>
#include <unistd.h>

#include <map>
#include <string>
#include <vector>

#include "model.hpp"

namespace dcb {

// Kernels are compiled in groups so that a run only pays for what it launches: the patch kernels
// (hot path) eagerly, the element-per-thread kernels, the CSR fill and the facet kernels on first use.
enum class JitGroup { All, Patch, Element, Csr, Skeleton, Structured, StructuredQ1, Tile, TileQ1 };
// full translation unit (defines + model source + kernel_args.h + assembly.cuh + entry points)
std::string jit_source(const Model& model, const std::string& defines = "", JitGroup group = JitGroup::All);
// #defines derived from model.assembly.b200.* (patch geometry)
std::string jit_defines(const Model& model);
// compile with an on-disk cache next to the library (or $DCB_JIT_CACHE), keyed by source + NVRTC version
std::vector<char> jit_compile_cached(const std::string& source);
// drop the cache entry of a source (a cached blob the driver refuses to load is recompiled once)
void jit_cache_evict(const std::string& source);
// compile for sm_100a; works without a GPU.  `log` receives the NVRTC log.
std::vector<char> jit_compile(const std::string& source, std::string* log, bool ptx = false);

class JitModule {
 public:
  JitModule() = default;
  ~JitModule();
  JitModule(const JitModule&) = delete;
  JitModule& operator=(const JitModule&) = delete;
  void load(const std::vector<char>& cubin);
  cudaKernel_t kernel(const std::string& name);
  bool loaded() const { return lib_ != nullptr; }

 private:
  cudaLibrary_t lib_ = nullptr;
  std::map<std::string, cudaKernel_t> cache_;
};

template <class Args>
inline void jit_launch(cudaKernel_t k, unsigned grid, unsigned block, size_t smem, cudaStream_t s, Args& a) {
  void* params[] = {(void*)&a};
  DCB_CUDA(cudaLaunchKernel((const void*)k, dim3(grid), dim3(block), params, smem, s));
}

}  // namespace dcb
