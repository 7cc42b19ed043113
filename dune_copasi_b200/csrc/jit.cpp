#include "jit.hpp"

#include <dlfcn.h>
#include <nvrtc.h>
#include <fcntl.h>
#include <sys/stat.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace dcb {

// generated at build time from kernels/kernel_args.h and kernels/assembly.cuh (embedded_sources.cpp)
extern const char* kKernelArgsSource;
extern const char* kAssemblySource;
extern const char* kStructuredSource;
extern const char* kQ1Source;
extern const char* kTileSource;

std::string jit_source(const Model& model, const std::string& defines, JitGroup group) {
  std::ostringstream o;
  o << defines << model.cuda_source();
  o << kKernelArgsSource << "\n" << kAssemblySource << "\n";
  o << "// ---- entry points -------------------------------------------------------------------\n";
  const bool all = group == JitGroup::All;
  for (int c = 0; c < model.ncomp(); ++c) {
    if (model.comp_nspec[c] == 0) continue;
    if (all || group == JitGroup::Element) {
      o << "extern \"C\" __global__ void __launch_bounds__(128, DC_ELEM_MINB) dc_k_residual_volume_" << c
        << "(DcVolArgs a) { dc_residual_volume<" << c << ">(a); }\n";
      o << "extern \"C\" __global__ void __launch_bounds__(128, DC_ELEM_MINB) dc_k_jacobian_apply_volume_" << c
        << "(DcVolArgs a) { dc_jacobian_apply_volume<" << c << ">(a); }\n";
      o << "extern \"C\" __global__ void __launch_bounds__(128, DC_ELEM_MINB) dc_k_jacobian_apply_volume_nomask_" << c
        << "(DcVolArgs a) { dc_jacobian_apply_volume<" << c << ", true>(a); }\n";
      o << "extern \"C\" __global__ void __launch_bounds__(128) dc_k_bdiag_volume_" << c
        << "(DcVolArgs a) { dc_jacobian_volume<" << c << ", 1>(a); }\n";
    }
    if (all || group == JitGroup::Csr) {
      o << "extern \"C\" __global__ void __launch_bounds__(128, DC_CSR_MINB) dc_k_jacobian_volume_" << c
        << "(DcVolArgs a) { dc_jacobian_volume<" << c << ", 0>(a); }\n";
      if (!model.numerical_jacobian && !model.has_extended_terms(c))
        o << "extern \"C\" __global__ void __launch_bounds__(64) dc_k_jacobian_gather_" << c
          << "(DcVolArgs a) { dc_jacobian_gather<" << c << ">(a); }\n";
    }
    if (all || group == JitGroup::Patch) {
      o << "extern \"C\" __global__ void __launch_bounds__(DC_PATCH_THREADS, DC_PATCH_MINB) dc_k_patch_residual_" << c
        << "(DcPatchArgs a) { dc_patch_kernel<" << c << ", 0>(a); }\n";
      o << "extern \"C\" __global__ void __launch_bounds__(DC_PATCH_THREADS, DC_PATCH_MINB) dc_k_patch_apply_" << c
        << "(DcPatchArgs a) { dc_patch_kernel<" << c << ", 1>(a); }\n";
      o << "extern \"C\" __global__ void __launch_bounds__(DC_PATCH_THREADS, DC_PATCH_MINB) dc_k_patch_bdiag_" << c
        << "(DcPatchArgs a) { dc_patch_kernel<" << c << ", 2>(a); }\n";
    }
  }
  // the implicit-geometry kernels serve grids that one compartment covers entirely: models with
  // species in several compartments never launch them, so they are not generated (compile time)
  int with_species = 0;
  for (int c = 0; c < model.ncomp(); ++c) with_species += model.comp_nspec[c] > 0;
  const bool structured_possible = with_species == 1;
  if (all || group == JitGroup::Structured) {
    o << kStructuredSource << "\n";
    for (int c = 0; c < model.ncomp(); ++c) {
      if (!structured_possible || model.comp_nspec[c] == 0 || !model.diffusion_is_constant(c)) continue;
      const char* names[4] = {"residual", "apply", "bdiag", "diag"};
      for (int mode = 0; mode < 4; ++mode)
        o << "extern \"C\" __global__ void __launch_bounds__(DC_STRUCT_THREADS, DC_STRUCT_MINB) dc_k_struct_" << names[mode] << "_" << c
          << "(DcStructArgs a) { dc_structured_kernel<" << c << ", " << mode << ">(a); }\n";
      o << "extern \"C\" __global__ void __launch_bounds__(DC_STRUCT_THREADS, DC_STRUCT_MINB) dc_k_struct_apply_scaled_" << c
        << "(DcStructArgs a) { dc_structured_kernel<" << c << ", 1, true>(a); }\n"
        << "extern \"C\" __global__ void __launch_bounds__(DC_STRUCT_THREADS, DC_STRUCT_MINB) dc_k_struct_apply_nomask_" << c
        << "(DcStructArgs a) { dc_structured_kernel<" << c << ", 1, false, true>(a); }\n";
      for (int mode = 0; mode < 2; ++mode)
        o << "extern \"C\" __global__ void __launch_bounds__(DC_STRUCT_THREADS, DC_STRUCT_MINB) dc_k_struct_march_" << names[mode] << "_" << c
          << "(DcStructArgs a) { dc_structured_march_kernel<" << c << ", " << mode << ">(a); }\n";
    }
  }
  if (all || group == JitGroup::StructuredQ1) {
    // Q1 cells of a structured lattice (BASELINE configs[3]; not a reference element type)
    if (!all) o << kStructuredSource << "\n";   // shared drivers (per cell / marching)
    o << kQ1Source << "\n";
    for (int c = 0; c < model.ncomp(); ++c) {
      if (!structured_possible || model.comp_nspec[c] == 0 || !model.diffusion_is_constant(c) || model.has_extended_terms(c)) continue;
      const char* names[5] = {"residual", "apply", "bdiag", "diag", "csr"};
      for (int mode = 0; mode < 5; ++mode)
        o << "extern \"C\" __global__ void __launch_bounds__(DC_STRUCT_THREADS, " << (mode == 4 ? "2" : "DC_STRUCT_MINB") << ") dc_k_q1_"
          << names[mode] << "_" << c << "(DcStructArgs a) { dc_q1_kernel<" << c << ", " << mode << ">(a); }\n";
      o << "extern \"C\" __global__ void __launch_bounds__(DC_STRUCT_THREADS, DC_STRUCT_MINB) dc_k_q1_apply_scaled_" << c
        << "(DcStructArgs a) { dc_q1_kernel<" << c << ", 1, true>(a); }\n"
        << "extern \"C\" __global__ void __launch_bounds__(DC_STRUCT_THREADS, DC_STRUCT_MINB) dc_k_q1_apply_nomask_" << c
        << "(DcStructArgs a) { dc_q1_kernel<" << c << ", 1, false, true>(a); }\n";
      for (int mode = 0; mode < 2; ++mode)
        o << "extern \"C\" __global__ void __launch_bounds__(DC_STRUCT_THREADS, DC_STRUCT_MINB) dc_k_q1_march_"
          << names[mode] << "_" << c << "(DcStructArgs a) { dc_q1_march_kernel<" << c << ", " << mode << ">(a); }\n";
    }
  }
  if (group == JitGroup::Tile || group == JitGroup::TileQ1) {
    // tile-marching drivers (kernels/assembly_tile.cuh) around the same cell functions; compiled on first use
    const bool q1 = group == JitGroup::TileQ1;
    if (q1) o << "#define DC_TILE_Q1 1\n";
    o << kStructuredSource << "\n";
    if (q1) o << kQ1Source << "\n";
    o << kTileSource << "\n";
    for (int c = 0; c < model.ncomp(); ++c) {
      if (!structured_possible || model.comp_nspec[c] == 0 || !model.diffusion_is_constant(c)) continue;
      if (q1 && model.has_extended_terms(c)) continue;
      const char* names[2] = {"residual", "apply"};
      for (int mode = 0; mode < 2; ++mode)
        o << "extern \"C\" __global__ void __launch_bounds__(DC_TILE_THREADS, DC_TILE_MINB) dc_k_tile_" << (q1 ? "q1_" : "")
          << names[mode] << "_" << c << "(DcTileArgs A) { dc_tile_" << (q1 ? "q1_" : "") << "kernel<" << c << ", " << mode << ">(A); }\n";
    }
  }
  if (all || group == JitGroup::Skeleton) {
    // One launch covers the facet lists of every directional compartment pair: the lists are small
    // (interfaces), so separate launches would be a chain of latency-bound kernels.  Blocks
    // [first[p], first[p+1]) work on pair p.
    size_t np = model.outflow_pairs().size();
    if (np) {
      o << "struct DcFacetArgsAll { DcFacetArgs a[" << np << "]; int first[" << np + 1 << "]; };\n";
      const char* names[4] = {"residual", "jacobian", "apply", "bdiag"};
      for (int mode = 0; mode < 4; ++mode) {
        o << "extern \"C\" __global__ void __launch_bounds__(64) dc_k_skeleton_" << names[mode] << "(DcFacetArgsAll A) {\n";
        for (size_t p = 0; p < np; ++p) {
          o << "  if ((int)blockIdx.x < A.first[" << p + 1 << "]) { ";
          if (mode == 0) o << "dc_skeleton_residual<" << p << ">(A.a[" << p << "]);";
          else o << "dc_skeleton_jacobian<" << p << ", " << mode - 1 << ">(A.a[" << p << "]);";
          o << " return; }\n";
        }
        o << "}\n";
      }
    }
  }
  return o.str();
}

std::string jit_defines(const Model& model) {
  const PTree& acfg = model.cfg.sub("model.assembly.b200");
  int th = acfg.get("patch_threads", 256), minb = acfg.get("patch_min_blocks", 3);
  if (th < 32 || th > 1024 || th % 32) fail("model.assembly.b200.patch_threads must be a multiple of 32 in [32,1024]");
  if (minb < 1 || minb > 8) fail("model.assembly.b200.patch_min_blocks out of range");
  // 32 threads x 12 resident CTAs measured best on B200 (same 12 warps per SM as 64 x 6, finer
  // grained: -4 % on the apply kernels; fewer registers spill, more registers starve the fp64 pipe)
  int sth = acfg.get("struct_threads", 32), sminb = acfg.get("struct_min_blocks", 12);
  if (sth < 32 || sth > 1024 || sth % 32) fail("model.assembly.b200.struct_threads must be a multiple of 32 in [32,1024]");
  if (sminb < 1 || sminb > 16) fail("model.assembly.b200.struct_min_blocks out of range");
  // CSR fill is latency bound (binary searches + fp64 atomics): 8 resident CTAs of 128 threads
  // measured 16 % faster than 4 on B200 despite the spills (profiles/r01_csr_fill_128_ncu.txt)
  int cminb = acfg.get("csr_min_blocks", 8);
  if (cminb < 1 || cminb > 16) fail("model.assembly.b200.csr_min_blocks out of range");
  // element-per-thread residual / apply (128 threads): resident CTAs per SM the compiler has to leave room for
  // structured per-cell driver: L2 prefetch of the first-touch vertex row one plane ahead (assembly_structured.cuh)
  const int host_vol = acfg.get("host_vol", false) ? 1 : 0;   // simplex volume from the host instead of one fp64 division per thread
  const int sprefetch = acfg.get("struct_prefetch", false) ? 1 : 0;   // measured neutral on B200: off
  int eminb = acfg.get("elem_min_blocks", 4);   // 4 x 128 threads: 128 registers, measured -2 % on the cell model against 164 registers x 3
  if (eminb < 1 || eminb > 16) fail("model.assembly.b200.elem_min_blocks out of range");
  // tile-marching drivers: 32 x (tile_w * tile_r) cells per CTA in 3-D (tile_w warps, tile_r cell rows each)
  int ns_max = 1;
  for (int c = 0; c < model.ncomp(); ++c) ns_max = std::max(ns_max, model.comp_nspec[c]);
  int tw = acfg.get("tile_w", 4), tr = acfg.get("tile_r", 1);
  int tminb = acfg.get("tile_min_blocks", ns_max <= 2 ? 3 : ns_max <= 4 ? 2 : 1);
  if (tw < 1 || tw > 32 || tr < 1 || tr > 16) fail("model.assembly.b200.tile_w / tile_r out of range");
  if (tminb < 1 || tminb > 16) fail("model.assembly.b200.tile_min_blocks out of range");
  return "#define DC_TILE_W " + std::to_string(tw) + "\n#define DC_TILE_R " + std::to_string(tr) +
         "\n#define DC_TILE_MINB " + std::to_string(tminb) + "\n#define DC_CSR_MINB " + std::to_string(cminb) + "\n#define DC_ELEM_MINB " + std::to_string(eminb) + "\n#define DC_STRUCT_PREFETCH " + std::to_string(sprefetch) + "\n#define DC_HOST_VOL " + std::to_string(host_vol) + "\n#define DC_PATCH_THREADS " + std::to_string(th) + "\n#define DC_PATCH_MINB " + std::to_string(minb) +
         "\n#define DC_STRUCT_THREADS " + std::to_string(sth) + "\n#define DC_STRUCT_MINB " + std::to_string(sminb) + "\n";
}

std::vector<char> jit_compile(const std::string& source, std::string* log, bool ptx) {
  nvrtcProgram prog;
  if (nvrtcCreateProgram(&prog, source.c_str(), "dune_copasi_b200_model.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS)
    fail("nvrtcCreateProgram failed");
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "--fmad=true",
                        "-default-device"};
  nvrtcResult res = nvrtcCompileProgram(prog, 5, opts);
  size_t log_size = 0;
  nvrtcGetProgramLogSize(prog, &log_size);
  std::string l(log_size, '\0');
  if (log_size) nvrtcGetProgramLog(prog, l.data());
  if (log) *log = l;
  if (res != NVRTC_SUCCESS) {
    nvrtcDestroyProgram(&prog);
    fail("NVRTC compilation of the model kernels failed (", nvrtcGetErrorString(res), "):\n", l);
  }
  std::vector<char> out;
  size_t n = 0;
  if (ptx) {
    nvrtcGetPTXSize(prog, &n);
    out.resize(n);
    nvrtcGetPTX(prog, out.data());
  } else {
    nvrtcGetCUBINSize(prog, &n);
    out.resize(n);
    nvrtcGetCUBIN(prog, out.data());
  }
  nvrtcDestroyProgram(&prog);
  return out;
}

namespace {
uint64_t fnv1a(const std::string& s) {
  uint64_t h = 1469598103934665603ULL;
  for (unsigned char c : s) { h ^= c; h *= 1099511628211ULL; }
  return h;
}
std::string cache_dir() {
  if (const char* e = std::getenv("DCB_JIT_CACHE")) return e;
  Dl_info info;
  if (dladdr((void*)&fnv1a, &info) && info.dli_fname) {
    std::string p = info.dli_fname;
    auto slash = p.rfind('/');
    return (slash == std::string::npos ? std::string(".") : p.substr(0, slash)) + "/_jit_cache";
  }
  return "/tmp/dune_copasi_b200_jit_cache";
}
}  // namespace

void jit_cache_evict(const std::string& source) {
  int major = 0, minor = 0;
  nvrtcVersion(&major, &minor);
  char name[64];
  snprintf(name, sizeof name, "%016llx_%d_%d.cubin", (unsigned long long)fnv1a(source), major, minor);
  std::remove((cache_dir() + "/" + name).c_str());
}

std::vector<char> jit_compile_cached(const std::string& source) {
  int major = 0, minor = 0;
  nvrtcVersion(&major, &minor);
  char name[64];
  snprintf(name, sizeof name, "%016llx_%d_%d.cubin",
           (unsigned long long)fnv1a(source), major, minor);
  const std::string dir = cache_dir(), path = dir + "/" + name;
  {
    std::ifstream f(path, std::ios::binary);
    if (f) {
      std::vector<char> bin((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
      if (!bin.empty()) {
        utimensat(AT_FDCWD, path.c_str(), nullptr, 0);   // mark as in use: build() prunes entries no build has touched
        return bin;
      }
    }
  }
  std::string log;
  std::vector<char> bin = jit_compile(source, &log, false);
  mkdir(dir.c_str(), 0755);
  const std::string tmp = path + ".tmp" + std::to_string((long long)getpid());
  bool written = false;
  {
    std::ofstream f(tmp, std::ios::binary);
    if (f) {
      f.write(bin.data(), (std::streamsize)bin.size());
      f.close();
      written = f.good();
    }
  }
  // atomic publish of a complete file only (a short write -- disk full, quota -- must not become a cache entry)
  if (!written || std::rename(tmp.c_str(), path.c_str()) != 0) std::remove(tmp.c_str());
  return bin;
}

JitModule::~JitModule() {
  if (lib_) cudaLibraryUnload(lib_);
}

void JitModule::load(const std::vector<char>& cubin) {
  require_device();
  if (lib_) { cudaLibraryUnload(lib_); lib_ = nullptr; cache_.clear(); }
  DCB_CUDA(cudaLibraryLoadData(&lib_, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
}

cudaKernel_t JitModule::kernel(const std::string& name) {
  auto it = cache_.find(name);
  if (it != cache_.end()) return it->second;
  cudaKernel_t k;
  cudaError_t e = cudaLibraryGetKernel(&k, lib_, name.c_str());
  if (e != cudaSuccess) fail("kernel '", name, "' not found in the JIT module: ", cudaGetErrorString(e));
  cache_[name] = k;
  return k;
}

}  // namespace dcb
