// Micro-benchmark behind the scatter layout of the structured kernels: how does the L2 atomic unit of a B200
// charge fp64 `RED`s -- per element or per 32-byte sector?  Pattern A: every warp instruction touches 32 half
// sectors (one species of 32 consecutive vertices, stride 16 B); pattern B: the same elements as 8 full sectors per
// instruction (256 contiguous bytes).  Both add 4 times to every entry of a 2 x 257^3 vector, as the apply kernel does.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o red_sectors red_sectors.cu && ./red_sectors
#include <cstdio>
#include <cuda_runtime.h>

template <int PATTERN>
__global__ void __launch_bounds__(32) k(double* y, long long nvert, int rowlen) {
  const long long w = blockIdx.x;            // one warp per 32 consecutive vertices
  const int lane = threadIdx.x;
  const long long v0 = w * 32;
  if (v0 + 32 > nvert) return;
#pragma unroll
  for (int rep = 0; rep < 4; ++rep) {
    long long vb = v0 + (rep & 1) * rowlen + (rep >> 1) * (long long)rowlen * rowlen;
    if (vb + 32 > nvert) vb = v0;
    if (PATTERN == 0) {
      atomicAdd(&y[(vb + lane) * 2 + 0], 1.0);
      atomicAdd(&y[(vb + lane) * 2 + 1], 1.0);
    } else {
      atomicAdd(&y[vb * 2 + lane], 1.0);
      atomicAdd(&y[vb * 2 + 32 + lane], 1.0);
    }
  }
}

int main() {
  const int rowlen = 257;
  const long long nvert = (long long)rowlen * rowlen * rowlen;
  double* y;
  cudaMalloc(&y, nvert * 2 * sizeof(double));
  cudaMemset(y, 0, nvert * 2 * sizeof(double));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int blocks = (int)(nvert / 32);
  for (int pat = 0; pat < 2; ++pat)
    for (int it = 0; it < 4; ++it) {
      cudaMemsetAsync(y, 0, nvert * 2 * sizeof(double));
      cudaEventRecord(a);
      if (pat == 0) k<0><<<blocks, 32>>>(y, nvert, rowlen); else k<1><<<blocks, 32>>>(y, nvert, rowlen);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      if (it) printf("pattern %c: %.3f ms, %.1f G RED elements/s\n", pat ? 'B' : 'A', ms, nvert * 8.0 / ms / 1e6);
    }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
