// sdv_core_common.cuh — what the self-contained device cores (sdv_select_core.cuh, sdv_lidar_core.cuh) share: launch macros (CUDA, or the host emulation of tests/emu),
// a block-wide exclusive scan, a grow-only scratch allocator.  Depends on nothing but the CUDA runtime.
#pragma once
#include <stdint.h>
#include <math.h>
#include <vector>
#include <algorithm>
#include <string>
#include <limits.h>
#include <string.h>
#ifndef SDV_EMU
#include <cuda_runtime.h>
#define SDV_LAUNCH(kern, grid, block, st, ...)      kern<<<grid, block, 0, st>>>(__VA_ARGS__)
#define SDV_LAUNCH_SYNC(kern, grid, block, st, ...) kern<<<grid, block, 0, st>>>(__VA_ARGS__)
#define SDV_DEVCONST static __constant__
#define SDV_DYN_SMEM(T, name) extern __shared__ __align__(16) unsigned char name##_raw_[]; T* name = reinterpret_cast<T*>(name##_raw_)
#define SDV_SET_SMEM(kern, bytes) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
#define SDV_LAUNCH_SYNC_SMEM(kern, grid, block, smem, st, ...) kern<<<grid, block, smem, st>>>(__VA_ARGS__)
#endif

namespace sdv { namespace sel {

__device__ __forceinline__ int imin_(int a, int b) { return a < b ? a : b; }
// exclusive prefix sum of a block's values through shared memory (Hillis-Steele); every thread of the CTA must call it
__device__ __forceinline__ int block_excl_scan(int v, int* sm /* 2*blockDim */, int& total) {
  const int t = threadIdx.x, n = blockDim.x; int cur = 0;
  sm[t] = v; __syncthreads();
  for (int d = 1; d < n; d <<= 1) { const int x = sm[cur*n + t] + (t >= d ? sm[cur*n + t - d] : 0); sm[(cur^1)*n + t] = x; cur ^= 1; __syncthreads(); }
  const int incl = sm[cur*n + t]; total = sm[cur*n + n-1]; __syncthreads();
  return incl - v;
}

#define SEL_CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = std::string(#call) + " -> " + cudaGetErrorString(e_); return -1; } } while (0)

struct Scratch {                                             // grow-only device buffer carved into aligned pieces
  char* p = nullptr; size_t cap = 0, used = 0;
  int reserve(size_t bytes, cudaStream_t st) { if (bytes <= cap) return 0; cudaStreamSynchronize(st); if (p) cudaFree(p); p = nullptr; cap = 0; if (cudaMalloc((void**)&p, bytes + bytes/4) != cudaSuccess) return -1; cap = bytes + bytes/4; return 0; }
  void reset() { used = 0; }
  template <class T> T* take(size_t n) { used = (used + 255) & ~(size_t)255; T* r = (T*)(p + used); used += n*sizeof(T); return r; }
  static size_t need(size_t n, size_t sz) { return ((n*sz + 255) & ~(size_t)255) + 256; }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};


}}  // namespace sdv::sel
