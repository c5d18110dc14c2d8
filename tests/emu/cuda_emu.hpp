// cuda_emu.hpp — TEST INFRASTRUCTURE: a minimal host emulation of the CUDA subset used by sdv-loam_b200/csrc/sdv_select_core.cuh, so that the REAL kernel
// source and its host orchestration (scratch layout, pass loop, launch sequence) run on the CPU build container — which has no GPU — against the oracle
// (tests/test_select_emu_cpu.py).  Not part of the product; the product compiles the same header with nvcc for sm_100a.
//   * kernels without barriers: every CUDA thread runs to completion, one after the other (launch(false, ...))
//   * kernels with __syncthreads: blockDim OS threads per launch walk the blocks together, barriers are pthread barriers (launch(true, ...))
//   * atomics are GCC __atomic builtins; __shared__ becomes a static local (blocks run one at a time); memory calls map to malloc / memcpy / memset
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <climits>
#include <thread>
#include <vector>
#include <functional>
#include <pthread.h>

struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
typedef int cudaError_t; typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 1; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __ldg(p) (*(p))
#define __ldcg(p) (*(p))

namespace emu {
extern thread_local dim3 t_threadIdx, t_blockIdx; extern dim3 g_blockDim, g_gridDim; extern pthread_barrier_t g_bar, g_bar_warp0; extern int g_or_flag[2];
#ifdef SDV_EMU_IMPL
thread_local dim3 t_threadIdx, t_blockIdx; dim3 g_blockDim, g_gridDim; pthread_barrier_t g_bar, g_bar_warp0; int g_or_flag[2];
#endif
static inline void launch(bool barriers, dim3 grid, dim3 block, const std::function<void()>& body) {
  g_blockDim = block; g_gridDim = grid;
  if (!barriers) {
    for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) { t_blockIdx = dim3(bx, by);
      for (unsigned tx = 0; tx < block.x; tx++) { t_threadIdx = dim3(tx); body(); } }
    return;
  }
  pthread_barrier_init(&g_bar, nullptr, block.x); pthread_barrier_init(&g_bar_warp0, nullptr, block.x < 32 ? block.x : 32); g_or_flag[0] = g_or_flag[1] = 0;
  std::vector<std::thread> th;
  for (unsigned tx = 0; tx < block.x; tx++) th.emplace_back([&, tx] {
    t_threadIdx = dim3(tx);
    for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) { t_blockIdx = dim3(bx, by); body(); pthread_barrier_wait(&g_bar); } });
  for (auto& t : th) t.join();
  pthread_barrier_destroy(&g_bar); pthread_barrier_destroy(&g_bar_warp0);
}
}
#define threadIdx emu::t_threadIdx
#define blockIdx emu::t_blockIdx
#define blockDim emu::g_blockDim
#define gridDim emu::g_gridDim
static inline void __syncthreads() { pthread_barrier_wait(&emu::g_bar); }
// only warp 0 of a block may call it in this emulation (the kernels that use it keep their warp-synchronous part in warp 0)
static inline void __syncwarp() { if (emu::t_threadIdx.x < 32) pthread_barrier_wait(&emu::g_bar_warp0); }
static inline int __syncthreads_or(int p) {          // two flags used alternately would race with a fast thread's next call; three barriers keep it simple
  if (p) __atomic_store_n(&emu::g_or_flag[0], 1, __ATOMIC_SEQ_CST);
  pthread_barrier_wait(&emu::g_bar); int r = __atomic_load_n(&emu::g_or_flag[0], __ATOMIC_SEQ_CST); pthread_barrier_wait(&emu::g_bar);
  if (emu::t_threadIdx.x == 0) emu::g_or_flag[0] = 0;
  pthread_barrier_wait(&emu::g_bar); return r;
}
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicMin(int* p, int v) { int o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (o > v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
static inline int atomicMax(int* p, int v) { int o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (o < v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
using std::isfinite;
#define SDV_LAUNCH(kern, grid, block, st, ...)      emu::launch(false, grid, block, [&] { kern(__VA_ARGS__); })
#define SDV_LAUNCH_SYNC(kern, grid, block, st, ...) emu::launch(true,  grid, block, [&] { kern(__VA_ARGS__); })
#define SDV_DEVCONST static const
static inline int __float_as_int(float f) { int b; memcpy(&b, &f, 4); return b; }
static inline float __int_as_float(int b) { float f; memcpy(&f, &b, 4); return f; }
static inline unsigned long long atomicOr(unsigned long long* p, unsigned long long v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
typedef void* cudaEvent_t;
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
namespace emu { extern unsigned char* g_dyn;
#ifdef SDV_EMU_IMPL
unsigned char* g_dyn = nullptr;
#endif
}
static inline unsigned int atomicCAS(unsigned int* p, unsigned int cmp, unsigned int val) { __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
#define SDV_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(emu::g_dyn)
#define SDV_SET_SMEM(kern, bytes) 0
#define SDV_LAUNCH_SYNC_SMEM(kern, grid, block, smem, st, ...) do { emu::g_dyn = (unsigned char*)calloc((smem) + 16, 1); emu::launch(true, grid, block, [&] { kern(__VA_ARGS__); }); free(emu::g_dyn); emu::g_dyn = nullptr; } while (0)
