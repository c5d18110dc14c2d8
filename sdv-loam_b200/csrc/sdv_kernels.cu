// sdv_kernels.cu — hand-written sm_100a kernels of the tracker path (compiled with --fmad=false).
//
//   pyr_grad_kernel / pyr_down_kernel      FrameHessian::makeImages           HessianBlocks.cpp:107-167
//   coarse_res_gs_kernel                   CoarseTracker::calcRes + calcGSSSE  CoarseTracker.cpp:486-634, 427-484 (one fused pass)
//   track_cluster_kernel                   CoarseTracker::trackNewestCoarse    CoarseTracker.cpp:662-838 (whole coarse-to-fine LM
//                                          device-resident: one thread-block cluster per call, DSMEM all-reduce, per-CTA redundant solve)
//   cd_* kernels                           CoarseTracker::makeCoarseDepthL0    CoarseTracker.cpp:258-425
#include "sdv_kernels.cuh"
#include "sdv_warp_solve.cuh"
#include <cooperative_groups.h>
#include <cstdlib>
namespace cg = cooperative_groups;

namespace sdv {

// ================================================================================================ pyramid
// Batched over frames (blockIdx.y = frame of the batch): one launch per level for the whole batch.
// gradient + pack of one level from a planar intensity image (HessianBlocks.cpp:147-165).  Flat-index neighbours on
// purpose: at x=0 / x=w-1 the reference reads across the row boundary (idx±1), and so do we.
struct PyrBatch { const void* src; float* I0; float* scratch; float4* out; int flags; int pad; };   // per frame: level-0 input (float/u8; may equal I0), level-0 plane, planar scratch, base of levels >= 1; flags bit0: photometric response applies (exposure > 0)
static_assert(sizeof(PyrBatch) == sizeof(PyrBatchHost), "PyrBatch mirrors PyrBatchHost");

template <typename T> __device__ __forceinline__ float px_load(const T* p, int i);
template <> __device__ __forceinline__ float px_load<float>(const float* p, int i) { return __ldg(p + i); }
template <> __device__ __forceinline__ float px_load<unsigned char>(const unsigned char* p, int i) { return (float)__ldg(p + i); }

__device__ __forceinline__ float4 grad_texel(const float* __restrict__ I, int idx, int w, int h) {
  float c = __ldg(I + idx); float dx = 0.f, dy = 0.f, ab = 0.f;
  if (idx >= w && idx < w*(h-1)) {
    dx = 0.5f*(__ldg(I + idx+1) - __ldg(I + idx-1));
    dy = 0.5f*(__ldg(I + idx+w) - __ldg(I + idx-w));
    if (!isfinite(dx)) dx = 0;
    if (!isfinite(dy)) dy = 0;
    ab = dx*dx + dy*dy;
  }
  return make_float4(c, dx, dy, ab);
}
// packed {I,dx,dy,|grad|^2} texels of one level >= 1 from its planar intensity in scratch
__global__ void __launch_bounds__(256) pyr_grad_kernel(const PyrBatch* __restrict__ batch, size_t scratch_off, size_t out_off, int w, int h) {
  const PyrBatch b = batch[blockIdx.y];
  const float* I = b.scratch + scratch_off; float4* out = b.out + out_off; const int n = w*h;
  for (int idx = blockIdx.x*blockDim.x + threadIdx.x; idx < n; idx += gridDim.x*blockDim.x) out[idx] = grad_texel(I, idx, w, h);
}
// level-0 packed texels on demand (keyframes entering the BA window, debugging read-back)
__global__ void __launch_bounds__(256) pyr_grad0_kernel(const float* __restrict__ I, float4* __restrict__ out, int w, int h) {
  const int n = w*h;
  for (int idx = blockIdx.x*blockDim.x + threadIdx.x; idx < n; idx += gridDim.x*blockDim.x) out[idx] = grad_texel(I, idx, w, h);
}
// 2x2 box filter of intensities (HessianBlocks.cpp:137-145): 0.25f*(((a+b)+c)+d).  Level 1 reads the level-0 input (float or mono8) and,
// when that input is not already the frame's own plane, also materialises the level-0 float plane in the same pass.
template <typename T>
__global__ void __launch_bounds__(256) pyr_down_kernel(const PyrBatch* __restrict__ batch, int from_src, size_t src_off, size_t dst_off, int wl, int hl, int wlm1) {
  const PyrBatch b = batch[blockIdx.y];
  const T* Iprev = from_src ? reinterpret_cast<const T*>(b.src) : reinterpret_cast<const T*>(b.scratch + src_off);
  float* I = b.scratch + dst_off;
  float* I0 = (from_src && reinterpret_cast<const void*>(b.I0) != b.src) ? b.I0 : nullptr;
  const int n = wl*hl;
  for (int idx = blockIdx.x*blockDim.x + threadIdx.x; idx < n; idx += gridDim.x*blockDim.x) {
    int y = idx / wl, x = idx - y*wl;
    int bi = 2*x + 2*y*wlm1;
    float a0 = px_load<T>(Iprev, bi), a1 = px_load<T>(Iprev, bi+1), a2 = px_load<T>(Iprev, bi+wlm1), a3 = px_load<T>(Iprev, bi+wlm1+1);
    I[idx] = 0.25f * (((a0 + a1) + a2) + a3);
    if (I0) { I0[bi] = a0; I0[bi+1] = a1; I0[bi+wlm1] = a2; I0[bi+wlm1+1] = a3; }
  }
}

// Undistort::undistort<unsigned char> (util/Undistort.cpp:341-435) fused with PhotometricUndistorter::processFrame (:177-214): one rectified level-0 pixel from the
// RAW mono8 wire image.  remapX/remapY are the tables of Undistort::readFromFile (:842-886; -1 = outside), the four taps go through the photometric stage
// (factor*v, or G[v] (* vignetteMapInv) with a response calibration and exposure > 0) and are blended in the reference's order (fmad off):
//   xxyy*src[1+wOrg] + (yy-xxyy)*src[wOrg] + (xx-xxyy)*src[1] + (1-xx-yy+xxyy)*src[0]
__device__ __forceinline__ float undist_tap(const unsigned char* __restrict__ raw, int i, const UndistortDev& U, bool photo) {
  const unsigned char v = __ldg(raw + i);
  if (!photo) return U.factor * (float)v;
  float r = __ldg(U.G + v);
  if (U.vignette) r *= __ldg(U.vignette + i);
  return r;
}
__device__ __forceinline__ float undist_px(const unsigned char* __restrict__ raw, int idx, const UndistortDev& U, bool photo) {
  float xx = __ldg(U.remapX + idx), yy = __ldg(U.remapY + idx);
  if (xx < 0) return 0.f;
  const int xxi = (int)xx, yyi = (int)yy;
  xx -= xxi; yy -= yyi;
  const float xxyy = xx*yy;
  const int o = xxi + yyi*U.wOrg;
  const float s0 = undist_tap(raw, o, U, photo), s1 = undist_tap(raw, o+1, U, photo), sw = undist_tap(raw, o+U.wOrg, U, photo), sw1 = undist_tap(raw, o+1+U.wOrg, U, photo);
  return xxyy*sw1 + (yy-xxyy)*sw + (xx-xxyy)*s1 + (1-xx-yy+xxyy)*s0;
}
// raw ingest, levels > 1: level-0 plane + level 1 in one pass (one thread = one level-1 pixel = a 2x2 quad of rectified pixels)
__global__ void __launch_bounds__(256) pyr_down_remap_kernel(const PyrBatch* __restrict__ batch, UndistortDev U, size_t dst_off, int wl, int hl, int wlm1) {
  const PyrBatch b = batch[blockIdx.y];
  const unsigned char* raw = reinterpret_cast<const unsigned char*>(b.src); const bool photo = (b.flags & 1) && U.G;
  float* I = b.scratch + dst_off; float* I0 = b.I0;
  const int n = wl*hl;
  for (int idx = blockIdx.x*blockDim.x + threadIdx.x; idx < n; idx += gridDim.x*blockDim.x) {
    int y = idx / wl, x = idx - y*wl;
    int bi = 2*x + 2*y*wlm1;
    float a0 = undist_px(raw, bi, U, photo), a1 = undist_px(raw, bi+1, U, photo), a2 = undist_px(raw, bi+wlm1, U, photo), a3 = undist_px(raw, bi+wlm1+1, U, photo);
    I[idx] = 0.25f * (((a0 + a1) + a2) + a3);
    *reinterpret_cast<float2*>(I0 + bi) = make_float2(a0, a1); *reinterpret_cast<float2*>(I0 + bi + wlm1) = make_float2(a2, a3);
  }
}
__global__ void __launch_bounds__(256) pyr_remap0_kernel(const PyrBatch* __restrict__ batch, UndistortDev U, int n) {
  const PyrBatch b = batch[blockIdx.y]; const bool photo = (b.flags & 1) && U.G;
  for (int idx = blockIdx.x*blockDim.x + threadIdx.x; idx < n; idx += gridDim.x*blockDim.x) b.I0[idx] = undist_px(reinterpret_cast<const unsigned char*>(b.src), idx, U, photo);
}

size_t pyramid_scratch_floats(int w, int h, int levels) { size_t n = 0; for (int l = 1; l < levels; l++) n += (size_t)(w>>l)*(h>>l); return n + 4; }

// batch_dev: nframes PyrBatch descriptors in device memory.  src_u8: level-0 input is mono8 (sensor_msgs/Image wire format) instead of float.
// Requires even w,h whenever levels > 1 (pyrLevelsUsed only halves even sizes, globalCalib.cpp:24).
void launch_pyramid_batch(const void* batch_dev_v, int nframes, bool src_u8, const size_t* lvl_off, int w, int h, int levels, cudaStream_t st, const UndistortDev* und) {
  const PyrBatch* batch_dev = reinterpret_cast<const PyrBatch*>(batch_dev_v);
  if (nframes <= 0) return;
  size_t soff = 0, prev_soff = 0; int wl = w, hl = h;
  int cap = (148*16 + nframes - 1)/nframes; if (cap < 4) cap = 4;
  for (int l = 0; l + 1 < levels; l++) {
    int wn = wl>>1, hn = hl>>1; int gn = (wn*hn + 255)/256; if (gn > cap) gn = cap;
    dim3 g2(gn, nframes);
    if (l == 0 && und) pyr_down_remap_kernel<<<g2, 256, 0, st>>>(batch_dev, *und, soff, wn, hn, wl);
    else if (l == 0 && src_u8) pyr_down_kernel<unsigned char><<<g2, 256, 0, st>>>(batch_dev, 1, 0, soff, wn, hn, wl);
    else pyr_down_kernel<float><<<g2, 256, 0, st>>>(batch_dev, l == 0, prev_soff, soff, wn, hn, wl);
    pyr_grad_kernel<<<g2, 256, 0, st>>>(batch_dev, soff, lvl_off[l+1], wn, hn);
    prev_soff = soff; soff += (size_t)wn*hn; wl = wn; hl = hn;
  }
}
// levels == 1 or odd sizes: the level-0 plane still has to be materialised from a foreign / mono8 source
template <typename T> __global__ void pyr_copy0_kernel(const PyrBatch* __restrict__ batch, int n) {
  const PyrBatch b = batch[blockIdx.y]; if (reinterpret_cast<const void*>(b.I0) == b.src) return;
  for (int idx = blockIdx.x*blockDim.x + threadIdx.x; idx < n; idx += gridDim.x*blockDim.x) b.I0[idx] = px_load<T>(reinterpret_cast<const T*>(b.src), idx);
}
void launch_pyramid_copy0(const void* batch_dev_v, int nframes, bool src_u8, int w, int h, cudaStream_t st, const UndistortDev* und) {
  const PyrBatch* batch_dev = reinterpret_cast<const PyrBatch*>(batch_dev_v); if (nframes <= 0) return;
  dim3 g((w*h + 255)/256 > 1024 ? 1024 : (w*h + 255)/256, nframes);
  if (und) pyr_remap0_kernel<<<g, 256, 0, st>>>(batch_dev, *und, w*h);
  else if (src_u8) pyr_copy0_kernel<unsigned char><<<g, 256, 0, st>>>(batch_dev, w*h); else pyr_copy0_kernel<float><<<g, 256, 0, st>>>(batch_dev, w*h);
}
void launch_pyramid_level0_texels(const float* I0, float4* out, int w, int h, cudaStream_t st) {
  int n = w*h; int grid = (n + 255)/256; if (grid > 148*16) grid = 148*16;
  pyr_grad0_kernel<<<grid, 256, 0, st>>>(I0, out, w, h);
}

__global__ void unpack_level_kernel(const float4* __restrict__ in, float* dI3, float* ab, int n) {
  int i = blockIdx.x*blockDim.x + threadIdx.x; if (i >= n) return;
  float4 t = in[i];
  if (dI3) { dI3[3*i] = t.x; dI3[3*i+1] = t.y; dI3[3*i+2] = t.z; }
  if (ab) ab[i] = t.w;
}
void launch_unpack_level(const float4* in, float* dI3, float* ab, int n, cudaStream_t st) {
  unpack_level_kernel<<<(n+255)/256, 256, 0, st>>>(in, dI3, ab, n);
}

// ================================================================================================ reductions
// Deterministic CTA reduction of the 51 per-thread partial sums: transpose through shared memory, each warp owns rows,
// 8 serial adds + xor-butterfly in double.  Fixed order => bit-reproducible run to run.
template <int THREADS>
__device__ __forceinline__ void block_reduce_acc(const float (&acc)[kNAcc], float* red, double* out) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int k = 0; k < kNAcc; k++) red[k*THREADS + tid] = acc[k];
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31; constexpr int NW = THREADS/32;
  for (int k = warp; k < kNAcc; k += NW) {
    double s = 0;
#pragma unroll
    for (int j = 0; j < NW; j++) s += (double)red[k*THREADS + lane + 32*j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[k] = s;
  }
  __syncthreads();
}

// ================================================================================================ step-wise fused calcRes+calcGSSSE
// grid-stride over the reference cloud; per-block partials -> global; the last block to finish (ticket) sums the block
// partials in block order and writes the 51 totals.  One launch per calcRes.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) coarse_res_gs_kernel(const float4* __restrict__ pts, int n, const float4* __restrict__ img, const float* __restrict__ I0,
                                                               LevelGeom g, EvalParams ep, double* __restrict__ partials,
                                                               unsigned int* __restrict__ ticket, double* __restrict__ totals) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* red = reinterpret_cast<float*>(smem_raw);
  __shared__ double bsum[kNAcc];
  __shared__ bool is_last;
  float acc[kNAcc];
#pragma unroll
  for (int k = 0; k < kNAcc; k++) acc[k] = 0.f;
  {
    const int stride = gridDim.x*THREADS; int i = blockIdx.x*THREADS + threadIdx.x;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 pc = (i < n) ? __ldg(pts + i) : zero4;
    float4 pn = (i + stride < n) ? __ldg(pts + i + stride) : zero4;
    if (i < n) prefetch_taps(pc, g, ep, img);
    for (; i < n; i += stride) {
      const float4 pn2 = (i + 2*stride < n) ? __ldg(pts + i + 2*stride) : zero4;
      if (i + stride < n) prefetch_taps(pn, g, ep, img);
      eval_point(pc, i, g, ep, img, I0, acc);
      pc = pn; pn = pn2;
    }
  }
  block_reduce_acc<THREADS>(acc, red, bsum);
  if (threadIdx.x < kNAcc) partials[(size_t)blockIdx.x*kNAcc + threadIdx.x] = bsum[threadIdx.x];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) { unsigned int t = atomicAdd(ticket, 1u); is_last = (t == gridDim.x - 1); }
  __syncthreads();
  if (is_last) {
    __threadfence();
    if (threadIdx.x < kNAcc) {
      double s = 0; for (unsigned int b = 0; b < gridDim.x; b++) s += __ldcg(partials + (size_t)b*kNAcc + threadIdx.x);
      totals[threadIdx.x] = s;
    }
    if (threadIdx.x == 0) *ticket = 0;
  }
}

constexpr int kStepThreads = 256;
int step_kernel_max_grid() { return 148*4; }
void launch_coarse_res_gs(const float4* pts, int n, const float4* img, const float* I0, const LevelGeom& g, const EvalParams& ep,
                          double* partials, unsigned int* ticket, double* totals, cudaStream_t st) {
  size_t smem = (size_t)kNAcc*kStepThreads*sizeof(float);                // opt-in size set per device by kernels_init_device()
  int grid = (n + kStepThreads - 1)/kStepThreads; if (grid < 1) grid = 1; if (grid > step_kernel_max_grid()) grid = step_kernel_max_grid();
  coarse_res_gs_kernel<kStepThreads><<<grid, kStepThreads, smem, st>>>(pts, n, img, I0, g, ep, partials, ticket, totals);
}

// ================================================================================================ device-resident LM
struct Ctl {                               // per-CTA copy of the LM state (every CTA of the cluster computes it redundantly and identically)
  SE3d cur; double a_cur, b_cur;           // refToNew_current, aff_g2l_current
  SE3d cand; double a_cand, b_cand;        // refToNew_new, aff_g2l_new
  double H[64], b[8];
  double resOld[6], resNew[6];
  double lastRes[5], flow[3];
  float lambda;
  int flag;
  EvalParams ep;
};

constexpr int kMaxCluster = 16;

// THREADS x MINB chosen so that several jobs are resident per SM: one job's (single-warp, latency-bound) LM control step
// overlaps the point sweeps of the others.  <128,4> is the batched-throughput configuration (cluster size 1);
// <256,1> with a cluster of 8-16 CTAs is the low-latency single-sequence configuration.
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) track_cluster_v1_kernel(TrackJob* __restrict__ jobs, const TrackConst* __restrict__ tc_g) {
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x;
  TrackJob& J = jobs[blockIdx.x / C];

  extern __shared__ __align__(16) unsigned char smem_raw[];
  float*  red    = reinterpret_cast<float*>(smem_raw);                                   // [kNAcc][THREADS]
  double* gather = reinterpret_cast<double*>(smem_raw + (size_t)kNAcc*THREADS*sizeof(float)); // [2][kMaxCluster][kNAcc]
  double* bsum   = gather + 2*kMaxCluster*kNAcc;                                           // [kNAcc]
  double* tot    = bsum + kNAcc;                                                           // [kNAcc]
  __shared__ Ctl ctl;
  __shared__ TrackConst tc;
  for (int i = tid; i < (int)(sizeof(TrackConst)/4); i += THREADS) reinterpret_cast<int*>(&tc)[i] = reinterpret_cast<const int*>(tc_g)[i];
  if (tid == 0) {
    ctl.cur = se3_from7(J.T); ctl.a_cur = J.ab[0]; ctl.b_cur = J.ab[1];
    for (int i = 0; i < 5; i++) ctl.lastRes[i] = nan("");
    for (int i = 0; i < 3; i++) ctl.flow[i] = 1000.0;
  }
  // NOTE: keep these statistics as per-thread local arrays.  Moving them to shared memory changed ptxas' allocation under the
  // 128-register cap (1 spill in the sweep) and cost 45 % of the kernel's throughput (measured, round 1).
  long long evals[kLevels]; int iters[kLevels], accs[kLevels];
#pragma unroll
  for (int i = 0; i < kLevels; i++) { evals[i] = 0; iters[i] = 0; accs[i] = 0; }
  __syncthreads();
  if (C > 1) cluster.sync();                                // every CTA of the cluster must be running before the first distributed-shared-memory write (racecheck)

  int evalCount = 0;
  // all threads: evaluate the residual + GN system at ctl.ep for level `lvl`; totals (identical in every CTA) land in tot[]
  auto eval = [&](int lvl) {
    float acc[kNAcc];
#pragma unroll
    for (int k = 0; k < kNAcc; k++) acc[k] = 0.f;
    const float4* __restrict__ pts = J.pts[lvl]; const int n = J.npts[lvl];
    const float4* __restrict__ img = (lvl == 0) ? nullptr : J.img[lvl]; const float* __restrict__ I0 = (lvl == 0) ? J.img0 : nullptr;
    const LevelGeom& g = tc.geom[lvl];
    {                                                      // 3-deep software pipeline: point k+2 in flight, taps of k+1 prefetched, k evaluated
      const int stride = C*THREADS; int i = rank*THREADS + tid;
      const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 pc = (i < n) ? __ldg(pts + i) : zero4;
      float4 pn = (i + stride < n) ? __ldg(pts + i + stride) : zero4;
      if (i < n) prefetch_taps(pc, g, ctl.ep, img);
      for (; i < n; i += stride) {
        const float4 pn2 = (i + 2*stride < n) ? __ldg(pts + i + 2*stride) : zero4;
        if (i + stride < n) prefetch_taps(pn, g, ctl.ep, img);
        eval_point(pc, i, g, ctl.ep, img, I0, acc);
        pc = pn; pn = pn2;
      }
    }
    block_reduce_acc<THREADS>(acc, red, bsum);
    if (C > 1) {
      const int buf = evalCount & 1;
      for (int k = tid; k < kNAcc*C; k += THREADS) {
        int r = k / kNAcc, e = k - r*kNAcc;
        double* dst = cluster.map_shared_rank(gather, r);
        dst[(buf*kMaxCluster + rank)*kNAcc + e] = bsum[e];
      }
      cluster.sync();
      if (tid < kNAcc) { double s = 0; for (int r = 0; r < C; r++) s += gather[(buf*kMaxCluster + r)*kNAcc + tid]; tot[tid] = s; }
    } else {
      if (tid < kNAcc) tot[tid] = bsum[tid];
    }
    __syncthreads();
    evalCount++; evals[lvl] += n;
  };

  const int maxIterations[5] = {10,20,50,50,50};           // CoarseTracker.cpp:679
  const float lambdaExtrapolationLimit = 0.001f;            // :680
  bool haveRepeated = false, aborted = false;
  for (int lvl = J.coarsest; lvl >= 0; lvl--) {
    float levelCutoffRepeat = 1;
    if (tid == 0) make_eval_params(ctl.cur, ctl.a_cur, ctl.b_cur, J.refExposure, J.newExposure, J.ref_a, J.ref_b, tc.geom[lvl], lvl,
                                   tc.coarseCutoffTH*levelCutoffRepeat, tc.huberTH, ctl.ep);
    __syncthreads();
    eval(lvl);
    if (tid == 0) { finalize_res(tot, ctl.resOld); ctl.flag = (ctl.resOld[5] > 0.6 && levelCutoffRepeat < 50); }
    __syncthreads();
    while (ctl.flag) {                                      // :694-701
      levelCutoffRepeat *= 2;
      __syncthreads();
      if (tid == 0) make_eval_params(ctl.cur, ctl.a_cur, ctl.b_cur, J.refExposure, J.newExposure, J.ref_a, J.ref_b, tc.geom[lvl], lvl,
                                     tc.coarseCutoffTH*levelCutoffRepeat, tc.huberTH, ctl.ep);
      __syncthreads();
      eval(lvl);
      if (tid == 0) { finalize_res(tot, ctl.resOld); ctl.flag = (ctl.resOld[5] > 0.6 && levelCutoffRepeat < 50); }
      __syncthreads();
    }
    if (tid == 0) { finalize_gs(tot, ctl.H, ctl.b); ctl.lambda = 0.01f; }

    for (int iteration = 0; iteration < maxIterations[lvl]; iteration++) {
      iters[lvl]++;
      double incn = 0;
      if (tid < 32) {                                        // warp 0: propose the LM step (CoarseTracker.cpp:722-765)
        __syncwarp();                                        // ctl.H / ctl.b / ctl.lambda were written by lane 0 (finalize_gs) — order them before the other lanes' reads (racecheck)
        const int r = tid & 7;
        const float lambda = ctl.lambda;
        const bool fixA = tc.affineOptModeA < 0, fixB = tc.affineOptModeB < 0;
        const int nv = (!fixA && !fixB) ? 8 : ((fixA && fixB) ? 6 : 7);
        const bool stitch = fixA && !fixB;                   // fix a only: b takes slot 6 (:736-748)
        const int rr = (stitch && r == 6) ? 7 : r;
        double a[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int cc = (stitch && j == 6) ? 7 : j;
          double hv = ctl.H[rr*8 + cc];
          if (rr == cc) hv *= (1 + lambda);
          a[j] = (r < nv && j < nv) ? hv : ((r == j) ? 1.0 : 0.0);
        }
        double rhs = (r < nv) ? -ctl.b[rr] : 0.0;
        const double x = warp_ldlt_solve8(a, rhs);
        double inc[8];
#pragma unroll
        for (int j = 0; j < 8; j++) inc[j] = __shfl_sync(0xffffffffu, x, j, 8);
        if (fixA && fixB) { inc[6] = 0; inc[7] = 0; }
        else if (!fixA && fixB) { inc[7] = 0; }
        else if (stitch) { inc[7] = inc[6]; inc[6] = 0; }
        float extrapFac = 1;
        if (lambda < lambdaExtrapolationLimit) extrapFac = sqrtf(sqrtf(lambdaExtrapolationLimit / lambda));
#pragma unroll
        for (int i = 0; i < 8; i++) inc[i] *= extrapFac;
        double incScaled[8];
#pragma unroll
        for (int i = 0; i < 3; i++) incScaled[i] = inc[i]*1.0f;        // SCALE_XI_ROT   (:755)
#pragma unroll
        for (int i = 3; i < 6; i++) incScaled[i] = inc[i]*0.5f;        // SCALE_XI_TRANS (:756)
        incScaled[6] = inc[6]*10.0f; incScaled[7] = inc[7]*1000.0f;    // SCALE_A, SCALE_B
        double s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) s += incScaled[i];
        if (!isfinite(s)) {
#pragma unroll
          for (int i = 0; i < 8; i++) incScaled[i] = 0;
        }
        const SE3d cand = se3_mul(se3_exp(incScaled), ctl.cur);
        const double a_cand = ctl.a_cur + incScaled[6], b_cand = ctl.b_cur + incScaled[7];
        EvalParams ep;
        make_eval_params(cand, a_cand, b_cand, J.refExposure, J.newExposure, J.ref_a, J.ref_b, tc.geom[lvl], lvl,
                         tc.coarseCutoffTH*levelCutoffRepeat, tc.huberTH, ep);
#pragma unroll
        for (int i = 0; i < 8; i++) incn += inc[i]*inc[i];
        incn = sqrt(incn);
        __syncwarp();
        if (tid == 0) { ctl.cand = cand; ctl.a_cand = a_cand; ctl.b_cand = b_cand; ctl.ep = ep; }
      }
      __syncthreads();
      eval(lvl);
      if (tid == 0) {
        finalize_res(tot, ctl.resNew);
        bool accept = (ctl.resNew[0]/ctl.resNew[1]) < (ctl.resOld[0]/ctl.resOld[1]);
        if (accept) {
          finalize_gs(tot, ctl.H, ctl.b);
          for (int i = 0; i < 6; i++) ctl.resOld[i] = ctl.resNew[i];
          ctl.cur = ctl.cand; ctl.a_cur = ctl.a_cand; ctl.b_cur = ctl.b_cand;
          ctl.lambda *= 0.5f;
        } else {
          ctl.lambda *= 4;
          if (ctl.lambda < lambdaExtrapolationLimit) ctl.lambda = lambdaExtrapolationLimit;
        }
        ctl.flag = (accept ? 1 : 0) | ((!(incn > 1e-3)) ? 2 : 0);
      }
      __syncthreads();
      const int f = ctl.flag;
      __syncthreads();
      if (f & 1) accs[lvl]++;
      if (f & 2) break;
    }
    if (tid == 0) {
      ctl.lastRes[lvl] = sqrtf((float)(ctl.resOld[0]/ctl.resOld[1]));
      ctl.flow[0] = ctl.resOld[2]; ctl.flow[1] = ctl.resOld[3]; ctl.flow[2] = ctl.resOld[4];
      ctl.flag = (ctl.lastRes[lvl] > 1.5*J.minRes[lvl]) ? 1 : 0;
    }
    __syncthreads();
    const int ab = ctl.flag;
    __syncthreads();
    if (ab) { aborted = true; break; }
    if (levelCutoffRepeat > 1 && !haveRepeated) { lvl++; haveRepeated = true; }
  }

  if (rank == 0 && tid == 0) {
    int good = 0;
    if (!aborted) {
      se3_to7(ctl.cur, J.T);
      double a_out = ctl.a_cur, b_out = ctl.b_cur;
      good = 1;
      if ((tc.affineOptModeA != 0 && (fabsf((float)a_out) > 1.2f)) || (tc.affineOptModeB != 0 && (fabsf((float)b_out) > 200))) good = 0;
      if (good) {
        double rel[2]; aff_from_to(J.refExposure, J.newExposure, J.ref_a, J.ref_b, a_out, b_out, rel);
        float r0 = (float)rel[0], r1 = (float)rel[1];
        if ((tc.affineOptModeA == 0 && (fabsf(logf(r0)) > 1.5f)) || (tc.affineOptModeB == 0 && (fabsf(r1) > 200))) good = 0;
        if (good) { if (tc.affineOptModeA < 0) a_out = 0; if (tc.affineOptModeB < 0) b_out = 0; }
      }
      J.ab[0] = a_out; J.ab[1] = b_out;
    }
    J.good = good;
    for (int i = 0; i < 5; i++) J.lastRes[i] = ctl.lastRes[i];
    for (int i = 0; i < 3; i++) J.flow[i] = ctl.flow[i];
    for (int i = 0; i < kLevels; i++) { J.point_evals[i] = evals[i]; J.iterations[i] = iters[i]; J.accepts[i] = accs[i]; }
  }
  if (C > 1) cluster.sync();                                // keep every CTA's shared memory alive until all remote writes/reads are done
}


// ================================================================================================ device-resident LM, asynchronous data path (round 2)
// Same algorithm and the same point -> thread mapping and summation order as track_cluster_v1_kernel (results are bit-identical), but no
// load ever stalls a warp on HBM:
//   * the reference cloud streams through shared memory in chunks: one elected thread issues cp.async.bulk (TMA bulk copy, SASS UBLKCP) per
//     2 KB block of float4 points, completion is signalled on an mbarrier (expect_tx), two chunks in flight;
//   * the bilinear footprint of point k+1 is gathered by per-thread cp.async (SASS LDGSTS) into a private shared-memory slot while point k
//     is accumulated — the gather needs no registers and is complete one full iteration later;
//   * the level-0 flow indicators (every 32nd point, CoarseTracker.cpp:538-566) are a dense pre-pass (all lanes busy) instead of a divergent branch
//     that cost one lane's 6 divisions per warp iteration;
//   * the LM control flow is a state machine around ONE instance of the sweep (a third of the code size of v1).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if defined(SDV_MBAR_SPIN)          // non-blocking test_wait poll
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
#elif defined(SDV_MBAR_HINT_NS)     // try_wait with an explicit suspend-time hint
  asm volatile("{\n\t.reg .pred p;\n\tSDV_WAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t@p bra SDV_DONE_%=;\n\tbra SDV_WAIT_%=;\n\tSDV_DONE_%=:\n\t}"
               :: "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)SDV_MBAR_HINT_NS) : "memory");
#else
  asm volatile("{\n\t.reg .pred p;\n\tSDV_WAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra SDV_DONE_%=;\n\tbra SDV_WAIT_%=;\n\tSDV_DONE_%=:\n\t}"
               :: "r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t s, const void* g) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(s), "l"(g) : "memory"); }
__device__ __forceinline__ void cp_async4(uint32_t s, const void* g)  { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(s), "l"(g) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1()  { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0()  { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct PtState { float u, v, nid, dx, dy, col; bool ok; };     // what the accumulate stage needs of a projected point (its taps are in the thread's smem slot)

// calcRes, first half (CoarseTracker.cpp:525-575): projection + bounds, then launch the gather of the 2x2 (level 0: 12-tap) footprint into `slot`.
template <int THREADS, bool LVL0>
__device__ __forceinline__ PtState stage_project(const float4 p, const LevelGeom& g, const EvalParams& ep, const float4* __restrict__ img, const float* __restrict__ I0, uint32_t slot) {
  PtState s; const float x = p.x, y = p.y, id = p.z; s.col = p.w;
  float pt0 = ((ep.RKi[0]*x + ep.RKi[1]*y) + ep.RKi[2]*1.0f) + ep.t[0]*id;
  float pt1 = ((ep.RKi[3]*x + ep.RKi[4]*y) + ep.RKi[5]*1.0f) + ep.t[1]*id;
  float pt2 = ((ep.RKi[6]*x + ep.RKi[7]*y) + ep.RKi[8]*1.0f) + ep.t[2]*id;
  s.u = pt0 / pt2; s.v = pt1 / pt2;
  const float Ku = g.fx*s.u + g.cx, Kv = g.fy*s.v + g.cy;
  s.nid = id / pt2;
  s.ok = (Ku > 2 && Kv > 2 && Ku < (float)(g.w-3) && Kv < (float)(g.h-3) && s.nid > 0);
  s.dx = 0.f; s.dy = 0.f;
  if (s.ok) {
    const int ix = (int)Ku, iy = (int)Kv, w = g.w;
    s.dx = Ku - ix; s.dy = Kv - iy;
    if (LVL0) {                                             // planar intensity: rows -1..2 (gradients are formed in stage_accumulate exactly like makeImages does)
      const float* b = I0 + ix + iy*w;
      cp_async4(slot + 0*THREADS*4, b - w);       cp_async4(slot + 1*THREADS*4, b - w + 1);
      cp_async4(slot + 2*THREADS*4, b - 1);       cp_async4(slot + 3*THREADS*4, b);         cp_async4(slot + 4*THREADS*4, b + 1);     cp_async4(slot + 5*THREADS*4, b + 2);
      cp_async4(slot + 6*THREADS*4, b + w - 1);   cp_async4(slot + 7*THREADS*4, b + w);     cp_async4(slot + 8*THREADS*4, b + w + 1); cp_async4(slot + 9*THREADS*4, b + w + 2);
      cp_async4(slot + 10*THREADS*4, b + 2*w);    cp_async4(slot + 11*THREADS*4, b + 2*w + 1);
    } else {
      const float4* bp = img + ix + iy*w;
      cp_async16(slot + 0*THREADS*16, bp); cp_async16(slot + 1*THREADS*16, bp + 1); cp_async16(slot + 2*THREADS*16, bp + w); cp_async16(slot + 3*THREADS*16, bp + w + 1);
    }
  }
  return s;
}
// calcRes second half + calcGSSSE for one point whose taps have landed (CoarseTracker.cpp:576-601, 442-466).  Expression trees identical to eval_point.
template <int THREADS, bool LVL0>
__device__ __forceinline__ void stage_accumulate(const PtState s, const LevelGeom& g, const EvalParams& ep, const unsigned char* slot, float (&acc)[kNAcc]) {
  if (!s.ok) return;
  float4 p00, p10, p01, p11;
  if (LVL0) {
    const float* t = reinterpret_cast<const float*>(slot);
    float a_m1_0 = t[0*THREADS], a_m1_1 = t[1*THREADS];
    float a_0_m1 = t[2*THREADS], a_0_0 = t[3*THREADS], a_0_1 = t[4*THREADS], a_0_2 = t[5*THREADS];
    float a_1_m1 = t[6*THREADS], a_1_0 = t[7*THREADS], a_1_1 = t[8*THREADS], a_1_2 = t[9*THREADS];
    float a_2_0 = t[10*THREADS], a_2_1 = t[11*THREADS];
    p00.x = a_0_0; p00.y = grad_guard(0.5f*(a_0_1 - a_0_m1)); p00.z = grad_guard(0.5f*(a_1_0 - a_m1_0));
    p10.x = a_0_1; p10.y = grad_guard(0.5f*(a_0_2 - a_0_0));  p10.z = grad_guard(0.5f*(a_1_1 - a_m1_1));
    p01.x = a_1_0; p01.y = grad_guard(0.5f*(a_1_1 - a_1_m1)); p01.z = grad_guard(0.5f*(a_2_0 - a_0_0));
    p11.x = a_1_1; p11.y = grad_guard(0.5f*(a_1_2 - a_1_0));  p11.z = grad_guard(0.5f*(a_2_1 - a_0_1));
  } else {
    const float4* t = reinterpret_cast<const float4*>(slot);
    p00 = t[0*THREADS]; p10 = t[1*THREADS]; p01 = t[2*THREADS]; p11 = t[3*THREADS];
  }
  const float dx = s.dx, dy = s.dy, dxdy = dx*dy, u = s.u, v = s.v, new_idepth = s.nid, refColor = s.col;
  float w11 = dxdy, w01 = dy-dxdy, w10 = dx-dxdy, w00 = 1-dx-dy+dxdy;
  float hit0 = ((w11*p11.x + w01*p01.x) + w10*p10.x) + w00*p00.x;
  float hit1 = ((w11*p11.y + w01*p01.y) + w10*p10.y) + w00*p00.y;
  float hit2 = ((w11*p11.z + w01*p01.z) + w10*p10.z) + w00*p00.z;
  if (!isfinite(hit0)) return;
  float residual = hit0 - (ep.aLL*refColor + ep.bLL);
  float ar = fabsf(residual);
  float hw = ar < ep.huber ? 1.0f : ep.huber / ar;
  acc[kIdxNE] += 1.0f;
  if (ar > ep.cutoff) { acc[kIdxE] += ep.maxEnergy; acc[kIdxNSat] += 1.0f; return; }
  acc[kIdxE] += hw*residual*residual*(2-hw);
  float dxf = hit1*g.fx, dyf = hit2*g.fy;
  float J[9];
  J[0] = new_idepth*dxf;
  J[1] = new_idepth*dyf;
  J[2] = 0.0f - new_idepth*(u*dxf + v*dyf);
  J[3] = 0.0f - ((u*v)*dxf + dyf*(1.0f + v*v));
  J[4] = (u*v)*dyf + dxf*(1.0f + u*u);
  J[5] = u*dyf - v*dxf;
  J[6] = ep.aLL*(ep.b0 - refColor);
  J[7] = -1.0f;
  J[8] = residual;
  int k = 0;
#pragma unroll
  for (int r = 0; r < 9; r++) {
    float Jw = J[r]*hw;
#pragma unroll
    for (int c = r; c < 9; c++) { acc[k] = fmaf(Jw, J[c], acc[k]); k++; }
  }
}
// flow indicators of one level-0 point with index % 32 == 0 (CoarseTracker.cpp:538-566); same expressions as eval_point
__device__ __forceinline__ void flow_point(const float4 p, const LevelGeom& g, const EvalParams& ep, float (&acc)[kNAcc]) {
  const float x = p.x, y = p.y, id = p.z;
  float pt0 = ((ep.RKi[0]*x + ep.RKi[1]*y) + ep.RKi[2]*1.0f) + ep.t[0]*id;
  float pt1 = ((ep.RKi[3]*x + ep.RKi[4]*y) + ep.RKi[5]*1.0f) + ep.t[1]*id;
  float pt2 = ((ep.RKi[6]*x + ep.RKi[7]*y) + ep.RKi[8]*1.0f) + ep.t[2]*id;
  float u = pt0 / pt2, v = pt1 / pt2;
  float Ku = g.fx*u + g.cx, Kv = g.fy*v + g.cy;
  float k0 = (g.Ki[0]*x + g.Ki[1]*y) + g.Ki[2]*1.0f, k1 = (g.Ki[3]*x + g.Ki[4]*y) + g.Ki[5]*1.0f, k2 = (g.Ki[6]*x + g.Ki[7]*y) + g.Ki[8]*1.0f;
  float r0 = (ep.RKi[0]*x + ep.RKi[1]*y) + ep.RKi[2]*1.0f, r1 = (ep.RKi[3]*x + ep.RKi[4]*y) + ep.RKi[5]*1.0f, r2 = (ep.RKi[6]*x + ep.RKi[7]*y) + ep.RKi[8]*1.0f;
  float a0 = k0 + ep.t[0]*id, a1 = k1 + ep.t[1]*id, a2 = k2 + ep.t[2]*id;
  float KuT = g.fx*(a0/a2) + g.cx, KvT = g.fy*(a1/a2) + g.cy;
  float b0_ = k0 - ep.t[0]*id, b1_ = k1 - ep.t[1]*id, b2_ = k2 - ep.t[2]*id;
  float KuT2 = g.fx*(b0_/b2_) + g.cx, KvT2 = g.fy*(b1_/b2_) + g.cy;
  float c0 = r0 - ep.t[0]*id, c1 = r1 - ep.t[1]*id, c2 = r2 - ep.t[2]*id;
  float Ku3 = g.fx*(c0/c2) + g.cx, Kv3 = g.fy*(c1/c2) + g.cy;
  acc[kIdxFlowT]  += (KuT-x)*(KuT-x) + (KvT-y)*(KvT-y);
  acc[kIdxFlowT]  += (KuT2-x)*(KuT2-x) + (KvT2-y)*(KvT2-y);
  acc[kIdxFlowRT] += (Ku-x)*(Ku-x) + (Kv-y)*(Kv-y);
  acc[kIdxFlowRT] += (Ku3-x)*(Ku3-x) + (Kv3-y)*(Kv3-y);
  acc[kIdxFlowN]  += 2.0f;
}

#ifdef SDV_TRACK_PROFILE
__device__ long long g_track_prof[16];
#ifndef SDV_PROF_MASK
#define SDV_PROF_MASK 0xffff
#endif
#define SDV_PROF_T(var) long long var = ((SDV_PROF_MASK >> __COUNTER__) & 1) ? clock64() : 0
#define SDV_PROF_ADD(slot, t0, t1) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_track_prof[slot] += (t1) - (t0); } while (0)
#else
#define SDV_PROF_T(var) do {} while (0)
#define SDV_PROF_ADD(slot, t0, t1) do {} while (0)
#endif
#ifndef SDV_TRACK_STAGGER_DEFAULT_NS
#define SDV_TRACK_STAGGER_DEFAULT_NS 0
#endif
constexpr int kPtChunk = 4;                                 // sweep iterations (blocks of THREADS points) per TMA-staged chunk.  Measured (B200, 592/1184 jobs): 2 -> -7 %, 8 -> -35 % (the
                                                            // larger buffers cost the 4th resident CTA); per-WARP streams with __syncwarp instead of the block barrier -> -5 % (warps drift apart)

struct Ctl2 {                                               // per-CTA LM state (every CTA of a cluster computes it redundantly and identically)
  SE3d cur; double a_cur, b_cur;
  SE3d cand; double a_cand, b_cand;
  double H[64], b[8];
  double resOld[6], resNew[6];
  double lastRes[5], flow[3];
  double incn;
  float lambda, lcr;                                        // LM damping, levelCutoffRepeat
  int lvl, mode, iteration, haveRepeated, aborted, done;
  long long evals[kLevels]; int iters[kLevels], accs[kLevels];
  EvalParams ep;
};
__constant__ unsigned char kAccRow[45] = {0,0,0,0,0,0,0,0,0, 1,1,1,1,1,1,1,1, 2,2,2,2,2,2,2, 3,3,3,3,3,3, 4,4,4,4,4, 5,5,5,5, 6,6,6, 7,7, 8};
__constant__ unsigned char kAccCol[45] = {0,1,2,3,4,5,6,7,8, 1,2,3,4,5,6,7,8, 2,3,4,5,6,7,8, 3,4,5,6,7,8, 4,5,6,7,8, 5,6,7,8, 6,7,8, 7,8, 8};

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) track_cluster_kernel(TrackJob* __restrict__ jobs, const TrackConst* __restrict__ tc_g, unsigned stagger_ns) {
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x;
  TrackJob& J = jobs[blockIdx.x / C];
  // Phase stagger: the MINB jobs co-resident on an SM run the same program on similar data and would reach their single-warp LM control steps
  // together, leaving the SM idle; CTAs are dealt round-robin over the SMs, so wave j of the grid (blockIdx / #SMs-worth) starts j*stagger_ns late
  // and one job's control step overlaps the other jobs' sweeps.
  if (stagger_ns && C == 1) {                               // only the first resident set is delayed; CTAs that start later are desynchronised by their predecessors
    unsigned nsm; asm("mov.u32 %0, %%nsmid;" : "=r"(nsm));
    const unsigned wave = (blockIdx.x < nsm*MINB) ? (blockIdx.x / nsm) % MINB : 0u;
    for (unsigned k = 0; k < wave; k++) __nanosleep(stagger_ns);
  }

  // dynamic smem: [ tap ring 2 x 4 x THREADS x 16 B | point chunks 2 x kPtChunk x THREADS x 16 B ] aliased by the reduction scratch [kNAcc][THREADS] floats,
  // then the DSMEM gather buffers and the reduced sums
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr size_t kRingBytes = (size_t)2*4*THREADS*16, kPtsBytes = (size_t)2*kPtChunk*THREADS*16, kRedBytes = (size_t)kNAcc*THREADS*sizeof(float);
  constexpr size_t kStageBytes = (kRingBytes + kPtsBytes) > kRedBytes ? (kRingBytes + kPtsBytes) : kRedBytes;
  unsigned char* ring = smem_raw;
  float4* pbuf   = reinterpret_cast<float4*>(smem_raw + kRingBytes);
  float*  red    = reinterpret_cast<float*>(smem_raw);
  double* gather = reinterpret_cast<double*>(smem_raw + kStageBytes);                    // [2][C][kNAcc] — sized by the launch's cluster size (13 KB at 16 CTAs, 0.8 KB at 1)
  double* bsum   = gather + 2*C*kNAcc;                                                   // [kNAcc]
  double* tot    = bsum + kNAcc;                                                         // [kNAcc]
  __shared__ Ctl2 ctl;
  __shared__ TrackConst tc;
  __shared__ __align__(8) uint64_t full_bar[2];
  for (int i = tid; i < (int)(sizeof(TrackConst)/4); i += THREADS) reinterpret_cast<int*>(&tc)[i] = reinterpret_cast<const int*>(tc_g)[i];
  if (tid == 0) {
    ctl.cur = se3_from7(J.T); ctl.a_cur = J.ab[0]; ctl.b_cur = J.ab[1];
    for (int i = 0; i < 5; i++) ctl.lastRes[i] = nan("");
    for (int i = 0; i < 3; i++) ctl.flow[i] = 1000.0;
    for (int i = 0; i < kLevels; i++) { ctl.evals[i] = 0; ctl.iters[i] = 0; ctl.accs[i] = 0; }
    ctl.lvl = J.coarsest; ctl.mode = 0; ctl.iteration = 0; ctl.haveRepeated = 0; ctl.aborted = 0; ctl.done = 0; ctl.lcr = 1.0f; ctl.lambda = 0.01f; ctl.incn = 0;
    mbar_init(&full_bar[0], 1); mbar_init(&full_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) make_eval_params(ctl.cur, ctl.a_cur, ctl.b_cur, J.refExposure, J.newExposure, J.ref_a, J.ref_b, tc.geom[ctl.lvl], ctl.lvl,
                                 tc.coarseCutoffTH*ctl.lcr, tc.huberTH, ctl.ep);
  __syncthreads();
  if (C > 1) cluster.sync();                                // every CTA of the cluster must be running before the first distributed-shared-memory write

  int evalCount = 0;
  uint32_t bar_phase = 0;                                   // bit b = parity the next wait on full_bar[b] expects
  const int stride = C*THREADS;
  const uint32_t slot_base = smem_u32(ring) + tid*16;       // level >= 1: [slot][tap][tid] float4 ; level 0 reuses the bytes as [slot][tap 0..11][tid] float
  const uint32_t slot_base0 = smem_u32(ring) + tid*4;
  constexpr uint32_t kSlotStride = 4*THREADS*16;

  while (true) {
    const int lvl = ctl.lvl; int n_eval = 0;
    // ------------------------------------------------------------------ one calcRes + calcGSSSE sweep at ctl.ep
    float acc[kNAcc];
#pragma unroll
    for (int k = 0; k < kNAcc; k++) acc[k] = 0.f;
    {
      const float4* __restrict__ pts = J.pts[lvl]; const int n = J.npts[lvl];
      const float4* __restrict__ img = (lvl == 0) ? nullptr : J.img[lvl]; const float* __restrict__ I0 = J.img0;
      const LevelGeom& g = tc.geom[lvl];
      const int K = (n + stride - 1)/stride;                // sweep iterations (uniform over the CTA)
      const int NC = (K + kPtChunk - 1)/kPtChunk;           // point chunks
      auto issue_chunk = [&](int c) {                       // thread 0: TMA bulk copies of chunk c's (<= kPtChunk) blocks into buffer c&1
        const int b = c & 1; uint32_t bytes = 0; int cnt[kPtChunk];
#pragma unroll
        for (int j = 0; j < kPtChunk; j++) {
          const int base = (c*kPtChunk + j)*stride + rank*THREADS; int m = n - base; m = m < 0 ? 0 : (m > THREADS ? THREADS : m);
          cnt[j] = m; bytes += (uint32_t)m*16u;
        }
        mbar_expect_tx(&full_bar[b], bytes);
#pragma unroll
        for (int j = 0; j < kPtChunk; j++) if (cnt[j] > 0)
          bulk_g2s(pbuf + (b*kPtChunk + j)*THREADS, pts + (size_t)(c*kPtChunk + j)*stride + rank*THREADS, (uint32_t)cnt[j]*16u, &full_bar[b]);
      };
      SDV_PROF_T(tp0);
#ifndef SDV_PTS_LDG
      if (tid == 0) { if (NC > 0) issue_chunk(0); if (NC > 1) issue_chunk(1); }
#endif
      SDV_PROF_T(tp1); SDV_PROF_ADD(0, tp0, tp1);
      if (lvl == 0) {                                       // dense flow-indicator pre-pass: point indices 0, 32, 64, ...
        for (int j = rank*THREADS + tid; 32*j < n; j += stride) flow_point(__ldg(pts + 32*j), g, ctl.ep, acc);
      }
      PtState sN; sN.ok = false; sN.u = sN.v = sN.nid = sN.dx = sN.dy = sN.col = 0.f;
      auto project = [&](int m) {                           // stage A for sweep iteration m (its chunk has landed)
        const int i = m*stride + rank*THREADS + tid;
        PtState s; s.ok = false; s.u = s.v = s.nid = s.dx = s.dy = s.col = 0.f;
        if (i < n) {
#ifdef SDV_PTS_LDG
          const float4 p = __ldg(pts + i);
#else
          const float4 p = pbuf[(((m/kPtChunk) & 1)*kPtChunk + (m % kPtChunk))*THREADS + tid];
#endif
          const uint32_t so = (uint32_t)(m & 1)*kSlotStride;
          s = (lvl == 0) ? stage_project<THREADS, true>(p, g, ctl.ep, img, I0, slot_base0 + so) : stage_project<THREADS, false>(p, g, ctl.ep, img, I0, slot_base + so);
        }
        return s;
      };
      SDV_PROF_T(tp2); SDV_PROF_ADD(1, tp1, tp2);
#ifdef SDV_PTS_LDG
      if (K > 0) sN = project(0);
#else
      if (K > 0) { mbar_wait(&full_bar[0], bar_phase & 1u); bar_phase ^= 1u; sN = project(0); }
#endif
      cp_async_commit();
      SDV_PROF_T(tp3); SDV_PROF_ADD(2, tp2, tp3);
      for (int k = 0; k < K; k++) {
        const PtState sC = sN;
        const int m = k + 1;
        if (m < K) {
#ifndef SDV_PTS_LDG
          if (m % kPtChunk == 0) {                          // entering chunk c: everyone is done reading chunk c-1, its buffer can take chunk c+1
            const int c = m / kPtChunk;
            __syncthreads();
            if (tid == 0 && c + 1 < NC) issue_chunk(c + 1);
            mbar_wait(&full_bar[c & 1], (bar_phase >> (c & 1)) & 1u); bar_phase ^= (1u << (c & 1));
          }
#endif
          sN = project(m);
        }
        cp_async_commit();
        cp_async_wait1();                                   // the gather of iteration k (committed one iteration ago) has landed in this thread's slot
        const unsigned char* slot = ring + (size_t)(k & 1)*kSlotStride;
        if (lvl == 0) stage_accumulate<THREADS, true>(sC, g, ctl.ep, slot + tid*4, acc);
        else          stage_accumulate<THREADS, false>(sC, g, ctl.ep, slot + tid*16, acc);
      }
      SDV_PROF_T(tp4); SDV_PROF_ADD(3, tp3, tp4);
      cp_async_wait0();
      __syncthreads();                                      // the reduction scratch aliases the tap ring / point chunks
      SDV_PROF_T(tp5); SDV_PROF_ADD(4, tp4, tp5);
      block_reduce_acc<THREADS>(acc, red, bsum);
      if (C > 1) {
        const int buf = evalCount & 1;
        for (int k = tid; k < kNAcc*C; k += THREADS) {
          int r = k / kNAcc, e = k - r*kNAcc;
          double* dst = cluster.map_shared_rank(gather, r);
          dst[(buf*C + rank)*kNAcc + e] = bsum[e];
        }
        cluster.sync();
        if (tid < kNAcc) { double s = 0; for (int r = 0; r < C; r++) s += gather[(buf*C + r)*kNAcc + tid]; tot[tid] = s; }
      } else {
        if (tid < kNAcc) tot[tid] = bsum[tid];
      }
      __syncthreads();
      evalCount++;
      n_eval = n;                                             // ctl.evals is updated by lane 0 inside the control block: a lane-0-only store HERE left warp 0 split 1 + 31 with
                                                            // nothing to reconverge it before the control step (ncu: every control instruction issued twice, 30 us per evaluation)
      SDV_PROF_T(tp6); SDV_PROF_ADD(5, tp5, tp6);
#ifdef SDV_TRACK_PROFILE
      if (blockIdx.x == 0 && tid == 0) g_track_prof[8] += 1;
#endif
    }
    SDV_PROF_T(tc0);
    // ------------------------------------------------------------------ LM control (warp 0), CoarseTracker.cpp:679-812 as a state machine
    // All 32 lanes execute the SAME instruction stream on the same values (SIMT: no extra cost) and only lane 0's stores to `ctl` are predicated.  A lane-0-only
    // branch in front of the warp-cooperative LDLT left the warp diverged at the shuffles (ncu: the solve executed twice per call, every __shfl_sync a rendezvous of
    // two groups): 30 us per evaluation, 0.45 ms per launch.
    if (tid < 32) {
      __syncwarp();                                          // converge warp 0 before anything warp-cooperative
      const bool L0 = (tid == 0);
      const float lambdaExtrapolationLimit = 0.001f;         // :680
      const int maxIt = (lvl == 0) ? 10 : ((lvl == 1) ? 20 : 50);   // maxIterations[] = {10,20,50,50,50}  (:679)
      const int mode = ctl.mode; float lcr = ctl.lcr; float lambda = ctl.lambda; int iteration = ctl.iteration;
      SE3d cur = ctl.cur; double a_cur = ctl.a_cur, b_cur = ctl.b_cur;
      double resOld[6];
#pragma unroll
      for (int i = 0; i < 6; i++) resOld[i] = ctl.resOld[i];
      bool propose = false, level_end = false, repeat_entry = false;
      __syncwarp();
      if (mode == 0) {                                       // level entry: resOld at the current pose (:690-702)
        finalize_res(tot, resOld);
        if (resOld[5] > 0.6 && lcr < 50) { lcr *= 2; repeat_entry = true; }
        else { finalize_gs_warp(tot, ctl.H, ctl.b, tid); lambda = 0.01f; iteration = 0; propose = true; }
      } else {                                               // candidate evaluated: accept / reject (:770-806)
        double resNew[6]; finalize_res(tot, resNew);
        const bool accept = (resNew[0]/resNew[1]) < (resOld[0]/resOld[1]);
        if (accept) {
          finalize_gs_warp(tot, ctl.H, ctl.b, tid);
#pragma unroll
          for (int i = 0; i < 6; i++) resOld[i] = resNew[i];
          cur = ctl.cand; a_cur = ctl.a_cand; b_cur = ctl.b_cand;
          lambda *= 0.5f;
        } else {
          lambda *= 4;
          if (lambda < lambdaExtrapolationLimit) lambda = lambdaExtrapolationLimit;
        }
        if (L0 && accept) ctl.accs[lvl]++;
        iteration++;
        if (!(ctl.incn > 1e-3) || iteration >= maxIt) level_end = true; else propose = true;
      }
      __syncwarp();                                          // ctl.H / ctl.b (lane 0's stores) before every lane reads them
      SDV_PROF_T(tq1); SDV_PROF_ADD(9, tc0, tq1);
      int next_lvl = lvl, next_mode = mode, done = 0, aborted = 0, haveRepeated = ctl.haveRepeated;
      double incn_out = ctl.incn; SE3d cand = cur; double a_cand = a_cur, b_cand = b_cur;
      bool new_ep = false; SE3d ep_pose = cur; double ep_a = a_cur, ep_b = b_cur; int ep_lvl = lvl; float ep_lcr = lcr;
      double lastRes_l = 0;
      if (repeat_entry) { new_ep = true; next_mode = 0; }
      if (level_end) {                                       // :808-822
        lastRes_l = sqrtf((float)(resOld[0]/resOld[1]));
        if (lastRes_l > 1.5*J.minRes[lvl]) { aborted = 1; done = 1; }
        else {
          int nl = lvl;
          if (lcr > 1 && !haveRepeated) { nl++; haveRepeated = 1; }
          nl--;
          if (nl < 0) done = 1;
          else { next_lvl = nl; next_mode = 0; lcr = 1.0f; new_ep = true; ep_lvl = nl; ep_lcr = 1.0f; }
        }
      }
      if (propose) {                                         // propose the LM step (:722-765) — warp-cooperative, every lane converged
        const int r = tid & 7;
        const bool fixA = tc.affineOptModeA < 0, fixB = tc.affineOptModeB < 0;
        const int nv = (!fixA && !fixB) ? 8 : ((fixA && fixB) ? 6 : 7);
        const bool stitch = fixA && !fixB;                   // fix a only: b takes slot 6 (:736-748)
        const int rr = (stitch && r == 6) ? 7 : r;
        double a[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int cc = (stitch && j == 6) ? 7 : j;
          double hv = ctl.H[rr*8 + cc];
          if (rr == cc) hv *= (1 + lambda);
          a[j] = (r < nv && j < nv) ? hv : ((r == j) ? 1.0 : 0.0);
        }
        double rhs = (r < nv) ? -ctl.b[rr] : 0.0;
        const double x = warp_ldlt_solve8(a, rhs);
        SDV_PROF_T(tq2); SDV_PROF_ADD(10, tq1, tq2);
        double inc[8];
#pragma unroll
        for (int j = 0; j < 8; j++) inc[j] = __shfl_sync(0xffffffffu, x, j, 8);
        if (fixA && fixB) { inc[6] = 0; inc[7] = 0; }
        else if (!fixA && fixB) { inc[7] = 0; }
        else if (stitch) { inc[7] = inc[6]; inc[6] = 0; }
        float extrapFac = 1;
        if (lambda < lambdaExtrapolationLimit) extrapFac = sqrtf(sqrtf(lambdaExtrapolationLimit / lambda));
#pragma unroll
        for (int i = 0; i < 8; i++) inc[i] *= extrapFac;
        double incScaled[8];
#pragma unroll
        for (int i = 0; i < 3; i++) incScaled[i] = inc[i]*1.0f;        // SCALE_XI_ROT   (:755)
#pragma unroll
        for (int i = 3; i < 6; i++) incScaled[i] = inc[i]*0.5f;        // SCALE_XI_TRANS (:756)
        incScaled[6] = inc[6]*10.0f; incScaled[7] = inc[7]*1000.0f;    // SCALE_A, SCALE_B
        double ssum = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) ssum += incScaled[i];
        if (!isfinite(ssum)) {
#pragma unroll
          for (int i = 0; i < 8; i++) incScaled[i] = 0;
        }
        double incn = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) incn += inc[i]*inc[i];
        incn_out = sqrt(incn);
        cand = se3_mul(se3_exp(incScaled), cur); a_cand = a_cur + incScaled[6]; b_cand = b_cur + incScaled[7];
        next_mode = 1; new_ep = true; ep_pose = cand; ep_a = a_cand; ep_b = b_cand;
        SDV_PROF_T(tq3); SDV_PROF_ADD(11, tq2, tq3);
      }
      SDV_PROF_T(tq4);
      EvalParams ep;
      if (new_ep) make_eval_params(ep_pose, ep_a, ep_b, J.refExposure, J.newExposure, J.ref_a, J.ref_b, tc.geom[ep_lvl], ep_lvl, tc.coarseCutoffTH*ep_lcr, tc.huberTH, ep);
      __syncwarp();                                          // every lane has read the old control state (racecheck: warp-level WAR on ctl.incn / ctl.haveRepeated otherwise)
      if (L0) {                                              // publish (plain predicated stores; nothing collective follows inside this block)
        ctl.cur = cur; ctl.a_cur = a_cur; ctl.b_cur = b_cur; ctl.cand = cand; ctl.a_cand = a_cand; ctl.b_cand = b_cand;
#pragma unroll
        for (int i = 0; i < 6; i++) ctl.resOld[i] = resOld[i];
        ctl.lambda = lambda; ctl.lcr = lcr; ctl.iteration = iteration; ctl.incn = incn_out; ctl.mode = next_mode; ctl.lvl = next_lvl;
        ctl.haveRepeated = haveRepeated; ctl.done = done; ctl.aborted = aborted;
        if (propose) ctl.iters[lvl]++;
        ctl.evals[lvl] += n_eval;
        if (level_end) { ctl.lastRes[lvl] = lastRes_l; ctl.flow[0] = resOld[2]; ctl.flow[1] = resOld[3]; ctl.flow[2] = resOld[4]; }
        if (new_ep) ctl.ep = ep;
      }
      SDV_PROF_T(tq5); SDV_PROF_ADD(12, tq4, tq5);
    }
    __syncthreads();
    SDV_PROF_T(tc1); SDV_PROF_ADD(6, tc0, tc1);
    if (ctl.done) break;
  }

  if (rank == 0 && tid == 0) {
    int good = 0;
    if (!ctl.aborted) {
      se3_to7(ctl.cur, J.T);
      double a_out = ctl.a_cur, b_out = ctl.b_cur;
      good = 1;
      if ((tc.affineOptModeA != 0 && (fabsf((float)a_out) > 1.2f)) || (tc.affineOptModeB != 0 && (fabsf((float)b_out) > 200))) good = 0;
      if (good) {
        double rel[2]; aff_from_to(J.refExposure, J.newExposure, J.ref_a, J.ref_b, a_out, b_out, rel);
        float r0 = (float)rel[0], r1 = (float)rel[1];
        if ((tc.affineOptModeA == 0 && (fabsf(logf(r0)) > 1.5f)) || (tc.affineOptModeB == 0 && (fabsf(r1) > 200))) good = 0;
        if (good) { if (tc.affineOptModeA < 0) a_out = 0; if (tc.affineOptModeB < 0) b_out = 0; }
      }
      J.ab[0] = a_out; J.ab[1] = b_out;
    }
    J.good = good;
    for (int i = 0; i < 5; i++) J.lastRes[i] = ctl.lastRes[i];
    for (int i = 0; i < 3; i++) J.flow[i] = ctl.flow[i];
    for (int i = 0; i < kLevels; i++) { J.point_evals[i] = ctl.evals[i]; J.iterations[i] = ctl.iters[i]; J.accepts[i] = ctl.accs[i]; }
  }
  if (C > 1) cluster.sync();                                // keep every CTA's shared memory alive until all remote writes/reads are done
}

template <int THREADS>
static size_t track_kernel_smem_v2(int cluster_size = kMaxCluster) {
  size_t stage = (size_t)2*4*THREADS*16 + (size_t)2*kPtChunk*THREADS*16, red = (size_t)kNAcc*THREADS*sizeof(float);
  return (stage > red ? stage : red) + (size_t)(2*cluster_size*kNAcc + 2*kNAcc)*sizeof(double);
}

template <int THREADS>
static size_t track_kernel_smem() { return (size_t)kNAcc*THREADS*sizeof(float) + (size_t)(2*kMaxCluster*kNAcc + 2*kNAcc)*sizeof(double); }

static bool track_use_v1() { static int v = -1; if (v < 0) { const char* e = getenv("SDV_TRACK_IMPL"); v = (e && e[0] == 'v' && e[1] == '1') ? 1 : 0; } return v == 1; }

static unsigned track_stagger_ns() { static long v = -1; if (v < 0) { const char* e = getenv("SDV_TRACK_STAGGER_NS"); v = e ? atol(e) : SDV_TRACK_STAGGER_DEFAULT_NS; if (v < 0) v = 0; } return (unsigned)v; }
template <typename Kern, typename... Extra>
static cudaError_t launch_track_kern(Kern kern, size_t smem, int threads, TrackJob* jobs_dev, int njobs, const TrackConst* tc_dev, int cluster_size, cudaStream_t st, Extra... extra) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(njobs*cluster_size)); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)cluster_size; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, jobs_dev, tc_dev, extra...);
}
template <int THREADS, int MINB>
static cudaError_t launch_track_t(TrackJob* jobs_dev, int njobs, const TrackConst* tc_dev, int cluster_size, cudaStream_t st) {
  if (track_use_v1()) return launch_track_kern(track_cluster_v1_kernel<THREADS, MINB>, track_kernel_smem<THREADS>(), THREADS, jobs_dev, njobs, tc_dev, cluster_size, st);
  // stagger only when the grid fills the co-resident slots of the chip (otherwise there is nothing to desynchronise)
  const unsigned stag = (cluster_size == 1 && njobs > 148) ? track_stagger_ns() : 0u;
  return launch_track_kern(track_cluster_kernel<THREADS, MINB>, track_kernel_smem_v2<THREADS>(cluster_size), THREADS, jobs_dev, njobs, tc_dev, cluster_size, st, stag);
}
template <typename Kern>
static cudaError_t track_attrs(Kern kern, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
}
// threads: 128 (throughput, 4 jobs resident per SM) or 256 (latency)
#ifndef SDV_TRACK_MINB
#define SDV_TRACK_MINB 4                    // v1 kernel: 3 (163 regs) -18 %, 5 (96 regs, spills) -16 %, 6 (80 regs) -45 % vs 4 (128 regs).  v2 kernel (gather buffers sized by the
                                            // cluster, so 5-6 CTAs fit in shared memory): 5 (96 regs, 236 B spilled) 0.506 at 1 480 jobs vs 4: 0.509 at 1 776 / 0.490 at 1 184; 6: 0.34
#endif
cudaError_t launch_track_cluster(TrackJob* jobs_dev, int njobs, const TrackConst* tc_dev, int cluster_size, int threads, cudaStream_t st) {
  if (threads == 256) return launch_track_t<256, 1>(jobs_dev, njobs, tc_dev, cluster_size, st);
  if (threads == 64)  return launch_track_t<64, 8>(jobs_dev, njobs, tc_dev, cluster_size, st);
  return launch_track_t<128, SDV_TRACK_MINB>(jobs_dev, njobs, tc_dev, cluster_size, st);
}

// Function attributes (opt-in dynamic shared memory, non-portable cluster sizes) belong to the CURRENT DEVICE's copy of a kernel: sdv_create calls this
// after cudaSetDevice, so a second context on another device of the same process gets them too (a process-wide "already set" flag did not).
cudaError_t kernels_init_device() {
  cudaError_t e = cudaFuncSetAttribute(coarse_res_gs_kernel<kStepThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kNAcc*kStepThreads*sizeof(float)));
  if (e != cudaSuccess) return e;
#define SDV_TRACK_ATTRS(T, M) \
  if ((e = track_attrs(track_cluster_v1_kernel<T, M>, track_kernel_smem<T>())) != cudaSuccess) return e; \
  if ((e = track_attrs(track_cluster_kernel<T, M>, track_kernel_smem_v2<T>())) != cudaSuccess) return e;
  SDV_TRACK_ATTRS(256, 1) SDV_TRACK_ATTRS(64, 8) SDV_TRACK_ATTRS(128, SDV_TRACK_MINB)
#undef SDV_TRACK_ATTRS
  return cudaSuccess;
}

#ifdef SDV_TRACK_PROFILE
extern "C" int sdv_debug_track_profile(long long* out16, int reset) {
  long long z[16] = {0};
  if (out16 && cudaMemcpyFromSymbol(out16, g_track_prof, sizeof(z)) != cudaSuccess) return -1;
  if (reset && cudaMemcpyToSymbol(g_track_prof, z, sizeof(z)) != cudaSuccess) return -1;
  return 0;
}
#endif
__global__ void h2d_words_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16) {
  for (size_t i = (size_t)blockIdx.x*blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x*blockDim.x) dst[i] = src[i];
}
void launch_h2d_words(void* dst_dev, const void* src_pinned, size_t bytes, cudaStream_t st) {      // bytes is rounded up to 16: both buffers are allocated with slack
  const size_t n16 = (bytes + 15)/16; if (!n16) return;
  int blocks = (int)((n16 + 255)/256); if (blocks > 296) blocks = 296;
  h2d_words_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<uint4*>(dst_dev), reinterpret_cast<const uint4*>(src_pinned), n16);
}

// ================================================================================================ makeCoarseDepthL0
// (a) splat in POINT ORDER per pixel: round r adds, for every pixel, the not-yet-added point of lowest index, which
//     reproduces the reference's sequential float `+=` order bit-for-bit even for colliding points (CoarseTracker.cpp:264-294).
__global__ void cd_owner_kernel(const float4* __restrict__ splats /*{pix as int bits, idepth*w, w, 0}*/, int n, const int* __restrict__ done, int* owner) {
  int i = blockIdx.x*blockDim.x + threadIdx.x; if (i >= n || done[i]) return;
  atomicMin(owner + __float_as_int(splats[i].x), i);
}
__global__ void cd_apply_kernel(const float4* __restrict__ splats, int n, int* done, int* owner, float* idepth, float* weightSums, int* remaining) {
  int i = blockIdx.x*blockDim.x + threadIdx.x; if (i >= n || done[i]) return;
  int pix = __float_as_int(splats[i].x);
  if (owner[pix] == i) { idepth[pix] += splats[i].y; weightSums[pix] += splats[i].z; done[i] = 1; }
  else atomicAdd(remaining, 1);
}
__global__ void cd_reset_owner_kernel(const float4* __restrict__ splats, int n, int* owner) {
  int i = blockIdx.x*blockDim.x + threadIdx.x; if (i >= n) return;
  owner[__float_as_int(splats[i].x)] = 0x7fffffff;
}
__global__ void cd_prep_kernel(const float* __restrict__ pts4, const int* __restrict__ round_half, int n, int w, float4* splats, int* done) {
  int i = blockIdx.x*blockDim.x + threadIdx.x; if (i >= n) return;
  float pu = pts4[4*i], pv = pts4[4*i+1], id = pts4[4*i+2], HdiF = pts4[4*i+3];
  int u = round_half[i] ? (int)(pu + 0.5f) : (int)pu;
  int v = round_half[i] ? (int)(pv + 0.5f) : (int)pv;
  float weight = sqrtf((float)(1e-3 / ((double)HdiF + 1e-12)));      // sqrtf(1e-3 / (HdiF+1e-12)) — double division, float sqrt (:273)
  splats[i] = make_float4(__int_as_float(u + w*v), id*weight, weight, 0.f);
  done[i] = 0;
}
// (b) sum-pool to the next level (:296-322)
__global__ void cd_pool_kernel(const float* __restrict__ id_lm, const float* __restrict__ ws_lm, float* id_l, float* ws_l, int wl, int hl, int wlm1) {
  int idx = blockIdx.x*blockDim.x + threadIdx.x; if (idx >= wl*hl) return;
  int y = idx / wl, x = idx - y*wl; int bidx = 2*x + 2*y*wlm1;
  id_l[idx] = ((id_lm[bidx] + id_lm[bidx+1]) + id_lm[bidx+wlm1]) + id_lm[bidx+wlm1+1];
  ws_l[idx] = ((ws_lm[bidx] + ws_lm[bidx+1]) + ws_lm[bidx+wlm1]) + ws_lm[bidx+wlm1+1];
}
// (c) 1-px dilation into empty cells, reading the pre-dilation weights (:324-375).  diag=1: 4 diagonal neighbours (levels 0,1),
//     diag=0: 4 axis neighbours.  The reference updates idepth in place but only ever reads cells with bak>0 and writes cells
//     with bak<=0, so reading the un-dilated idepth array is equivalent.
__global__ void cd_dilate_kernel(const float* __restrict__ id_in, const float* __restrict__ bak, float* id_out, float* ws_out, int w, int h, int diag) {
  int i = blockIdx.x*blockDim.x + threadIdx.x; int n = w*h; if (i >= n) return;
  float idv = id_in[i], wsv = bak[i];
  if (i >= w && i < n - w && bak[i] <= 0) {
    int o0, o1, o2, o3;
    if (diag) { o0 = 1+w; o1 = -1-w; o2 = w-1; o3 = -w+1; } else { o0 = 1; o1 = -1; o2 = w; o3 = -w; }
    float sum = 0, num = 0, numn = 0;
    // The reference reads one element past either end of the array for the first / last pixel of the loop (i = w: i-1-w = -1; i = n-w-1: i+1+w = n) —
    // undefined there; here such a neighbour counts as empty.  Both pixels lie in the 2 px border that the cloud emission drops, so no output depends on it.
#define CD_NB(o) { const int j = i + (o); if (j >= 0 && j < n && bak[j] > 0) { sum += id_in[j]; num += bak[j]; numn++; } }
    CD_NB(o0) CD_NB(o1) CD_NB(o2) CD_NB(o3)
#undef CD_NB
    if (numn > 0) { idv = sum/numn; wsv = num/numn; }
  }
  id_out[i] = idv; ws_out[i] = wsv;
}
// (d) normalise + raster-order compaction (:378-423): block counts -> exclusive scan -> scatter
constexpr int kScanThreads = 256, kScanItems = 4;            // 1024 pixels per block, contiguous per thread
__device__ __forceinline__ bool cd_valid(const float* id, const float* ws, const float4* ref, const float* ref0, int i, int w, int h, float& idn, float& col) {
  int y = i / w, x = i - y*w;
  if (x < 2 || x >= w-2 || y < 2 || y >= h-2) return false;
  if (!(ws[i] > 0)) return false;
  idn = id[i] / ws[i]; col = ref0 ? ref0[i] : ref[i].x;
  return isfinite(col) && (idn > 0);
}
__global__ void __launch_bounds__(kScanThreads) cd_count_kernel(const float* __restrict__ id, const float* __restrict__ ws, const float4* __restrict__ ref, const float* __restrict__ ref0, int w, int h, int* blockCounts) {
  __shared__ int s[kScanThreads/32];
  int base = (blockIdx.x*kScanThreads + threadIdx.x)*kScanItems; int c = 0, n = w*h;
  for (int j = 0; j < kScanItems; j++) { int i = base + j; float a, b; if (i < n && cd_valid(id, ws, ref, ref0, i, w, h, a, b)) c++; }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) { int t = 0; for (int k = 0; k < kScanThreads/32; k++) t += s[k]; blockCounts[blockIdx.x] = t; }
}
__global__ void cd_scan_kernel(int* blockCounts, int nblocks, int* total) {   // single block exclusive scan (nblocks <= a few thousand)
  __shared__ int carry;
  __shared__ int sh[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    int i = base + threadIdx.x; int v = (i < nblocks) ? blockCounts[i] : 0;
    sh[threadIdx.x] = v; __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) { int t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0; __syncthreads(); sh[threadIdx.x] += t; __syncthreads(); }
    int incl = sh[threadIdx.x]; int c = carry;
    if (i < nblocks) blockCounts[i] = c + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(kScanThreads) cd_emit_kernel(const float* __restrict__ id, const float* __restrict__ ws, const float4* __restrict__ ref, const float* __restrict__ ref0, int w, int h,
                                                              const int* __restrict__ blockOffsets, float4* out, int cap) {
  __shared__ int s[kScanThreads/32];
  int base = (blockIdx.x*kScanThreads + threadIdx.x)*kScanItems; int n = w*h;
  float idn[kScanItems], col[kScanItems]; bool ok[kScanItems]; int c = 0;
  for (int j = 0; j < kScanItems; j++) { int i = base + j; ok[j] = (i < n) && cd_valid(id, ws, ref, ref0, i, w, h, idn[j], col[j]); c += ok[j]; }
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5; int incl = c;
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) s[warp] = incl;
  __syncthreads();
  int woff = 0; for (int k = 0; k < warp; k++) woff += s[k];
  int pos = blockOffsets[blockIdx.x] + woff + incl - c;
  // cap = capacity of `out` (max_ref_points may make it smaller than the cloud): never write past it — the host compares the scanned total with cap afterwards
  for (int j = 0; j < kScanItems; j++) if (ok[j]) { int i = base + j; int y = i / w, x = i - y*w; if (pos < cap) out[pos] = make_float4((float)x, (float)y, idn[j], col[j]); pos++; }
}

int cd_num_blocks(int w, int h) { return (w*h + kScanThreads*kScanItems - 1)/(kScanThreads*kScanItems); }

void launch_cd_prep(const float* pts4, const int* round_half, int n, int w, float4* splats, int* done, cudaStream_t st) {
  if (n > 0) cd_prep_kernel<<<(n+255)/256, 256, 0, st>>>(pts4, round_half, n, w, splats, done);
}
void launch_cd_round(const float4* splats, int n, int* done, int* owner, float* idepth, float* ws, int* remaining, cudaStream_t st) {
  if (n <= 0) return;
  int g = (n+255)/256;
  cd_reset_owner_kernel<<<g, 256, 0, st>>>(splats, n, owner);
  cd_owner_kernel<<<g, 256, 0, st>>>(splats, n, done, owner);
  cd_apply_kernel<<<g, 256, 0, st>>>(splats, n, done, owner, idepth, ws, remaining);
}
void launch_cd_pool(const float* id_lm, const float* ws_lm, float* id_l, float* ws_l, int wl, int hl, int wlm1, cudaStream_t st) {
  cd_pool_kernel<<<(wl*hl+255)/256, 256, 0, st>>>(id_lm, ws_lm, id_l, ws_l, wl, hl, wlm1);
}
void launch_cd_dilate(const float* id_in, const float* bak, float* id_out, float* ws_out, int w, int h, int diag, cudaStream_t st) {
  cd_dilate_kernel<<<(w*h+255)/256, 256, 0, st>>>(id_in, bak, id_out, ws_out, w, h, diag);
}
void launch_cd_compact(const float* id, const float* ws, const float4* ref, const float* ref0, int w, int h, int* blockCounts, int* total, float4* out, int cap, cudaStream_t st) {
  int nb = cd_num_blocks(w, h);
  cd_count_kernel<<<nb, kScanThreads, 0, st>>>(id, ws, ref, ref0, w, h, blockCounts);
  cd_scan_kernel<<<1, 1024, 0, st>>>(blockCounts, nb, total);
  cd_emit_kernel<<<nb, kScanThreads, 0, st>>>(id, ws, ref, ref0, w, h, blockCounts, out, cap);
}

__global__ void pack_cloud_kernel(const float* u, const float* v, const float* id, const float* col, int n, float4* out) {
  int i = blockIdx.x*blockDim.x + threadIdx.x; if (i < n) out[i] = make_float4(u[i], v[i], id[i], col[i]);
}
void launch_pack_cloud(const float* u, const float* v, const float* id, const float* col, int n, float4* out, cudaStream_t st) {
  if (n > 0) pack_cloud_kernel<<<(n+255)/256, 256, 0, st>>>(u, v, id, col, n, out);
}

} // namespace sdv
