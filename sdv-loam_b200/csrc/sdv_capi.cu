// sdv_capi.cu — context + extern "C" boundary (include/sdv_b200.h).  Host logic only; all numerics run in the kernels of
// sdv_kernels.cu / sdv_ba_kernels.cu.  There is deliberately NO CPU fallback: every entry point fails with SDV_ERR_CUDA when
// the device path is unavailable.
#include "../../include/sdv_b200.h"
#include "sdv_ctx.cuh"
#include "sdv_refine.cuh"
#include <cstdio>
#include <cstring>
#include <cstdarg>

using namespace sdv;

namespace sdv {
int ctx_fail(sdv_ctx* c, int code, const char* fmt, ...) {
  if (c) { std::lock_guard<std::mutex> lk(c->mu_err); va_list ap; va_start(ap, fmt); vsnprintf(c->err, sizeof(c->err), fmt, ap); va_end(ap); }
  return code;
}
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ctx_fail(c, SDV_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)

extern "C" {

void sdv_default_settings(sdv_settings* s) {
  s->huberTH = 6; s->coarseCutoffTH = 20; s->affineOptModeA = 0; s->affineOptModeB = 0;
  s->outlierTH = 12*12; s->outlierTHSumComponent = 50*50; s->idepthFixPrior = 50*50;
  s->max_ref_points = 0; s->n_tracker_slots = 2; s->max_frames = 16; s->cluster_size = 1; s->track_threads = 128; s->max_kf_images = 0;
}

int sdv_pyr_levels(int w, int h) {               // util/globalCalib.cpp:22-30
  int wl = w, hl = h, used = 1;
  while (wl%2==0 && hl%2==0 && wl*hl > 5000 && used < SDV_PYR_LEVELS) { wl/=2; hl/=2; used++; }
  return used;
}

static thread_local char g_create_err[512] = "null context";   // message of the last failed sdv_create on this thread (there is no context to hold it)
const char* sdv_last_error(sdv_ctx* c) { return c ? c->err : g_create_err; }

static void make_geom(sdv_ctx* c, const sdv_calib* K) {   // CoarseTracker::makeK (CoarseTracker.cpp:77-106)
  float fx[SDV_PYR_LEVELS], fy[SDV_PYR_LEVELS], cx[SDV_PYR_LEVELS], cy[SDV_PYR_LEVELS];
  fx[0]=K->fx; fy[0]=K->fy; cx[0]=K->cx; cy[0]=K->cy;
  for (int l=1;l<c->levels;l++) {
    fx[l] = fx[l-1]*0.5; fy[l] = fy[l-1]*0.5;
    cx[l] = (cx[0]+0.5)/((int)1<<l) - 0.5; cy[l] = (cy[0]+0.5)/((int)1<<l) - 0.5;
  }
  for (int l=0;l<c->levels;l++) {
    LevelGeom& g = c->tc.geom[l]; g.w = c->w>>l; g.h = c->h>>l; g.fx=fx[l]; g.fy=fy[l]; g.cx=cx[l]; g.cy=cy[l];
    float Km[9] = {fx[l],0,cx[l], 0,fy[l],cy[l], 0,0,1}; inv3f(Km, g.Ki);
  }
}

// CoarseTracker::makeK(HCalib) / the CalibHessian the Reprojector reads (FullSystem::optimize moves the intrinsics at every keyframe): new K for every later call.
// Stream-ordered after the work already enqueued on the compute stream.
int sdv_set_calib(sdv_ctx* c, const sdv_calib* K) { SDV_GUARD_TRK(c);
  if (!c || !K) return SDV_ERR_ARG;
  if (!(K->fx > 0 && K->fy > 0)) return ctx_fail(c, SDV_ERR_ARG, "sdv_set_calib: focal lengths must be positive");
  CK(cudaSetDevice(c->device));
  make_geom(c, K);
  CK(cudaMemcpyAsync(c->tc_dev, &c->tc, sizeof(TrackConst), cudaMemcpyHostToDevice, c->st));   // pageable source: staged before the call returns
  rp_calib_changed(c);
  return SDV_OK;
}

static int create_impl(sdv_ctx* c, const sdv_calib* K, int w, int h, int levels, const sdv_settings* s_in, int device);
int sdv_create(const sdv_calib* K, int w, int h, int levels, const sdv_settings* s_in, int device, sdv_ctx** out) {
  if (out) *out = nullptr;
  if (!K || !out || w <= 0 || h <= 0 || levels < 1 || levels > SDV_PYR_LEVELS || device < 0) { snprintf(g_create_err, sizeof g_create_err, "sdv_create: bad argument"); return SDV_ERR_ARG; }
  int ndev = 0; if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) { snprintf(g_create_err, sizeof g_create_err, "sdv_create: no CUDA device %d (this library has no CPU fallback)", device); return SDV_ERR_CUDA; }
  sdv_ctx* c = new sdv_ctx();                              // value-initialised: every pointer/handle member starts null, so sdv_destroy is safe on a partially built context
  c->err[0] = 0;
  int rc = create_impl(c, K, w, h, levels, s_in, device);
  if (rc != SDV_OK) { snprintf(g_create_err, sizeof g_create_err, "%s", c->err[0] ? c->err : "sdv_create failed"); sdv_destroy(c); return rc; }
  *out = c; return SDV_OK;
}
static int create_impl(sdv_ctx* c, const sdv_calib* K, int w, int h, int levels, const sdv_settings* s_in, int device) {
  sdv_settings s; if (s_in) s = *s_in; else sdv_default_settings(&s);
  if (s.n_tracker_slots < 1) s.n_tracker_slots = 2;
  if (s.max_frames < 2) s.max_frames = 2;
  if (s.cluster_size <= 0) s.cluster_size = 1;
  if (s.track_threads != 64 && s.track_threads != 256) s.track_threads = 128;
  if (s.cluster_size > 16) s.cluster_size = 16;
  c->set = s; c->device = device; c->w = w; c->h = h; c->levels = levels;
  CK(cudaSetDevice(device));
  CK(kernels_init_device()); CK(refine_init_device());      // per-device function attributes (opt-in shared memory, cluster sizes)
  CK(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&c->st_ba, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&c->st_in, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&c->st_cp, cudaStreamNonBlocking));
  for (int i=0;i<sdv_ctx::kIngRing;i++) CK(cudaEventCreateWithFlags(&c->ev_ing[i], cudaEventDisableTiming)); for (int i=0;i<2;i++) CK(cudaEventCreateWithFlags(&c->ev_cp[i], cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming)); c->ingest_pending = false; c->launches = 0;
  for (int i=0;i<2;i++) { c->pyr_batch_dev[i] = nullptr; c->pyr_batch_host[i] = nullptr; }
  CK(cudaEventCreate(&c->ev0)); CK(cudaEventCreate(&c->ev1)); CK(cudaEventCreate(&c->ba_ev0)); CK(cudaEventCreate(&c->ba_ev1)); CK(cudaEventCreateWithFlags(&c->ev_xdom, cudaEventDisableTiming));
  memset(&c->tc, 0, sizeof(c->tc));
  c->tc.levels = levels; c->tc.huberTH = s.huberTH; c->tc.coarseCutoffTH = s.coarseCutoffTH;
  c->tc.affineOptModeA = s.affineOptModeA; c->tc.affineOptModeB = s.affineOptModeB;
  make_geom(c, K);
  CK(cudaMalloc(&c->tc_dev, sizeof(TrackConst)));
  CK(cudaMemcpy(c->tc_dev, &c->tc, sizeof(TrackConst), cudaMemcpyHostToDevice));
  // frame pool: level-0 intensity plane + packed texels of levels >= 1 per frame; packed level-0 texels come from a small keyframe pool
  size_t texels = 0; c->lvl_off[0] = 0; for (int l=1;l<levels;l++) { c->lvl_off[l] = texels; texels += (size_t)(w>>l)*(h>>l); }
  c->frame_texels = texels;
  c->frames.resize(s.max_frames);
  for (auto& f : c->frames) {
    f.used = false; f.adopted = false; f.lvl0_slot = -1; f.base = nullptr; f.ingest_seq = 0;
    CK(cudaMalloc(&f.I0_own, (size_t)w*h*sizeof(float))); f.I0 = f.I0_own;
    if (texels) CK(cudaMalloc(&f.base, texels*sizeof(float4)));
    f.lvl[0] = nullptr; for (int l=1;l<levels;l++) f.lvl[l] = f.base + c->lvl_off[l];
  }
  { int nkf = s.max_kf_images > 0 ? s.max_kf_images : SDV_MAX_FRAMES_WINDOW + 4; if (nkf > s.max_frames) nkf = s.max_frames;
    c->lvl0_pool.resize(nkf); for (int i=0;i<nkf;i++) { CK(cudaMalloc(&c->lvl0_pool[i], (size_t)w*h*sizeof(float4))); c->lvl0_free.push_back(nkf-1-i); } }
  // tracker slots
  c->slots.resize(s.n_tracker_slots);
  for (auto& t : c->slots) {
    for (int l=0;l<levels;l++) {
      size_t cap = (size_t)(w>>l)*(h>>l); if (s.max_ref_points > 0 && (size_t)s.max_ref_points < cap) cap = s.max_ref_points;
      t.cap[l] = (int)cap; t.npts[l] = 0; CK(cudaMalloc(&t.pts[l], cap*sizeof(float4)));
    }
    t.ref_frame = ~0ull; t.has_totals = false;
  }
  // step-kernel reduction buffers
  CK(cudaMalloc(&c->partials, (size_t)step_kernel_max_grid()*kNAcc*sizeof(double)));
  CK(cudaMalloc(&c->ticket, sizeof(unsigned int))); CK(cudaMemset(c->ticket, 0, sizeof(unsigned int)));
  CK(cudaMalloc(&c->totals_dev, kNAcc*sizeof(double)));
  CK(cudaMallocHost(&c->totals_host, kNAcc*sizeof(double)));
  // coarse-depth work buffers
  for (int l=0;l<levels;l++) {
    size_t n = (size_t)(w>>l)*(h>>l);
    CK(cudaMalloc(&c->cd_id[l], n*sizeof(float))); CK(cudaMalloc(&c->cd_ws[l], n*sizeof(float)));
    CK(cudaMalloc(&c->cd_id2[l], n*sizeof(float))); CK(cudaMalloc(&c->cd_ws2[l], n*sizeof(float)));
  }
  CK(cudaMalloc(&c->cd_owner, (size_t)w*h*sizeof(int)));
  CK(cudaMalloc(&c->cd_counts, (size_t)(cd_num_blocks(w,h)+1)*sizeof(int)));
  CK(cudaMalloc(&c->cd_scalars, 8*sizeof(int)));
  CK(cudaMallocHost(&c->cd_scalars_host, 8*sizeof(int)));
  c->cd_cap = 0; c->cd_pts4 = nullptr; c->cd_round = nullptr; c->cd_splats = nullptr; c->cd_done = nullptr;
  c->jobs_cap = 0; c->jobs_dev = nullptr; c->jobs_host = nullptr;
  c->stage_cap = 0; c->last_ms = 0;
  return SDV_OK;
}

void sdv_destroy(sdv_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->st_cp) cudaStreamSynchronize(c->st_cp); if (c->st_in) cudaStreamSynchronize(c->st_in); if (c->st) cudaStreamSynchronize(c->st);
  for (auto& f : c->frames) { cudaFree(f.base); cudaFree(f.I0_own); }
  for (auto p : c->lvl0_pool) cudaFree(p);
  for (auto& t : c->slots) for (int l=0;l<c->levels;l++) cudaFree(t.pts[l]);
  for (int l=0;l<c->levels;l++) { cudaFree(c->cd_id[l]); cudaFree(c->cd_ws[l]); cudaFree(c->cd_id2[l]); cudaFree(c->cd_ws2[l]); }
  cudaFree(c->cd_owner); cudaFree(c->cd_counts); cudaFree(c->cd_scalars); cudaFreeHost(c->cd_scalars_host);
  cudaFree(c->cd_pts4); cudaFree(c->cd_round); cudaFree(c->cd_splats); cudaFree(c->cd_done);
  for (int i=0;i<2;i++) { cudaFree(c->pyr_batch_dev[i]); cudaFreeHost(c->pyr_batch_host[i]); } cudaFree(c->partials); cudaFree(c->ticket); cudaFree(c->totals_dev); cudaFreeHost(c->totals_host);
  cudaFree(c->tc_dev); cudaFree(c->jobs_dev); cudaFreeHost(c->jobs_host);
  for (int i=0;i<2;i++) { for (auto p : c->stage[i]) cudaFree(p); cudaFree(c->stage_u8[i]); }
  cudaFree(c->und_buf); cudaFree(c->trace_dev);
  for (int i=0;i<sdv_ctx::kIngRing;i++) if (c->ev_ing[i]) cudaEventDestroy(c->ev_ing[i]); for (int i=0;i<2;i++) if (c->ev_cp[i]) cudaEventDestroy(c->ev_cp[i]); if (c->st_cp) cudaStreamDestroy(c->st_cp);
  cudaFree(c->refine_dev); cudaFreeHost(c->refine_host);
  rp_destroy(c);
  sel_destroy(c);
  lidar_destroy(c);
  ba_destroy(c);
  if (c->ba_ev0) cudaEventDestroy(c->ba_ev0); if (c->ba_ev1) cudaEventDestroy(c->ba_ev1); if (c->ev_xdom) cudaEventDestroy(c->ev_xdom); if (c->st_ba) cudaStreamDestroy(c->st_ba);
  if (c->ev0) cudaEventDestroy(c->ev0); if (c->ev1) cudaEventDestroy(c->ev1); if (c->ev_in) cudaEventDestroy(c->ev_in); if (c->st) cudaStreamDestroy(c->st); if (c->st_in) cudaStreamDestroy(c->st_in);
  cudaGetLastError();                                      // a partially built context may have produced benign errors above: do not leave them sticky
  delete c;
}

int sdv_sync(sdv_ctx* c) { SDV_GUARD_TRK(c); if (!c) return SDV_ERR_ARG; CK(cudaSetDevice(c->device)); CK(cudaStreamSynchronize(c->st_cp)); CK(cudaStreamSynchronize(c->st_in)); CK(cudaStreamSynchronize(c->st)); CK(cudaStreamSynchronize(c->st_ba)); return SDV_OK; }
long long sdv_launch_count(sdv_ctx* c) { return c ? c->launches.load() : 0LL; }
int sdv_track_job_bytes(void) { return (int)sizeof(TrackJob); }
int sdv_debug_gs_entry_rc(int k) { int r, c; if (k < 0 || k >= kNH) return -1; gs_entry_rc(k, r, c); return r*16 + c; }   // test hook (host evaluation of the device index walk)
float sdv_last_kernel_ms(sdv_ctx* c) { return c ? c->last_ms : 0.f; }

// ------------------------------------------------------------------------------------------------ frames
static FrameDev* find_frame(sdv_ctx* c, uint64_t id) { auto it = c->frame_index.find(id); return it == c->frame_index.end() ? nullptr : &c->frames[it->second]; }

static int ensure_stage(sdv_ctx* c, int n, int par) {
  size_t fl = (size_t)c->w*c->h + pyramid_scratch_floats(c->w, c->h, c->levels);
  while ((int)c->stage[par].size() < n) { float* p = nullptr; CK(cudaMalloc(&p, fl*sizeof(float))); c->stage[par].push_back(p); }
  if (n > c->stage_cap) {
    CK(cudaStreamSynchronize(c->st_in));
    for (int i=0;i<2;i++) { cudaFree(c->pyr_batch_dev[i]); cudaFreeHost(c->pyr_batch_host[i]);
      CK(cudaMalloc(&c->pyr_batch_dev[i], (size_t)n*sizeof(PyrBatchHost) + 16)); CK(cudaMallocHost(&c->pyr_batch_host[i], (size_t)n*sizeof(PyrBatchHost) + 16)); }
    c->stage_cap = n;
  }
  return SDV_OK;
}

static void frame_drop_lvl0(sdv_ctx* c, FrameDev& f) { if (f.lvl0_slot >= 0) { c->lvl0_free.push_back(f.lvl0_slot); f.lvl0_slot = -1; } f.lvl[0] = nullptr; }

// kind bit0: mono8 input, bit1: input already in device memory, bit2: adopt the (device, float) buffer as the frame's level-0 plane,
// bit3: RAW mono8 input of the camera's native size, rectified through the tables of sdv_set_undistort inside the level-0 kernel.
// Everything is enqueued on the ingest stream; the compute stream picks it up through ev_in at the next tracker / BA call, so an
// upload overlaps the tracking of the previous batch.  Level 0 stays a planar intensity image: a pinned-host float upload lands
// directly in frame storage and only levels >= 1 are built (gradients of level 0 are formed on the fly by the consumers).
static int frame_ingest(sdv_ctx* c, int n, const uint64_t* frames, const void* const* imgs, const float* exposures, int kind) {
  if (!c || n < 0 || !frames || !imgs) return SDV_ERR_ARG;
  if (n == 0) return SDV_OK;
  CK(cudaSetDevice(c->device));
  const long long seq = c->ingest_seq + 1; const int par = (int)(seq & 1);
  // staging buffers + descriptors of this parity were last used by ingest seq-2: it must have drained before the host rewrites the pinned
  // descriptors (the device-side order copy(seq) after pyramid(seq-2) is enforced on the copy stream below)
  if (seq > 2) CK(cudaEventSynchronize(c->ev_ing[(seq-2) % sdv_ctx::kIngRing]));
  int rc = ensure_stage(c, n, par); if (rc) return rc;
  const bool raw = (kind & 8), u8 = (kind & 1) || raw, dev = (kind & 2), adopt = (kind & 4) && dev && !u8;
  if (raw && !c->has_und) return ctx_fail(c, SDV_ERR_STATE, "raw upload before sdv_set_undistort");
  const size_t px = (size_t)c->w*c->h, px_in = raw ? (size_t)c->und.wOrg*c->und.hOrg : px;     // bytes of one mono8 input image
  if (u8 && !dev && (size_t)n*px_in > c->stage_u8_cap[par]) { CK(cudaStreamSynchronize(c->st_in)); cudaFree(c->stage_u8[par]); c->stage_u8[par] = nullptr;
    c->stage_u8_cap[par] = (size_t)n*px_in; CK(cudaMalloc(&c->stage_u8[par], c->stage_u8_cap[par])); }
  c->cp_dst.clear(); c->cp_src.clear(); c->cp_sz.clear();
  if (c->levels > 1 && ((c->w | c->h) & 1)) return ctx_fail(c, SDV_ERR_ARG, "pyramid needs even image sizes");
  PyrBatchHost* desc = c->pyr_batch_host[par];
  for (int k=0;k<n;k++) {
    int idx = -1;
    auto it = c->frame_index.find(frames[k]);
    if (it != c->frame_index.end()) { idx = it->second;
      if (frame_pinned(c, frames[k])) return ctx_fail(c, SDV_ERR_STATE, "frame %llu is referenced by a resident BA window / map slot and cannot be re-uploaded", (unsigned long long)frames[k]); }
    else { for (size_t i=0;i<c->frames.size();i++) if (!c->frames[i].used) { idx = (int)i; break; } }
    if (idx < 0) return ctx_fail(c, SDV_ERR_CAPACITY, "frame pool exhausted (max_frames=%d)", (int)c->frames.size());
    FrameDev& f = c->frames[idx]; f.used = true; f.id = frames[k]; f.exposure = exposures ? exposures[k] : 1.0f; c->frame_index[frames[k]] = idx;
    frame_drop_lvl0(c, f); f.ingest_seq = seq;
    f.adopted = adopt; f.I0 = adopt ? const_cast<float*>(reinterpret_cast<const float*>(imgs[k])) : f.I0_own;
    PyrBatchHost& b = desc[k];
    b.scratch = c->stage[par][k] + px; b.out = f.base; b.I0 = f.I0; b.flags = (f.exposure > 0.f) ? 1 : 0; b.pad = 0;
    if (dev) b.src = imgs[k];
    else if (u8) { unsigned char* d8 = c->stage_u8[par] + (size_t)k*px_in; b.src = d8;
      if (!c->cp_dst.empty() && (unsigned char*)c->cp_src.back() + c->cp_sz.back() == (const unsigned char*)imgs[k] && (unsigned char*)c->cp_dst.back() + c->cp_sz.back() == d8) c->cp_sz.back() += px_in;   // adjacent in host memory: one copy
      else { c->cp_dst.push_back(d8); c->cp_src.push_back(const_cast<void*>(imgs[k])); c->cp_sz.push_back(px_in); } }
    else { b.src = f.I0; c->cp_dst.push_back(f.I0); c->cp_src.push_back(const_cast<void*>(imgs[k])); c->cp_sz.push_back(px*sizeof(float)); }
  }
  if (!c->cp_dst.empty()) {                                               // all H2D copies of the batch in ONE runtime call, on the copy stream: PCIe stays busy while the
    if (seq > 2) CK(cudaStreamWaitEvent(c->st_cp, c->ev_ing[(seq-2) % sdv_ctx::kIngRing], 0));   // previous batch's pyramid runs on st_in
    bool done = false;
    if (c->cp_dst.size() > 1 && !c->no_batch_copy) {
      cudaMemcpyAttributes at; memset(&at, 0, sizeof(at)); at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
      size_t aidx = 0, fail = 0;
      cudaError_t e = cudaMemcpyBatchAsync(c->cp_dst.data(), c->cp_src.data(), c->cp_sz.data(), c->cp_dst.size(), &at, &aidx, 1, &fail, c->st_cp);
      if (e == cudaSuccess) done = true; else { cudaGetLastError(); c->no_batch_copy = true; }
    }
    if (!done) for (size_t i=0;i<c->cp_dst.size();i++) CK(cudaMemcpyAsync(c->cp_dst[i], c->cp_src[i], c->cp_sz[i], cudaMemcpyHostToDevice, c->st_cp));
    c->cp_dst.clear(); c->cp_src.clear(); c->cp_sz.clear();
    CK(cudaEventRecord(c->ev_cp[par], c->st_cp)); CK(cudaStreamWaitEvent(c->st_in, c->ev_cp[par], 0));
  }
  launch_h2d_words(c->pyr_batch_dev[par], desc, (size_t)n*sizeof(PyrBatchHost), c->st_in);   // kernel copy: a cudaMemcpyAsync here becomes ready only after this batch's bulk copy and
  c->launches += 1;                                                                              // would queue behind the NEXT batch's bulk copy on the H2D engine
  const UndistortDev* und = raw ? &c->und : nullptr;
  if (c->levels > 1) { launch_pyramid_batch(c->pyr_batch_dev[par], n, u8, c->lvl_off, c->w, c->h, c->levels, c->st_in, und); c->launches += 2*(c->levels - 1); }
  else if (u8 || (dev && !adopt)) { launch_pyramid_copy0(c->pyr_batch_dev[par], n, u8, c->w, c->h, c->st_in, und); c->launches += 1; }
  CK(cudaGetLastError());
  CK(cudaEventRecord(c->ev_ing[seq % sdv_ctx::kIngRing], c->st_in)); c->ingest_seq = seq; c->ingest_pending = true;
  return SDV_OK;
}
}  // extern "C"
namespace sdv {
int join_ingest_upto(sdv_ctx* c, long long seq) {  // the compute stream waits for ingest calls <= seq (st_in is in order, so one event covers all earlier ones)
  if (seq > c->ingest_seq) seq = c->ingest_seq;
  if (seq <= c->seq_waited) return SDV_OK;
  if (seq > c->ingest_seq - sdv_ctx::kIngRing) CK(cudaStreamWaitEvent(c->st, c->ev_ing[seq % sdv_ctx::kIngRing], 0));   // older ones have drained (frame_ingest syncs on seq-2)
  c->seq_waited = seq;
  return SDV_OK;
}
int join_ingest(sdv_ctx* c) { c->ingest_pending = false; return join_ingest_upto(c, c->ingest_seq); }
int ensure_lvl0(sdv_ctx* c, FrameDev& f) {        // FrameHessian::dI of a keyframe: packed level-0 texels, built once on demand
  if (f.lvl[0]) return SDV_OK;
  if (c->lvl0_free.empty()) return ctx_fail(c, SDV_ERR_CAPACITY, "keyframe level-0 image pool exhausted (max_kf_images=%d)", (int)c->lvl0_pool.size());
  int rc = join_ingest(c); if (rc) return rc;
  f.lvl0_slot = c->lvl0_free.back(); c->lvl0_free.pop_back(); f.lvl[0] = c->lvl0_pool[f.lvl0_slot];
  launch_pyramid_level0_texels(f.I0, f.lvl[0], c->w, c->h, c->st); c->launches += 1;
  CK(cudaGetLastError());
  return SDV_OK;
}
}
extern "C" {

int sdv_frame_upload_batch(sdv_ctx* c, int n, const uint64_t* frames, const float* const* imgs, const float* exposures) { SDV_GUARD_TRK(c);
  return frame_ingest(c, n, frames, reinterpret_cast<const void* const*>(imgs), exposures, 0);
}
int sdv_frame_upload_batch_u8(sdv_ctx* c, int n, const uint64_t* frames, const uint8_t* const* imgs, const float* exposures) { SDV_GUARD_TRK(c);
  return frame_ingest(c, n, frames, reinterpret_cast<const void* const*>(imgs), exposures, 1);
}
// Undistort (util/Undistort.cpp) as data: the tables the reference builds once per calibration file, kept on the device for the raw-image ingest
int sdv_set_undistort(sdv_ctx* c, int w_org, int h_org, const float* remapX, const float* remapY, float factor, const float* G256, const float* vignette_inv) { SDV_GUARD_TRK(c);
  if (!c || w_org < 2 || h_org < 2 || !remapX || !remapY) return SDV_ERR_ARG;
  if (vignette_inv && !G256) return ctx_fail(c, SDV_ERR_ARG, "a vignette map needs a response function (PhotometricUndistorter::processFrame applies it to G[v] only)");
  CK(cudaSetDevice(c->device));
  const size_t px = (size_t)c->w*c->h, po = (size_t)w_org*h_org;
  for (size_t i=0;i<px;i++) {                                              // Undistort.cpp:871 tests iy against wOrg-1: a table built that way may read below the raw image
    const float x = remapX[i], y = remapY[i];
    if (x < 0) continue;
    if (!(x > 0 && y > 0 && x < w_org-1 && y < h_org-1)) return ctx_fail(c, SDV_ERR_ARG, "remap table entry %zu (%g,%g) reads outside the %dx%d raw image", i, x, y, w_org, h_org);
  }
  CK(cudaStreamSynchronize(c->st_in));                                     // an ingest in flight may still read the previous tables
  cudaFree(c->und_buf); c->und_buf = nullptr; c->has_und = false;
  const size_t nfl = 2*px + (G256 ? 256 : 0) + (vignette_inv ? po : 0);
  CK(cudaMalloc(&c->und_buf, nfl*sizeof(float)));
  float* p = c->und_buf;
  CK(cudaMemcpy(p, remapX, px*sizeof(float), cudaMemcpyHostToDevice)); c->und.remapX = p; p += px;
  CK(cudaMemcpy(p, remapY, px*sizeof(float), cudaMemcpyHostToDevice)); c->und.remapY = p; p += px;
  c->und.G = nullptr; c->und.vignette = nullptr;
  if (G256) { CK(cudaMemcpy(p, G256, 256*sizeof(float), cudaMemcpyHostToDevice)); c->und.G = p; p += 256; }
  if (vignette_inv) { CK(cudaMemcpy(p, vignette_inv, po*sizeof(float), cudaMemcpyHostToDevice)); c->und.vignette = p; p += po; }
  c->und.wOrg = w_org; c->und.hOrg = h_org; c->und.factor = factor; c->has_und = true;
  return SDV_OK;
}
int sdv_frame_upload_batch_raw_u8(sdv_ctx* c, int n, const uint64_t* frames, const uint8_t* const* raw_imgs, const float* exposures) { SDV_GUARD_TRK(c);
  return frame_ingest(c, n, frames, reinterpret_cast<const void* const*>(raw_imgs), exposures, 8);
}
int sdv_frame_build_batch_dev(sdv_ctx* c, int n, const uint64_t* frames, const void* const* imgs_dev, int fmt, const float* exposures) { SDV_GUARD_TRK(c);
  if (fmt < 0 || fmt > 2) return SDV_ERR_ARG;
  return frame_ingest(c, n, frames, imgs_dev, exposures, 2 | (fmt == 1 ? 1 : 0) | (fmt == 2 ? 4 : 0));
}
int sdv_frame_upload(sdv_ctx* c, uint64_t frame, const float* img, float exposure) { SDV_GUARD_TRK(c);
  const float* imgs[1] = {img}; return sdv_frame_upload_batch(c, 1, &frame, imgs, &exposure);
}
int sdv_frame_release(sdv_ctx* c, uint64_t frame) { SDV_GUARD_TRK(c);
  if (!c) return SDV_ERR_ARG;
  auto it = c->frame_index.find(frame); if (it == c->frame_index.end()) return ctx_fail(c, SDV_ERR_NOFRAME, "unknown frame %llu", (unsigned long long)frame);
  if (frame_pinned(c, frame)) return ctx_fail(c, SDV_ERR_STATE, "frame %llu is referenced by a resident BA window / map slot: replace or clear that window / map first", (unsigned long long)frame);
  FrameDev& f = c->frames[it->second]; f.used = false; frame_drop_lvl0(c, f); f.adopted = false; f.I0 = f.I0_own; c->frame_index.erase(it); return SDV_OK;
}
int sdv_frame_download(sdv_ctx* c, uint64_t frame, int lvl, float* dI3_out, float* abs_out) { SDV_GUARD_TRK(c);
  if (!c || lvl < 0 || lvl >= c->levels) return SDV_ERR_ARG;
  CK(cudaSetDevice(c->device));
  FrameDev* f = find_frame(c, frame); if (!f) return ctx_fail(c, SDV_ERR_NOFRAME, "unknown frame %llu", (unsigned long long)frame);
  { int rcj = join_ingest(c); if (rcj) return rcj; }
  if (lvl == 0) { int rc0 = ensure_lvl0(c, *f); if (rc0) return rc0; }
  int n = (c->w>>lvl)*(c->h>>lvl); float *d3 = nullptr, *da = nullptr;
  if (dI3_out) CK(cudaMalloc(&d3, (size_t)3*n*sizeof(float)));
  if (abs_out) CK(cudaMalloc(&da, (size_t)n*sizeof(float)));
  launch_unpack_level(f->lvl[lvl], d3, da, n, c->st);
  if (d3) CK(cudaMemcpyAsync(dI3_out, d3, (size_t)3*n*sizeof(float), cudaMemcpyDeviceToHost, c->st));
  if (da) CK(cudaMemcpyAsync(abs_out, da, (size_t)n*sizeof(float), cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  cudaFree(d3); cudaFree(da);
  return SDV_OK;
}

// ------------------------------------------------------------------------------------------------ tracker reference
int sdv_tracker_set_cloud(sdv_ctx* c, int slot, uint64_t ref_frame, int lvl, int n, const float* u, const float* v,
                          const float* idepth, const float* color, double ref_a, double ref_b) { SDV_GUARD_TRK(c);
  if (!c || slot < 0 || slot >= (int)c->slots.size() || lvl < 0 || lvl >= c->levels || n < 0) return SDV_ERR_ARG;
  CK(cudaSetDevice(c->device));
  TrackerSlot& t = c->slots[slot];
  if (n > t.cap[lvl]) return ctx_fail(c, SDV_ERR_CAPACITY, "cloud of %d points exceeds capacity %d", n, t.cap[lvl]);
  FrameDev* f = find_frame(c, ref_frame); if (!f) return ctx_fail(c, SDV_ERR_NOFRAME, "unknown ref frame");
  { int rcj = join_ingest(c); if (rcj) return rcj; }
  float* tmp = nullptr; CK(cudaMalloc(&tmp, (size_t)4*(n+1)*sizeof(float)));
  const float* src[4] = {u, v, idepth, color};
  for (int k=0;k<4;k++) CK(cudaMemcpyAsync(tmp + (size_t)k*n, src[k], (size_t)n*sizeof(float), cudaMemcpyHostToDevice, c->st));
  launch_pack_cloud(tmp, tmp+n, tmp+2*(size_t)n, tmp+3*(size_t)n, n, t.pts[lvl], c->st);
  CK(cudaStreamSynchronize(c->st)); cudaFree(tmp);
  t.npts[lvl] = n; t.ref_frame = ref_frame; t.ref_a = ref_a; t.ref_b = ref_b; t.refExposure = f->exposure; t.has_totals = false;
  return SDV_OK;
}

int sdv_tracker_get_cloud(sdv_ctx* c, int slot, int lvl, int* n_out, float* u, float* v, float* idepth, float* color) { SDV_GUARD_TRK(c);
  if (!c || slot < 0 || slot >= (int)c->slots.size() || lvl < 0 || lvl >= c->levels || !n_out) return SDV_ERR_ARG;
  CK(cudaSetDevice(c->device));
  TrackerSlot& t = c->slots[slot]; int n = t.npts[lvl]; *n_out = n;
  if (!u || n == 0) return SDV_OK;
  std::vector<float4> h(n);
  CK(cudaMemcpyAsync(h.data(), t.pts[lvl], (size_t)n*sizeof(float4), cudaMemcpyDeviceToHost, c->st)); CK(cudaStreamSynchronize(c->st));
  for (int i=0;i<n;i++) { u[i]=h[i].x; v[i]=h[i].y; idepth[i]=h[i].z; color[i]=h[i].w; }
  return SDV_OK;
}

int sdv_tracker_set_ref(sdv_ctx* c, int slot, uint64_t ref_frame, int n, const float* pts4, const int32_t* round_half,
                        float /*unused*/, double ref_a, double ref_b) { SDV_GUARD_TRK(c);
  if (!c || slot < 0 || slot >= (int)c->slots.size() || n < 0 || (n > 0 && (!pts4 || !round_half))) return SDV_ERR_ARG;
  CK(cudaSetDevice(c->device));
  TrackerSlot& t = c->slots[slot];
  FrameDev* f = find_frame(c, ref_frame); if (!f) return ctx_fail(c, SDV_ERR_NOFRAME, "unknown ref frame");
  for (int i=0;i<n;i++) {                                   // the reference would write out of bounds here; we refuse instead
    int u = round_half[i] ? (int)(pts4[4*i]+0.5f) : (int)pts4[4*i], v = round_half[i] ? (int)(pts4[4*i+1]+0.5f) : (int)pts4[4*i+1];
    if (u < 0 || v < 0 || u >= c->w || v >= c->h) return ctx_fail(c, SDV_ERR_ARG, "splat %d (%d,%d) outside the image", i, u, v);
  }
  if (n > c->cd_cap) {
    cudaFree(c->cd_pts4); cudaFree(c->cd_round); cudaFree(c->cd_splats); cudaFree(c->cd_done);
    c->cd_cap = n + n/2 + 1024;
    CK(cudaMalloc(&c->cd_pts4, (size_t)4*c->cd_cap*sizeof(float))); CK(cudaMalloc(&c->cd_round, (size_t)c->cd_cap*sizeof(int)));
    CK(cudaMalloc(&c->cd_splats, (size_t)c->cd_cap*sizeof(float4))); CK(cudaMalloc(&c->cd_done, (size_t)c->cd_cap*sizeof(int)));
  }
  { int rcj = join_ingest(c); if (rcj) return rcj; }
  const int w = c->w, h = c->h;
  CK(cudaMemsetAsync(c->cd_id[0], 0, (size_t)w*h*sizeof(float), c->st));
  CK(cudaMemsetAsync(c->cd_ws[0], 0, (size_t)w*h*sizeof(float), c->st));
  if (n > 0) {
    CK(cudaMemcpyAsync(c->cd_pts4, pts4, (size_t)4*n*sizeof(float), cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(c->cd_round, round_half, (size_t)n*sizeof(int), cudaMemcpyHostToDevice, c->st));
    launch_cd_prep(c->cd_pts4, c->cd_round, n, w, c->cd_splats, c->cd_done, c->st);
    for (int round = 0; round < n; round++) {               // usually 1-3 rounds: one per multiplicity of colliding splats; at most n (every splat on one pixel)
      CK(cudaMemsetAsync(c->cd_scalars, 0, sizeof(int), c->st));
      launch_cd_round(c->cd_splats, n, c->cd_done, c->cd_owner, c->cd_id[0], c->cd_ws[0], c->cd_scalars, c->st);
      CK(cudaMemcpyAsync(c->cd_scalars_host, c->cd_scalars, sizeof(int), cudaMemcpyDeviceToHost, c->st));
      CK(cudaStreamSynchronize(c->st));
      if (c->cd_scalars_host[0] == 0) break;
    }
  }
  for (int l=1;l<c->levels;l++) launch_cd_pool(c->cd_id[l-1], c->cd_ws[l-1], c->cd_id[l], c->cd_ws[l], w>>l, h>>l, w>>(l-1), c->st);
  for (int l=0;l<c->levels;l++) {
    launch_cd_dilate(c->cd_id[l], c->cd_ws[l], c->cd_id2[l], c->cd_ws2[l], w>>l, h>>l, l < 2 ? 1 : 0, c->st);
    launch_cd_compact(c->cd_id2[l], c->cd_ws2[l], l == 0 ? nullptr : f->lvl[l], l == 0 ? f->I0 : nullptr, w>>l, h>>l, c->cd_counts, c->cd_scalars + 1 + l, t.pts[l], t.cap[l], c->st);
  }
  CK(cudaMemcpyAsync(c->cd_scalars_host, c->cd_scalars, 8*sizeof(int), cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  CK(cudaGetLastError());
  for (int l=0;l<c->levels;l++) if (c->cd_scalars_host[1+l] > t.cap[l]) {
    // the emit kernel never wrote past cap[l], but the slot's clouds are now a truncated mix: invalidate the slot instead of leaving it half-updated
    t.ref_frame = ~0ull; t.has_totals = false; for (int k=0;k<c->levels;k++) t.npts[k] = 0;
    return ctx_fail(c, SDV_ERR_CAPACITY, "level %d cloud (%d) exceeds capacity %d (max_ref_points); slot %d has no reference now", l, c->cd_scalars_host[1+l], t.cap[l], slot);
  }
  for (int l=0;l<c->levels;l++) t.npts[l] = c->cd_scalars_host[1+l];
  t.ref_frame = ref_frame; t.ref_a = ref_a; t.ref_b = ref_b; t.refExposure = f->exposure; t.has_totals = false;
  return SDV_OK;
}

// ------------------------------------------------------------------------------------------------ calcRes / calcGSSSE
int sdv_tracker_calc_res(sdv_ctx* c, int slot, uint64_t new_frame, int lvl, const double T[7], double a, double b, float cutoffTH, double rs_out[6]) { SDV_GUARD_TRK(c);
  if (!c || slot < 0 || slot >= (int)c->slots.size() || lvl < 0 || lvl >= c->levels || !T || !rs_out) return SDV_ERR_ARG;
  CK(cudaSetDevice(c->device));
  TrackerSlot& t = c->slots[slot];
  FrameDev* f = find_frame(c, new_frame); if (!f) return ctx_fail(c, SDV_ERR_NOFRAME, "unknown new frame");
  if (t.ref_frame == ~0ull) return ctx_fail(c, SDV_ERR_STATE, "tracker slot %d has no reference", slot);
  EvalParams ep; make_eval_params(se3_from7(T), a, b, t.refExposure, f->exposure, t.ref_a, t.ref_b, c->tc.geom[lvl], lvl, cutoffTH, c->tc.huberTH, ep);
  { int rcj = join_ingest(c); if (rcj) return rcj; }
  CK(cudaEventRecord(c->ev0, c->st));
  c->launches += 1;
  launch_coarse_res_gs(t.pts[lvl], t.npts[lvl], lvl == 0 ? nullptr : f->lvl[lvl], lvl == 0 ? f->I0 : nullptr, c->tc.geom[lvl], ep, c->partials, c->ticket, c->totals_dev, c->st);
  CK(cudaEventRecord(c->ev1, c->st));
  CK(cudaMemcpyAsync(c->totals_host, c->totals_dev, kNAcc*sizeof(double), cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st)); CK(cudaGetLastError());
  CK(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  memcpy(t.totals, c->totals_host, sizeof(t.totals)); t.has_totals = true;
  finalize_res(t.totals, rs_out);
  return SDV_OK;
}
int sdv_tracker_calc_gs(sdv_ctx* c, int slot, int /*lvl*/, double H[64], double b[8]) { SDV_GUARD_TRK(c);
  if (!c || slot < 0 || slot >= (int)c->slots.size() || !H || !b) return SDV_ERR_ARG;
  TrackerSlot& t = c->slots[slot];
  if (!t.has_totals) return ctx_fail(c, SDV_ERR_STATE, "calc_gs before calc_res on slot %d", slot);
  finalize_gs(t.totals, H, b);
  return SDV_OK;
}

// ------------------------------------------------------------------------------------------------ trackNewestCoarse
int sdv_tracker_track_batch(sdv_ctx* c, int n, const int32_t* slots, const uint64_t* new_frames, double* T_io, double* ab_io, int coarsest,
                            const double* minRes, double* lastRes, double* flow, int32_t* good, sdv_track_stats* stats) { SDV_GUARD_TRK(c);
  if (!c || n <= 0 || !slots || !new_frames || !T_io || !ab_io || coarsest < 0 || coarsest >= c->levels || coarsest >= 5) return SDV_ERR_ARG;
  CK(cudaSetDevice(c->device));
  if (n > c->jobs_cap) {
    cudaFree(c->jobs_dev); cudaFreeHost(c->jobs_host); c->jobs_cap = n;
    CK(cudaMalloc(&c->jobs_dev, (size_t)n*sizeof(TrackJob) + 16)); CK(cudaMallocHost(&c->jobs_host, (size_t)n*sizeof(TrackJob) + 16));
  }
  long long need_seq = 0;
  for (int k=0;k<n;k++) {
    if (slots[k] < 0 || slots[k] >= (int)c->slots.size()) return SDV_ERR_ARG;
    TrackerSlot& t = c->slots[slots[k]];
    FrameDev* f = find_frame(c, new_frames[k]); if (!f) return ctx_fail(c, SDV_ERR_NOFRAME, "unknown new frame (job %d)", k);
    if (f->ingest_seq > need_seq) need_seq = f->ingest_seq;
    if (t.ref_frame == ~0ull) return ctx_fail(c, SDV_ERR_STATE, "tracker slot %d has no reference", slots[k]);
    TrackJob& J = c->jobs_host[k]; memset(&J, 0, sizeof(J));
    J.img0 = f->I0;
    for (int l=0;l<c->levels;l++) { J.img[l] = (l == 0) ? nullptr : f->lvl[l]; J.pts[l] = t.pts[l]; J.npts[l] = t.npts[l]; }
    J.refExposure = t.refExposure; J.newExposure = f->exposure; J.ref_a = t.ref_a; J.ref_b = t.ref_b;
    for (int i=0;i<7;i++) J.T[i] = T_io[7*k+i];
    J.ab[0] = ab_io[2*k]; J.ab[1] = ab_io[2*k+1];
    for (int i=0;i<5;i++) J.minRes[i] = minRes ? minRes[5*k+i] : nan("");
    J.coarsest = coarsest;
  }
  { int rcj = join_ingest_upto(c, need_seq); if (rcj) return rcj; }     // only the uploads that built THESE frames: batch k+1's upload keeps streaming
  c->launches += 2;                                                    // descriptor copy kernel + track_cluster_kernel
  launch_h2d_words(c->jobs_dev, c->jobs_host, (size_t)n*sizeof(TrackJob), c->st);      // kernel copy: not queued behind the next batch's image upload
  CK(cudaEventRecord(c->ev0, c->st));
  CK(launch_track_cluster(c->jobs_dev, n, c->tc_dev, c->set.cluster_size, c->set.track_threads, c->st));
  CK(cudaEventRecord(c->ev1, c->st));
  CK(cudaMemcpyAsync(c->jobs_host, c->jobs_dev, (size_t)n*sizeof(TrackJob), cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st)); CK(cudaGetLastError());
  CK(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  for (int k=0;k<n;k++) {
    const TrackJob& J = c->jobs_host[k];
    for (int i=0;i<7;i++) T_io[7*k+i] = J.T[i];
    ab_io[2*k] = J.ab[0]; ab_io[2*k+1] = J.ab[1];
    if (lastRes) for (int i=0;i<5;i++) lastRes[5*k+i] = J.lastRes[i];
    if (flow) for (int i=0;i<3;i++) flow[3*k+i] = J.flow[i];
    if (good) good[k] = J.good;
    if (stats) for (int l=0;l<SDV_PYR_LEVELS;l++) { stats[k].point_evals[l] = J.point_evals[l]; stats[k].iterations[l] = J.iterations[l]; stats[k].accepts[l] = J.accepts[l]; }
  }
  return SDV_OK;
}
int sdv_tracker_track(sdv_ctx* c, int slot, uint64_t new_frame, double T_io[7], double ab_io[2], int coarsest, const double minRes[5],
                      double lastRes[5], double flow[3], int* good, sdv_track_stats* stats) { SDV_GUARD_TRK(c);
  int32_t s = slot, g = 0;
  int rc = sdv_tracker_track_batch(c, 1, &s, &new_frame, T_io, ab_io, coarsest, minRes, lastRes, flow, &g, stats);
  if (good) *good = g; return rc;
}

} // extern "C"
