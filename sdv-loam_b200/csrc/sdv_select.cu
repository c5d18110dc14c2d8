// sdv_select.cu — C-ABI of the candidate management at keyframe rate (SURVEY.md §8f rank 4 + the caller half of rank 2): binds the engine of sdv_select_core.cuh
// (kernels + host orchestration; see the header for the reference map and the design) to the context's resident frames and tracker-domain stream.
//   sdv_selector_init / _potential / _get_map   PixelSelector::PixelSelector, currentPotential, FullSystem::selectionMap      PixelSelector2.cpp:11-26, FullSystem.cpp:180-186
//   sdv_selector_make_hists                     PixelSelector::makeHists                                                     PixelSelector2.cpp:47-106
//   sdv_selector_make_maps_batch                PixelSelector::makeMapsFromLidar / makeMaps                                  PixelSelector2.cpp:354-449 / :108-200
//   sdv_make_new_traces_batch                   FullSystem::makeNewTraces (+ shiTomasiScore, setMask, ImmaturePoint ctor)    FullSystem.cpp:1261-1356, 1540-1583
//   sdv_activate_select_batch                   CoarseDistanceMap::makeDistanceMap + the candidate walk of activatePointsMT  CoarseTracker.cpp:1139-1282, FullSystem.cpp:600-671
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include <stdlib.h>
#include "sdv_ctx.cuh"
#include "sdv_select_core.cuh"

using namespace sdv;
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ctx_fail(c, SDV_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)
static_assert(sizeof(sel::NewTrace) == sizeof(sdv_new_trace) && sizeof(sel::ImmPt) == sizeof(sdv_immature_pt), "engine records mirror the C-ABI structs");

namespace sdv {
struct SelState { sel::SelEngine eng; std::vector<sel::SelectorSlot> slots; sel::Scratch aux; };
void sel_destroy(sdv_ctx* c) {
  SelState* s = (SelState*)c->sel; if (!s) return;
  for (auto& sl : s->slots) if (sl.mapD) cudaFree(sl.mapD);
  s->eng.io.release(); s->aux.release(); s->eng.destroy(); delete s; c->sel = nullptr;
}
static int sel_frame(sdv_ctx* c, uint64_t id, sel::FrameImg& F, long long& need_seq, const char* who) {
  auto it = c->frame_index.find(id); if (it == c->frame_index.end()) return ctx_fail(c, SDV_ERR_NOFRAME, "%s: unknown frame %llu", who, (unsigned long long)id);
  const FrameDev& f = c->frames[it->second]; if (f.ingest_seq > need_seq) need_seq = f.ingest_seq;
  F.I0 = f.I0; F.L1 = f.lvl[1]; F.L2 = f.lvl[2]; return SDV_OK;
}
static int sel_fail(sdv_ctx* c, SelState* s, int rc, const char* who) { return ctx_fail(c, rc == -2 ? SDV_ERR_CAPACITY : SDV_ERR_CUDA, "%s: %s", who, s->eng.err.c_str()); }
}  // namespace sdv
#define SEL_STATE(who) SelState* s = (SelState*)c->sel; if (!s) return ctx_fail(c, SDV_ERR_STATE, who ": sdv_selector_init has not been called")

extern "C" {

int sdv_selector_init(sdv_ctx* c, const uint8_t* random_pattern, int n_slots) { SDV_GUARD_TRK(c);
  if (!c || !random_pattern || n_slots < 1) return SDV_ERR_ARG;
  if (c->levels < 3) return ctx_fail(c, SDV_ERR_ARG, "selector_init: the selector reads pyramid levels 0..2 (context has %d levels)", c->levels);
  CK(cudaSetDevice(c->device));
  sel_destroy(c);
  SelState* s = new SelState(); c->sel = s;
  sel::SelSet S; S.minGradHistCut = 0.5f; S.minGradHistAdd = 3; S.gradDownweightPerLevel = 0.75f; S.selectDirectionDistribution = 1;            // util/settings.cpp:119-122
  S.outlierTH = c->set.outlierTH; S.outlierTHSumComponent = c->set.outlierTHSumComponent; S.overallEnergyTHWeight = 1;
  if (s->eng.init(c->w, c->h, S, random_pattern, c->st)) { int rc = sel_fail(c, s, -1, "selector_init"); sel_destroy(c); return rc; }
  s->slots.resize(n_slots);
  if (const char* e = getenv("SDV_WALK_THREADS")) { const int t = atoi(e); if (t == 32 || t == 64 || t == 128 || t == 256) s->eng.walk_threads = t; }   // tuning knobs
  if (const char* e = getenv("SDV_FUSE_MAP")) s->eng.fuse_map = atoi(e) != 0;
  return SDV_OK;
}
int sdv_selector_potential(sdv_ctx* c, int slot, int set_to, int* out) { SDV_GUARD_TRK(c);
  if (!c) return SDV_ERR_ARG; SEL_STATE("selector_potential");
  if (slot < 0 || slot >= (int)s->slots.size()) return ctx_fail(c, SDV_ERR_ARG, "selector_potential: slot %d of %d", slot, (int)s->slots.size());
  if (set_to > 0) s->slots[slot].currentPotential = set_to;
  if (out) *out = s->slots[slot].currentPotential;
  return SDV_OK;
}
int sdv_selector_get_map(sdv_ctx* c, int slot, uint8_t* out) { SDV_GUARD_TRK(c);
  if (!c || !out) return SDV_ERR_ARG; SEL_STATE("selector_get_map");
  if (slot < 0 || slot >= (int)s->slots.size()) return ctx_fail(c, SDV_ERR_ARG, "selector_get_map: slot %d of %d", slot, (int)s->slots.size());
  CK(cudaSetDevice(c->device));
  const size_t wh = (size_t)c->w*c->h;
  if (!s->slots[slot].mapD) { memset(out, 0, wh); return SDV_OK; }
  CK(cudaMemcpyAsync(out, s->slots[slot].mapD, wh, cudaMemcpyDeviceToHost, c->st)); CK(cudaStreamSynchronize(c->st));
  return SDV_OK;
}
int sdv_selector_make_hists(sdv_ctx* c, uint64_t frame, float* ths_out, float* thsSmoothed_out) { SDV_GUARD_TRK(c);
  if (!c) return SDV_ERR_ARG; SEL_STATE("selector_make_hists");
  CK(cudaSetDevice(c->device));
  sel::FrameImg F; long long need = 0; { int rc = sel_frame(c, frame, F, need, "selector_make_hists"); if (rc) return rc; }
  { int rc = join_ingest_upto(c, need); if (rc) return rc; }
  const size_t tf = s->eng.ths_floats(); if (s->aux.reserve(2*sel::Scratch::need(tf, 4), c->st)) return ctx_fail(c, SDV_ERR_CUDA, "selector_make_hists: scratch");
  s->aux.reset(); float* a = s->aux.take<float>(tf); float* b = s->aux.take<float>(tf);
  CK(cudaMemsetAsync(a, 0, tf*4, c->st)); CK(cudaMemsetAsync(b, 0, tf*4, c->st));
  sel::HistJob J{F.I0, a, b}; const long long l0 = s->eng.launches;
  if (s->eng.make_hists(1, &J)) return sel_fail(c, s, -1, "selector_make_hists");
  c->launches += s->eng.launches - l0;
  const int n = (c->w/32)*(c->h/32);
  if (ths_out) CK(cudaMemcpyAsync(ths_out, a, n*4, cudaMemcpyDeviceToHost, c->st));
  if (thsSmoothed_out) CK(cudaMemcpyAsync(thsSmoothed_out, b, n*4, cudaMemcpyDeviceToHost, c->st));
  CK(cudaStreamSynchronize(c->st));
  return SDV_OK;
}

int sdv_selector_make_maps_batch(sdv_ctx* c, int n, const int32_t* slots, const uint64_t* frames, const int32_t* cloud_begin, const double* cloud3, const float* density,
                                 const int32_t* recursions_left, const float* th_factor, uint8_t* maps_out, int32_t* num_have_out) { SDV_GUARD_TRK(c);
  if (!c || n < 0 || (n && (!slots || !frames || !density || !recursions_left || !th_factor || !maps_out))) return SDV_ERR_ARG;
  if (n == 0) return SDV_OK; SEL_STATE("selector_make_maps");
  const bool lidar = cloud_begin != nullptr; if (lidar && !cloud3) return SDV_ERR_ARG;
  if (lidar) { if (cloud_begin[0] != 0) return ctx_fail(c, SDV_ERR_ARG, "selector_make_maps: cloud_begin[0] must be 0");
    for (int j = 0; j < n; j++) if (cloud_begin[j+1] < cloud_begin[j]) return ctx_fail(c, SDV_ERR_ARG, "selector_make_maps: cloud_begin is not ascending at job %d", j); }
  for (int j = 0; j < n; j++) { if (slots[j] < 0 || slots[j] >= (int)s->slots.size()) return ctx_fail(c, SDV_ERR_ARG, "selector_make_maps: slot %d of %d", slots[j], (int)s->slots.size());
    for (int k = 0; k < j; k++) if (slots[k] == slots[j]) return ctx_fail(c, SDV_ERR_ARG, "selector_make_maps: slot %d appears twice in one batch", slots[j]); }
  CK(cudaSetDevice(c->device));
  const size_t wh = (size_t)c->w*c->h, tf = s->eng.ths_floats(); long long need = 0;
  std::vector<sel::FrameImg> F(n); for (int j = 0; j < n; j++) { int rc = sel_frame(c, frames[j], F[j], need, "selector_make_maps"); if (rc) return rc; }
  { int rc = join_ingest_upto(c, need); if (rc) return rc; }
  size_t bytes = 1024; for (int j = 0; j < n; j++) { const int m = lidar ? cloud_begin[j+1] - cloud_begin[j] : 0; bytes += 2*sel::Scratch::need(tf, 4) + sel::Scratch::need(3*(size_t)m, 8) + sel::Scratch::need(lidar ? (size_t)std::max(m, 1) : wh, 1); }
  if (s->aux.reserve(bytes, c->st)) return ctx_fail(c, SDV_ERR_CUDA, "selector_make_maps: scratch");
  s->aux.reset(); std::vector<sel::MapsJobHost> M(n); std::vector<sel::HistJob> Hj(n);
  for (int j = 0; j < n; j++) { const int m = lidar ? cloud_begin[j+1] - cloud_begin[j] : 0;
    float* a = s->aux.take<float>(tf); float* b = s->aux.take<float>(tf); CK(cudaMemsetAsync(a, 0, tf*4, c->st)); CK(cudaMemsetAsync(b, 0, tf*4, c->st));
    double* cl = nullptr; if (lidar) { cl = s->aux.take<double>(3*(size_t)std::max(m, 1)); if (m) CK(cudaMemcpyAsync(cl, cloud3 + 3*(size_t)cloud_begin[j], 3*(size_t)m*sizeof(double), cudaMemcpyHostToDevice, c->st)); }
    sel::SelectorSlot& sl = s->slots[slots[j]];
    if (!lidar && !sl.mapD) { CK(cudaMalloc((void**)&sl.mapD, wh)); CK(cudaMemsetAsync(sl.mapD, 0, wh, c->st)); }
    Hj[j] = sel::HistJob{F[j].I0, a, b};
    M[j].img = F[j]; M[j].thsSm = b; M[j].cloud_dev = cl; M[j].n = m; M[j].map = lidar ? s->aux.take<unsigned char>(std::max(m, 1)) : sl.mapD; M[j].density = density[j];
    M[j].recursionsLeft = recursions_left[j]; M[j].thFactor = th_factor[j]; M[j].currentPotential = &sl.currentPotential; M[j].numHaveSub = 0; M[j].passes = 0; }
  const long long l0 = s->eng.launches;
  if (s->eng.make_hists(n, Hj.data())) return sel_fail(c, s, -1, "selector_make_maps");
  if (s->eng.make_maps(M, lidar)) return sel_fail(c, s, -1, "selector_make_maps");
  c->launches += s->eng.launches - l0;
  size_t o = 0;
  for (int j = 0; j < n; j++) { const size_t sz = lidar ? (size_t)(cloud_begin[j+1] - cloud_begin[j]) : wh;
    if (sz) CK(cudaMemcpyAsync(maps_out + o, M[j].map, sz, cudaMemcpyDeviceToHost, c->st)); o += sz; if (num_have_out) num_have_out[j] = M[j].numHaveSub; }
  CK(cudaStreamSynchronize(c->st));
  return SDV_OK;
}

int sdv_make_new_traces_batch(sdv_ctx* c, int n, const int32_t* slots, const uint64_t* frames, const int32_t* cloud_begin, const double* cloud3, const float* density_lidar,
                              const float* density_dense, const int32_t* add_feature_point, int cap, sdv_new_trace* out, sdv_immature_pt* imm_out, int32_t* n_out, int32_t* num_points2) { SDV_GUARD_TRK(c);
  if (!c || n < 0 || (n && (!slots || !frames || !cloud_begin || !density_lidar || !density_dense || !add_feature_point || !out || !n_out || cap < 1))) return SDV_ERR_ARG;
  if (n == 0) return SDV_OK; SEL_STATE("make_new_traces");
  if (cloud_begin[0] != 0) return ctx_fail(c, SDV_ERR_ARG, "make_new_traces: cloud_begin[0] must be 0");
  for (int j = 0; j < n; j++) { if (cloud_begin[j+1] < cloud_begin[j]) return ctx_fail(c, SDV_ERR_ARG, "make_new_traces: cloud_begin is not ascending at job %d", j);
    if (cloud_begin[j+1] > cloud_begin[j] && !cloud3) return SDV_ERR_ARG;
    if (slots[j] < 0 || slots[j] >= (int)s->slots.size()) return ctx_fail(c, SDV_ERR_ARG, "make_new_traces: slot %d of %d", slots[j], (int)s->slots.size());
    for (int k = 0; k < j; k++) if (slots[k] == slots[j]) return ctx_fail(c, SDV_ERR_ARG, "make_new_traces: slot %d appears twice in one batch", slots[j]); }
  CK(cudaSetDevice(c->device));
  long long need = 0; std::vector<sel::SelEngine::NewTracesJob> J(n);
  for (int j = 0; j < n; j++) { int rc = sel_frame(c, frames[j], J[j].img, need, "make_new_traces"); if (rc) return rc;
    J[j].cloud_host = cloud3 ? cloud3 + 3*(size_t)cloud_begin[j] : nullptr; J[j].n = cloud_begin[j+1] - cloud_begin[j]; J[j].slot = &s->slots[slots[j]]; J[j].densityLidar = density_lidar[j];
    J[j].densityDense = density_dense[j]; J[j].addFeaturePoint = add_feature_point[j]; J[j].out_host = (sel::NewTrace*)out + (size_t)j*cap; J[j].imm_host = imm_out ? (sel::ImmPt*)imm_out + (size_t)j*cap : nullptr; J[j].cap = cap; }
  { int rc = join_ingest_upto(c, need); if (rc) return rc; }
  const long long l0 = s->eng.launches;
  CK(cudaEventRecord(c->ev0, c->st));
  { int rc = s->eng.make_new_traces(J); if (rc) return sel_fail(c, s, rc, "make_new_traces"); }
  CK(cudaEventRecord(c->ev1, c->st)); CK(cudaStreamSynchronize(c->st)); CK(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  c->launches += s->eng.launches - l0;
  for (int j = 0; j < n; j++) { n_out[j] = J[j].n_out; if (num_points2) { num_points2[2*j] = J[j].numPoints[0]; num_points2[2*j+1] = J[j].numPoints[1]; } }
  return SDV_OK;
}

int sdv_activate_select_batch(sdv_ctx* c, int n, const int32_t* host_begin, const int32_t* pt_begin, const float* KRKi9, const float* Kt3, const float* uvid,
                              const int32_t* cand_host_begin, const int32_t* cand_begin, const float* cKRKi9, const float* cKt3, const float* cand4, const float* min_act_dist,
                              int32_t* decision_out, float* dist_map_out) { SDV_GUARD_TRK(c);
  if (!c || n < 0 || (n && (!host_begin || !pt_begin || !cand_host_begin || !cand_begin || !min_act_dist))) return SDV_ERR_ARG;
  if (n == 0) return SDV_OK; SEL_STATE("activate_select");
  if (host_begin[0] != 0 || cand_host_begin[0] != 0) return ctx_fail(c, SDV_ERR_ARG, "activate_select: host_begin[0] and cand_host_begin[0] must be 0");
  // the engine wants, per sequence, CSR offsets that start at 0: rebase the global ones
  std::vector<std::vector<int>> pb(n), cb(n); std::vector<sel::SelEngine::ActJob> J(n); const size_t n1 = (size_t)(c->w >> 1)*(c->h >> 1);
  for (int j = 0; j < n; j++) {
    const int h0 = host_begin[j], h1 = host_begin[j+1], g0 = cand_host_begin[j], g1 = cand_host_begin[j+1];
    if (h1 < h0 || g1 < g0 || h1 - h0 > 16 || g1 - g0 > 16) return ctx_fail(c, SDV_ERR_ARG, "activate_select: sequence %d has %d source / %d candidate keyframes (0..16)", j, h1 - h0, g1 - g0);
    for (int k = h0; k <= h1; k++) { if (k > h0 && pt_begin[k] < pt_begin[k-1]) return ctx_fail(c, SDV_ERR_ARG, "activate_select: pt_begin is not ascending"); pb[j].push_back(pt_begin[k] - pt_begin[h0]); }
    for (int k = g0; k <= g1; k++) { if (k > g0 && cand_begin[k] < cand_begin[k-1]) return ctx_fail(c, SDV_ERR_ARG, "activate_select: cand_begin is not ascending"); cb[j].push_back(cand_begin[k] - cand_begin[g0]); }
    const int np = pb[j].back(), nc = cb[j].back();
    if ((np && (!KRKi9 || !Kt3 || !uvid)) || (nc && (!cKRKi9 || !cKt3 || !cand4 || !decision_out))) return SDV_ERR_ARG;
    sel::SelEngine::ActJob& a = J[j]; a.nHosts = h1 - h0; a.pt_begin = pb[j].data(); a.KRKi = KRKi9 ? KRKi9 + 9*(size_t)h0 : nullptr; a.Kt = Kt3 ? Kt3 + 3*(size_t)h0 : nullptr;
    a.uvid = uvid ? uvid + 3*(size_t)pt_begin[h0] : nullptr; a.nCandHosts = g1 - g0; a.cand_begin = cb[j].data(); a.cKRKi = cKRKi9 ? cKRKi9 + 9*(size_t)g0 : nullptr; a.cKt = cKt3 ? cKt3 + 3*(size_t)g0 : nullptr;
    a.cand4 = cand4 ? cand4 + 4*(size_t)cand_begin[g0] : nullptr; a.minActDist = min_act_dist[j]; a.decision_host = decision_out ? decision_out + cand_begin[g0] : nullptr;
    a.map_host = dist_map_out ? dist_map_out + (size_t)j*n1 : nullptr;
  }
  CK(cudaSetDevice(c->device));
  const long long l0 = s->eng.launches;
  CK(cudaEventRecord(c->ev0, c->st));
  { int rc = s->eng.activate(J); if (rc) return sel_fail(c, s, rc, "activate_select"); }
  CK(cudaEventRecord(c->ev1, c->st)); CK(cudaStreamSynchronize(c->st)); CK(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  if (s->eng.have_ev) c->last_ms = s->eng.last_kernel_ms;                                  // the 43 launches, copies excluded
  c->launches += s->eng.launches - l0;
  return SDV_OK;
}

}  // extern "C"
