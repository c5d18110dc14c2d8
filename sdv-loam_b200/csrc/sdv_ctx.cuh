// sdv_ctx.cuh — the context object behind the opaque sdv_ctx handle (host bookkeeping only).
#pragma once
#include <vector>
#include <unordered_map>
#include <mutex>
#include <atomic>
#include "sdv_kernels.cuh"
#include "../../include/sdv_b200.h"

namespace sdv {

struct FrameDev {                         // one FrameHessian's images on the device
  uint64_t id; float exposure; bool used;
  float* I0; float* I0_own; bool adopted;  // level-0 intensity plane (own storage, or an adopted caller buffer)
  float4* base; float4* lvl[kLevels];     // packed {I,dx,dy,|grad|^2} texels of levels >= 1 (lvl[0] = lazily built level-0 texels or nullptr)
  int lvl0_slot;                          // index into the keyframe level-0 texel pool, -1 if not built
  long long ingest_seq;                   // ingest call that (re)built this frame's pyramid
};

struct TrackerSlot {                      // one CoarseTracker instance (reference keeps two: FullSystem.h coarseTracker / coarseTracker_forNewKF)
  float4* pts[kLevels]; int npts[kLevels]; int cap[kLevels];
  uint64_t ref_frame; float refExposure; double ref_a, ref_b;   // lastRef, lastRef_aff_g2l
  double totals[kNAcc]; bool has_totals;                        // reduced sums of the last calc_res (what calcGSSSE reads from buf_warped_*)
};

struct BAState;                           // sdv_ba.cuh
struct RpState;                           // sdv_reproject.cu

} // namespace sdv

struct sdv_ctx {
  // Two call domains, as in the reference (FullSystem::trackMutex / mapMutex): tracker-slot, frame, map and policy entries run under mu_trk on stream `st`;
  // back-end (sdv_ba_*) entries run under mu_ba on stream `st_ba`.  A tracking thread and a mapping thread may therefore use ONE context concurrently.
  // Back-end entries that touch the frame table (set_window, marginalize_frame, clear) take mu_trk as well, always after mu_ba (tracker entries never take mu_ba).
  std::recursive_mutex mu_trk, mu_ba; std::mutex mu_err;
  int device; cudaStream_t st, st_in, st_cp, st_ba; cudaEvent_t ev0, ev1, ev_in, ba_ev0, ba_ev1, ev_xdom; bool ingest_pending; std::atomic<long long> launches;
  float ba_last_ms = 0.f;
  // ingest pipeline: H2D copies on st_cp, pyramids on st_in, one completion event per ingest call (ring); a frame remembers the ingest that built it
  static constexpr int kIngRing = 8; cudaEvent_t ev_ing[kIngRing], ev_cp[2]; long long ingest_seq = 0, seq_waited = 0;
  int w, h, levels; sdv_settings set;
  sdv::TrackConst tc; sdv::TrackConst* tc_dev;
  size_t lvl_off[sdv::kLevels]; size_t frame_texels;
  std::vector<sdv::FrameDev> frames; std::unordered_map<uint64_t,int> frame_index;
  // frames whose device images are referenced by a resident BA window (BAHeader::frames[].img0) or map slot (MapDev::hostI0): refcount per handle.
  // A pinned frame can be neither released nor re-uploaded (SDV_ERR_STATE) — its pool storage would be handed to another frame under the reader.
  std::unordered_map<uint64_t,int> pins;
  unsigned char* stage_u8[2] = {nullptr, nullptr}; size_t stage_u8_cap[2] = {0, 0};   // contiguous mono8 staging per parity (adjacent host images coalesce into one copy)
  std::vector<float*> stage[2]; int stage_cap; sdv::PyrBatchHost* pyr_batch_dev[2]; sdv::PyrBatchHost* pyr_batch_host[2];     // double-buffered by ingest parity
  sdv::UndistortDev und = {}; bool has_und = false; float* und_buf = nullptr;       // sdv_set_undistort: remap tables (+ response / vignette) of the raw-image ingest
  std::vector<float4*> lvl0_pool; std::vector<int> lvl0_free;
  std::vector<void*> cp_dst, cp_src; std::vector<size_t> cp_sz; bool no_batch_copy = false;
  std::vector<sdv::TrackerSlot> slots;
  double* partials; unsigned int* ticket; double* totals_dev; double* totals_host;
  float *cd_id[sdv::kLevels], *cd_ws[sdv::kLevels], *cd_id2[sdv::kLevels], *cd_ws2[sdv::kLevels];
  int* cd_owner; int* cd_counts; int* cd_scalars; int* cd_scalars_host;
  int cd_cap; float* cd_pts4; int* cd_round; float4* cd_splats; int* cd_done;
  int jobs_cap; sdv::TrackJob* jobs_dev; sdv::TrackJob* jobs_host;
  float last_ms;
  void* refine_dev = nullptr; void* refine_host = nullptr; size_t refine_cap = 0;      // staging of sdv_tracker_struct_pose_batch
  void* trace_dev = nullptr; size_t trace_cap = 0;                                  // scratch of the immature-point calls (sdv_trace.cu)
  void* lidar = nullptr;                        // sdv::LidarState: engine of the LiDAR front-end (sdv_lidar.cu)
  void* sel = nullptr;                          // sdv::SelState: PixelSelector slots + engine of the candidate management (sdv_select.cu)
  sdv::RpState* rp = nullptr;                   // map slots + scratch of the Reprojector path (sdv_reproject.cu)
  sdv::BAState* ba = nullptr;                   // selected back-end window
  std::vector<sdv::BAState*> ba_windows; void* ba_wins_dev = nullptr; void* ba_wins_host = nullptr; int ba_wins_cap = 0;
  // the fixed Gauss-Newton launch schedule of sdv_ba_optimize_batch as a CUDA graph, keyed by what the launches depend on
  cudaGraphExec_t ba_graph = nullptr; const void* bag_wins = nullptr; int bag_n = 0, bag_maxP = 0, bag_maxR = 0, bag_its = 0;
  char err[512];
};

namespace sdv {
void rp_destroy(sdv_ctx* c);
void sel_destroy(sdv_ctx* c);
void lidar_destroy(sdv_ctx* c);
void rp_calib_changed(sdv_ctx* c);             // the Reprojector constants (K, K^-1) are rebuilt from the context calibration at the next call
int ctx_fail(sdv_ctx* c, int code, const char* fmt, ...);
void ba_destroy(sdv_ctx* c);
int  ensure_lvl0(sdv_ctx* c, FrameDev& f);     // build the packed level-0 texels of a frame on demand (keyframes / read-back)
inline void frame_pin(sdv_ctx* c, uint64_t id) { c->pins[id]++; }
inline void frame_unpin(sdv_ctx* c, uint64_t id) { auto it = c->pins.find(id); if (it != c->pins.end() && --it->second <= 0) c->pins.erase(it); }
inline bool frame_pinned(const sdv_ctx* c, uint64_t id) { return c->pins.find(id) != c->pins.end(); }
int  join_ingest(sdv_ctx* c);                  // compute stream waits for every ingest enqueued so far
int  join_ingest_upto(sdv_ctx* c, long long seq);   // ... for ingest calls <= seq only (frames carry their ingest_seq)
}
// entry-point guards (a NULL context falls through to the entry's own argument check)
#define SDV_GUARD_TRK(c) std::unique_lock<std::recursive_mutex> lk_trk_; if (c) lk_trk_ = std::unique_lock<std::recursive_mutex>((c)->mu_trk)
#define SDV_GUARD_BA(c)  std::unique_lock<std::recursive_mutex> lk_ba_;  if (c) lk_ba_  = std::unique_lock<std::recursive_mutex>((c)->mu_ba)
