/* sdv_b200.h — C-ABI of the B200-native SDV-LOAM tracking / optimisation hot path.
 *
 * The reference (ZikangYuan/SDV-LOAM) has NO plugin/FFI layer: the hot path sits behind C++ member
 * functions called by FullSystem.  Each entry point below names the reference member it replaces (file:line relative
 * to /root/reference/src); INTEGRATION.md shows the shim classes a maintainer adds to keep those signatures.
 *
 * Conventions
 *   - return 0 on success, negative sdv_status on error; never throws; sdv_last_error() gives the message.
 *   - NaN/inf in energies propagate unchanged (the reference signals tracking loss through them,
 *     FullSystem.cpp:862-867, FullSystemOptimize.cpp:472-476).
 *   - all pointers are HOST pointers owned by the caller and are consumed before the call returns, except the
 *     `_dev` variants which take device pointers of the context's device.
 *   - poses: double T[7] = {qw,qx,qy,qz, tx,ty,tz} (Sophus SE3d storage order, unit quaternion).
 *   - a context is bound to one CUDA device; calls on one context must not overlap (one FullSystem = one context;
 *     the reference serialises the same calls under trackMutex / mapMutex, FullSystem.h:277-323).
 */
#ifndef SDV_B200_H
#define SDV_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SDV_PYR_LEVELS 6            /* util/settings.h:25 PYR_LEVELS */
#define SDV_MAX_FRAMES_WINDOW 8     /* setting_maxFrames=7 (settings.cpp:56); 8 for the stress config */

typedef enum {
  SDV_OK = 0, SDV_ERR_ARG = -1, SDV_ERR_CUDA = -2, SDV_ERR_NOFRAME = -3, SDV_ERR_CAPACITY = -4, SDV_ERR_STATE = -5
} sdv_status;

typedef struct sdv_ctx sdv_ctx;

typedef struct { float fx, fy, cx, cy; } sdv_calib;          /* CalibHessian::value_scaledf, HessianBlocks.h:260-358 */

typedef struct {                                             /* util/settings.cpp globals that reach the hot path */
  float huberTH;                /* setting_huberTH = 6            settings.cpp:101 */
  float coarseCutoffTH;         /* setting_coarseCutoffTH = 20    settings.cpp:112 */
  float affineOptModeA;         /* <0 fix, >=0 optimise           settings.cpp:93 ; mode=1 -> 0 (main.cpp:453) */
  float affineOptModeB;         /*                                settings.cpp:94 ; mode=1 -> 0 (main.cpp:454) */
  float outlierTH;              /* setting_outlierTH = 144        settings.cpp:64 */
  float outlierTHSumComponent;  /* = 2500                         settings.cpp:65 */
  float idepthFixPrior;         /* = 2500                         settings.cpp */
  int   max_ref_points;         /* capacity of a tracker reference cloud per level (0 = w*h) */
  int   n_tracker_slots;        /* reference uses 2 (coarseTracker, coarseTracker_forNewKF; FullSystem.h); batched mode: 2 x sequences */
  int   max_frames;             /* device-resident frame pyramids kept alive at once */
  int   cluster_size;           /* CTAs per thread-block cluster per trackNewestCoarse call in the device-resident LM kernel
                                   (default 1 = batched-throughput mode; 8..16 = low-latency single-sequence mode) */
  int   track_threads;          /* threads per CTA of that kernel: 128 (default, 4 jobs resident per SM), 64 or 256 */
  int   max_kf_images;          /* pool of packed level-0 {I,dx,dy} images, built on demand for keyframes entering the BA window
                                   (0 = SDV_MAX_FRAMES_WINDOW + 4); tracked-only frames never materialise them */
} sdv_settings;

void sdv_default_settings(sdv_settings* s);

/* one per FullSystem (or one per GPU in batched mode) */
int  sdv_create(const sdv_calib* K, int w, int h, int levels, const sdv_settings* s, int device, sdv_ctx** out);
void sdv_destroy(sdv_ctx* c);
const char* sdv_last_error(sdv_ctx* c);
int  sdv_pyr_levels(int w, int h);                           /* pyrLevelsUsed rule, util/globalCalib.cpp:22-30 */
int  sdv_sync(sdv_ctx* c);                                   /* drain the context's stream */

/* Threading: a context may be used by TWO host threads at once, split the way the reference splits its work — every sdv_ba_* entry (mapping
 * thread, FullSystem::mapMutex) on one side, every other entry (tracking thread, trackMutex) on the other.  Each side is serialised internally and runs on its own
 * stream; calls of the same side from several threads are serialised by the library. */
/* CoarseTracker::makeK(CalibHessian*) (CoarseTracker.cpp:77-106) and the CalibHessian Reprojector::reprojectMap reads (Reprojector.cpp:117-124): the bundle
 * adjustment optimises the intrinsics at every keyframe, and setCoarseTrackingRef re-derives the tracker's per-level K from them.  Replaces the calibration given
 * to sdv_create for every later tracker / reprojection / structPoseEstimation call of this context (the BA windows carry their own: sdv_ba_set_window). */
int  sdv_set_calib(sdv_ctx* c, const sdv_calib* K);

/* ---- frames: FrameHessian::makeImages(float* color, CalibHessian*)  FullSystem/HessianBlocks.cpp:107-167 ------------
 * builds the pyramid on the device, keyed by a caller-chosen handle: level 0 is kept as the planar intensity image (its {dx,dy} are
 * formed on the fly, bit-identically, by the tracker; packed level-0 texels are built lazily when the frame enters a BA window),
 * levels >= 1 as packed {I,dx,dy,absSquaredGrad} texels. */
int  sdv_frame_upload(sdv_ctx* c, uint64_t frame, const float* img_wh, float exposure);
int  sdv_frame_upload_batch(sdv_ctx* c, int n, const uint64_t* frames, const float* const* imgs_wh, const float* exposures);
/* Uploads are ASYNCHRONOUS on a dedicated ingest stream so that they overlap the tracking of the previous batch: pinned host
 * buffers must stay valid until the next call that synchronises (sdv_sync, any tracker/BA call, sdv_frame_download).
 * _u8: level-0 input in the sensor wire format (sensor_msgs/Image mono8, what src/main.cpp:537-560 receives) — the
 * u8->float conversion of the ingest (DatasetReader.h:152-155, Undistort crop without photometric calibration) is fused
 * into the level-0 kernel; 4x less PCIe traffic.   _dev: level-0 images already in device memory (kernel-only timing). */
int  sdv_frame_upload_batch_u8(sdv_ctx* c, int n, const uint64_t* frames, const uint8_t* const* imgs_wh, const float* exposures);
/* RAW ingest — Undistort::undistort<unsigned char> (util/Undistort.cpp:341-435) + PhotometricUndistorter::processFrame (:177-214), what
 * src/main.cpp:537-560 runs on every sensor_msgs/Image before FullSystem::addActiveFrame.  sdv_set_undistort hands over, once per calibration,
 * the tables the reference's Undistort object holds after readFromFile (:842-886): remapX/remapY (w*h floats of the RECTIFIED size given to
 * sdv_create, source coordinates in the raw w_org x h_org image, -1 = outside), the scalar `factor` of the uncalibrated path, and optionally the
 * inverse response G[256] and the inverse vignette (w_org*h_org) of the photometric calibration (both NULL for KITTI: no pcalib).  A response is
 * applied to frames uploaded with exposure > 0 only (:185).  _raw_u8 then takes native-size mono8 images; rectification, photometric stage and the
 * level-0/level-1 pyramid build are ONE kernel.  Entries that would read outside the raw image are refused (SDV_ERR_ARG). */
int  sdv_set_undistort(sdv_ctx* c, int w_org, int h_org, const float* remapX, const float* remapY, float factor, const float* G256, const float* vignette_inv);
int  sdv_frame_upload_batch_raw_u8(sdv_ctx* c, int n, const uint64_t* frames, const uint8_t* const* raw_imgs_worg_horg, const float* exposures);
int  sdv_frame_build_batch_dev(sdv_ctx* c, int n, const uint64_t* frames, const void* const* imgs_dev, int fmt, const float* exposures);
/* fmt: 0 float (copied into frame storage), 1 mono8, 2 float ADOPTED as the frame's level-0 plane (zero copy: the buffer must stay
 * valid and unmodified until sdv_frame_release / the handle is re-uploaded — for producers that already write into device memory). */
int  sdv_frame_release(sdv_ctx* c, uint64_t frame);
/* test hook: copy one level back as AoS {I,dx,dy} (w*h*3 floats) and absSquaredGrad (w*h floats); either may be NULL */
int  sdv_frame_download(sdv_ctx* c, uint64_t frame, int lvl, float* dI3_out, float* abs_out);

/* ---- tracker reference: CoarseTracker::makeK :77-106 is folded into sdv_create ------------------------------------
 * CoarseTracker::setCoarseTrackingRef / setCTRefForFirstFrame  CoarseTracker.cpp:636-660 -> makeCoarseDepthL0 :258-425.
 * The PointHessian graph walk (:264-294) stays on the host: the caller passes one row per splat, {u,v,idepth,HdiF},
 * with round_half[i]=0 for `int(u)` (newest-KF sensor points, :270-271) or 1 for `int(u+0.5f)` (:116-117, :285-286). */
int  sdv_tracker_set_ref(sdv_ctx* c, int slot, uint64_t ref_frame, int n, const float* pts4, const int32_t* round_half,
                         float ref_exposure_unused, double ref_a, double ref_b);
/* direct injection / readback of one level's reference cloud pc_u,pc_v,pc_idepth,pc_color (tests, stress config) */
int  sdv_tracker_set_cloud(sdv_ctx* c, int slot, uint64_t ref_frame, int lvl, int n, const float* u, const float* v,
                           const float* idepth, const float* color, double ref_a, double ref_b);
int  sdv_tracker_get_cloud(sdv_ctx* c, int slot, int lvl, int* n, float* u, float* v, float* idepth, float* color);

/* ---- tracker residual + Gauss-Newton system (one fused kernel launch) ---------------------------------------------
 * Vec6 CoarseTracker::calcRes(int lvl, const SE3& refToNew, AffLight aff_g2l, float cutoffTH)   CoarseTracker.cpp:486-634
 * void CoarseTracker::calcGSSSE(int lvl, Mat88& H_out, Vec8& b_out, const SE3&, AffLight)        CoarseTracker.cpp:427-484
 * calc_res evaluates residuals AND accumulates the 9x9 system of the same pose in one pass; calc_gs returns the
 * (scaled, /n) H,b of the last calc_res on that slot, exactly the state calcGSSSE would read from buf_warped_*. */
int  sdv_tracker_calc_res(sdv_ctx* c, int slot, uint64_t new_frame, int lvl, const double T[7], double a, double b,
                          float cutoffTH, double rs_out[6]);
int  sdv_tracker_calc_gs(sdv_ctx* c, int slot, int lvl, double H88_out[64], double b8_out[8]);

/* ---- device-resident coarse-to-fine LM ------------------------------------------------------------------------------
 * bool CoarseTracker::trackNewestCoarse(FrameHessian*, SE3& lastToNew_out, AffLight& aff_g2l_out, int coarsestLvl,
 *                                       Vec5 minResForAbort, Output3DWrapper*)                  CoarseTracker.cpp:662-838
 * T_io/ab_io are in-out like lastToNew_out/aff_g2l_out (written only when the call is not aborted, as in the reference);
 * lastRes = CoarseTracker::lastResiduals, flow = lastFlowIndicators (CoarseTracker.h:57-65). */
typedef struct {
  int64_t point_evals[SDV_PYR_LEVELS];   /* sum of pc_n[lvl] over all calcRes evaluations (roofline accounting) */
  int32_t iterations[SDV_PYR_LEVELS];
  int32_t accepts[SDV_PYR_LEVELS];
} sdv_track_stats;
int  sdv_tracker_track(sdv_ctx* c, int slot, uint64_t new_frame, double T_io[7], double ab_io[2], int coarsest,
                       const double minResForAbort[5], double lastRes[5], double flow[3], int* good, sdv_track_stats* stats);
/* batched mode (north_star: independent sequences / Monte-Carlo re-runs): n independent trackNewestCoarse calls in ONE
 * launch, one thread-block cluster each.  Arrays are n x the single-call shapes. */
int  sdv_tracker_track_batch(sdv_ctx* c, int n, const int32_t* slots, const uint64_t* new_frames, double* T_io, double* ab_io,
                             int coarsest, const double* minResForAbort, double* lastRes, double* flow, int32_t* good,
                             sdv_track_stats* stats);
/* ---- map reprojection + direct feature alignment: the Reprojector of FullSystem::trackNewCoarse / makeKeyFrame
 *   void Reprojector::reprojectMap(FrameHessian* frame, std::vector<std::pair<PointHessian*, Eigen::Vector2d>>& overlap_pts)     Reprojector.cpp:117-156
 *   void Reprojector::backprojectMap(FrameHessian* ref_frame, FrameHessian* frame, overlap_pts&)                                   :158-185
 *   with reprojectPoint :600-616, reprojectCell :198-233, findMatchDirect :235-292, getWarpMatrixAffine :14-37, getBestSearchLevel :39-51,
 *   warpAffine :53-86, align1D :346-455, align2D :457-560 (options_: find_match_direct = true, align_max_iter = 10; grid cell 25 px).
 * sdv_map_set     makes the active map of one sequence resident (per keyframe): the window's keyframes (frameHessians_ order: device frame
 *                 handle, shell->camToWorld, shell->aff_g2l {a,b}) and their ACTIVE PointHessians grouped by host in that order
 *                 (u, v, idepth, host index, type 0 CORNER / 1 EDGELET).  `slot` shares the index space of the tracker slots.
 * sdv_reproject_map_batch   n independent reprojectMap calls in one launch sequence.  Per job: map slot, target frame handle + camToWorld +
 *                 aff_g2l, cur_kf_index (index of the target in the map's keyframes or -1), only_host (-1: reprojectMap over all keyframes in
 *                 close_kfs order; h: backprojectMap of keyframe h's points), backup (Reprojector::backup, selects the reference keyframe when
 *                 there are <= 2 keyframes, :242-250).  cell_order[n_cols*n_rows] is the grid visiting order (the reference shuffles it with
 *                 rand(), :111; NULL = identity); max_matches = (int)(0.8*setting_desiredImmatureDensity).
 *                 Outputs per job k, in visiting order: n_out[k] matches, out_pt[k*cells + i] = index into the slot's points,
 *                 out_px[(k*cells + i)*2] = aligned pixel (it->px). */
typedef struct { float u, v, idepth; int32_t host; int32_t type; } sdv_map_pt;
int  sdv_reproject_grid(sdv_ctx* c, int* n_cols, int* n_rows);
int  sdv_map_set(sdv_ctx* c, int slot, int nH, const uint64_t* host_frames, const double* host_T7, const double* host_ab, int nP, const sdv_map_pt* pts);
/* A resident map slot (and a resident BA window, below) REFERENCES the device images of its keyframes: while it does, sdv_frame_release /
 * re-uploading one of those handles returns SDV_ERR_STATE instead of handing the storage to another frame under the reader.  sdv_map_set on
 * the same slot replaces the references; sdv_map_clear drops them (the FrameHessian destructor path of a marginalised keyframe). */
int  sdv_map_clear(sdv_ctx* c, int slot);
int  sdv_reproject_map_batch(sdv_ctx* c, int n_jobs, const int32_t* slots, const uint64_t* cur_frames, const double* cur_T7, const double* cur_ab,
                             const int32_t* cur_kf_index, const int32_t* only_host, const int32_t* backup, const int32_t* cell_order, int max_matches,
                             int32_t* n_out, int32_t* out_pt, double* out_px);

/* ---- semi-direct pose refinement on matched map points
 * bool CoarseTracker::structPoseEstimation(SE3& curToWorld, std::vector<std::pair<PointHessian*, Eigen::Vector2d>>& overlap_pts)
 *                                                                                              CoarseTracker.cpp:949-1007
 *   with calculateRes :840-871, calculateWeight :873-887, calcHandb :889-947 (called from FullSystem::trackNewCoarse, FullSystem.cpp:483-488).
 * One overlap point = the PointHessian fields the function reads (u, v, idepth, host) + the matched pixel (it->second cast to float).
 * `host` indexes the job's host_T7 rows = host->shell->camToWorld of the distinct host keyframes.  curToWorld_io is in-out like the
 * reference argument (written only by accepted steps).  res = final mean squared reprojection error (resOld), iterations/accepts = loop
 * statistics.  Batched: job k owns pts[pt_begin[k] .. pt_begin[k+1]) and host_T7 rows [host_begin[k] .. host_begin[k+1]) (<= 16 hosts). */
typedef struct { float u, v, idepth; int32_t host; float obs_x, obs_y; } sdv_overlap_pt;
int  sdv_tracker_struct_pose(sdv_ctx* c, int n, const sdv_overlap_pt* pts, int nH, const double* host_T7, double curToWorld_io[7],
                             float* res, int* iterations, int* accepts);
int  sdv_tracker_struct_pose_batch(sdv_ctx* c, int n_jobs, const int32_t* pt_begin, const sdv_overlap_pt* pts, const int32_t* host_begin,
                                   const double* host_T7, double* curToWorld_io, float* res, int32_t* iterations, int32_t* accepts);
/* ---- fused per-frame refinement (the tail of FullSystem::trackNewCoarse, FullSystem.cpp:482-488): reprojectMap(fh) -> structPoseEstimation
 * without leaving the device.  curToWorld_io: fh->shell->camToWorld from the photometric tracker in, refined pose out; cur_ab = the tracked
 * aff_g2l.  n_matches = overlap_pts.size(), res/iterations/accepts as in sdv_tracker_struct_pose_batch. */
int  sdv_tracker_refine_batch(sdv_ctx* c, int n_jobs, const int32_t* slots, const uint64_t* cur_frames, double* curToWorld_io, const double* cur_ab,
                              const int32_t* cell_order, int max_matches, int32_t* n_matches, float* res, int32_t* iterations, int32_t* accepts);
/* ---- Vec4 FullSystem::trackNewCoarse(FrameHessian* fh)                                                   FullSystem.cpp:283-500 (§8 a4)
 * The per-frame policy for a running system (allFrameHistory.size() > 2): 31 motion hypotheses (:334-395), the re-track loop over
 * trackNewestCoarse with the achievedRes abort vector (:410-462), fallback (:464-470), pose composition (:474-479), reprojectMap +
 * structPoseEstimation (:481-488).  Batched over n sequences: try i is one track_batch launch over the sequences still searching.
 * The caller keeps FullSystem's per-sequence state (allFrameHistory shells, lastCoarseRMSE, firstCoarseRMSE) and passes what the function reads. */
typedef struct {
  /* in */
  int32_t  slot;                 /* tracker slot == map slot of the sequence (coarseTracker / the active window) */
  int32_t  poses_valid;          /* 1: slast->poseValid && sprelast->poseValid && lastF->shell->poseValid (:390) -> 31 hypotheses; 0: -> identity only (:391-394);
                                    2: the second frame of a sequence (allFrameHistory.size() == 2, :299-331) -> identity + 52 pure rotations, no history needed */
  uint64_t frame;                /* device handle of fh */
  double   sprelast_c2w[7], slast_c2w[7], lastF_c2w[7];   /* camToWorld of allFrameHistory[size-3], [size-2], coarseTracker->lastRef->shell */
  double   aff_last[2];          /* slast->aff_g2l {a,b} */
  /* in-out */
  double   lastCoarseRMSE[5];    /* FullSystem::lastCoarseRMSE in, achievedRes out (:472) */
  /* out */
  double   camToWorld[7];        /* fh->shell->camToWorld after structPoseEstimation */
  double   camToTrackingRef[7];  /* fh->shell->camToTrackingRef (:490-491) */
  double   aff_g2l[2];           /* fh->shell->aff_g2l */
  double   flow[3];              /* flowVecs (the Vec4 result is {lastCoarseRMSE[0], flow[0..2]}) */
  float    refine_res;
  int32_t  have_one_good, tries, n_matches, refine_iterations, refine_accepts;
} sdv_track_new_coarse_io;
int  sdv_track_new_coarse_batch(sdv_ctx* c, int n, sdv_track_new_coarse_io* io, const int32_t* cell_order, int max_matches);
/* lastF_2_fh_tries[i] of FullSystem.cpp:346-394 for one job (host arithmetic only, no device needed); n_tries = 31, or 1 when the pose history is invalid */
int  sdv_track_hypothesis(const sdv_track_new_coarse_io* io, int i, double T7_out[7], int* n_tries);

/* device time of the last track / track_batch / calc_res launch in milliseconds (CUDA events on the context stream) */
float sdv_last_kernel_ms(sdv_ctx* c);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
long long sdv_launch_count(sdv_ctx* c);
/* bytes copied H2D and again D2H per trackNewestCoarse job (the job descriptor carries inputs and results) */
int  sdv_track_job_bytes(void);


/* ---- immature points: candidate construction and epipolar tracing
 *   ImmaturePoint::ImmaturePoint(int u, int v, FrameHessian* host, float type, CalibHessian*)               FullSystem/ImmaturePoint.cpp:8-36
 *   ImmaturePointStatus ImmaturePoint::traceOn(FrameHessian* frame, const Mat33f& hostToFrame_KRKi, const Vec3f& hostToFrame_Kt,
 *                                              const Vec2f& hostToFrame_affine, CalibHessian*, bool)           FullSystem/ImmaturePoint.cpp:50-352
 *   called for every immature point of every active keyframe by FullSystem::traceNewCoarse                     FullSystem/FullSystem.cpp:519-552
 * The record mirrors the members of ImmaturePoint the two functions read or write (ImmaturePoint.h:33-78); it stays with the caller (the reference keeps the
 * points in host->immaturePoints), the device works on a copy per call.  lastTraceStatus: 0 GOOD, 1 OOB, 2 OUTLIER, 3 SKIPPED, 4 BADCONDITION, 5 UNINITIALIZED. */
typedef struct {
  float u, v, idepth_min, idepth_max;
  float color[8], weights[8], gradH[4];           /* gradH: Mat22f row-major */
  float energyTH, quality, lastTraceUV[2], lastTracePixelInterval;
  int32_t lastTraceStatus;
} sdv_immature_pt;
/* constructor for n candidates at integer pixels uv[2n] of a resident keyframe (>= 3 px from the border, else SDV_ERR_ARG) */
int  sdv_immature_init(sdv_ctx* c, uint64_t host_frame, int n, const int32_t* uv, sdv_immature_pt* out);
/* traceOn for n_groups (host keyframe, traced frame) pairs in ONE launch: group g traces pts_io[pt_begin[g] .. pt_begin[g+1]) against frames[g] with
 * KRKi9[9g..] = K R K^-1 (row-major), Kt3[3g..] = K t of hostToNew and aff2[2g..] = AffLight::fromToVecExposure(host, new) — what traceNewCoarse computes per host
 * (:532-538).  Groups may belong to different sequences (batched mode).  pts_io is updated in place; status_out (n points) may be NULL. */
int  sdv_immature_trace_batch(sdv_ctx* c, int n_groups, const uint64_t* frames, const int32_t* pt_begin, const float* KRKi9, const float* Kt3, const float* aff2,
                              sdv_immature_pt* pts_io, int32_t* status_out);
/* Activation — PointHessian* FullSystem::optimizeImmaturePoint(ImmaturePoint*, int minObs, ImmaturePointTemporaryResidual*)   FullSystemOptPoint.cpp:18-183
 * over ImmaturePoint::linearizeResidual (ImmaturePoint.cpp:410-476): Gauss-Newton on the inverse depth of a candidate against every other keyframe of the window.
 * Group g = the candidates pts[pt_begin[g] .. pt_begin[g+1]) of ONE host keyframe; its targets are entries tgt_begin[g] .. tgt_begin[g+1]) of target_frames / pre14
 * (frameHessians without the host, window order; pre14 = host->targetPrecalc[target]: PRE_RTll[9] row-major, PRE_tTll[3], PRE_aff_mode[2]); calib6[6g..] = fxl fyl cxl cyl
 * fxli fyli of the CalibHessian.  Outputs per candidate: status 0 = not well constrained (stays immature; the function's `return 0`), -1 = outlier / non-finite
 * (deleted), 1 = activated with idepth_out (setIdepth/setIdepthZero) and res_state_out[k*res_stride + i] = final state of the temporary residual to target i
 * (0 IN -> a PointFrameResidual is created, 1 OOB, 2 OUTLIER; -1 = unused).  The PointHessian / PointFrameResidual objects stay with the caller. */
int  sdv_immature_optimize_batch(sdv_ctx* c, int n_groups, const int32_t* pt_begin, const int32_t* tgt_begin, const uint64_t* target_frames, const float* pre14, const float* calib6,
                                 int min_obs, const sdv_immature_pt* pts, const uint8_t* is_from_sensor, int res_stride, int32_t* status_out, float* idepth_out, int32_t* res_state_out);

/* frames + calibration + marginalisation prior:  EnergyFunctional::insertFrame :365-398, setAdjointsF :21-71, HM/bM :88-89.
 * state10 / state_zero10 = FrameHessian::state / state_zero (Vec10: 6 pose, a, b, 2 unused; HessianBlocks.h:141-175),
 * T_evalPT7 = worldToCam_evalPT, calib_value_scaled = CalibHessian::value_scaled {fx,fy,cx,cy}; HM (dim x dim, row-major), bM may be NULL. */
/* =====================================================================================================================
 * Sliding-window back-end: EnergyFunctional + FullSystem::optimize on a FLATTENED window.
 * The reference's pointer graph (FrameHessian -> PointHessian -> PointFrameResidual and the EF* mirrors) stays on the host;
 * the caller passes it in the reference's own iteration order — frames = ef->frames, points = ef->allPoints (contiguous per
 * host frame, EnergyFunctional.cpp:761-782), residuals grouped per point in residualsAll order (res_begin is the CSR row
 * pointer).  This is the makeIDX / insertFrame / insertPoint / insertResidual surface (EnergyFunctional.h:51-72) in one call.
 * ===================================================================================================================== */
int  sdv_ba_set_window(sdv_ctx* c, int nF, const uint64_t* frame_ids, const double* T_evalPT7, const double* state10, const double* state_zero10,
                       const float* ab_exposure, const int32_t* frameID, const float* frameEnergyTH, const double calib_value_scaled[4],
                       const double* HM, const double* bM);
/* CalibHessian::value_zero (HessianBlocks.h:273-289) of a live system: sdv_ba_set_window takes value_zero = value (a fresh CalibHessian); after the first bundle adjustment the two
 * differ and value - value_zero is what the marginalisation prior acts on (EnergyFunctional.cpp:144).  Optional; between sdv_ba_set_window and sdv_ba_set_points.  Units of
 * CalibHessian::value (= value_scaled / SCALE_F resp. SCALE_C).  Oracle side: orc_ba_set_calib_zero, exercised on live windows by tests/test_sequence_ba.py. */
int  sdv_ba_set_calib_zero(sdv_ctx* c, const double value_zero[4]);
/* points + residuals: PointHessian {u,v,idepth,idepth_zero,color[8],weights[8],hasDepthPrior,isFromSensor} (HessianBlocks.h:361-465),
 * PointFrameResidual {host,target,hasMatcher,matcher,isNew} (Residuals.h:30-83).  Runs setPrecalcValues (FullSystem.cpp:1358-1368). */
int  sdv_ba_set_points(sdv_ctx* c, int nP, const float* uv, const float* idepth, const float* idepth_zero, const float* color8, const float* weights8,
                       const int32_t* host, const int32_t* hasDepthPrior, const int32_t* isFromSensor, const int32_t* res_begin,
                       int nR, const int32_t* r_point, const int32_t* r_host, const int32_t* r_target, const int32_t* r_hasMatcher,
                       const float* r_matcher, const int32_t* r_isNew);
/* Vec3 FullSystem::linearizeAll(bool fixLinearization)  FullSystemOptimize.cpp:99-159 -> PointFrameResidual::linearize Residuals.cpp:60-224
 * (+ setNewFrameEnergyTH :63-97; with fix: applyRes + the isNew / toRemove bookkeeping of linearizeAll_Reductor :23-55) */
int  sdv_ba_reset_oob(sdv_ctx* c);                                 /* PointFrameResidual::resetOOB on every residual (optimize :361-362) */
int  sdv_ba_linearize(sdv_ctx* c, int fix, double* energy_out);
int  sdv_ba_apply_res(sdv_ctx* c);                                 /* applyRes_Reductor :57-61 -> Residuals.cpp:252-274, EFResidual::takeDataF */
int  sdv_ba_energy(sdv_ctx* c, double* EL_out, double* EM_out);    /* EnergyFunctional::calcLEnergyF_MT :333-350, calcMEnergyF :284-293 */
/* void EnergyFunctional::solveSystemF(int iteration, double lambda, CalibHessian*)  EnergyFunctional.cpp:650-759
 * (accumulateAF/LF/SCF, stitch, diag-scaled LDLT, orthogonalize for iteration>=2, resubstituteF_MT); x_out[CPARS+6nF] = ef->lastX */
int  sdv_ba_solve(sdv_ctx* c, int iteration, double lambda, double* x_out);
int  sdv_ba_backup(sdv_ctx* c);                                    /* FullSystem::backupState  FullSystemOptimize.cpp:255-300 */
int  sdv_ba_step(sdv_ctx* c, float stepfac, int load_backup, int* canbreak_out);   /* doStepFromBackup :165-250 / loadSateBackup :303-321 */
/* float FullSystem::optimize(int mnumOptIts)  FullSystemOptimize.cpp:344-502 — the whole GN loop incl. the final re-anchoring and
 * linearizeAll(true).  Returns sqrt(lastEnergy / resInA) like the reference. */
int  sdv_ba_optimize(sdv_ctx* c, int mnumOptIts, float* rmse_out, int32_t* iterations_out, int32_t* accepts_out);
float sdv_ba_last_kernel_ms(sdv_ctx* c);                        /* device time of the last sdv_ba_optimize / sdv_ba_optimize_batch (back-end stream) */
/* Batched mode: a context holds any number of independent windows (one per resident sequence).  sdv_ba_select picks the window the
 * set_ / get_ / step-wise calls address (default 0); sdv_ba_optimize_batch runs FullSystem::optimize on n windows at once — every
 * kernel is launched once for all windows and the accept/reject/break decisions are taken on the device. */
int  sdv_ba_select(sdv_ctx* c, int window);
int  sdv_ba_clear(sdv_ctx* c);               /* empty the selected window and drop its references to keyframe images (see sdv_map_clear) */
int  sdv_ba_optimize_batch(sdv_ctx* c, int n, const int32_t* windows, int mnumOptIts, float* rmse_out, int32_t* iterations_out, int32_t* accepts_out);
/* read-back of what the reference leaves in FrameHessian/CalibHessian, PointHessian/EFPoint, PointFrameResidual/EFResidual, EnergyFunctional */
int  sdv_ba_get_frames(sdv_ctx* c, double* T_evalPT7, double* state10, double* step10, float* frameEnergyTH, double* PRE_worldToCam7,
                       double calib_value[4], double calib_step[4]);
int  sdv_ba_get_points(sdv_ctx* c, float* idepth, float* step, float* HdiF, float* bdSumF, float* maxRelBaseline, int32_t* numGoodResiduals, float* idepth_hessian);
/* J24 per residual: {resF[2], Jpdxi[0][6], Jpdxi[1][6], Jpdc[0][4], Jpdc[1][4], Jpdd[2]} = the live part of RawResidualJacobian.h:7-36;
 * energies3 = {state_energy, state_NewEnergy, state_NewEnergyWithOutlier}; states: 0 IN, 1 OOB, 2 OUTLIER (Residuals.h:21) */
int  sdv_ba_get_residuals(sdv_ctx* c, int32_t* state_state, int32_t* state_NewState, float* energies3, int32_t* isActive, float* J24, float* efJ24,
                          float* JpJdF8, float* center3, int32_t* toRemove);
/* ---- keyframe hand-over: the numeric part of FullSystem::makeKeyFrame after optimize()      FullSystem.cpp:1152-1171
 * sdv_ba_flag_points       FullSystem::flagPointsForRemoval :764-797 for the points the host selected
 *                          (selected[p] = (ph->isOOB(..) || host->flaggedForMarginalization) && ph->isInlierNew(), graph bookkeeping):
 *                          resetOOB + linearize + applyRes(true) + EFResidual::fixLinearizationF (EnergyFunctionalStructs.cpp:46-55) per residual;
 *                          status[p] = 0 untouched / 1 PS_DROP / 2 PS_MARGINALIZE (idepth_hessian > setting_minIdepthH_marg).
 * sdv_ba_marginalize_points EnergyFunctional::marginalizePointsF, EnergyFunctional.cpp:514-576 (addPoint<2>, addPoint(p,false), stitchDouble,
 *                          HM += margWeightFac*(M-Msc)); status==NULL uses the device-resident result of sdv_ba_flag_points.  The accumulation
 *                          order is the order of the status==2 points in the flattened list (= allPointsToMarg when flattened after dropPointsF).
 *                          M, Mb, Msc, Mbsc can be read through sdv_ba_get_system (HA, bA, Hsc, bsc slots).
 * sdv_ba_marginalize_frame EnergyFunctional::marginalizeFrame, EnergyFunctional.cpp:434-512: Schur-eliminates frame idx from (HM,bM) and drops it
 *                          from the window; points/residuals must be re-set (sdv_ba_set_points) before the window is used again.
 * sdv_ba_get_prior         reads (HM,bM) of the current window, dim = 4 + 6*nF (row-major). */
int  sdv_ba_flag_points(sdv_ctx* c, const int32_t* selected, int32_t* status_out);
int  sdv_ba_marginalize_points(sdv_ctx* c, const int32_t* status);
int  sdv_ba_marginalize_frame(sdv_ctx* c, int idx);
int  sdv_ba_get_prior(sdv_ctx* c, int* dim, double* HM, double* bM);
int  sdv_ba_get_linearized(sdv_ctx* c, float* res_toZero2, int32_t* isLinearized);
int  sdv_ba_get_system(sdv_ctx* c, double* HA, double* bA, double* Hsc, double* bsc, double* lastHS, double* lastbS);
int  sdv_ba_get_precalc(sdv_ctx* c, int host, int target, float out27[27], double adHost36[36], double adTarget36[36], float adHTdelta6[6]);

/* =====================================================================================================================
 * Candidate management at keyframe rate (SURVEY.md §8f rank 4 and the caller half of rank 2).  Tracker call domain.
 *   PixelSelector                         FullSystem/PixelSelector2.cpp:11-26 (ctor), :47-106 makeHists, :354-622 makeMapsFromLidar / selectFromLidar,
 *                                         :108-352 makeMaps / select
 *   FullSystem::makeNewTraces             FullSystem/FullSystem.cpp:1273-1356 (+ shiTomasiScore :1540-1583, setMask :1261-1271, ImmaturePoint ctor)
 *   CoarseDistanceMap, activatePointsMT   FullSystem/CoarseTracker.cpp:1139-1282, FullSystem/FullSystem.cpp:600-671
 * A selector SLOT stands for one PixelSelector / FullSystem::selectionMap pair, i.e. one resident sequence: it carries currentPotential (initially 3) and the
 * persistent monocular selection map (w*h bytes, values 0 / 1 / 2 / 4 — the reference's float map; zero at birth where the reference's is uninitialised).
 * random_pattern: the w*h bytes of PixelSelector::randomPattern — the caller generates them exactly like the constructor does (srand(3141592); rand() & 0xFF), so the
 * library never touches the process-wide rand() stream.  Reads of thsSmoothed past its (h/32) rows — the reference indexes its uninitialised tail for the last
 * image rows — see zeros here.  LiDAR pixels are rows {Ku, Kv, depth} of doubles, as main.cpp:810-848 pushes them into FullSystem::qCloudPixel. */
typedef struct { float u, v, my_type, score, idepth_fromSensor; int32_t isFromSensor, type; } sdv_new_trace;   /* type: 0 CORNER, 1 EDGELET (ImmaturePoint.h), -1 not assigned (monocular points) */
int  sdv_selector_init(sdv_ctx* c, const uint8_t* random_pattern, int n_slots);
int  sdv_selector_potential(sdv_ctx* c, int slot, int set_to /* <= 0: leave */, int* out);
int  sdv_selector_get_map(sdv_ctx* c, int slot, uint8_t* out_wh);
/* void PixelSelector::makeHists(const FrameHessian*): ths / thsSmoothed of the (w/32) x (h/32) blocks */
int  sdv_selector_make_hists(sdv_ctx* c, uint64_t frame, float* ths_out, float* thsSmoothed_out);
/* int PixelSelector::makeMapsFromLidar(fh, map_out, density, recursionsLeft, plot, thFactor, vCloudPixel) for n (slot, frame) pairs at once when cloud_begin != NULL
 * (job j owns cloud rows cloud_begin[j] .. cloud_begin[j+1]; maps_out = the n maps back to back, one byte per cloud row), and
 * int PixelSelector::makeMaps(fh, map_out, density, recursionsLeft, plot, thFactor) when cloud_begin == NULL (maps_out = n maps of w*h bytes; the slot's persistent map
 * is the one written).  makeHists runs first, like `if(fh != gradHistFrame) makeHists(fh)`.  num_have_out[j] = the return value; the slot's currentPotential is updated. */
int  sdv_selector_make_maps_batch(sdv_ctx* c, int n, const int32_t* slots, const uint64_t* frames, const int32_t* cloud_begin, const double* cloud3, const float* density,
                                  const int32_t* recursions_left, const float* th_factor, uint8_t* maps_out, int32_t* num_have_out);
/* void FullSystem::makeNewTraces(FrameHessian* newFrame, float*) for n new keyframes (one per sequence / slot) in one call.  density_lidar[j] =
 * ((float)lidarArea/(float)imageArea) * setting_desiredImmatureDensity (:1287-1290), density_dense[j] = setting_desiredImmatureDensity (:1293), add_feature_point[j] =
 * FullSystem::addFeaturePoint (off: the slot's map of an EARLIER keyframe is walked again, as in the reference).  Per job up to `cap` points come back in creation order
 * (LiDAR points in cloud order, then monocular points in raster order): out[j*cap ..] and, if imm_out != NULL, the constructed ImmaturePoint records imm_out[j*cap ..];
 * n_out[j] = how many; num_points2[2j..] = {numPointLidar, numPointMonocular}.  SDV_ERR_CAPACITY when cap is too small. */
int  sdv_make_new_traces_batch(sdv_ctx* c, int n, const int32_t* slots, const uint64_t* frames, const int32_t* cloud_begin, const double* cloud3, const float* density_lidar,
                               const float* density_dense, const int32_t* add_feature_point, int cap, sdv_new_trace* out, sdv_immature_pt* imm_out, int32_t* n_out, int32_t* num_points2);
/* The selection half of void FullSystem::activatePointsMT() for n sequences: CoarseDistanceMap::makeDistanceMap from the ACTIVE points of the other keyframes, then the
 * candidate walk :600-671 with addIntoDistFinal after every accept.  Sequence j owns source keyframes host_begin[j] .. host_begin[j+1] (their points: uvid rows
 * pt_begin[k] .. pt_begin[k+1] = {u, v, idepth_scaled}; KRKi9 / Kt3 per keyframe = K[1] R K[0]^-1, K[1] t of host -> newest, FullSystem.cpp:606-608) and candidate
 * keyframes cand_host_begin[j] .. (their candidates: cand4 rows {u, v, 0.5f*(idepth_max+idepth_min), my_type} that passed the deletion / canActivate tests of
 * :618-645, which are pointer-graph bookkeeping).  decision_out per candidate: 1 = handed to optimizeImmaturePoint (sdv_immature_optimize_batch), 0 = too close to
 * existing points, -1 = projects outside the level-1 image (deleted).  dist_map_out (optional): n maps of (w/2)*(h/2) floats = fwdWarpedIDDistFinal after the walk. */
int  sdv_activate_select_batch(sdv_ctx* c, int n, const int32_t* host_begin, const int32_t* pt_begin, const float* KRKi9, const float* Kt3, const float* uvid,
                               const int32_t* cand_host_begin, const int32_t* cand_begin, const float* cKRKi9, const float* cKt3, const float* cand4, const float* min_act_dist,
                               int32_t* decision_out, float* dist_map_out);

/* =====================================================================================================================
 * LiDAR front-end of the node (second half): void lidarCloudHandler(const sensor_msgs::PointCloud2ConstPtr&)   src/main.cpp:785-858
 *   -> projectPointCloud :563-607, groundRemoval :609-655, cloudSegmentation / labelComponents :657-783, projection into the image :806-849.  Tracker call domain.
 * sdv_lidar_init: the sensor constants of main.cpp:103-108 (N_SCAN, Horizon_SCAN, ang_res_x, ang_res_y, ang_bottom, groundScanInd; Velodyne-64: 64, 1800, 0.2, 0.427, 24.9, 50).
 * sdv_lidar_handler_batch: n raw sweeps (one per resident sequence), sweep j = XYZI rows sweep_begin[j] .. sweep_begin[j+1] of xyzi (the decoded PointCloud2, what
 * pcl::fromROSMsg yields; rows with a non-finite coordinate are dropped like pcl::removeNaNFromPointCloud does).  Per sweep: Rlc9 (row-major) / tlc3 = FullSystem::Rlc, tlc
 * (LiDAR -> camera), K4 = FullSystem::fx fy cx cy, lrud_io = FullSystem::left, right, up, down (the running pixel box; 10000, -1, 10000, -1 at start).  The image
 * size is the context's.  Out, per sweep: up to cap rows {Ku, Kv, depth} (doubles, the vCloudPixel pushed into FullSystem::qCloudPixel, same order) at
 * cloud3_out[3*j*cap ..], n_out[j] rows, add_feature_point_out[j] = FullSystem::addFeaturePoint (ground ratio > 0.8), stats_out[2j..] = {numGround, size of
 * segmentedCloud} (may be NULL).  The rows feed sdv_make_new_traces_batch and the LiDAR-depth splats of sdv_tracker_set_ref. */
int  sdv_lidar_init(sdv_ctx* c, int n_scan, int horizon_scan, float ang_res_x, float ang_res_y, float ang_bottom, int ground_scan_ind);
int  sdv_lidar_handler_batch(sdv_ctx* c, int n, const int32_t* sweep_begin, const float* xyzi, const double* Rlc9, const double* tlc3, const float* K4, int32_t* lrud_io, int cap,
                             double* cloud3_out, int32_t* n_out, int32_t* add_feature_point_out, int32_t* stats_out);

#ifdef __cplusplus
}
#endif
#endif /* SDV_B200_H */
