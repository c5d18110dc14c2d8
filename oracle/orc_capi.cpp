// ORACLE — TEST INFRASTRUCTURE ONLY.  Parity status: orc_math.hpp.
// Flat C entry points over the oracle classes so tests/ and bench.py (cpu_baseline / --impl reference) can
// drive them through ctypes.  Pose layout everywhere: double T[7] = {qw,qx,qy,qz, tx,ty,tz}.
#include "orc_tracker.hpp"
#include <chrono>

using namespace orc;

static SE3 se3_from(const double T[7]) { SE3 s; s.q = Quat{T[0],T[1],T[2],T[3]}; s.t = Vec3d{{T[4],T[5],T[6]}}; return s; }
static void se3_to(const SE3& s, double T[7]) { T[0]=s.q.w; T[1]=s.q.x; T[2]=s.q.y; T[3]=s.q.z; T[4]=s.t.v[0]; T[5]=s.t.v[1]; T[6]=s.t.v[2]; }

extern "C" {

// ---- math KATs
void orc_se3_exp(const double a[6], double T[7]) { se3_to(SE3::exp(a), T); }
void orc_se3_log(const double T[7], double a[6]) { se3_from(T).log(a); }
void orc_se3_mul(const double A[7], const double B[7], double C[7]) { se3_to(se3_from(A)*se3_from(B), C); }
void orc_se3_inv(const double A[7], double C[7]) { se3_to(se3_from(A).inverse(), C); }
void orc_se3_rot(const double A[7], double R[9]) { Mat33d M = se3_from(A).rotationMatrix(); for(int i=0;i<3;i++) for(int j=0;j<3;j++) R[i*3+j]=M.m[i][j]; }
void orc_se3_adj(const double A[7], double Ad[36]) { double M[6][6]; se3_from(A).Adj(M); for(int i=0;i<6;i++) for(int j=0;j<6;j++) Ad[i*6+j]=M[i][j]; }
void orc_se3_from_rt(const double R[9], const double t[3], double T[7]) {
  Mat33d M; for(int i=0;i<3;i++) for(int j=0;j<3;j++) M.m[i][j]=R[i*3+j];
  SE3 s = SE3::fromQuatT(qfrommat(M), Vec3d{{t[0],t[1],t[2]}}); se3_to(s,T);
}
void orc_ldlt_solve(int n, const double* A, const double* b, double* x) { ldlt_solve<64>(n, A, b, x); }
void orc_aff_from_to(float eF, float eT, double aF, double bF, double aT, double bT, double out[2]) {
  AffLight f; f.a=aF; f.b=bF; AffLight t; t.a=aT; t.b=bT; fromToVecExposure(eF,eT,f,t,out);
}
int orc_pyr_levels(int w, int h) { return pyrLevelsUsedFor(w,h); }

// ---- frames
void* orc_frame_create(const float* color, int w, int h, int levels, float exposure) {
  Frame* f = new Frame(); f->makeImages(color, w, h, levels); f->ab_exposure = exposure; return f;
}
void orc_frame_destroy(void* f) { delete (Frame*)f; }
const float* orc_frame_dI(void* f, int lvl) { return ((Frame*)f)->dIp[lvl].data(); }
const float* orc_frame_abs(void* f, int lvl) { return ((Frame*)f)->absSquaredGrad[lvl].data(); }

// ---- tracker
void* orc_tracker_create(int w, int h, int levels, float fx, float fy, float cx, float cy) {
  CoarseTracker* t = new CoarseTracker(); t->init(w,h,levels); t->makeK(fx,fy,cx,cy); return t;
}
void orc_tracker_destroy(void* t) { delete (CoarseTracker*)t; }
void orc_tracker_settings(void* t, float huberTH, float coarseCutoffTH, float affA, float affB) {
  CoarseTracker* T=(CoarseTracker*)t; T->set.huberTH=huberTH; T->set.coarseCutoffTH=coarseCutoffTH; T->set.affineOptModeA=affA; T->set.affineOptModeB=affB;
}
void orc_tracker_get_K(void* t, int lvl, float out[4]) { CoarseTracker* T=(CoarseTracker*)t; out[0]=T->fx[lvl]; out[1]=T->fy[lvl]; out[2]=T->cx[lvl]; out[3]=T->cy[lvl]; }
void orc_tracker_get_Ki(void* t, int lvl, float out[9]) { CoarseTracker* T=(CoarseTracker*)t; for(int i=0;i<3;i++) for(int j=0;j<3;j++) out[i*3+j]=T->Ki[lvl].m[i][j]; }
// pts: n x {u,v,idepth,HdiF} floats, round_half: n ints
void orc_tracker_set_ref(void* t, void* ref_frame, const float* pts, const int* round_half, int n, double ref_a, double ref_b) {
  std::vector<RefPoint> v(n);
  for (int i=0;i<n;i++) { v[i].u=pts[4*i]; v[i].v=pts[4*i+1]; v[i].idepth=pts[4*i+2]; v[i].HdiF=pts[4*i+3]; v[i].round_half=round_half[i]; }
  AffLight a; a.a=ref_a; a.b=ref_b;
  ((CoarseTracker*)t)->setCoarseTrackingRef((Frame*)ref_frame, v.data(), n, a);
}
void orc_tracker_set_cloud(void* t, void* ref_frame, int lvl, int n, const float* u, const float* v, const float* id, const float* color, double ref_a, double ref_b) {
  CoarseTracker* T=(CoarseTracker*)t; T->setRefCloud((Frame*)ref_frame, lvl, n, u, v, id, color); T->lastRef_aff_g2l.a=ref_a; T->lastRef_aff_g2l.b=ref_b;
}
int orc_tracker_cloud_n(void* t, int lvl) { return ((CoarseTracker*)t)->pc_n[lvl]; }
void orc_tracker_get_cloud(void* t, int lvl, float* u, float* v, float* id, float* color) {
  CoarseTracker* T=(CoarseTracker*)t; int n=T->pc_n[lvl];
  for (int i=0;i<n;i++) { u[i]=T->pc_u[lvl][i]; v[i]=T->pc_v[lvl][i]; id[i]=T->pc_idepth[lvl][i]; color[i]=T->pc_color[lvl][i]; }
}
void orc_tracker_calc_res(void* t, void* new_frame, int lvl, const double T7[7], double a, double b, float cutoffTH, double rs[6]) {
  CoarseTracker* T=(CoarseTracker*)t; T->newFrame=(Frame*)new_frame; AffLight aff; aff.a=a; aff.b=b;
  T->calcRes(lvl, se3_from(T7), aff, cutoffTH, rs);
}
int orc_tracker_warped_n(void* t) { return ((CoarseTracker*)t)->buf_warped_n; }
// out: 8 x n floats {idepth,u,v,dx,dy,residual,weight,refColor}
void orc_tracker_get_warped(void* t, float* out) {
  CoarseTracker* T=(CoarseTracker*)t; int n=T->buf_warped_n;
  const std::vector<float>* bufs[8] = {&T->buf_warped_idepth,&T->buf_warped_u,&T->buf_warped_v,&T->buf_warped_dx,&T->buf_warped_dy,&T->buf_warped_residual,&T->buf_warped_weight,&T->buf_warped_refColor};
  for (int k=0;k<8;k++) for (int i=0;i<n;i++) out[k*n+i]=(*bufs[k])[i];
}
void orc_tracker_calc_gs(void* t, int lvl, const double T7[7], double a, double b, double H[64], double bb[8]) {
  AffLight aff; aff.a=a; aff.b=b; ((CoarseTracker*)t)->calcGSSSE(lvl, H, bb, se3_from(T7), aff);
}
// returns good flag; stats: evals[6] (point evaluations per level), its[6], accepts[6]
int orc_tracker_track(void* t, void* new_frame, double T_io[7], double ab_io[2], int coarsest, const double minRes[5],
                      double lastRes[5], double flow[3], long long* evals, int* its, int* accepts) {
  CoarseTracker* T=(CoarseTracker*)t; SE3 s = se3_from(T_io); AffLight aff; aff.a=ab_io[0]; aff.b=ab_io[1];
  bool good = T->trackNewestCoarse((Frame*)new_frame, s, aff, coarsest, minRes);
  se3_to(s, T_io); ab_io[0]=aff.a; ab_io[1]=aff.b;
  for (int i=0;i<5;i++) lastRes[i]=T->lastResiduals[i];
  for (int i=0;i<3;i++) flow[i]=T->lastFlowIndicators[i];
  if (evals) for (int i=0;i<PYR_LEVELS;i++) evals[i]=T->evals[i];
  if (its) for (int i=0;i<PYR_LEVELS;i++) its[i]=T->iterations[i];
  if (accepts) for (int i=0;i<PYR_LEVELS;i++) accepts[i]=T->accepts[i];
  return good ? 1 : 0;
}

} // extern "C"

// ================================================================================================ back-end (BA) window
#include "orc_ba.hpp"
extern "C" {
void* orc_ba_create(int w, int h) { BAWindow* b = new BAWindow(); b->w=w; b->h=h; return b; }
void orc_ba_destroy(void* p) { delete (BAWindow*)p; }
void orc_ba_set_calib(void* p, const double vs[4]) { ((BAWindow*)p)->setCalibScaled(vs); }
// CalibHessian::value_zero of a LIVE system stays at the initial intrinsics (HessianBlocks.h:287) while value moves with every bundle adjustment
void orc_ba_set_calib_zero(void* p, const double vz[4]) { BAWindow* b=(BAWindow*)p; for (int i=0;i<4;i++) { b->c_value_zero[i]=vz[i]; b->c_vmvz[i]=b->c_value[i]-vz[i]; } }
void orc_ba_add_frame(void* p, void* img, const double T_eval[7], const double state[10], const double state_zero[10], float ab_exposure, int frameID, float frameEnergyTH) {
  BAWindow* b=(BAWindow*)p; BAFrame f; f.worldToCam_evalPT = se3_from(T_eval);
  for (int i=0;i<10;i++) { f.state[i]=state[i]; f.state_zero[i]=state_zero[i]; f.state_backup[i]=state[i]; f.step[i]=0; }
  f.ab_exposure=ab_exposure; f.frameID=frameID; f.frameEnergyTH=frameEnergyTH; f.img=(Frame*)img; b->frames.push_back(f);
}
void orc_ba_set_points(void* p, int nP, const float* uv, const float* idepth, const float* idepth_zero, const float* color, const float* weights,
                       const int* host, const int* hasDepthPrior, const int* isFromSensor, const int* res_begin) {
  BAWindow* b=(BAWindow*)p; b->points.assign(nP, BAPoint());
  for (int i=0;i<nP;i++) { BAPoint& q=b->points[i]; q.u=uv[2*i]; q.v=uv[2*i+1]; q.idepth=idepth[i]; q.idepth_zero=idepth_zero[i];
    for (int k=0;k<8;k++) { q.color[k]=color[8*i+k]; q.weights[k]=weights[8*i+k]; }
    q.host=host[i]; q.hasDepthPrior=hasDepthPrior[i]; q.isFromSensor=isFromSensor[i]; q.res_begin=res_begin[i]; q.res_end=res_begin[i+1]; }
}
void orc_ba_set_residuals(void* p, int nR, const int* point, const int* host, const int* target, const int* hasMatcher, const float* matcher, const int* isNew) {
  BAWindow* b=(BAWindow*)p; b->res.assign(nR, BARes());
  for (int i=0;i<nR;i++) { BARes& r=b->res[i]; std::memset(&r.J,0,sizeof(RawJ)); std::memset(&r.efJ,0,sizeof(RawJ)); std::memset(r.JpJdF,0,sizeof(r.JpJdF));
    std::memset(r.centerProjectedTo,0,sizeof(r.centerProjectedTo)); std::memset(r.projectedTo,0,sizeof(r.projectedTo));
    r.point=point[i]; r.host=host[i]; r.target=target[i]; r.hasMatcher=hasMatcher[i]; r.matcher[0]=matcher[2*i]; r.matcher[1]=matcher[2*i+1]; r.isNew=isNew[i]; }
}
void orc_ba_set_prior(void* p, const double* HM, const double* bM) { BAWindow* b=(BAWindow*)p; int n=b->dim(); b->HM.assign(HM,HM+(size_t)n*n); b->bM.assign(bM,bM+n); }
void orc_ba_init(void* p) { ((BAWindow*)p)->init(); }
void orc_ba_reset_oob(void* p) { for (auto& r : ((BAWindow*)p)->res) { r.state_NewEnergy=r.state_energy=0; r.state_NewState=RS_OUTLIER; r.state_state=RS_IN; } }
double orc_ba_linearize_all(void* p, int fix) { return ((BAWindow*)p)->linearizeAll(fix!=0); }
void orc_ba_apply_res(void* p) { BAWindow* b=(BAWindow*)p; for (auto& r : b->res) b->applyRes(r); }
double orc_ba_energy_L(void* p) { return ((BAWindow*)p)->calcLEnergy(); }
double orc_ba_energy_M(void* p) { return ((BAWindow*)p)->calcMEnergy(); }
// J layout per residual: 24 floats {resF[2], Jpdxi[0][6], Jpdxi[1][6], Jpdc[0][4], Jpdc[1][4], Jpdd[2]}
static void packJ(const RawJ& J, float* o) { o[0]=J.resF[0]; o[1]=J.resF[1]; for(int i=0;i<6;i++){o[2+i]=J.Jpdxi[0][i]; o[8+i]=J.Jpdxi[1][i];} for(int i=0;i<4;i++){o[14+i]=J.Jpdc[0][i]; o[18+i]=J.Jpdc[1][i];} o[22]=J.Jpdd[0]; o[23]=J.Jpdd[1]; }
void orc_ba_get_residuals(void* p, int* state_state, int* state_NewState, double* energies3, int* isActive, float* J24, float* efJ24, float* JpJdF8, float* center3, int* toRemove) {
  BAWindow* b=(BAWindow*)p; int n=(int)b->res.size();
  for (int i=0;i<n;i++) { const BARes& r=b->res[i]; state_state[i]=r.state_state; state_NewState[i]=r.state_NewState;
    energies3[3*i]=r.state_energy; energies3[3*i+1]=r.state_NewEnergy; energies3[3*i+2]=r.state_NewEnergyWithOutlier; isActive[i]=r.isActive;
    packJ(r.J, J24+24*i); packJ(r.efJ, efJ24+24*i); for(int k=0;k<8;k++) JpJdF8[8*i+k]=r.JpJdF[k]; for(int k=0;k<3;k++) center3[3*i+k]=r.centerProjectedTo[k]; toRemove[i]=r.toRemove; }
}
void orc_ba_accumulate(void* p, double* HA, double* bA, double* Hsc, double* bsc) {
  BAWindow* b=(BAWindow*)p; std::vector<double> h1,b1,h2,b2; b->accumulateA(h1,b1); b->accumulateSC(h2,b2);
  std::copy(h1.begin(),h1.end(),HA); std::copy(b1.begin(),b1.end(),bA); std::copy(h2.begin(),h2.end(),Hsc); std::copy(b2.begin(),b2.end(),bsc);
}
void orc_ba_solve(void* p, int iteration, double lambda, double* x, double* HS, double* bS) {
  BAWindow* b=(BAWindow*)p; b->solveSystem(iteration, lambda); int n=b->dim();
  std::copy(b->lastX.begin(), b->lastX.end(), x); if (HS) std::copy(b->lastHS.begin(), b->lastHS.end(), HS); if (bS) std::copy(b->lastbS.begin(), b->lastbS.end(), bS); (void)n;
}
void orc_ba_backup(void* p) { ((BAWindow*)p)->backupState(); }
int  orc_ba_do_step(void* p, float stepfac) { return ((BAWindow*)p)->doStepFromBackup(stepfac) ? 1 : 0; }
void orc_ba_load_backup(void* p) { ((BAWindow*)p)->loadStateBackup(); }
void orc_ba_get_points(void* p, float* idepth, float* step, float* HdiF, float* bdSumF, float* maxRelBaseline, int* numGood, float* idepth_hessian) {
  BAWindow* b=(BAWindow*)p; int n=(int)b->points.size();
  for (int i=0;i<n;i++) { const BAPoint& q=b->points[i]; idepth[i]=q.idepth; step[i]=q.step; HdiF[i]=q.HdiF; bdSumF[i]=q.bdSumF; maxRelBaseline[i]=q.maxRelBaseline; numGood[i]=q.numGoodResiduals; idepth_hessian[i]=q.idepth_hessian; }
}
void orc_ba_get_frames(void* p, double* T_eval7, double* state10, double* step10, float* frameEnergyTH, double* PRE_w2c7) {
  BAWindow* b=(BAWindow*)p; int n=b->nF();
  for (int i=0;i<n;i++) { const BAFrame& f=b->frames[i]; se3_to(f.worldToCam_evalPT, T_eval7+7*i); se3_to(f.PRE_worldToCam, PRE_w2c7+7*i);
    for (int k=0;k<10;k++) { state10[10*i+k]=f.state[k]; step10[10*i+k]=f.step[k]; } frameEnergyTH[i]=f.frameEnergyTH; }
}
void orc_ba_get_calib(void* p, double value[4], double step[4]) { BAWindow* b=(BAWindow*)p; for (int i=0;i<4;i++) { value[i]=b->c_value[i]; step[i]=b->c_step[i]; } }
void orc_ba_get_precalc(void* p, int host, int target, float* out /*9 KRKi, 3 Kt, 9 R0, 3 t0, 2 aff, 1 b0 = 27*/, double* adH36, double* adT36, float* adHTdelta6) {
  BAWindow* b=(BAWindow*)p; int n=b->nF(); const Precalc& c=b->precalc[(size_t)host*n+target];
  for(int i=0;i<3;i++) for(int j=0;j<3;j++) { out[i*3+j]=c.PRE_KRKiTll.m[i][j]; out[12+i*3+j]=c.PRE_RTll_0.m[i][j]; }
  for(int i=0;i<3;i++) { out[9+i]=c.PRE_KtTll.v[i]; out[21+i]=c.PRE_tTll_0.v[i]; } out[24]=c.PRE_aff_mode[0]; out[25]=c.PRE_aff_mode[1]; out[26]=c.PRE_b0_mode;
  int idx=host+target*n; for(int i=0;i<36;i++) { adH36[i]=b->adHost[(size_t)idx*36+i]; adT36[i]=b->adTarget[(size_t)idx*36+i]; } for(int i=0;i<6;i++) adHTdelta6[i]=b->adHTdeltaF[(size_t)idx*6+i];
}
float orc_ba_optimize(void* p, int its, int* stats2) { BAWindow* b=(BAWindow*)p; float r=b->optimize(its); if (stats2) { stats2[0]=b->opt_iterations; stats2[1]=b->opt_accepts; } return r; }
long long orc_ba_linearize_calls(void* p) { return ((BAWindow*)p)->linearize_calls; }
}

// ---- keyframe hand-over (marginalisation)
extern "C" {
void orc_ba_flag_points(void* p, const int* selected, int* status) { ((BAWindow*)p)->flagPointsForRemoval(selected, status); }
void orc_ba_marginalize_points(void* p, const int* status, double* M, double* Mb, double* Msc, double* Mbsc) {
  BAWindow* b=(BAWindow*)p; b->marginalizePointsF(status); size_t n=b->dim();
  if (M) std::copy(b->margM.begin(), b->margM.end(), M); if (Mb) std::copy(b->margMb.begin(), b->margMb.end(), Mb);
  if (Msc) std::copy(b->margMsc.begin(), b->margMsc.end(), Msc); if (Mbsc) std::copy(b->margMbsc.begin(), b->margMbsc.end(), Mbsc); (void)n; }
void orc_ba_marginalize_frame(void* p, int idx) { ((BAWindow*)p)->marginalizeFrame(idx); }
int  orc_ba_dim(void* p) { return ((BAWindow*)p)->dim(); }
void orc_ba_get_prior(void* p, double* HM, double* bM) { BAWindow* b=(BAWindow*)p; std::copy(b->HM.begin(), b->HM.end(), HM); std::copy(b->bM.begin(), b->bM.end(), bM); }
void orc_ba_get_res_to_zero(void* p, float* r2, int* isLin) { BAWindow* b=(BAWindow*)p; for (size_t i=0;i<b->res.size();i++) { r2[2*i]=b->res[i].res_toZeroF[0]; r2[2*i+1]=b->res[i].res_toZeroF[1]; isLin[i]=b->res[i].isLinearized; } }
}

// ---- timing loop of the CPU arm (bench.py cpu_baseline / --impl reference): n_frames x { FrameHessian::makeImages ; CoarseTracker::trackNewestCoarse } in ONE call, so
// that a host thread spends its time in this code and not in the Python interpreter (the GIL is released for the whole call).  imgs: n_imgs level-0 images that are
// tracked in turn starting at `start`; inits7: one initial guess per frame.  Stops early after budget_s seconds (<= 0: no budget).  Returns the frames done.
#include <time.h>
extern "C" int orc_bench_track_loop(void* t, int w, int h, int levels, const float* const* imgs, int n_imgs, int start, int n_frames, const double* inits7, double budget_s,
                                    double* last_T7, int* good_count) {
  CoarseTracker* T = (CoarseTracker*)t; timespec t0; clock_gettime(CLOCK_MONOTONIC, &t0);
  const double nanv = std::nan(""); const double minRes[5] = {nanv, nanv, nanv, nanv, nanv};
  int done = 0, good = 0;
  Frame fk;                                       // one frame object per thread, buffers reused (the reference allocates a FrameHessian per frame; with one process and
                                                  // many threads the page faults of fresh 10 MB pyramids serialise on the kernel's mmap lock, which is not the path under test)
  for (int f = 0; f < n_frames; f++) {
    fk.makeImages(imgs[(start + f) % n_imgs], w, h, levels);
    SE3 s = se3_from(inits7 + 7*(size_t)f); AffLight aff; aff.a = 0; aff.b = 0;
    if (T->trackNewestCoarse(&fk, s, aff, levels-1, minRes)) good++;
    if (last_T7) se3_to(s, last_T7);
    done++;
    if (budget_s > 0) { timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1); if ((t1.tv_sec - t0.tv_sec) + 1e-9*(t1.tv_nsec - t0.tv_nsec) > budget_s) break; }
  }
  if (good_count) *good_count = good;
  return done;
}
