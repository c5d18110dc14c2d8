// ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
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
