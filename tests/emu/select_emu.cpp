// select_emu.cpp — TEST INFRASTRUCTURE: sdv-loam_b200/csrc/sdv_select_core.cuh (the real kernels + host engine of the candidate-management path) compiled for the
// host on top of tests/emu/cuda_emu.hpp, with flat C entry points for tests/test_select_emu_cpu.py.  "Device" pointers are host pointers here.
#define SDV_EMU 1
#define SDV_EMU_IMPL 1
#include "cuda_emu.hpp"
#include "../../sdv-loam_b200/csrc/sdv_select_core.cuh"
#include "../../sdv-loam_b200/csrc/sdv_lidar_core.cuh"
using namespace sdv::sel;

extern "C" {
void* emu_engine_create(int w, int h, const unsigned char* rp, int dirDist) {
  SelEngine* e = new SelEngine(); SelSet S; S.minGradHistCut = 0.5f; S.minGradHistAdd = 3; S.gradDownweightPerLevel = 0.75f; S.selectDirectionDistribution = dirDist;
  S.outlierTH = 12*12; S.outlierTHSumComponent = 50*50; S.overallEnergyTHWeight = 1; e->init(w, h, S, rp, nullptr);
  e->walk_threads = 64;                               // one OS thread per CUDA thread here: keep the emulated CTA small (the result does not depend on the CTA size)
  return e; }
void emu_engine_destroy(void* e) { ((SelEngine*)e)->destroy(); delete (SelEngine*)e; }
void emu_engine_fuse_map(void* e, int f) { ((SelEngine*)e)->fuse_map = f != 0; }
const char* emu_engine_error(void* e) { return ((SelEngine*)e)->err.c_str(); }
void emu_engine_max_scratch(void* e, long long b) { ((SelEngine*)e)->max_scratch = (size_t)b; }
int emu_make_hists(void* ep, const float* I0, float* ths_out, float* thsSm_out) {
  SelEngine* e = (SelEngine*)ep; size_t tf = e->ths_floats(); std::vector<float> a(tf, 0.f), b(tf, 0.f); HistJob J{I0, a.data(), b.data()};
  int rc = e->make_hists(1, &J); int n = (e->w/32)*(e->h/32); memcpy(ths_out, a.data(), n*4); memcpy(thsSm_out, b.data(), n*4); return rc; }
// nj identical-shape jobs in one batch (different potentials / densities) to exercise the batched paths: arrays of length nj
int emu_make_maps(void* ep, int nj, const float* I0, const float4* L1, const float4* L2, const double* cloud3, int n, const float* density, const int* rec, const float* thFactor, int* pot_io,
                  unsigned char* maps_out, int* numHaveSub, int* passes) {
  SelEngine* e = (SelEngine*)ep; size_t tf = e->ths_floats(); std::vector<float> a(tf, 0.f), b(tf, 0.f); HistJob J{I0, a.data(), b.data()};
  if (e->make_hists(1, &J)) return -1;
  const bool lidar = cloud3 != nullptr; const size_t msz = lidar ? (size_t)std::max(n, 1) : (size_t)e->w*e->h;
  std::vector<MapsJobHost> M(nj);
  for (int j = 0; j < nj; j++) { M[j].img = FrameImg{I0, L1, L2}; M[j].thsSm = b.data(); M[j].cloud_dev = cloud3; M[j].n = n; M[j].map = maps_out + j*msz; M[j].density = density[j]; M[j].recursionsLeft = rec[j];
    M[j].thFactor = thFactor[j]; M[j].currentPotential = pot_io + j; M[j].numHaveSub = 0; M[j].passes = 0; }
  int rc = e->make_maps(M, lidar);
  for (int j = 0; j < nj; j++) { numHaveSub[j] = M[j].numHaveSub; passes[j] = M[j].passes; }
  return rc;
}
void* emu_slot_create() { return new SelectorSlot(); }
void emu_slot_destroy(void* s) { SelectorSlot* S = (SelectorSlot*)s; if (S->mapD) free(S->mapD); delete S; }
void emu_slot_set_potential(void* s, int p) { ((SelectorSlot*)s)->currentPotential = p; }
int  emu_slot_get_potential(void* s) { return ((SelectorSlot*)s)->currentPotential; }
void emu_slot_get_map(void* s, unsigned char* out, int wh) { SelectorSlot* S = (SelectorSlot*)s; if (S->mapD) memcpy(out, S->mapD, wh); else memset(out, 0, wh); }
int emu_new_trace_bytes() { return (int)sizeof(NewTrace); }
int emu_imm_bytes() { return (int)sizeof(ImmPt); }
int emu_make_new_traces(void* ep, int nj, void** slots, const float* const* I0, const float4* const* L1, const float4* const* L2, const double* const* cloud3, const int* n, const float* densL, const float* densD,
                        const int* add, void* out, void* imm, int cap, int* n_out, int* numPoints2, int* passes2) {
  SelEngine* e = (SelEngine*)ep; std::vector<SelEngine::NewTracesJob> J(nj);
  for (int j = 0; j < nj; j++) { J[j].img = FrameImg{I0[j], L1[j], L2[j]}; J[j].cloud_host = cloud3[j]; J[j].n = n[j]; J[j].slot = (SelectorSlot*)slots[j]; J[j].densityLidar = densL[j]; J[j].densityDense = densD[j];
    J[j].addFeaturePoint = add[j]; J[j].out_host = (NewTrace*)out + (size_t)j*cap; J[j].imm_host = (ImmPt*)imm + (size_t)j*cap; J[j].cap = cap; }
  int rc = e->make_new_traces(J);
  for (int j = 0; j < nj; j++) { n_out[j] = J[j].n_out; numPoints2[2*j] = J[j].numPoints[0]; numPoints2[2*j+1] = J[j].numPoints[1]; passes2[2*j] = J[j].passes[0]; passes2[2*j+1] = J[j].passes[1]; }
  return rc;
}
int emu_activate(void* ep, int nHosts, const int* pt_begin, const float* KRKi, const float* Kt, const float* uvid, int nCandHosts, const int* cand_begin, const float* cKRKi, const float* cKt,
                 const float* cand4, float minActDist, int* decision, float* map_out, int copies) {
  SelEngine* e = (SelEngine*)ep; std::vector<SelEngine::ActJob> J(copies);
  const int nc = nCandHosts ? cand_begin[nCandHosts] : 0; const size_t n1 = (size_t)(e->w >> 1)*(e->h >> 1);
  for (int j = 0; j < copies; j++) J[j] = SelEngine::ActJob{nHosts, pt_begin, KRKi, Kt, uvid, nCandHosts, cand_begin, cKRKi, cKt, cand4, minActDist, decision ? decision + (size_t)j*nc : nullptr, map_out ? map_out + j*n1 : nullptr};
  return e->activate(J);
}
long long emu_engine_launches(void* e) { return ((SelEngine*)e)->launches; }
// ---- LiDAR front-end (sdv_lidar_core.cuh)
float emu_atan2f(float y, float x) { return sdv::lidar::emu_lib_atan2f(y, x); }
void* emu_lidar_create(int n_scan, int horizon, float ang_res_x, float ang_res_y, float ang_bottom, int groundScanInd) { sdv::lidar::LidarEngine* e = new sdv::lidar::LidarEngine(); e->init(n_scan, horizon, ang_res_x, ang_res_y, ang_bottom, groundScanInd, nullptr); return e; }
void emu_lidar_destroy(void* e) { ((sdv::lidar::LidarEngine*)e)->destroy(); delete (sdv::lidar::LidarEngine*)e; }
const char* emu_lidar_error(void* e) { return ((sdv::lidar::LidarEngine*)e)->err.c_str(); }
// nj sweeps in one batch: xyzi pointers / sizes per sweep, shared extrinsics; lrud (4 per sweep) in/out; out3: nj x cap x 3; res: per sweep {n_out, numGround, n_segmented, addFeaturePoint}
int emu_lidar_handle(void* ep, int nj, const float* const* xyzi, const int* n, const double* R9, const double* t3, const float* K4, int w, int h, int* lrud, double* out3, int cap, int* res) {
  sdv::lidar::LidarEngine* e = (sdv::lidar::LidarEngine*)ep; std::vector<sdv::lidar::LidarEngine::Sweep> S(nj);
  for (int j = 0; j < nj; j++) { S[j].xyzi_host = xyzi[j]; S[j].n = n[j]; for (int k = 0; k < 9; k++) S[j].R[k] = R9[k]; for (int k = 0; k < 3; k++) S[j].t[k] = t3[k]; for (int k = 0; k < 4; k++) { S[j].K[k] = K4[k]; S[j].lrud[k] = lrud[4*j+k]; }
    S[j].w = w; S[j].h = h; S[j].out3_host = out3 + (size_t)j*cap*3; S[j].cap = cap; }
  int rc = e->handle(S);
  for (int j = 0; j < nj; j++) { for (int k = 0; k < 4; k++) lrud[4*j+k] = S[j].lrud[k]; res[4*j] = S[j].n_out; res[4*j+1] = S[j].numGround; res[4*j+2] = S[j].n_segmented; res[4*j+3] = S[j].addFeaturePoint; }
  return rc;
}
}
