// sdv_math.cuh — host/device math substrate of the B200 path (product code; independent of oracle/).
//
// What it provides, and the reference behaviour each piece reproduces (file:line under /root/reference):
//   * SE3 with unit-quaternion storage and Sophus 0.9a exp/log/compose/inverse/Adj semantics
//       thirdparty/Sophus/sophus/se3.hpp:131-139,162-172,407-430,560-585 ; so3.hpp:196-202,343-369,491-531
//   * AffLight::fromToVecExposure                       src/util/NumType.h:149-158
//   * pivoted LDLT solve (the Eigen `ldlt().solve()` call sites: CoarseTracker.cpp:724, EnergyFunctional.cpp:743)
//   * 3x3 cofactor inverse (Eigen `K.inverse()`, CoarseTracker.cpp:100)
// Pose layout across the C-ABI: double T[7] = {qw,qx,qy,qz, tx,ty,tz}.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define SDV_HD __host__ __device__ __forceinline__
#else
#define SDV_HD inline
#endif

namespace sdv {

struct Quat { double w, x, y, z; };
struct SE3d { Quat q; double t[3]; };

SDV_HD SE3d se3_identity() { SE3d s; s.q = Quat{1,0,0,0}; s.t[0]=s.t[1]=s.t[2]=0; return s; }
SDV_HD SE3d se3_from7(const double* T) { SE3d s; s.q = Quat{T[0],T[1],T[2],T[3]}; s.t[0]=T[4]; s.t[1]=T[5]; s.t[2]=T[6]; return s; }
SDV_HD void se3_to7(const SE3d& s, double* T) { T[0]=s.q.w; T[1]=s.q.x; T[2]=s.q.y; T[3]=s.q.z; T[4]=s.t[0]; T[5]=s.t[1]; T[6]=s.t[2]; }

SDV_HD Quat qmul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w*b.w - a.x*b.x - a.y*b.y - a.z*b.z;
  r.x = a.w*b.x + a.x*b.w + a.y*b.z - a.z*b.y;
  r.y = a.w*b.y + a.y*b.w + a.z*b.x - a.x*b.z;
  r.z = a.w*b.z + a.z*b.w + a.x*b.y - a.y*b.x;
  return r;
}
SDV_HD Quat qnormalize(const Quat& q) {
  double len = sqrt(q.x*q.x + q.y*q.y + q.z*q.z + q.w*q.w);
  Quat r; r.w=q.w/len; r.x=q.x/len; r.y=q.y/len; r.z=q.z/len; return r;
}
SDV_HD void cross3(const double* a, const double* b, double* c) {
  c[0]=a[1]*b[2]-a[2]*b[1]; c[1]=a[2]*b[0]-a[0]*b[2]; c[2]=a[0]*b[1]-a[1]*b[0];
}
SDV_HD void qrot(const Quat& q, const double* v, double* out) {
  double qv[3]={q.x,q.y,q.z}, uv[3], c[3];
  cross3(qv, v, uv); uv[0]+=uv[0]; uv[1]+=uv[1]; uv[2]+=uv[2];
  cross3(qv, uv, c);
  out[0]=v[0]+q.w*uv[0]+c[0]; out[1]=v[1]+q.w*uv[1]+c[1]; out[2]=v[2]+q.w*uv[2]+c[2];
}
SDV_HD void qmat(const Quat& q, double* R /*row-major 3x3*/) {
  double tx=2*q.x, ty=2*q.y, tz=2*q.z;
  double twx=tx*q.w, twy=ty*q.w, twz=tz*q.w;
  double txx=tx*q.x, txy=ty*q.x, txz=tz*q.x;
  double tyy=ty*q.y, tyz=tz*q.y, tzz=tz*q.z;
  R[0]=1-(tyy+tzz); R[1]=txy-twz; R[2]=txz+twy;
  R[3]=txy+twz; R[4]=1-(txx+tzz); R[5]=tyz-twx;
  R[6]=txz-twy; R[7]=tyz+twx; R[8]=1-(txx+tyy);
}
SDV_HD void hat3(const double* w, double* O) {
  O[0]=0; O[1]=-w[2]; O[2]=w[1]; O[3]=w[2]; O[4]=0; O[5]=-w[0]; O[6]=-w[1]; O[7]=w[0]; O[8]=0;
}
SDV_HD void mm3(const double* A, const double* B, double* C) {
  for (int i=0;i<3;i++) for (int j=0;j<3;j++) C[i*3+j] = (A[i*3]*B[j] + A[i*3+1]*B[3+j]) + A[i*3+2]*B[6+j];
}
SDV_HD SE3d se3_mul(const SE3d& a, const SE3d& b) {
  SE3d r; double rt[3]; qrot(a.q, b.t, rt);
  r.t[0]=a.t[0]+rt[0]; r.t[1]=a.t[1]+rt[1]; r.t[2]=a.t[2]+rt[2];
  r.q = qnormalize(qmul(a.q, b.q)); return r;
}
SDV_HD SE3d se3_inv(const SE3d& a) {
  SE3d r; r.q = Quat{a.q.w,-a.q.x,-a.q.y,-a.q.z}; double m[3]={-a.t[0],-a.t[1],-a.t[2]}; qrot(r.q, m, r.t); return r;
}
SDV_HD SE3d se3_exp(const double* a /*[upsilon;omega]*/) {
  const double eps = 1e-10;
  const double* om = a+3;
  double theta_sq = om[0]*om[0] + om[1]*om[1] + om[2]*om[2];
  double theta = sqrt(theta_sq), half = 0.5*theta, imag, real;
  if (theta < eps) { double p4 = theta_sq*theta_sq;
    imag = 0.5 - (1.0/48.0)*theta_sq + (1.0/3840.0)*p4; real = 1.0 - 0.5*theta_sq + (1.0/384.0)*p4; }
  else { imag = sin(half)/theta; real = cos(half); }
  SE3d r; r.q = qnormalize(Quat{real, imag*om[0], imag*om[1], imag*om[2]});
  double Om[9], Om2[9], V[9]; hat3(om, Om); mm3(Om, Om, Om2);
  if (theta < eps) qmat(r.q, V);
  else { double c1 = (1.0-cos(theta))/theta_sq, c2 = (theta-sin(theta))/(theta_sq*theta);
    for (int i=0;i<9;i++) V[i] = ((i%4==0) ? 1.0 : 0.0) + c1*Om[i] + c2*Om2[i]; }
  for (int i=0;i<3;i++) r.t[i] = (V[i*3]*a[0] + V[i*3+1]*a[1]) + V[i*3+2]*a[2];
  return r;
}
SDV_HD void se3_log(const SE3d& s, double* out) {
  const double eps = 1e-10; const double pi = 3.14159265358979323846;
  double sqn = s.q.x*s.q.x + s.q.y*s.q.y + s.q.z*s.q.z, n = sqrt(sqn), w = s.q.w, f;
  if (n < eps) { double sw = w*w; f = 2.0/w - 2.0*sqn/(w*sw); }
  else if (fabs(w) < eps) f = (w > 0 ? pi/n : -pi/n);
  else f = 2.0*atan(n/w)/n;
  double theta = f*n; double om[3] = {f*s.q.x, f*s.q.y, f*s.q.z};
  double Om[9], Om2[9]; hat3(om, Om); mm3(Om, Om, Om2);
  double c = (fabs(theta) < eps) ? (1.0/12.0) : (1.0 - theta/(2.0*tan(theta/2.0)))/(theta*theta);
  for (int i=0;i<3;i++) {
    double v0 = ((i==0)?1.0:0.0) - 0.5*Om[i*3]   + c*Om2[i*3];
    double v1 = ((i==1)?1.0:0.0) - 0.5*Om[i*3+1] + c*Om2[i*3+1];
    double v2 = ((i==2)?1.0:0.0) - 0.5*Om[i*3+2] + c*Om2[i*3+2];
    out[i] = (v0*s.t[0] + v1*s.t[1]) + v2*s.t[2];
  }
  out[3]=om[0]; out[4]=om[1]; out[5]=om[2];
}
SDV_HD void se3_adj(const SE3d& s, double* A /*6x6 row-major*/) {
  double R[9], H[9], TR[9]; qmat(s.q, R); hat3(s.t, H); mm3(H, R, TR);
  for (int i=0;i<3;i++) for (int j=0;j<3;j++) { A[i*6+j]=R[i*3+j]; A[(i+3)*6+j+3]=R[i*3+j]; A[i*6+j+3]=TR[i*3+j]; A[(i+3)*6+j]=0; }
}

// AffLight::fromToVecExposure (NumType.h:149-158)
SDV_HD void aff_from_to(float exposureF, float exposureT, double aF, double bF, double aT, double bT, double* out) {
  if (exposureF==0 || exposureT==0) exposureT = exposureF = 1;
  double a = exp(aT-aF) * exposureT / exposureF;
  out[0] = a; out[1] = bT - a*bF;
}

// cofactor 3x3 inverse in float (the arithmetic of Eigen's fixed-size inverse)
SDV_HD void inv3f(const float* A, float* R) {
#define SDV_COF(i,j) (A[((i+1)%3)*3+((j+1)%3)]*A[((i+2)%3)*3+((j+2)%3)] - A[((i+1)%3)*3+((j+2)%3)]*A[((i+2)%3)*3+((j+1)%3)])
  float c00=SDV_COF(0,0), c10=SDV_COF(1,0), c20=SDV_COF(2,0);
  float det = (c00*A[0] + c10*A[3]) + c20*A[6];
  float invdet = 1.0f/det;
  R[0]=c00*invdet; R[1]=c10*invdet; R[2]=c20*invdet;
  R[3]=SDV_COF(0,1)*invdet; R[4]=SDV_COF(1,1)*invdet; R[5]=SDV_COF(2,1)*invdet;
  R[6]=SDV_COF(0,2)*invdet; R[7]=SDV_COF(1,2)*invdet; R[8]=SDV_COF(2,2)*invdet;
#undef SDV_COF
}

// Pivoted (diagonal) lower LDLT + solve, double, n <= MAXN.  A row-major n x n with leading dimension lda.
template <int MAXN>
SDV_HD void ldlt_solve(int n, const double* Ain, int lda, const double* b, double* x) {
  double A[MAXN*MAXN]; int perm[MAXN]; double tmp[MAXN]; double y[MAXN];
  for (int i=0;i<n;i++) for (int j=0;j<n;j++) A[i*MAXN+j] = Ain[i*lda+j];
  for (int k=0;k<n;k++) {
    int piv = k; double big = fabs(A[k*MAXN+k]);
    for (int i=k+1;i<n;i++) { double a = fabs(A[i*MAXN+i]); if (a > big) { big=a; piv=i; } }
    perm[k] = piv;
    if (piv != k) {
      int s = n-piv-1; double sw;
      for (int j=0;j<k;j++) { sw=A[k*MAXN+j]; A[k*MAXN+j]=A[piv*MAXN+j]; A[piv*MAXN+j]=sw; }
      for (int i=0;i<s;i++) { sw=A[(piv+1+i)*MAXN+k]; A[(piv+1+i)*MAXN+k]=A[(piv+1+i)*MAXN+piv]; A[(piv+1+i)*MAXN+piv]=sw; }
      sw=A[k*MAXN+k]; A[k*MAXN+k]=A[piv*MAXN+piv]; A[piv*MAXN+piv]=sw;
      for (int i=k+1;i<piv;i++) { sw=A[i*MAXN+k]; A[i*MAXN+k]=A[piv*MAXN+i]; A[piv*MAXN+i]=sw; }
    }
    int rs = n-k-1;
    if (k > 0) {
      for (int j=0;j<k;j++) tmp[j] = A[j*MAXN+j]*A[k*MAXN+j];
      double s=0; for (int j=0;j<k;j++) s += A[k*MAXN+j]*tmp[j];
      A[k*MAXN+k] -= s;
      for (int i=0;i<rs;i++) { double s2=0; for (int j=0;j<k;j++) s2 += A[(k+1+i)*MAXN+j]*tmp[j]; A[(k+1+i)*MAXN+k] -= s2; }
    }
    double akk = A[k*MAXN+k];
    if (rs > 0 && fabs(akk) > 0) for (int i=0;i<rs;i++) A[(k+1+i)*MAXN+k] /= akk;
  }
  for (int i=0;i<n;i++) y[i]=b[i];
  for (int k=0;k<n;k++) { double sw=y[k]; y[k]=y[perm[k]]; y[perm[k]]=sw; }
  for (int i=0;i<n;i++) { double s=y[i]; for (int j=0;j<i;j++) s -= A[i*MAXN+j]*y[j]; y[i]=s; }
  double dmax=0; for (int i=0;i<n;i++) { double d=fabs(A[i*MAXN+i]); if (d>dmax) dmax=d; }
  double tol = dmax*2.220446049250313e-16; if (tol < 1.0/1.7976931348623157e308) tol = 1.0/1.7976931348623157e308;
  for (int i=0;i<n;i++) { double d=A[i*MAXN+i]; y[i] = (fabs(d) > tol) ? y[i]/d : 0.0; }
  for (int i=n-1;i>=0;i--) { double s=y[i]; for (int j=i+1;j<n;j++) s -= A[j*MAXN+i]*y[j]; y[i]=s; }
  for (int k=n-1;k>=0;k--) { double sw=y[k]; y[k]=y[perm[k]]; y[perm[k]]=sw; }
  for (int i=0;i<n;i++) x[i]=y[i];
}

} // namespace sdv
