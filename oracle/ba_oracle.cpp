// ba_oracle.cpp -- CPU oracle for the bundle-adjustment hot path.
//
// *** TEST INFRASTRUCTURE, NOT PRODUCT CODE. ***  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library, and only
// as the checker / the CPU baseline.  The shipped path (libba_cuda.so) never links,
// loads or calls anything in oracle/.
//
// What it restates (citations relative to the reference tree /root/reference):
//   * the residual functors
//       Model A  ReprojectionError            Test1_BundleAdjustment/bundle_adjustmenter.cpp:106-148
//       Model B  TargetCameraReprojectionError            Main_Calibration/bundle_adjustment.h:56-132
//                BaseCameraReprojectionError              bundle_adjustment.h:134-205
//                TargetCameraBaseMarkerReprojectionError  bundle_adjustment.h:207-276
//                BaseCameraBaseMarkerReprojectionError    bundle_adjustment.h:278-343
//                (and the Test2 two-functor forms, Test2_BundleAdjustment/bundle_adjustmenter.cpp:217-366)
//     evaluated with forward-mode dual numbers ("Jets") exactly as
//     ceres::AutoDiffCostFunction does, so the Jacobian is the derivative of the same
//     expression sequence the reference differentiates;
//   * the functor dispatch of BAManager::StartBA (bundle_adjustment_manager.cpp:21-88) and of
//     Test2_BundleAdjustment/main.cpp:64-97;
//   * the solve the reference delegates to a THIRD-PARTY dependency that is not under
//     /root/reference: Ceres Solver 1.14.0 (README.md:17; prebuilt ceres.lib,
//     PropertySheet_Release.props:11), call sites bundle_adjustment_manager.cpp:90-94,
//     Test1_BundleAdjustment/main.cpp:82-86, Test2_BundleAdjustment/main.cpp:99-103:
//     trust-region Levenberg-Marquardt, Jacobi scaling, Marquardt diagonal, Schur
//     elimination + dense Cholesky (DENSE_SCHUR), restated from Ceres 1.14's published
//     algorithm incl. ceres::AngleAxisRotatePoint (rotation.h) and the
//     truncated conjugate-gradient rule of ITERATIVE_SCHUR for BAL-sized problems;
//   * BALProblem::getPoint3dCoordinates (Main_Calibration/bundle_adjustment.cpp:89-130),
//     the numeric part of BAManager::Write (bundle_adjustment_manager.cpp:121,135-149) and of
//     ReprojectionCheck::Reproject (reprojection_check.cpp:69,81,100-101).
//
// Parity pinning: tests/test_oracle_golden.py checks this oracle against the
// reference's two committed 17-digit outputs (hongo/Camera_Transform.xml,
// test2/Camera_Transform.xml) and the 6-digit point3d / Extrinsics files.  Model A has
// no committed reference output ("parity unpinned by reference artefacts" for Model A;
// its arithmetic shares every building block with the pinned Model B path).
//
// Three independent linear-solve paths (dense normal equations, Schur + dense Cholesky,
// Schur + PCG) are provided so the oracle can check itself.
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/ba_cuda.h"  // option / row / summary PODs only

namespace {

// ---------------------------------------------------------------------------------
// Forward-mode dual numbers, same arithmetic rules as ceres::Jet (jet.h).
// ---------------------------------------------------------------------------------
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0.0) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  explicit Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; v[k] = 1.0; }
};
template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f) {
  Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  // jet.h: g_a_inverse = 1/g.a; f_a_by_g_a = f.a*g_a_inverse; v = (f.v - f_a_by_g_a*g.v)*g_a_inverse
  const double gi = 1.0 / g.a; const double q = f.a * gi;
  Jet<N> h; h.a = q; for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - q * g.v[i]) * gi; return h; }
template <int N> inline Jet<N> operator+(const Jet<N>& f, double s) { Jet<N> h = f; h.a += s; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, double s) { Jet<N> h = f; h.a -= s; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, double s) {
  Jet<N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
template <int N> inline Jet<N> sqrt(const Jet<N>& f) {
  const double t = std::sqrt(f.a); const double two_a_inverse = 1.0 / (2.0 * t);
  Jet<N> h; h.a = t; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * two_a_inverse; return h; }
template <int N> inline Jet<N> cos(const Jet<N>& f) {
  const double s = -std::sin(f.a); Jet<N> h; h.a = std::cos(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = s * f.v[i]; return h; }
template <int N> inline Jet<N> sin(const Jet<N>& f) {
  const double c = std::cos(f.a); Jet<N> h; h.a = std::sin(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i]; return h; }
inline double scalar_of(double x) { return x; }
template <int N> inline double scalar_of(const Jet<N>& x) { return x.a; }
template <typename T> inline T make_const(double s);
template <> inline double make_const<double>(double s) { return s; }
#define BA_JET_CONST(N) template <> inline Jet<N> make_const<Jet<N>>(double s) { return Jet<N>(s); }
BA_JET_CONST(6) BA_JET_CONST(9) BA_JET_CONST(12) BA_JET_CONST(18)
using std::sqrt; using std::cos; using std::sin;

// ceres::AngleAxisRotatePoint (Ceres 1.14 rotation.h), restated; call sites
// bundle_adjustment.h:97,103,109,176,182,247,253,320; bundle_adjustment.cpp:114,119;
// Test1 bundle_adjustmenter.cpp:126.  Alias-safe for result == pt like the original.
template <typename T>
inline void AngleAxisRotatePoint(const T aa[3], const T pt[3], T result[3]) {
  const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (scalar_of(theta2) > std::numeric_limits<double>::epsilon()) {
    const T theta = sqrt(theta2);
    const T costheta = cos(theta);
    const T sintheta = sin(theta);
    const T theta_inverse = make_const<T>(1.0) / theta;
    const T w[3] = {aa[0] * theta_inverse, aa[1] * theta_inverse, aa[2] * theta_inverse};
    const T w_cross_pt[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2],
                             w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (make_const<T>(1.0) - costheta);
    const T r0 = pt[0] * costheta + w_cross_pt[0] * sintheta + w[0] * tmp;
    const T r1 = pt[1] * costheta + w_cross_pt[1] * sintheta + w[1] * tmp;
    const T r2 = pt[2] * costheta + w_cross_pt[2] * sintheta + w[2] * tmp;
    result[0] = r0; result[1] = r1; result[2] = r2;
  } else {
    const T w_cross_pt[3] = {aa[1] * pt[2] - aa[2] * pt[1], aa[2] * pt[0] - aa[0] * pt[2],
                             aa[0] * pt[1] - aa[1] * pt[0]};
    const T r0 = pt[0] + w_cross_pt[0];
    const T r1 = pt[1] + w_cross_pt[1];
    const T r2 = pt[2] + w_cross_pt[2];
    result[0] = r0; result[1] = r1; result[2] = r2;
  }
}

// Model A functor body, Test1_BundleAdjustment/bundle_adjustmenter.cpp:122-141.
template <typename T>
inline void ModelAResidual(const double intr[4], double ox, double oy, const T* camera,
                           const T* point, T* residuals) {
  T p[3];
  AngleAxisRotatePoint(camera, point, p);
  p[0] = p[0] + camera[3];
  p[1] = p[1] + camera[4];
  p[2] = p[2] + camera[5];
  const T xp = make_const<T>(intr[0]) * p[0] / p[2] + intr[2];
  const T yp = make_const<T>(intr[1]) * p[1] / p[2] + intr[3];
  residuals[0] = xp - ox;
  residuals[1] = yp - oy;
}

// Model B functor body; cam / marker may be NULL = "that transform is not applied and
// the block is not a parameter" (the 4 functors of bundle_adjustment.h:74-125,
// 153-198, 226-269, 297-336 differ only in which of the three transforms they chain).
template <typename T>
inline void ModelBResidual(const double intr[4], const double* obs8, double half_side,
                           const T* cam, const T* frame, const T* marker, T* residuals) {
  const double mp[4][3] = {{-half_side, half_side, 0.0}, {half_side, half_side, 0.0},
                           {half_side, -half_side, 0.0}, {-half_side, -half_side, 0.0}};
  for (int i = 0; i < 4; ++i) {
    T p[3] = {make_const<T>(mp[i][0]), make_const<T>(mp[i][1]), make_const<T>(mp[i][2])};
    if (marker) {  // coordinate on base marker
      AngleAxisRotatePoint(marker, p, p);
      p[0] = p[0] + marker[3]; p[1] = p[1] + marker[4]; p[2] = p[2] + marker[5];
    }
    AngleAxisRotatePoint(frame, p, p);  // coordinate on base camera
    p[0] = p[0] + frame[3]; p[1] = p[1] + frame[4]; p[2] = p[2] + frame[5];
    if (cam) {  // coordinate on target camera
      AngleAxisRotatePoint(cam, p, p);
      p[0] = p[0] + cam[3]; p[1] = p[1] + cam[4]; p[2] = p[2] + cam[5];
    }
    const T xp = make_const<T>(intr[0]) * p[0] / p[2] + intr[2];
    const T yp = make_const<T>(intr[1]) * p[1] / p[2] + intr[3];
    residuals[2 * i] = xp - obs8[2 * i];
    residuals[2 * i + 1] = yp - obs8[2 * i + 1];
  }
}

// ---------------------------------------------------------------------------------
// Problem description shared by both models: every residual block touches at most one
// eliminated block ("e", Model A: point, Model B: frame) and at most two kept blocks
// ("f", Model A: camera; Model B: camera and marker).
// ---------------------------------------------------------------------------------
struct Problem {
  int model = 0;         // 0 = A, 1 = B
  int de = 3;            // e-block size
  int rdim = 2;          // residuals per residual block
  int64_t nb = 0;        // residual blocks
  int64_t ne = 0, nf = 0;
  // residual block -> blocks (in e-sorted order); original index kept in perm
  std::vector<int64_t> perm;      // sorted position -> caller's observation index
  std::vector<int32_t> be, bf0, bf1;
  std::vector<int64_t> e_ptr;     // CSR over sorted residual blocks
  std::vector<int64_t> e_off, f_off;  // offsets into the caller's parameter array
  std::vector<uint8_t> e_active, f_active;
  // model data
  const double* obs = nullptr;    // 2 or 8 per residual block (caller order)
  std::vector<double> intr;       // 4 per camera
  int intr_stride = 4;
  std::vector<int32_t> cam_of;    // camera index per residual block (caller order)
  double half_side = 0.0;
  int64_t n_params = 0;
  // incidences (e,f): Model A one per observation, Model B aggregated per (frame, f-block)
  std::vector<int64_t> inc_ptr;   // per e-block
  std::vector<int32_t> inc_f;     // f of incidence
  std::vector<int64_t> b_inc0, b_inc1;  // residual block -> incidence index (or -1)
  // per f: incidences and residual blocks
  std::vector<int64_t> f_inc_ptr, f_inc;      // CSR f -> incidence ids
  std::vector<int64_t> f_blk_ptr, f_blk;      // CSR f -> sorted residual block ids
  std::vector<int64_t> inc_e;                 // incidence -> e
  int threads = 1;
  // robust loss on every residual block (ba_cuda_options.loss_function / loss_scale); 0 = none, what the reference passes
  // (NULL at bundle_adjustment_manager.cpp:38,51,68,82)
  int loss = 0;
  double loss_a = 1.0;
};

// ceres::HuberLoss / ceres::CauchyLoss (loss_function.cc): rho(s), rho'(s); both have rho'' <= 0, so Ceres' Corrector
// (corrector.cc) scales residuals and Jacobian rows of the block by sqrt(rho') and nothing else.
inline void loss_eval(int type, double a, double s, double* rho, double* rho1) {
  *rho = s; *rho1 = 1.0;
  if (type == 1) {
    const double b = a * a;
    if (s > b) { const double r = std::sqrt(s); *rho = 2.0 * a * r - b; *rho1 = std::max(std::numeric_limits<double>::min(), a / r); }
  } else if (type == 2) {
    const double b = a * a, sum = 1.0 + s / b, inv = 1.0 / sum;
    *rho = b * std::log(sum); *rho1 = std::max(std::numeric_limits<double>::min(), inv);
  }
}

struct Eval {
  std::vector<double> r, Je, Jf0, Jf1;  // per sorted residual block: rdim, rdim*de, rdim*6, rdim*6
};

const double* intr_of(const Problem& P, int cam) { return P.intr.data() + (int64_t)P.intr_stride * cam; }

// Evaluates residual (and Jacobian) of sorted residual block b at parameters x.
inline void eval_block(const Problem& P, const double* x, int64_t b, bool want_j, double* r,
                       double* Je, double* Jf0, double* Jf1) {
  const int64_t o = P.perm[b];
  const int cam = P.cam_of[o];
  const double* K = intr_of(P, cam);
  if (P.model == 0) {
    const double* c = x + P.f_off[P.bf0[b]];
    const double* pt = x + P.e_off[P.be[b]];
    if (!want_j) {
      ModelAResidual<double>(K, P.obs[2 * o], P.obs[2 * o + 1], c, pt, r);
      return;
    }
    Jet<9> jc[6], jp[3], jr[2];
    for (int i = 0; i < 6; ++i) jc[i] = Jet<9>(c[i], i);
    for (int i = 0; i < 3; ++i) jp[i] = Jet<9>(pt[i], 6 + i);
    ModelAResidual<Jet<9>>(K, P.obs[2 * o], P.obs[2 * o + 1], jc, jp, jr);
    for (int k = 0; k < 2; ++k) {
      r[k] = jr[k].a;
      for (int i = 0; i < 6; ++i) Jf0[k * 6 + i] = jr[k].v[i];
      for (int i = 0; i < 3; ++i) Je[k * 3 + i] = jr[k].v[6 + i];
    }
    return;
  }
  const double* fr = x + P.e_off[P.be[b]];
  const double* c = P.bf0[b] >= 0 ? x + P.f_off[P.bf0[b]] : nullptr;
  const double* m = P.bf1[b] >= 0 ? x + P.f_off[P.bf1[b]] : nullptr;
  const double* ob = P.obs + 8 * o;
  if (!want_j) {
    ModelBResidual<double>(K, ob, P.half_side, c, fr, m, r);
    return;
  }
  // Jet layout: [frame 0..5 | cam 6..11 | marker 12..17]; unused parts stay zero.
  Jet<18> jf[6], jc[6], jm[6], jr[8];
  for (int i = 0; i < 6; ++i) jf[i] = Jet<18>(fr[i], i);
  if (c) for (int i = 0; i < 6; ++i) jc[i] = Jet<18>(c[i], 6 + i);
  if (m) for (int i = 0; i < 6; ++i) jm[i] = Jet<18>(m[i], 12 + i);
  ModelBResidual<Jet<18>>(K, ob, P.half_side, c ? jc : nullptr, jf, m ? jm : nullptr, jr);
  for (int k = 0; k < 8; ++k) {
    r[k] = jr[k].a;
    for (int i = 0; i < 6; ++i) {
      Je[k * 6 + i] = jr[k].v[i];
      Jf0[k * 6 + i] = c ? jr[k].v[6 + i] : 0.0;
      Jf1[k * 6 + i] = m ? jr[k].v[12 + i] : 0.0;
    }
  }
}

double evaluate(const Problem& P, const double* x, Eval* ev) {
  const int rd = P.rdim, de = P.de;
  double cost = 0.0;
  if (ev) {
    ev->r.resize(P.nb * rd); ev->Je.resize(P.nb * rd * de);
    ev->Jf0.resize(P.nb * rd * 6);
    ev->Jf1.resize(P.model == 1 ? P.nb * rd * 6 : 0);
  }
#pragma omp parallel for schedule(static) reduction(+ : cost) num_threads(P.threads)
  for (int64_t b = 0; b < P.nb; ++b) {
    double r[8], je[48], j0[48], j1[48];
    if (ev) {
      eval_block(P, x, b, true, r, je, j0, j1);
      if (P.loss != 0) {   // Corrector::CorrectResiduals / CorrectJacobian with alpha = 0
        double s = 0.0, rho, rho1;
        for (int k = 0; k < rd; ++k) s += r[k] * r[k];
        loss_eval(P.loss, P.loss_a, s, &rho, &rho1);
        const double w = std::sqrt(rho1);
        for (int k = 0; k < rd * de; ++k) je[k] *= w;
        for (int k = 0; k < rd * 6; ++k) { j0[k] *= w; if (P.model == 1) j1[k] *= w; }
        for (int k = 0; k < rd; ++k) r[k] *= w;
        cost += rho;
        std::memcpy(&ev->r[b * rd], r, sizeof(double) * rd);
        std::memcpy(&ev->Je[b * rd * de], je, sizeof(double) * rd * de);
        std::memcpy(&ev->Jf0[b * rd * 6], j0, sizeof(double) * rd * 6);
        if (P.model == 1) std::memcpy(&ev->Jf1[b * rd * 6], j1, sizeof(double) * rd * 6);
        continue;
      }
      std::memcpy(&ev->r[b * rd], r, sizeof(double) * rd);
      std::memcpy(&ev->Je[b * rd * de], je, sizeof(double) * rd * de);
      std::memcpy(&ev->Jf0[b * rd * 6], j0, sizeof(double) * rd * 6);
      if (P.model == 1) std::memcpy(&ev->Jf1[b * rd * 6], j1, sizeof(double) * rd * 6);
    } else {
      eval_block(P, x, b, false, r, nullptr, nullptr, nullptr);
    }
    double s = 0.0;
    for (int k = 0; k < rd; ++k) s += r[k] * r[k];
    if (P.loss != 0) { double rho, rho1; loss_eval(P.loss, P.loss_a, s, &rho, &rho1); s = rho; }
    cost += s;
  }
  return 0.5 * cost;
}

// ---------------------------------------------------------------------------------
// Small dense helpers (row-major).
// ---------------------------------------------------------------------------------
// In-place lower Cholesky A = L L^T of an n x n SPD matrix; false when not PD.
bool cholesky_lower(double* A, int64_t n) {
  for (int64_t j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int64_t k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    const double l = std::sqrt(d);
    A[j * n + j] = l;
    const double inv = 1.0 / l;
#pragma omp parallel for schedule(static) if (n - j > 256)
    for (int64_t i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      const double* ai = A + i * n; const double* aj = A + j * n;
      for (int64_t k = 0; k < j; ++k) s -= ai[k] * aj[k];
      A[i * n + j] = s * inv;
    }
  }
  return true;
}
void chol_solve(const double* L, int64_t n, double* b) {
  for (int64_t i = 0; i < n; ++i) {
    double s = b[i];
    for (int64_t k = 0; k < i; ++k) s -= L[i * n + k] * b[k];
    b[i] = s / L[i * n + i];
  }
  for (int64_t i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int64_t k = i + 1; k < n; ++k) s -= L[k * n + i] * b[k];
    b[i] = s / L[i * n + i];
  }
}
template <int D>
inline bool chol_small(double* A) {  // lower, in place
  for (int j = 0; j < D; ++j) {
    double d = A[j * D + j];
    for (int k = 0; k < j; ++k) d -= A[j * D + k] * A[j * D + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    const double l = std::sqrt(d);
    A[j * D + j] = l;
    for (int i = j + 1; i < D; ++i) {
      double s = A[i * D + j];
      for (int k = 0; k < j; ++k) s -= A[i * D + k] * A[j * D + k];
      A[i * D + j] = s / l;
    }
  }
  return true;
}
// x <- L^-1 x
template <int D> inline void fwd_small(const double* L, double* x) {
  for (int i = 0; i < D; ++i) { double s = x[i]; for (int k = 0; k < i; ++k) s -= L[i * D + k] * x[k]; x[i] = s / L[i * D + i]; }
}
// x <- L^-T x
template <int D> inline void bwd_small(const double* L, double* x) {
  for (int i = D - 1; i >= 0; --i) { double s = x[i]; for (int k = i + 1; k < D; ++k) s -= L[k * D + i] * x[k]; x[i] = s / L[i * D + i]; }
}

// ---------------------------------------------------------------------------------
// Linear solvers for   (Js^T Js + D^2) y = Js^T r      (Ceres solves this and negates).
// ---------------------------------------------------------------------------------
struct Scaled {
  // Jacobian scaled by the Jacobi column scaling, same block layout as Eval.
  const Eval* ev; const Problem* P;
};

struct LinearWork {
  std::vector<double> De, Df;           // lm diagonal per e / f scalar column
  std::vector<double> ye, yf;           // solution (scaled space)
  int pcg_iterations = 0;
  bool ok = false;
  // Schur workspace
  std::vector<double> L, z, Y, S, rhs;
  // block-sparse S (full pattern), rows = f
  std::vector<int64_t> s_ptr; std::vector<int32_t> s_col; std::vector<double> s_val;
  std::vector<int64_t> compact;  // f -> compact active index or -1
  int64_t n_active_f = 0;
};

void build_compact(const Problem& P, LinearWork* W) {
  if (!W->compact.empty()) return;
  W->compact.assign(P.nf, -1);
  int64_t k = 0;
  for (int64_t f = 0; f < P.nf; ++f) if (P.f_active[f]) W->compact[f] = k++;
  W->n_active_f = k;
}

// (0) dense normal equations over all active columns.
bool solve_dense_normal(const Problem& P, const Eval& ev, LinearWork* W) {
  build_compact(P, W);
  const int de = P.de, rd = P.rdim;
  std::vector<int64_t> ecomp(P.ne, -1);
  int64_t n = 6 * W->n_active_f;
  for (int64_t e = 0; e < P.ne; ++e) if (P.e_active[e]) { ecomp[e] = n; n += de; }
  std::vector<double> H((size_t)n * n, 0.0), g(n, 0.0);
  for (int64_t b = 0; b < P.nb; ++b) {
    const double* blocks[3] = {&ev.Je[b * rd * de], &ev.Jf0[b * rd * 6],
                               P.model == 1 ? &ev.Jf1[b * rd * 6] : nullptr};
    const int64_t col0[3] = {ecomp[P.be[b]], P.bf0[b] >= 0 ? 6 * W->compact[P.bf0[b]] : -1,
                             P.bf1[b] >= 0 ? 6 * W->compact[P.bf1[b]] : -1};
    const int width[3] = {de, 6, 6};
    const double* r = &ev.r[b * rd];
    for (int A = 0; A < 3; ++A) {
      if (col0[A] < 0 || !blocks[A]) continue;
      for (int i = 0; i < width[A]; ++i) {
        double s = 0.0;
        for (int k = 0; k < rd; ++k) s += blocks[A][k * width[A] + i] * r[k];
        g[col0[A] + i] += s;
      }
      for (int B = 0; B < 3; ++B) {
        if (col0[B] < 0 || !blocks[B]) continue;
        for (int i = 0; i < width[A]; ++i)
          for (int j = 0; j < width[B]; ++j) {
            double s = 0.0;
            for (int k = 0; k < rd; ++k) s += blocks[A][k * width[A] + i] * blocks[B][k * width[B] + j];
            H[(col0[A] + i) * n + col0[B] + j] += s;
          }
      }
    }
  }
  for (int64_t f = 0; f < P.nf; ++f) if (P.f_active[f])
    for (int i = 0; i < 6; ++i) { const int64_t c = 6 * W->compact[f] + i; H[c * n + c] += W->Df[6 * f + i] * W->Df[6 * f + i]; }
  for (int64_t e = 0; e < P.ne; ++e) if (P.e_active[e])
    for (int i = 0; i < de; ++i) { const int64_t c = ecomp[e] + i; H[c * n + c] += W->De[de * e + i] * W->De[de * e + i]; }
  if (!cholesky_lower(H.data(), n)) return false;
  chol_solve(H.data(), n, g.data());
  W->yf.assign(6 * P.nf, 0.0); W->ye.assign(de * P.ne, 0.0);
  for (int64_t f = 0; f < P.nf; ++f) if (P.f_active[f]) for (int i = 0; i < 6; ++i) W->yf[6 * f + i] = g[6 * W->compact[f] + i];
  for (int64_t e = 0; e < P.ne; ++e) if (P.e_active[e]) for (int i = 0; i < de; ++i) W->ye[de * e + i] = g[ecomp[e] + i];
  return true;
}

// Schur phase E: per e-block M = E^T E + De^2 = L L^T, z = L^-1 E^T r, Y_i = W_i L^-T.
template <int DE>
bool schur_phase_e(const Problem& P, const Eval& ev, LinearWork* W) {
  const int rd = P.rdim;
  const int64_t n_inc = (int64_t)P.inc_f.size();
  W->L.assign(P.ne * DE * DE, 0.0); W->z.assign(P.ne * DE, 0.0); W->Y.assign(n_inc * 6 * DE, 0.0);
  bool ok = true;
#pragma omp parallel for schedule(dynamic, 64) num_threads(P.threads)
  for (int64_t e = 0; e < P.ne; ++e) {
    if (!P.e_active[e]) continue;
    double M[DE * DE] = {0}, g[DE] = {0};
    for (int64_t b = P.e_ptr[e]; b < P.e_ptr[e + 1]; ++b) {
      const double* Je = &ev.Je[b * rd * DE]; const double* r = &ev.r[b * rd];
      for (int i = 0; i < DE; ++i) {
        for (int j = 0; j < DE; ++j) { double s = 0; for (int k = 0; k < rd; ++k) s += Je[k * DE + i] * Je[k * DE + j]; M[i * DE + j] += s; }
        double s = 0; for (int k = 0; k < rd; ++k) s += Je[k * DE + i] * r[k]; g[i] += s;
      }
      const double* Jf[2] = {&ev.Jf0[b * rd * 6], P.model == 1 ? &ev.Jf1[b * rd * 6] : nullptr};
      const int64_t inc[2] = {P.b_inc0[b], P.b_inc1[b]};
      for (int a = 0; a < 2; ++a) {
        if (inc[a] < 0) continue;
        double* Wm = &W->Y[inc[a] * 6 * DE];  // accumulate W = Jf^T Je first
        for (int i = 0; i < 6; ++i) for (int j = 0; j < DE; ++j) {
          double s = 0; for (int k = 0; k < rd; ++k) s += Jf[a][k * 6 + i] * Je[k * DE + j]; Wm[i * DE + j] += s; }
      }
    }
    for (int i = 0; i < DE; ++i) M[i * DE + i] += W->De[DE * e + i] * W->De[DE * e + i];
    if (!chol_small<DE>(M)) {
#pragma omp atomic write
      ok = false;
      continue;
    }
    std::memcpy(&W->L[e * DE * DE], M, sizeof(M));
    fwd_small<DE>(M, g);
    std::memcpy(&W->z[e * DE], g, sizeof(g));
    for (int64_t i = P.inc_ptr[e]; i < P.inc_ptr[e + 1]; ++i) {
      double* Wm = &W->Y[i * 6 * DE];
      for (int rrow = 0; rrow < 6; ++rrow) fwd_small<DE>(M, Wm + rrow * DE);  // row * L^-T == (L^-1 row^T)^T
    }
  }
  return ok;
}

// Schur phase F for one kept block row f: S[f, *] and rhs[f].  add(f2, 6x6 block).
template <int DE, typename AddFn>
inline void schur_row(const Problem& P, const Eval& ev, const LinearWork& W, int64_t f, double rhs[6], AddFn add) {
  const int rd = P.rdim;
  for (int i = 0; i < 6; ++i) rhs[i] = 0.0;
  double blk[36];
  // F^T F and F^T r
  for (int64_t q = P.f_blk_ptr[f]; q < P.f_blk_ptr[f + 1]; ++q) {
    const int64_t b = P.f_blk[q];
    const bool first = (P.bf0[b] == f);
    const double* Jme = first ? &ev.Jf0[b * rd * 6] : &ev.Jf1[b * rd * 6];
    const double* r = &ev.r[b * rd];
    for (int i = 0; i < 6; ++i) { double s = 0; for (int k = 0; k < rd; ++k) s += Jme[k * 6 + i] * r[k]; rhs[i] += s; }
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double s = 0; for (int k = 0; k < rd; ++k) s += Jme[k * 6 + i] * Jme[k * 6 + j]; blk[i * 6 + j] = s; }
    add(f, blk);
    if (P.model == 1) {
      const int64_t other = first ? P.bf1[b] : P.bf0[b];
      if (other >= 0) {
        const double* Jo = first ? &ev.Jf1[b * rd * 6] : &ev.Jf0[b * rd * 6];
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) { double s = 0; for (int k = 0; k < rd; ++k) s += Jme[k * 6 + i] * Jo[k * 6 + j]; blk[i * 6 + j] = s; }
        add(other, blk);
      }
    }
  }
  // - Y_i Y_j^T over incidences sharing the e-block; rhs -= Y_i z_e
  for (int64_t q = P.f_inc_ptr[f]; q < P.f_inc_ptr[f + 1]; ++q) {
    const int64_t i = P.f_inc[q]; const int64_t e = P.inc_e[i];
    const double* Yi = &W.Y[i * 6 * DE]; const double* z = &W.z[e * DE];
    for (int a = 0; a < 6; ++a) { double s = 0; for (int k = 0; k < DE; ++k) s += Yi[a * DE + k] * z[k]; rhs[a] -= s; }
    for (int64_t j = P.inc_ptr[e]; j < P.inc_ptr[e + 1]; ++j) {
      const double* Yj = &W.Y[j * 6 * DE];
      for (int a = 0; a < 6; ++a) for (int c = 0; c < 6; ++c) { double s = 0; for (int k = 0; k < DE; ++k) s += Yi[a * DE + k] * Yj[c * DE + k]; blk[a * 6 + c] = -s; }
      add(P.inc_f[j], blk);
    }
  }
  double d[36] = {0};
  for (int i = 0; i < 6; ++i) d[i * 6 + i] = W.Df[6 * f + i] * W.Df[6 * f + i];
  add(f, d);
}

template <int DE>
void back_substitute(const Problem& P, LinearWork* W) {
  W->ye.assign(DE * P.ne, 0.0);
#pragma omp parallel for schedule(dynamic, 64) num_threads(P.threads)
  for (int64_t e = 0; e < P.ne; ++e) {
    if (!P.e_active[e]) continue;
    double t[DE];
    for (int k = 0; k < DE; ++k) t[k] = W->z[e * DE + k];
    for (int64_t i = P.inc_ptr[e]; i < P.inc_ptr[e + 1]; ++i) {
      const double* Yi = &W->Y[i * 6 * DE]; const double* yf = &W->yf[6 * P.inc_f[i]];
      for (int k = 0; k < DE; ++k) { double s = 0; for (int a = 0; a < 6; ++a) s += Yi[a * DE + k] * yf[a]; t[k] -= s; }
    }
    bwd_small<DE>(&W->L[e * DE * DE], t);
    for (int k = 0; k < DE; ++k) W->ye[DE * e + k] = t[k];
  }
}

// (1) Schur + dense Cholesky == Ceres DENSE_SCHUR.
template <int DE>
bool solve_schur_dense(const Problem& P, const Eval& ev, LinearWork* W) {
  build_compact(P, W);
  if (!schur_phase_e<DE>(P, ev, W)) return false;
  const int64_t n = 6 * W->n_active_f;
  W->S.assign((size_t)n * n, 0.0); W->rhs.assign(n, 0.0);
#pragma omp parallel for schedule(dynamic, 1) num_threads(P.threads)
  for (int64_t f = 0; f < P.nf; ++f) {
    if (!P.f_active[f]) continue;
    const int64_t r0 = 6 * W->compact[f];
    double rhs[6];
    schur_row<DE>(P, ev, *W, f, rhs, [&](int64_t f2, const double* blk) {
      const int64_t c0 = 6 * W->compact[f2];
      for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) W->S[(r0 + i) * n + c0 + j] += blk[i * 6 + j];
    });
    for (int i = 0; i < 6; ++i) W->rhs[r0 + i] = rhs[i];
  }
  if (!cholesky_lower(W->S.data(), n)) return false;
  chol_solve(W->S.data(), n, W->rhs.data());
  W->yf.assign(6 * P.nf, 0.0);
  for (int64_t f = 0; f < P.nf; ++f) if (P.f_active[f]) for (int i = 0; i < 6; ++i) W->yf[6 * f + i] = W->rhs[6 * W->compact[f] + i];
  back_substitute<DE>(P, W);
  return true;
}

// (2) Schur + block-Jacobi PCG on the explicit block-sparse S, stopping rule of
// Ceres 1.14 ConjugateGradientsSolver as driven by IterativeSchurComplementSolver
// (x0 = 0, q_tolerance = eta, r_tolerance = -1, residual reset every 10 iterations).
template <int DE>
bool solve_schur_pcg(const Problem& P, const Eval& ev, const ba_cuda_options& opt, LinearWork* W) {
  build_compact(P, W);
  if (!schur_phase_e<DE>(P, ev, W)) return false;
  if (W->s_ptr.empty()) {  // pattern: f -> sorted unique f2 (full, both triangles)
    std::vector<std::vector<int32_t>> rows(P.nf);
#pragma omp parallel for schedule(dynamic, 16) num_threads(P.threads)
    for (int64_t f = 0; f < P.nf; ++f) {
      if (!P.f_active[f]) continue;
      auto& row = rows[f];
      row.push_back((int32_t)f);
      for (int64_t q = P.f_inc_ptr[f]; q < P.f_inc_ptr[f + 1]; ++q) {
        const int64_t e = P.inc_e[P.f_inc[q]];
        for (int64_t j = P.inc_ptr[e]; j < P.inc_ptr[e + 1]; ++j) row.push_back(P.inc_f[j]);
      }
      if (P.model == 1)
        for (int64_t q = P.f_blk_ptr[f]; q < P.f_blk_ptr[f + 1]; ++q) {
          const int64_t b = P.f_blk[q];
          if (P.bf0[b] >= 0) row.push_back(P.bf0[b]);
          if (P.bf1[b] >= 0) row.push_back(P.bf1[b]);
        }
      std::sort(row.begin(), row.end());
      row.erase(std::unique(row.begin(), row.end()), row.end());
    }
    W->s_ptr.assign(P.nf + 1, 0);
    for (int64_t f = 0; f < P.nf; ++f) W->s_ptr[f + 1] = W->s_ptr[f] + (int64_t)rows[f].size();
    W->s_col.resize(W->s_ptr[P.nf]);
    for (int64_t f = 0; f < P.nf; ++f) std::copy(rows[f].begin(), rows[f].end(), W->s_col.begin() + W->s_ptr[f]);
  }
  W->s_val.assign((size_t)W->s_ptr[P.nf] * 36, 0.0);
  const int64_t n = 6 * P.nf;
  std::vector<double> b(n, 0.0), Minv((size_t)P.nf * 36, 0.0);
  bool ok = true;
#pragma omp parallel for schedule(dynamic, 1) num_threads(P.threads)
  for (int64_t f = 0; f < P.nf; ++f) {
    if (!P.f_active[f]) continue;
    double rhs[6];
    const int32_t* cb = &W->s_col[W->s_ptr[f]]; const int32_t* ce = &W->s_col[W->s_ptr[f + 1]];
    schur_row<DE>(P, ev, *W, f, rhs, [&](int64_t f2, const double* blk) {
      const int64_t slot = W->s_ptr[f] + (std::lower_bound(cb, ce, (int32_t)f2) - cb);
      double* dst = &W->s_val[slot * 36];
      for (int i = 0; i < 36; ++i) dst[i] += blk[i];
    });
    for (int i = 0; i < 6; ++i) b[6 * f + i] = rhs[i];
    // block-Jacobi preconditioner = inverse of the diagonal block of S
    const int64_t slot = W->s_ptr[f] + (std::lower_bound(cb, ce, (int32_t)f) - cb);
    double Ld[36]; std::memcpy(Ld, &W->s_val[slot * 36], sizeof(Ld));
    if (!chol_small<6>(Ld)) {
#pragma omp atomic write
      ok = false;
      continue;
    }
    for (int c = 0; c < 6; ++c) {
      double col[6] = {0}; col[c] = 1.0; fwd_small<6>(Ld, col); bwd_small<6>(Ld, col);
      for (int i = 0; i < 6; ++i) Minv[f * 36 + i * 6 + c] = col[i];
    }
  }
  if (!ok) return false;
  auto spmv = [&](const std::vector<double>& x, std::vector<double>& y) {
#pragma omp parallel for schedule(dynamic, 16) num_threads(P.threads)
    for (int64_t f = 0; f < P.nf; ++f) {
      double acc[6] = {0};
      for (int64_t s = W->s_ptr[f]; s < W->s_ptr[f + 1]; ++s) {
        const double* B = &W->s_val[s * 36]; const double* xv = &x[6 * W->s_col[s]];
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) acc[i] += B[i * 6 + j] * xv[j];
      }
      for (int i = 0; i < 6; ++i) y[6 * f + i] = acc[i];
    }
  };
  auto dot = [&](const std::vector<double>& a, const std::vector<double>& c) { double s = 0; for (int64_t i = 0; i < n; ++i) s += a[i] * c[i]; return s; };
  std::vector<double> x(n, 0.0), r(b), p(n, 0.0), z(n, 0.0), tmp(n, 0.0);
  W->pcg_iterations = 0;
  const double norm_b = std::sqrt(dot(b, b));
  W->yf.assign(n, 0.0);
  if (norm_b == 0.0) { back_substitute<DE>(P, W); return true; }
  const double tol_r = opt.pcg_r_tolerance * norm_b;
  double rho = 1.0;
  double Q0 = 0.0;  // -x.(b + r) with x = 0
  bool solved_ok = true;
  for (int it = 1;; ++it) {
    W->pcg_iterations = it;
    for (int64_t f = 0; f < P.nf; ++f) {
      const double* Mi = &Minv[f * 36];
      for (int i = 0; i < 6; ++i) { double s = 0; for (int j = 0; j < 6; ++j) s += Mi[i * 6 + j] * r[6 * f + j]; z[6 * f + i] = s; }
    }
    const double last_rho = rho;
    rho = dot(r, z);
    if (rho == 0.0 || !std::isfinite(rho)) { solved_ok = false; break; }
    if (it == 1) p = z;
    else {
      const double beta = rho / last_rho;
      if (beta == 0.0 || !std::isfinite(beta)) { solved_ok = false; break; }
      for (int64_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
    }
    std::vector<double>& q = z;
    spmv(p, q);
    const double pq = dot(p, q);
    if (pq <= 0.0 || !std::isfinite(pq)) break;  // NO_CONVERGENCE: keep the current x
    const double alpha = rho / pq;
    if (!std::isfinite(alpha)) { solved_ok = false; break; }
    for (int64_t i = 0; i < n; ++i) x[i] += alpha * p[i];
    if (it % opt.pcg_residual_reset_period == 0) {
      spmv(x, tmp);
      for (int64_t i = 0; i < n; ++i) r[i] = b[i] - tmp[i];
    } else {
      for (int64_t i = 0; i < n; ++i) r[i] -= alpha * q[i];
    }
    double Q1 = 0.0;
    for (int64_t i = 0; i < n; ++i) Q1 += x[i] * (b[i] + r[i]);
    Q1 = -Q1;
    const double zeta = it * (Q1 - Q0) / Q1;
    if (zeta < opt.pcg_eta && it >= opt.pcg_min_iterations) break;
    Q0 = Q1;
    const double norm_r = std::sqrt(dot(r, r));
    if (norm_r <= tol_r && it >= opt.pcg_min_iterations) break;
    if (it >= opt.pcg_max_iterations) break;
  }
  if (!solved_ok) return false;
  W->yf = x;
  back_substitute<DE>(P, W);
  return true;
}

// ---------------------------------------------------------------------------------
// Trust-region Levenberg-Marquardt, Ceres 1.14 TrustRegionMinimizer +
// LevenbergMarquardtStrategy restated.
// ---------------------------------------------------------------------------------
struct LMResult {
  std::vector<ba_cuda_iteration> rows;
  ba_cuda_summary summary;
};

void column_sq_norms(const Problem& P, const Eval& ev, std::vector<double>& ne2, std::vector<double>& nf2) {
  const int de = P.de, rd = P.rdim;
  ne2.assign(de * P.ne, 0.0); nf2.assign(6 * P.nf, 0.0);
  for (int64_t b = 0; b < P.nb; ++b) {
    const double* Je = &ev.Je[b * rd * de];
    for (int i = 0; i < de; ++i) { double s = 0; for (int k = 0; k < rd; ++k) s += Je[k * de + i] * Je[k * de + i]; ne2[de * P.be[b] + i] += s; }
    if (P.bf0[b] >= 0) { const double* J = &ev.Jf0[b * rd * 6];
      for (int i = 0; i < 6; ++i) { double s = 0; for (int k = 0; k < rd; ++k) s += J[k * 6 + i] * J[k * 6 + i]; nf2[6 * P.bf0[b] + i] += s; } }
    if (P.model == 1 && P.bf1[b] >= 0) { const double* J = &ev.Jf1[b * rd * 6];
      for (int i = 0; i < 6; ++i) { double s = 0; for (int k = 0; k < rd; ++k) s += J[k * 6 + i] * J[k * 6 + i]; nf2[6 * P.bf1[b] + i] += s; } }
  }
}

void gradient(const Problem& P, const Eval& ev, std::vector<double>& ge, std::vector<double>& gf) {
  const int de = P.de, rd = P.rdim;
  ge.assign(de * P.ne, 0.0); gf.assign(6 * P.nf, 0.0);
  for (int64_t b = 0; b < P.nb; ++b) {
    const double* r = &ev.r[b * rd];
    const double* Je = &ev.Je[b * rd * de];
    for (int i = 0; i < de; ++i) { double s = 0; for (int k = 0; k < rd; ++k) s += Je[k * de + i] * r[k]; ge[de * P.be[b] + i] += s; }
    if (P.bf0[b] >= 0) { const double* J = &ev.Jf0[b * rd * 6];
      for (int i = 0; i < 6; ++i) { double s = 0; for (int k = 0; k < rd; ++k) s += J[k * 6 + i] * r[k]; gf[6 * P.bf0[b] + i] += s; } }
    if (P.model == 1 && P.bf1[b] >= 0) { const double* J = &ev.Jf1[b * rd * 6];
      for (int i = 0; i < 6; ++i) { double s = 0; for (int k = 0; k < rd; ++k) s += J[k * 6 + i] * r[k]; gf[6 * P.bf1[b] + i] += s; } }
  }
}

void scale_columns(const Problem& P, Eval* ev, const std::vector<double>& se, const std::vector<double>& sf) {
  const int de = P.de, rd = P.rdim;
#pragma omp parallel for schedule(static) num_threads(P.threads)
  for (int64_t b = 0; b < P.nb; ++b) {
    double* Je = &ev->Je[b * rd * de];
    for (int k = 0; k < rd; ++k) for (int i = 0; i < de; ++i) Je[k * de + i] *= se[de * P.be[b] + i];
    if (P.bf0[b] >= 0) { double* J = &ev->Jf0[b * rd * 6]; for (int k = 0; k < rd; ++k) for (int i = 0; i < 6; ++i) J[k * 6 + i] *= sf[6 * P.bf0[b] + i]; }
    if (P.model == 1 && P.bf1[b] >= 0) { double* J = &ev->Jf1[b * rd * 6]; for (int k = 0; k < rd; ++k) for (int i = 0; i < 6; ++i) J[k * 6 + i] *= sf[6 * P.bf1[b] + i]; }
  }
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// linear_solver: 0 dense normal equations, 1 Schur + dense Cholesky, 2 Schur + PCG
void minimize(Problem& P, double* x, const ba_cuda_options& opt, int linear_solver, LMResult* out) {
  const int de = P.de, rd = P.rdim;
  const double t_start = now_s();
  ba_cuda_summary& S = out->summary;
  std::memset(&S, 0, sizeof(S));
  out->rows.clear();
  S.num_residuals = P.nb * rd;
  int64_t nfree = 0;
  for (int64_t e = 0; e < P.ne; ++e) nfree += P.e_active[e] ? de : 0;
  int64_t naf = 0;
  for (int64_t f = 0; f < P.nf; ++f) { nfree += P.f_active[f] ? 6 : 0; naf += P.f_active[f] ? 1 : 0; }
  S.num_free_parameters = nfree;
  S.rcs_dim = (int32_t)(6 * naf);
  S.rcs_solver_used = linear_solver == 2 ? BA_RCS_PCG : BA_RCS_DENSE_CHOLESKY;

  Eval ev;
  LinearWork W;
  std::vector<double> se(de * P.ne, 1.0), sf(6 * P.nf, 1.0), ge, gf, ne2, nf2, diag_e, diag_f;
  std::vector<double> cand(x, x + P.n_params), delta(P.n_params, 0.0);

  auto active_param_loop = [&](auto&& fn) {  // fn(param offset, is_e, block, i)
    for (int64_t f = 0; f < P.nf; ++f) if (P.f_active[f]) for (int i = 0; i < 6; ++i) fn(P.f_off[f] + i, false, f, i);
    for (int64_t e = 0; e < P.ne; ++e) if (P.e_active[e]) for (int i = 0; i < de; ++i) fn(P.e_off[e] + i, true, e, i);
  };

  double x_cost = 0.0, gmax = 0.0, gnorm = 0.0;
  auto eval_gradient_and_jacobian = [&](bool first) {
    double t0 = now_s();
    x_cost = evaluate(P, x, &ev);
    S.num_jacobian_evaluations++;
    gradient(P, ev, ge, gf);
    if (opt.jacobi_scaling) {
      if (first) {
        column_sq_norms(P, ev, ne2, nf2);
        for (size_t i = 0; i < se.size(); ++i) se[i] = 1.0 / (1.0 + std::sqrt(ne2[i]));
        for (size_t i = 0; i < sf.size(); ++i) sf[i] = 1.0 / (1.0 + std::sqrt(nf2[i]));
      }
      scale_columns(P, &ev, se, sf);
    }
    // |x - Plus(x, -g)| as TrustRegionMinimizer::EvaluateGradientAndJacobian computes it
    gmax = 0.0; double g2 = 0.0;
    active_param_loop([&](int64_t off, bool is_e, int64_t blk, int i) {
      const double g = is_e ? ge[de * blk + i] : gf[6 * blk + i];
      const double d = x[off] - (x[off] + (-g));
      gmax = std::max(gmax, std::fabs(d)); g2 += d * d;
    });
    gnorm = std::sqrt(g2);
    S.ms_jacobian += (now_s() - t0) * 1e3;
  };

  double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int num_invalid = 0;
  double iter_t0 = now_s();

  auto finalize = [&](ba_cuda_iteration row) -> bool {  // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (row.step_is_successful) S.num_successful_steps++; else S.num_unsuccessful_steps++;
    row.trust_region_radius = radius;
    row.iteration_time_s = now_s() - iter_t0;
    out->rows.push_back(row);
    if (opt.minimizer_progress_to_stdout)
      std::printf("%4d % 14.6e % 10.2e % 10.2e % 10.2e % 10.2e % 10.2e %4d\n", row.iteration, row.cost, row.cost_change,
                  row.gradient_max_norm, row.step_norm, row.relative_decrease, row.trust_region_radius, row.linear_solver_iterations);
    if (row.iteration >= opt.max_num_iterations) { S.termination_type = BA_NO_CONVERGENCE; S.termination_reason = BA_REASON_MAX_ITERATIONS; return false; }
    if (row.step_is_successful && row.gradient_max_norm <= opt.gradient_tolerance) { S.termination_type = BA_CONVERGENCE; S.termination_reason = BA_REASON_GRADIENT_TOLERANCE; return false; }
    if (row.trust_region_radius <= opt.min_trust_region_radius) { S.termination_type = BA_CONVERGENCE; S.termination_reason = BA_REASON_MIN_TRUST_REGION_RADIUS; return false; }
    return true;
  };

  // IterationZero
  eval_gradient_and_jacobian(true);
  S.initial_cost = x_cost;
  ba_cuda_iteration row; std::memset(&row, 0, sizeof(row));
  row.iteration = 0; row.step_is_valid = 1; row.step_is_successful = 1; row.cost = x_cost;
  row.gradient_max_norm = gmax; row.gradient_norm = gnorm;
  bool go = finalize(row);

  while (go) {
    iter_t0 = now_s();
    std::memset(&row, 0, sizeof(row));
    row.iteration = out->rows.back().iteration + 1;
    // LevenbergMarquardtStrategy::ComputeStep
    double t0 = now_s();
    if (!reuse_diagonal) {
      column_sq_norms(P, ev, diag_e, diag_f);
      for (auto& d : diag_e) d = std::min(std::max(d, opt.min_lm_diagonal), opt.max_lm_diagonal);
      for (auto& d : diag_f) d = std::min(std::max(d, opt.min_lm_diagonal), opt.max_lm_diagonal);
    }
    W.De.resize(diag_e.size()); W.Df.resize(diag_f.size());
    for (size_t i = 0; i < diag_e.size(); ++i) W.De[i] = std::sqrt(diag_e[i] / radius);
    for (size_t i = 0; i < diag_f.size(); ++i) W.Df[i] = std::sqrt(diag_f[i] / radius);
    bool ok;
    W.pcg_iterations = 0;
    if (linear_solver == 0) ok = solve_dense_normal(P, ev, &W);
    else if (linear_solver == 1) ok = de == 3 ? solve_schur_dense<3>(P, ev, &W) : solve_schur_dense<6>(P, ev, &W);
    else ok = de == 3 ? solve_schur_pcg<3>(P, ev, opt, &W) : solve_schur_pcg<6>(P, ev, opt, &W);
    S.num_linear_solves++;
    if (ok) {
      for (double v : W.ye) if (!std::isfinite(v)) ok = false;
      for (double v : W.yf) if (!std::isfinite(v)) ok = false;
    }
    reuse_diagonal = true;
    row.linear_solver_iterations = W.pcg_iterations;
    S.ms_rcs_solve += (now_s() - t0) * 1e3;
    double model_cost_change = 0.0;
    if (ok) {
      // step = -y ; model_cost_change = -(J step)^T (r + J step / 2)
      double mcc = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : mcc) num_threads(P.threads)
      for (int64_t b = 0; b < P.nb; ++b) {
        double m[8];
        for (int k = 0; k < rd; ++k) {
          double s = 0.0;
          const double* Je = &ev.Je[b * rd * de];
          for (int i = 0; i < de; ++i) s += Je[k * de + i] * -W.ye[de * P.be[b] + i];
          if (P.bf0[b] >= 0) { const double* J = &ev.Jf0[b * rd * 6]; for (int i = 0; i < 6; ++i) s += J[k * 6 + i] * -W.yf[6 * P.bf0[b] + i]; }
          if (P.model == 1 && P.bf1[b] >= 0) { const double* J = &ev.Jf1[b * rd * 6]; for (int i = 0; i < 6; ++i) s += J[k * 6 + i] * -W.yf[6 * P.bf1[b] + i]; }
          m[k] = s;
        }
        double s = 0.0;
        for (int k = 0; k < rd; ++k) s += m[k] * (ev.r[b * rd + k] + m[k] / 2.0);
        mcc += s;
      }
      model_cost_change = -mcc;
      row.step_is_valid = model_cost_change > 0.0;
    }
    if (!row.step_is_valid) {  // HandleInvalidStep
      if (++num_invalid >= opt.max_num_consecutive_invalid_steps) {
        S.termination_type = BA_FAILURE; S.termination_reason = BA_REASON_TOO_MANY_INVALID_STEPS; break;
      }
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      row.cost = x_cost; row.cost_change = 0.0;
      row.gradient_max_norm = out->rows.back().gradient_max_norm; row.gradient_norm = out->rows.back().gradient_norm;
      row.step_norm = 0.0; row.relative_decrease = 0.0;
      go = finalize(row);
      continue;
    }
    num_invalid = 0;
    std::fill(delta.begin(), delta.end(), 0.0);
    active_param_loop([&](int64_t off, bool is_e, int64_t blk, int i) {
      delta[off] = is_e ? -W.ye[de * blk + i] * se[de * blk + i] : -W.yf[6 * blk + i] * sf[6 * blk + i];
    });
    for (int64_t i = 0; i < P.n_params; ++i) cand[i] = x[i] + delta[i];
    t0 = now_s();
    const double cand_cost = evaluate(P, cand.data(), nullptr);
    S.num_cost_evaluations++;
    S.ms_cost += (now_s() - t0) * 1e3;
    // ParameterToleranceReached
    double xn2 = 0.0, sn2 = 0.0;
    active_param_loop([&](int64_t off, bool, int64_t, int) { xn2 += x[off] * x[off]; const double d = x[off] - cand[off]; sn2 += d * d; });
    row.step_norm = std::sqrt(sn2);
    if (row.step_norm <= opt.parameter_tolerance * (std::sqrt(xn2) + opt.parameter_tolerance)) {
      S.termination_type = BA_CONVERGENCE; S.termination_reason = BA_REASON_PARAMETER_TOLERANCE; break;
    }
    // FunctionToleranceReached
    row.cost_change = x_cost - cand_cost;
    if (std::fabs(row.cost_change) <= opt.function_tolerance * x_cost) {
      S.termination_type = BA_CONVERGENCE; S.termination_reason = BA_REASON_FUNCTION_TOLERANCE; break;
    }
    row.relative_decrease = row.cost_change / model_cost_change;
    if (row.relative_decrease > opt.min_relative_decrease) {  // HandleSuccessfulStep
      std::memcpy(x, cand.data(), sizeof(double) * P.n_params);
      eval_gradient_and_jacobian(false);
      row.step_is_successful = 1;
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * row.relative_decrease - 1.0, 3));
      radius = std::min(opt.max_trust_region_radius, radius);
      decrease_factor = 2.0; reuse_diagonal = false;
      row.cost = x_cost; row.gradient_max_norm = gmax; row.gradient_norm = gnorm;
    } else {  // HandleUnsuccessfulStep
      row.step_is_successful = 0;
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      row.cost = cand_cost;
    }
    go = finalize(row);
  }
  S.num_iterations = (int32_t)out->rows.size();
  S.final_cost = x_cost;
  S.total_time_s = now_s() - t_start;
}

// ---------------------------------------------------------------------------------
// Problem construction.
// ---------------------------------------------------------------------------------
void finish_structure(Problem& P, const std::vector<int32_t>& e_of, const std::vector<int32_t>& f0_of,
                      const std::vector<int32_t>& f1_of) {
  // stable counting sort of residual blocks by e-block
  P.e_ptr.assign(P.ne + 1, 0);
  for (int64_t o = 0; o < P.nb; ++o) P.e_ptr[e_of[o] + 1]++;
  for (int64_t e = 0; e < P.ne; ++e) P.e_ptr[e + 1] += P.e_ptr[e];
  P.perm.resize(P.nb);
  { std::vector<int64_t> cur(P.e_ptr.begin(), P.e_ptr.end() - 1);
    for (int64_t o = 0; o < P.nb; ++o) P.perm[cur[e_of[o]]++] = o; }
  P.be.resize(P.nb); P.bf0.resize(P.nb); P.bf1.resize(P.nb);
  P.e_active.assign(P.ne, 0); P.f_active.assign(P.nf, 0);
  for (int64_t b = 0; b < P.nb; ++b) {
    const int64_t o = P.perm[b];
    P.be[b] = e_of[o]; P.bf0[b] = f0_of[o]; P.bf1[b] = f1_of[o];
    P.e_active[P.be[b]] = 1;
    if (P.bf0[b] >= 0) P.f_active[P.bf0[b]] = 1;
    if (P.bf1[b] >= 0) P.f_active[P.bf1[b]] = 1;
  }
  // incidences
  P.inc_ptr.assign(P.ne + 1, 0); P.inc_f.clear(); P.inc_e.clear();
  P.b_inc0.assign(P.nb, -1); P.b_inc1.assign(P.nb, -1);
  std::vector<int64_t> slot(P.nf, -1);
  for (int64_t e = 0; e < P.ne; ++e) {
    const int64_t first = (int64_t)P.inc_f.size();
    for (int64_t b = P.e_ptr[e]; b < P.e_ptr[e + 1]; ++b) {
      const int32_t fs[2] = {P.bf0[b], P.bf1[b]};
      for (int a = 0; a < 2; ++a) {
        if (fs[a] < 0) continue;
        int64_t id;
        if (P.model == 0) { id = (int64_t)P.inc_f.size(); P.inc_f.push_back(fs[a]); P.inc_e.push_back(e); }
        else {
          if (slot[fs[a]] < first) { slot[fs[a]] = (int64_t)P.inc_f.size(); P.inc_f.push_back(fs[a]); P.inc_e.push_back(e); }
          id = slot[fs[a]];
        }
        (a == 0 ? P.b_inc0[b] : P.b_inc1[b]) = id;
      }
    }
    P.inc_ptr[e + 1] = (int64_t)P.inc_f.size();
  }
  const int64_t n_inc = (int64_t)P.inc_f.size();
  P.f_inc_ptr.assign(P.nf + 1, 0);
  for (int64_t i = 0; i < n_inc; ++i) P.f_inc_ptr[P.inc_f[i] + 1]++;
  for (int64_t f = 0; f < P.nf; ++f) P.f_inc_ptr[f + 1] += P.f_inc_ptr[f];
  P.f_inc.resize(n_inc);
  { std::vector<int64_t> cur(P.f_inc_ptr.begin(), P.f_inc_ptr.end() - 1);
    for (int64_t i = 0; i < n_inc; ++i) P.f_inc[cur[P.inc_f[i]]++] = i; }
  P.f_blk_ptr.assign(P.nf + 1, 0);
  for (int64_t b = 0; b < P.nb; ++b) { if (P.bf0[b] >= 0) P.f_blk_ptr[P.bf0[b] + 1]++; if (P.bf1[b] >= 0) P.f_blk_ptr[P.bf1[b] + 1]++; }
  for (int64_t f = 0; f < P.nf; ++f) P.f_blk_ptr[f + 1] += P.f_blk_ptr[f];
  P.f_blk.resize(P.f_blk_ptr[P.nf]);
  { std::vector<int64_t> cur(P.f_blk_ptr.begin(), P.f_blk_ptr.end() - 1);
    for (int64_t b = 0; b < P.nb; ++b) { if (P.bf0[b] >= 0) P.f_blk[cur[P.bf0[b]]++] = b; if (P.bf1[b] >= 0) P.f_blk[cur[P.bf1[b]]++] = b; } }
}

bool build_model_a(Problem& P, int32_t n_cam, int64_t n_pt, int64_t n_obs, const int32_t* cam_idx,
                   const int32_t* pt_idx, const double* obs_xy, const double* intr, int32_t intr_stride) {
  P.model = 0; P.de = 3; P.rdim = 2; P.nb = n_obs; P.ne = n_pt; P.nf = n_cam;
  P.obs = obs_xy; P.n_params = 6 * (int64_t)n_cam + 3 * n_pt;
  P.intr_stride = intr_stride ? 4 : 0;
  P.intr.assign(intr, intr + (intr_stride ? 4 * (int64_t)n_cam : 4));
  P.cam_of.assign(cam_idx, cam_idx + n_obs);
  P.e_off.resize(n_pt); P.f_off.resize(n_cam);
  for (int64_t i = 0; i < n_pt; ++i) P.e_off[i] = 6 * (int64_t)n_cam + 3 * i;
  for (int32_t c = 0; c < n_cam; ++c) P.f_off[c] = 6 * (int64_t)c;
  std::vector<int32_t> e_of(pt_idx, pt_idx + n_obs), f0(cam_idx, cam_idx + n_obs), f1(n_obs, -1);
  for (int64_t o = 0; o < n_obs; ++o) if (e_of[o] < 0 || e_of[o] >= n_pt || f0[o] < 0 || f0[o] >= n_cam) return false;
  finish_structure(P, e_of, f0, f1);
  return true;
}

// f-block index space of Model B: [0,C) cameras, [C,C+M) markers.
bool build_model_b(Problem& P, int32_t C, int32_t T, int32_t M, int64_t n, const int32_t* time_idx,
                   const int32_t* cam_idx, const int32_t* marker_idx, const double* obs8, const double* intr4,
                   double marker_side, int32_t fix_cam0, int32_t fix_marker0) {
  if (!fix_cam0) return false;
  P.model = 1; P.de = 6; P.rdim = 8; P.nb = n; P.ne = T; P.nf = C + M;
  P.obs = obs8; P.n_params = 6 * ((int64_t)C + T + M);
  P.intr_stride = 4; P.intr.assign(intr4, intr4 + 4 * (int64_t)C);
  P.cam_of.assign(cam_idx, cam_idx + n);
  P.half_side = marker_side / 2;
  P.e_off.resize(T); P.f_off.resize(C + M);
  for (int32_t t = 0; t < T; ++t) P.e_off[t] = 6 * ((int64_t)C + t);
  for (int32_t c = 0; c < C; ++c) P.f_off[c] = 6 * (int64_t)c;
  for (int32_t m = 0; m < M; ++m) P.f_off[C + m] = 6 * ((int64_t)C + T + m);
  std::vector<int32_t> e_of(n), f0(n), f1(n);
  for (int64_t o = 0; o < n; ++o) {
    if (time_idx[o] < 0 || time_idx[o] >= T || cam_idx[o] < 0 || cam_idx[o] >= C || marker_idx[o] < 0 || marker_idx[o] >= M) return false;
    e_of[o] = time_idx[o];
    f0[o] = cam_idx[o] == 0 ? -1 : cam_idx[o];                              // bundle_adjustment_manager.cpp:26
    f1[o] = (fix_marker0 && marker_idx[o] == 0) ? -1 : C + marker_idx[o];   // bundle_adjustment_manager.cpp:28,58
  }
  finish_structure(P, e_of, f0, f1);
  return true;
}

int run(Problem& P, double* params, const ba_cuda_options* options, int linear_solver, int n_threads,
        ba_cuda_summary* summary, ba_cuda_iteration* rows, int cap, int* n_rows) {
  ba_cuda_options opt;
  if (options) opt = *options; else ba_cuda_options_init(&opt);
#ifdef _OPENMP
  P.threads = n_threads > 0 ? n_threads : omp_get_max_threads();
#else
  P.threads = 1;
#endif
  P.loss = opt.loss_function; P.loss_a = opt.loss_scale;
  LMResult res;
  minimize(P, params, opt, linear_solver, &res);
  if (summary) *summary = res.summary;
  const int nr = (int)res.rows.size();
  if (n_rows) *n_rows = nr;
  if (rows) for (int i = 0; i < std::min(nr, cap); ++i) rows[i] = res.rows[i];
  return 0;
}

void write_eval(const Problem& P, const Eval& ev, double* residuals, double* jac) {
  const int rd = P.rdim, de = P.de;
  for (int64_t b = 0; b < P.nb; ++b) {
    const int64_t o = P.perm[b];
    if (residuals) for (int k = 0; k < rd; ++k) residuals[o * rd + k] = ev.r[b * rd + k];
    if (!jac) continue;
    if (P.model == 0) {
      double* J = jac + o * 18;
      for (int k = 0; k < 2; ++k) { for (int i = 0; i < 6; ++i) J[k * 6 + i] = ev.Jf0[b * 12 + k * 6 + i]; for (int i = 0; i < 3; ++i) J[12 + k * 3 + i] = ev.Je[b * 6 + k * 3 + i]; }
    } else {
      double* J = jac + o * 144;
      for (int i = 0; i < 48; ++i) { J[i] = ev.Jf0[b * 48 + i]; J[48 + i] = ev.Je[b * 48 + i]; J[96 + i] = ev.Jf1[b * 48 + i]; }
    }
    (void)de;
  }
}

}  // namespace

// =====================================================================================
// C interface (ctypes).  Layouts are those documented in include/ba_cuda.h.
// =====================================================================================
extern "C" {

// Defaults of Ceres 1.14 Solver::Options; the product library has its own copy.
void ba_cuda_options_init(ba_cuda_options* o) {
  std::memset(o, 0, sizeof(*o));
  o->max_num_iterations = 50; o->max_num_consecutive_invalid_steps = 5; o->jacobi_scaling = 1;
  o->rcs_solver = BA_RCS_AUTO; o->pcg_max_iterations = 500; o->pcg_min_iterations = 0;
  o->pcg_residual_reset_period = 10; o->minimizer_progress_to_stdout = 0;
  o->initial_trust_region_radius = 1e4; o->max_trust_region_radius = 1e16; o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3; o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32;
  o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8;
  o->pcg_eta = 1e-1; o->pcg_r_tolerance = -1.0;
  o->loss_function = 0; o->loss_scale = 1.0;
}
void ba_oracle_options_init(ba_cuda_options* o) { ba_cuda_options_init(o); }

int ba_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int ba_oracle_solve_model_a(int32_t n_cam, int64_t n_pt, int64_t n_obs, const int32_t* cam_idx, const int32_t* pt_idx,
                            const double* obs_xy, const double* intr, int32_t intr_stride, double* params,
                            const ba_cuda_options* options, int linear_solver, int n_threads, ba_cuda_summary* summary,
                            ba_cuda_iteration* rows, int cap, int* n_rows) {
  Problem P;
  if (!build_model_a(P, n_cam, n_pt, n_obs, cam_idx, pt_idx, obs_xy, intr, intr_stride)) return -1;
  return run(P, params, options, linear_solver, n_threads, summary, rows, cap, n_rows);
}

int ba_oracle_solve_model_b(int32_t C, int32_t T, int32_t M, int64_t n, const int32_t* time_idx, const int32_t* cam_idx,
                            const int32_t* marker_idx, const double* obs8, const double* intr4, double marker_side,
                            int32_t fix_cam0, int32_t fix_marker0, double* params, const ba_cuda_options* options,
                            int linear_solver, int n_threads, ba_cuda_summary* summary, ba_cuda_iteration* rows, int cap,
                            int* n_rows) {
  Problem P;
  if (!build_model_b(P, C, T, M, n, time_idx, cam_idx, marker_idx, obs8, intr4, marker_side, fix_cam0, fix_marker0)) return -1;
  return run(P, params, options, linear_solver, n_threads, summary, rows, cap, n_rows);
}

int ba_oracle_eval_model_a(int32_t n_cam, int64_t n_pt, int64_t n_obs, const int32_t* cam_idx, const int32_t* pt_idx,
                           const double* obs_xy, const double* intr, int32_t intr_stride, const double* params,
                           int n_threads, double* cost, double* residuals, double* jac) {
  Problem P;
  if (!build_model_a(P, n_cam, n_pt, n_obs, cam_idx, pt_idx, obs_xy, intr, intr_stride)) return -1;
  P.threads = n_threads > 0 ? n_threads : 1;
  Eval ev;
  const bool want = residuals || jac;
  const double c = evaluate(P, params, want ? &ev : nullptr);
  if (cost) *cost = c;
  if (want) write_eval(P, ev, residuals, jac);
  return 0;
}

int ba_oracle_eval_model_b(int32_t C, int32_t T, int32_t M, int64_t n, const int32_t* time_idx, const int32_t* cam_idx,
                           const int32_t* marker_idx, const double* obs8, const double* intr4, double marker_side,
                           int32_t fix_cam0, int32_t fix_marker0, const double* params, int n_threads, double* cost,
                           double* residuals, double* jac) {
  Problem P;
  if (!build_model_b(P, C, T, M, n, time_idx, cam_idx, marker_idx, obs8, intr4, marker_side, fix_cam0, fix_marker0)) return -1;
  P.threads = n_threads > 0 ? n_threads : 1;
  Eval ev;
  const bool want = residuals || jac;
  const double c = evaluate(P, params, want ? &ev : nullptr);
  if (cost) *cost = c;
  if (want) write_eval(P, ev, residuals, jac);
  return 0;
}

void ba_oracle_angle_axis_rotate_point(const double* aa, const double* pt, double* out) {
  AngleAxisRotatePoint<double>(aa, pt, out);
}

// cv::Rodrigues(rvec -> R) as BAManager::Write uses it (bundle_adjustment_manager.cpp:121):
// R = cos(t) I + (1 - cos t) k k^T + sin(t) [k]x, identity for t < DBL_EPSILON.
void ba_oracle_rodrigues(const double* rvec, double* R) {
  const double t = std::sqrt(rvec[0] * rvec[0] + rvec[1] * rvec[1] + rvec[2] * rvec[2]);
  if (t < DBL_EPSILON) { for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0; return; }
  const double c = std::cos(t), s = std::sin(t), c1 = 1.0 - c, it = 1.0 / t;
  const double k[3] = {rvec[0] * it, rvec[1] * it, rvec[2] * it};
  const double kk[9] = {k[0] * k[0], k[0] * k[1], k[0] * k[2], k[0] * k[1], k[1] * k[1], k[1] * k[2], k[0] * k[2], k[1] * k[2], k[2] * k[2]};
  const double kx[9] = {0, -k[2], k[1], k[2], 0, -k[0], -k[1], k[0], 0};
  for (int i = 0; i < 9; ++i) R[i] = c * ((i % 4 == 0) ? 1.0 : 0.0) + c1 * kk[i] + s * kx[i];
}

// Numeric part of BAManager::Write + BALProblem::getPoint3dCoordinates, Model B parameter layout.
void ba_oracle_model_b_outputs(int32_t C, int32_t T, int32_t M, int64_t n, const int32_t* time_idx,
                               const int32_t* marker_idx, const double* params, double marker_side, double* rot9,
                               double* inv12, double* corners) {
  (void)M;
  for (int32_t c = 0; c < C; ++c) {
    const double* cp = params + 6 * (int64_t)c;
    double R[9];
    ba_oracle_rodrigues(cp, R);
    if (rot9) std::memcpy(rot9 + 9 * (int64_t)c, R, sizeof(R));
    if (inv12) for (int row = 0; row < 3; ++row) {
      double* o = inv12 + 12 * (int64_t)c + 4 * row;
      o[0] = R[0 * 3 + row]; o[1] = R[1 * 3 + row]; o[2] = R[2 * 3 + row];
      o[3] = -(R[0 * 3 + row] * cp[3] + R[1 * 3 + row] * cp[4] + R[2 * 3 + row] * cp[5]);
    }
  }
  if (!corners) return;
  const double h = marker_side / 2;
  const double mp[4][3] = {{-h, h, 0}, {h, h, 0}, {h, -h, 0}, {-h, -h, 0}};
  for (int64_t o = 0; o < n; ++o) {
    const double* fr = params + 6 * ((int64_t)C + time_idx[o]);
    const double* mk = params + 6 * ((int64_t)C + T + marker_idx[o]);
    for (int j = 0; j < 4; ++j) {
      double p[3];
      AngleAxisRotatePoint<double>(mk, mp[j], p);
      p[0] += mk[3]; p[1] += mk[4]; p[2] += mk[5];
      AngleAxisRotatePoint<double>(fr, p, p);
      p[0] += fr[3]; p[1] += fr[4]; p[2] += fr[5];
      std::memcpy(corners + (o * 4 + j) * 3, p, sizeof(p));
    }
  }
}

// Numeric part of ReprojectionCheck::Reproject (reprojection_check.cpp:65-81,100-101):
// cv::projectPoints with zero distortion = R(rvec) X + t, pinhole; image points are float.
void ba_oracle_project_points_error(int64_t n_points, const double* xyz, const int32_t* cam_of_point, const double* rvec_tvec6,
                                    const double* intr4, const float* image_xy, double* sum_half_sq, double* rms,
                                    double* reprojected_xy) {
  double err = 0.0;
  for (int64_t i = 0; i < n_points; ++i) {
    const int c = cam_of_point[i];
    double R[9];
    ba_oracle_rodrigues(rvec_tvec6 + 6 * c, R);
    const double* t = rvec_tvec6 + 6 * c + 3; const double* X = xyz + 3 * i; const double* K = intr4 + 4 * c;
    const double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
    const double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
    const double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    const double iz = z != 0.0 ? 1.0 / z : 1.0;  // cv::projectPoints: z = z ? 1./z : 1; x *= z; y *= z
    const double u = (x * iz) * K[0] + K[2], v = (y * iz) * K[1] + K[3];
    if (reprojected_xy) { reprojected_xy[2 * i] = u; reprojected_xy[2 * i + 1] = v; }
    const double dx = (double)image_xy[2 * i] - u, dy = (double)image_xy[2 * i + 1] - v;
    err += (dx * dx + dy * dy) / 2;
  }
  if (sum_half_sq) *sum_half_sq = err;
  if (rms) *rms = std::pow((err * 2.0) / (n_points * 2.0), 0.5);
}

}  // extern "C"
