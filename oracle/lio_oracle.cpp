// ============================================================================================================
// TEST INFRASTRUCTURE ONLY — CPU ORACLE.  Never linked, imported or executed by the shipped GPU path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// Dependency-free restatement (no Eigen/Boost/ROS/PCL — none exist in this image) of the reference's per-scan
// FAST-LIO2 update.  Every function cites the reference file:line it follows (paths relative to /root/reference).
//
//   esti_plane                  include/common_lib.h:506-536          (Eigen colPivHouseholderQr restated, float)
//   h_share_model               src/laserMapping.cpp:1876-2004
//   update_iterated_dyn_share_modified   include/IKFoM_toolkit/esekfom/esekfom.hpp:1620-1938
//   state manifold ops          include/IKFoM_toolkit/mtk/types/SOn.hpp:233-297, types/S2.hpp:136-280,
//                               mtk/src/mtkmath.hpp:142-288, include/use-ikfom.hpp:21-30
//   map_incremental             src/laserMapping.cpp:1440-1496
//   lasermap_fov_segment        src/laserMapping.cpp:1136-1200
//
// PARITY PIN STATUS: the k-NN / map half is pinned by the reference's own ikd-Tree compiled unmodified
// (oracle/_ref/libikd_ref.so).  The ESIKF / plane-fit half is "parity unpinned": the reference ships no tests or
// golden vectors and needs Eigen (README.md:17, not vendored) which is absent here; this restatement is
// cross-validated against numpy/scipy in tests/test_oracle_math.py only.
//
// State layout used across this C ABI ("state26", doubles):
//   [0:3) pos  [3:7) rot (x,y,z,w — Eigen coeff order)  [7:11) offset_R_L_I (x,y,z,w)  [11:14) offset_T_L_I
//   [14:17) vel  [17:20) bg  [20:23) ba  [23:26) grav (S2, |g| = 9.809)
// Error-state layout (23): pos 0, rot 3, offR 6, offT 9, vel 12, bg 15, ba 18, grav 21(2)  (use-ikfom.hpp:21-30).
// ============================================================================================================
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

namespace {

constexpr int NS = 23;          // error-state dof
constexpr int NMATCH = 5;       // NUM_MATCH_POINTS, common_lib.h:149
constexpr double TOL = 1e-11;   // MTK::tolerance<double>(), mtkmath.hpp:121
constexpr double S2_LEN = 98090.0 / 10000.0;  // use-ikfom.hpp:8 -> S2.hpp:104

// ---------------------------------------------------------------- small dense helpers (row-major)
struct Q4 { double x, y, z, w; };

inline Q4 qmul(const Q4& a, const Q4& b) {  // Eigen quaternion product
  Q4 r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
inline Q4 qconj(const Q4& a) { return Q4{-a.x, -a.y, -a.z, a.w}; }

// Eigen QuaternionBase::_transformVector:  uv = 2*(q.vec x v);  v + w*uv + q.vec x uv
inline void qrot(const Q4& q, const double v[3], double out[3]) {
  double uv[3] = {q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0]};
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  double c[3] = {q.y * uv[2] - q.z * uv[1], q.z * uv[0] - q.x * uv[2], q.x * uv[1] - q.y * uv[0]};
  out[0] = v[0] + q.w * uv[0] + c[0];
  out[1] = v[1] + q.w * uv[1] + c[1];
  out[2] = v[2] + q.w * uv[2] + c[2];
}

// Eigen QuaternionBase::toRotationMatrix
inline void qmat(const Q4& q, double R[9]) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

inline void hat3(const double v[3], double H[9]) {  // mtkmath.hpp:176-183, SKEW_SYM_MATRX
  H[0] = 0;     H[1] = -v[2]; H[2] = v[1];
  H[3] = v[2];  H[4] = 0;     H[5] = -v[0];
  H[6] = -v[1]; H[7] = v[0];  H[8] = 0;
}
inline void mm3(const double A[9], const double B[9], double C[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += A[3 * i + k] * B[3 * k + j];
      C[3 * i + j] = s;
    }
}
inline void mv3(const double A[9], const double v[3], double o[3]) {
  for (int i = 0; i < 3; i++) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}
inline void tr3(const double A[9], double T[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) T[3 * i + j] = A[3 * j + i];
}

// mtkmath.hpp:142-174
inline void cos_sinc_sqrt(double x2, double& c, double& s) {
  static const double taylor_0_bound = 2.220446049250313e-16;  // boost epsilon<double>
  static const double taylor_2_bound = std::sqrt(taylor_0_bound);
  static const double taylor_n_bound = std::sqrt(taylor_2_bound);
  if (x2 >= taylor_n_bound) {
    double x = std::sqrt(x2);
    c = std::cos(x);
    s = std::sin(x) / x;
    return;
  }
  static const double inv[] = {1 / 3., 1 / 4., 1 / 5., 1 / 6., 1 / 7., 1 / 8., 1 / 9.};
  double cosi = 1., sinc = 1;
  double term = -1 / 2. * x2;
  for (int i = 0; i < 3; ++i) {
    cosi += term;
    term *= inv[2 * i];
    sinc += term;
    term *= -inv[2 * i + 1] * x2;
  }
  c = cosi;
  s = sinc;
}

// mtkmath.hpp:249-256  MTK::exp(result, vec, scale) -> returns w
inline double mtk_exp(double res[3], const double vec[3], double scale) {
  double norm2 = vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2];
  double c, s;
  cos_sinc_sqrt(scale * scale * norm2, c, s);
  double mult = s * scale;
  res[0] = mult * vec[0]; res[1] = mult * vec[1]; res[2] = mult * vec[2];
  return c;
}

// SOn.hpp:284-288  SO3::exp(dvec, scale=1)  -> quaternion with w=exp(..., scale/2)
inline Q4 so3_exp(const double v[3]) {
  double r[3];
  double w = mtk_exp(r, v, 0.5);
  return Q4{r[0], r[1], r[2], w};
}

// mtkmath.hpp:268-288 with scale=2, plus_minus_periodicity=true  (SOn.hpp:293-297)
inline void so3_log(const Q4& q, double res[3]) {
  double nv = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
  if (nv < TOL) nv = TOL;
  double s = 2.0 / nv * std::atan(nv / q.w);
  res[0] = s * q.x; res[1] = s * q.y; res[2] = s * q.z;
}

// mtkmath.hpp:235-247
inline void A_matrix(const double v[3], double res[9]) {
  double squaredNorm = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  double norm = std::sqrt(squaredNorm);
  for (int i = 0; i < 9; i++) res[i] = 0;
  res[0] = res[4] = res[8] = 1;
  if (norm < TOL) return;
  double H[9], HH[9];
  hat3(v, H);
  mm3(H, H, HH);
  double a = (1 - std::cos(norm)) / squaredNorm;
  double b = (1 - std::sin(norm) / norm) / squaredNorm;
  for (int i = 0; i < 9; i++) res[i] = res[i] + a * H[i] + b * HH[i];
}

// S2.hpp:187-247, S2_typ == 1 branch (:215-231).  Bx is 3x2 row-major.
inline void s2_Bx(const double vec[3], double Bx[6]) {
  const double len = S2_LEN;
  if (vec[0] + len > TOL) {
    Bx[0] = -vec[1];                                  Bx[1] = -vec[2];
    Bx[2] = len - vec[1] * vec[1] / (len + vec[0]);   Bx[3] = -vec[2] * vec[1] / (len + vec[0]);
    Bx[4] = -vec[2] * vec[1] / (len + vec[0]);        Bx[5] = len - vec[2] * vec[2] / (len + vec[0]);
    for (int i = 0; i < 6; i++) Bx[i] /= len;
  } else {
    for (int i = 0; i < 6; i++) Bx[i] = 0;
    Bx[3] = -1;  // res(1,1)
    Bx[4] = 1;   // res(2,0)
  }
}

// S2.hpp:259-264   Nx (2x3) = 1/len/len * Bx^T * hat(vec)
inline void s2_Nx_yy(const double vec[3], double Nx[6]) {
  double Bx[6], H[9];
  s2_Bx(vec, Bx);
  hat3(vec, H);
  double sc = 1 / S2_LEN / S2_LEN;
  double sBt[6];  // 2x3
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 3; j++) sBt[3 * i + j] = sc * Bx[2 * j + i];
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += sBt[3 * i + k] * H[3 * k + j];
      Nx[3 * i + j] = s;
    }
}

// S2.hpp:266-280.  NOTE the quirk: exp(..., scalar(1/2)) -> integer division -> scale 0 -> identity rotation.
inline void s2_Mx(const double vec[3], const double delta[2], double Mx[6]) {
  double Bx[6], H[9];
  s2_Bx(vec, Bx);
  hat3(vec, H);
  double dn = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1]);
  if (dn < TOL) {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 2; j++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += (-H[3 * i + k]) * Bx[2 * k + j];
        Mx[2 * i + j] = s;
      }
    return;
  }
  double Bu[3];
  for (int i = 0; i < 3; i++) Bu[i] = Bx[2 * i] * delta[0] + Bx[2 * i + 1] * delta[1];
  double ev[3];
  double ew = mtk_exp(ev, Bu, 0.0);  // scalar(1/2) == 0
  double E[9];
  qmat(Q4{ev[0], ev[1], ev[2], ew}, E);
  double A[9], At[9], T1[9], T2[9];
  A_matrix(Bu, A);
  tr3(A, At);
  for (int i = 0; i < 9; i++) E[i] = -E[i];
  mm3(E, H, T1);
  mm3(T1, At, T2);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 2; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += T2[3 * i + k] * Bx[2 * k + j];
      Mx[2 * i + j] = s;
    }
}

// S2.hpp:136-142 boxplus(delta, scale=1)
inline void s2_boxplus(double vec[3], const double delta[2]) {
  double Bx[6];
  s2_Bx(vec, Bx);
  double Bu[3];
  for (int i = 0; i < 3; i++) Bu[i] = Bx[2 * i] * delta[0] + Bx[2 * i + 1] * delta[1];
  double ev[3];
  double ew = mtk_exp(ev, Bu, 0.5);
  double R[9], o[3];
  qmat(Q4{ev[0], ev[1], ev[2], ew}, R);
  mv3(R, vec, o);
  vec[0] = o[0]; vec[1] = o[1]; vec[2] = o[2];
}

// S2.hpp:144-167 boxminus(res, other): this = vec, other = ovec
inline void s2_boxminus(const double vec[3], const double ovec[3], double res[2]) {
  double H[9], t[3];
  hat3(vec, H);
  mv3(H, ovec, t);
  double v_sin = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
  double v_cos = vec[0] * ovec[0] + vec[1] * ovec[1] + vec[2] * ovec[2];
  double theta = std::atan2(v_sin, v_cos);
  if (v_sin < TOL) {
    if (std::fabs(theta) > TOL) { res[0] = 3.1415926; res[1] = 0; }
    else { res[0] = 0; res[1] = 0; }
    return;
  }
  double Bx[6], Ho[9], u[3];
  s2_Bx(ovec, Bx);
  hat3(ovec, Ho);
  mv3(Ho, vec, u);
  double f = theta / v_sin;
  for (int i = 0; i < 2; i++) {
    // (theta/v_sin * Bx^T) * (hat(other)*vec)
    double s = 0;
    for (int k = 0; k < 3; k++) s += (f * Bx[2 * k + i]) * u[k];
    res[i] = s;
  }
}

// ---------------------------------------------------------------- state
struct State {
  double pos[3];
  Q4 rot;
  Q4 offR;
  double offT[3], vel[3], bg[3], ba[3], grav[3];
};

inline State load_state(const double* s) {
  State x;
  x.pos[0] = s[0]; x.pos[1] = s[1]; x.pos[2] = s[2];
  x.rot = Q4{s[3], s[4], s[5], s[6]};
  x.offR = Q4{s[7], s[8], s[9], s[10]};
  for (int i = 0; i < 3; i++) {
    x.offT[i] = s[11 + i]; x.vel[i] = s[14 + i]; x.bg[i] = s[17 + i]; x.ba[i] = s[20 + i]; x.grav[i] = s[23 + i];
  }
  return x;
}
inline void store_state(const State& x, double* s) {
  s[0] = x.pos[0]; s[1] = x.pos[1]; s[2] = x.pos[2];
  s[3] = x.rot.x; s[4] = x.rot.y; s[5] = x.rot.z; s[6] = x.rot.w;
  s[7] = x.offR.x; s[8] = x.offR.y; s[9] = x.offR.z; s[10] = x.offR.w;
  for (int i = 0; i < 3; i++) {
    s[11 + i] = x.offT[i]; s[14 + i] = x.vel[i]; s[17 + i] = x.bg[i]; s[20 + i] = x.ba[i]; s[23 + i] = x.grav[i];
  }
}

// build_manifold.hpp:188-190 (MTK_BOXPLUS over the entries, in declaration order)
inline void state_boxplus(State& x, const double d[NS]) {
  for (int i = 0; i < 3; i++) x.pos[i] += d[i];
  x.rot = qmul(x.rot, so3_exp(d + 3));
  x.offR = qmul(x.offR, so3_exp(d + 6));
  for (int i = 0; i < 3; i++) {
    x.offT[i] += d[9 + i]; x.vel[i] += d[12 + i]; x.bg[i] += d[15 + i]; x.ba[i] += d[18 + i];
  }
  s2_boxplus(x.grav, d + 21);
}
// build_manifold.hpp:194-196  x.boxminus(res, other)
inline void state_boxminus(const State& x, const State& o, double r[NS]) {
  for (int i = 0; i < 3; i++) r[i] = x.pos[i] - o.pos[i];
  so3_log(qmul(qconj(o.rot), x.rot), r + 3);
  so3_log(qmul(qconj(o.offR), x.offR), r + 6);
  for (int i = 0; i < 3; i++) {
    r[9 + i] = x.offT[i] - o.offT[i]; r[12 + i] = x.vel[i] - o.vel[i];
    r[15 + i] = x.bg[i] - o.bg[i];    r[18 + i] = x.ba[i] - o.ba[i];
  }
  s2_boxminus(x.grav, o.grav, r + 21);
}

// ---------------------------------------------------------------- dense n x n inverse (stand-in for Eigen .inverse():
// PartialPivLU for dynamic / >4 fixed sizes; esekfom.hpp:1744,1788,1808).  Row-major, Gauss-Jordan on [A|I] with
// partial pivoting.  Returns false if singular.
bool invert(const double* A, double* Ainv, int n) {
  std::vector<double> a(A, A + (size_t)n * n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) Ainv[(size_t)i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int c = 0; c < n; c++) {
    int piv = c;
    double best = std::fabs(a[(size_t)c * n + c]);
    for (int r = c + 1; r < n; r++) {
      double v = std::fabs(a[(size_t)r * n + c]);
      if (v > best) { best = v; piv = r; }
    }
    if (best == 0.0) return false;
    if (piv != c) {
      for (int j = 0; j < n; j++) {
        std::swap(a[(size_t)c * n + j], a[(size_t)piv * n + j]);
        std::swap(Ainv[(size_t)c * n + j], Ainv[(size_t)piv * n + j]);
      }
    }
    double d = 1.0 / a[(size_t)c * n + c];
    for (int j = 0; j < n; j++) { a[(size_t)c * n + j] *= d; Ainv[(size_t)c * n + j] *= d; }
    for (int r = 0; r < n; r++) {
      if (r == c) continue;
      double f = a[(size_t)r * n + c];
      if (f == 0.0) continue;
      for (int j = 0; j < n; j++) {
        a[(size_t)r * n + j] -= f * a[(size_t)c * n + j];
        Ainv[(size_t)r * n + j] -= f * Ainv[(size_t)c * n + j];
      }
    }
  }
  return true;
}

// P rows idx..idx+d-1  <- J * P rows ;  P cols idx.. <- P cols * J^T   (esekfom.hpp:1674-1680, 1696-1702)
template <int D>
inline void left_rows(double* P, int idx, const double* J /*DxD*/, const double* Src) {
  for (int c = 0; c < NS; c++) {
    double t[D];
    for (int i = 0; i < D; i++) {
      double s = 0;
      for (int k = 0; k < D; k++) s += J[D * i + k] * Src[(idx + k) * NS + c];
      t[i] = s;
    }
    for (int i = 0; i < D; i++) P[(idx + i) * NS + c] = t[i];
  }
}
template <int D>
inline void right_cols(double* P, int idx, const double* J /*DxD*/) {
  for (int r = 0; r < NS; r++) {
    double t[D];
    for (int j = 0; j < D; j++) {
      double s = 0;
      for (int k = 0; k < D; k++) s += P[r * NS + idx + k] * J[D * j + k];  // * J^T
      t[j] = s;
    }
    for (int j = 0; j < D; j++) P[r * NS + idx + j] = t[j];
  }
}

// ---------------------------------------------------------------- esti_plane<float>  (common_lib.h:506-536)
// Eigen::ColPivHouseholderQR<Matrix<float,5,3>>::compute + solve restated (ColPivHouseholderQR.h computeInPlace /
// _solve_impl, Householder.h makeHouseholder / applyHouseholderOnTheLeft of Eigen 3.3.x, README.md:17).
// Plain sequential float sums; Eigen's SIMD reduction order is not reproduced (parity unpinned, see header).
bool esti_plane_f(float pabcd[4], const float pts[15] /*5 x (x,y,z)*/, float threshold) {
  const int rows = 5, cols = 3, size = 3;
  float qr[5][3];
  for (int j = 0; j < 5; j++) { qr[j][0] = pts[3 * j]; qr[j][1] = pts[3 * j + 1]; qr[j][2] = pts[3 * j + 2]; }
  float b[5] = {-1.f, -1.f, -1.f, -1.f, -1.f};
  float hCoeffs[3];
  int perm[3] = {0, 1, 2};  // column permutation: position -> original column
  float normsUpdated[3], normsDirect[3];
  const float eps = 1.1920929e-07f;
  for (int k = 0; k < cols; k++) {
    float s = 0.f;
    for (int r = 0; r < rows; r++) s += qr[r][k] * qr[r][k];
    normsDirect[k] = std::sqrt(s);
    normsUpdated[k] = normsDirect[k];
  }
  float maxnorm = std::max(normsUpdated[0], std::max(normsUpdated[1], normsUpdated[2]));
  float th = maxnorm * eps;
  const float threshold_helper = (th * th) / float(rows);
  const float norm_downdate_threshold = std::sqrt(eps);
  int nonzero_pivots = size;
  float maxpivot = 0.f;
  for (int k = 0; k < size; k++) {
    int big = k;
    float bigv = normsUpdated[k];
    for (int j = k + 1; j < cols; j++)
      if (normsUpdated[j] > bigv) { bigv = normsUpdated[j]; big = j; }
    float biggest_col_sq_norm = bigv * bigv;
    if (nonzero_pivots == size && biggest_col_sq_norm < threshold_helper * float(rows - k)) nonzero_pivots = k;
    if (k != big) {
      for (int r = 0; r < rows; r++) std::swap(qr[r][k], qr[r][big]);
      std::swap(normsUpdated[k], normsUpdated[big]);
      std::swap(normsDirect[k], normsDirect[big]);
      std::swap(perm[k], perm[big]);
    }
    // makeHouseholderInPlace on qr[k..rows-1][k]
    float tailSqNorm = 0.f;
    for (int r = k + 1; r < rows; r++) tailSqNorm += qr[r][k] * qr[r][k];
    float c0 = qr[k][k];
    float tau, beta;
    const float tol = 1.17549435e-38f;  // numeric_limits<float>::min()
    if (tailSqNorm <= tol) {
      tau = 0.f;
      beta = c0;
      for (int r = k + 1; r < rows; r++) qr[r][k] = 0.f;
    } else {
      beta = std::sqrt(c0 * c0 + tailSqNorm);
      if (c0 >= 0.f) beta = -beta;
      float den = c0 - beta;
      for (int r = k + 1; r < rows; r++) qr[r][k] = qr[r][k] / den;
      tau = (beta - c0) / beta;
    }
    hCoeffs[k] = tau;
    qr[k][k] = beta;
    if (std::fabs(beta) > maxpivot) maxpivot = std::fabs(beta);
    // apply to the remaining columns: bottomRightCorner(rows-k, cols-k-1)
    if (tau != 0.f) {
      for (int j = k + 1; j < cols; j++) {
        float tmp = 0.f;
        for (int r = k + 1; r < rows; r++) tmp += qr[r][k] * qr[r][j];
        tmp += qr[k][j];
        qr[k][j] -= tau * tmp;
        for (int r = k + 1; r < rows; r++) qr[r][j] -= tau * qr[r][k] * tmp;
      }
    }
    // column-norm downdate
    for (int j = k + 1; j < cols; j++) {
      if (normsUpdated[j] != 0.f) {
        float temp = std::fabs(qr[k][j]) / normsUpdated[j];
        temp = (1.f + temp) * (1.f - temp);
        temp = temp < 0.f ? 0.f : temp;
        float ratio = normsUpdated[j] / normsDirect[j];
        float temp2 = temp * (ratio * ratio);
        if (temp2 <= norm_downdate_threshold) {
          float s = 0.f;
          for (int r = k + 1; r < rows; r++) s += qr[r][j] * qr[r][j];
          normsDirect[j] = std::sqrt(s);
          normsUpdated[j] = normsDirect[j];
        } else {
          normsUpdated[j] *= std::sqrt(temp);
        }
      }
    }
  }
  // solve: c = Q^T b (apply H_0, H_1, ... in order), back-substitute R x = c[0:nz], un-permute
  float x[3] = {0.f, 0.f, 0.f};
  if (nonzero_pivots > 0) {
    for (int k = 0; k < nonzero_pivots; k++) {
      float tau = hCoeffs[k];
      if (tau != 0.f) {
        float tmp = 0.f;
        for (int r = k + 1; r < rows; r++) tmp += qr[r][k] * b[r];
        tmp += b[k];
        b[k] -= tau * tmp;
        for (int r = k + 1; r < rows; r++) b[r] -= tau * qr[r][k] * tmp;
      }
    }
    float c[3] = {b[0], b[1], b[2]};
    // column-oriented back substitution (Eigen triangular_solve_vector, Upper, ColMajor)
    for (int i = nonzero_pivots - 1; i >= 0; i--) {
      c[i] = c[i] / qr[i][i];
      for (int r = 0; r < i; r++) c[r] -= c[i] * qr[r][i];
    }
    for (int i = 0; i < nonzero_pivots; i++) x[perm[i]] = c[i];
  }
  float n = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  pabcd[0] = x[0] / n;
  pabcd[1] = x[1] / n;
  pabcd[2] = x[2] / n;
  pabcd[3] = (float)(1.0 / n);  // common_lib.h:526 "1.0 / n": double division, stored to float
  for (int j = 0; j < NMATCH; j++) {
    // common_lib.h:530
    if (std::fabs(pabcd[0] * pts[3 * j] + pabcd[1] * pts[3 * j + 1] + pabcd[2] * pts[3 * j + 2] + pabcd[3]) > threshold)
      return false;
  }
  return true;
}

inline void body_to_world(const State& s, const float pb[3], float pw[3]) {
  // laserMapping.cpp:1894-1898 : double math, stored to float
  double p[3] = {(double)pb[0], (double)pb[1], (double)pb[2]};
  double a[3], g[3];
  qrot(s.offR, p, a);
  a[0] += s.offT[0]; a[1] += s.offT[1]; a[2] += s.offT[2];
  qrot(s.rot, a, g);
  pw[0] = (float)(g[0] + s.pos[0]);
  pw[1] = (float)(g[1] + s.pos[1]);
  pw[2] = (float)(g[2] + s.pos[2]);
}

}  // namespace

extern "C" {

// k-NN provider used by search passes. Must fill out_xyz[nq*5*3] (ascending distance), out_d2[nq*5], out_cnt[nq].
typedef void (*orc_knn5_fn)(void* map, const float* q_xyz, int nq, float* out_xyz, float* out_d2, int* out_cnt);

int orc_esti_plane(const float* nn15, float thr, float* pabcd) { return esti_plane_f(pabcd, nn15, thr) ? 1 : 0; }

void orc_transform(const double* state26, const float* body, int n, float* world) {
  State s = load_state(state26);
  for (int i = 0; i < n; i++) body_to_world(s, body + 3 * i, world + 3 * i);
}

// One measurement pass = h_share_model (laserMapping.cpp:1876-2004) with the k-NN already done by the caller
// when `search` != 0 (nbr/nbr_d2/nbr_cnt refreshed for THIS world cloud), else the cached ones.
//   sel[N]      in/out  point_selected_surf (persists across the passes of one scan, quirk a11-v)
//   normvec[N*4] out    (nx,ny,nz,pd2) for selected points
//   hx[M*12], h[M] out  compacted in index order;  returns M (effct_feat_num); 0 => valid=false
//   total_residual out  sum |pd2| over selected
int orc_residual_pass(const double* state26, const float* body, const float* world, int n, const float* nbr,
                      const float* nbr_d2, const int* nbr_cnt, int search, unsigned char* sel, int extrinsic_est_en,
                      float* normvec, double* hx, double* h, double* total_residual) {
  State s = load_state(state26);
  for (int i = 0; i < n; i++) {
    const float* pw = world + 3 * i;
    const float* pb = body + 3 * i;
    if (search) {
      // :1911  size<5 -> false ; d2[4] > 5 -> false
      sel[i] = (nbr_cnt[i] < NMATCH) ? 0 : (nbr_d2[5 * i + 4] > 5 ? 0 : 1);
    }
    if (!sel[i]) continue;
    sel[i] = 0;
    float pabcd[4];
    if (esti_plane_f(pabcd, nbr + 15 * i, 0.1f)) {
      float pd2 = pabcd[0] * pw[0] + pabcd[1] * pw[1] + pabcd[2] * pw[2] + pabcd[3];  // :1925
      double pbn = std::sqrt((double)pb[0] * pb[0] + (double)pb[1] * pb[1] + (double)pb[2] * pb[2]);
      float sc = (float)(1 - 0.9 * std::fabs(pd2) / std::sqrt(pbn));  // :1927 (sqrt of the norm — quirk a11-ii)
      if (sc > 0.9) {  // float promoted to double vs 0.9
        sel[i] = 1;
        normvec[4 * i + 0] = pabcd[0];
        normvec[4 * i + 1] = pabcd[1];
        normvec[4 * i + 2] = pabcd[2];
        normvec[4 * i + 3] = pd2;
      }
    }
  }
  // compaction :1944-1954
  int M = 0;
  double tot = 0;
  std::vector<int> idx;
  idx.reserve(n);
  for (int i = 0; i < n; i++)
    if (sel[i]) {
      idx.push_back(i);
      tot += (double)std::fabs(normvec[4 * i + 3]);  // res_last[i] = abs(pd2) (float), summed into a double
      M++;
    }
  *total_residual = tot;
  if (M < 1) return 0;
  // Jacobian rows :1967-2002
  double Rrot[9];
  (void)Rrot;
  for (int j = 0; j < M; j++) {
    int i = idx[j];
    double pbe[3] = {(double)body[3 * i], (double)body[3 * i + 1], (double)body[3 * i + 2]};
    double Hbe[9], Hp[9];
    hat3(pbe, Hbe);
    double pt[3];
    qrot(s.offR, pbe, pt);
    pt[0] += s.offT[0]; pt[1] += s.offT[1]; pt[2] += s.offT[2];
    hat3(pt, Hp);
    double nv[3] = {(double)normvec[4 * i], (double)normvec[4 * i + 1], (double)normvec[4 * i + 2]};
    double C[3], A[3];
    qrot(qconj(s.rot), nv, C);
    mv3(Hp, C, A);
    double* row = hx + (size_t)12 * j;
    row[0] = nv[0]; row[1] = nv[1]; row[2] = nv[2];
    row[3] = A[0]; row[4] = A[1]; row[5] = A[2];
    if (extrinsic_est_en) {
      double Rc[9], T[9], B[3];
      qmat(qconj(s.offR), Rc);
      mm3(Hbe, Rc, T);  // (point_be_crossmat * offset_R_L_I.conjugate()) * C
      mv3(T, C, B);
      row[6] = B[0]; row[7] = B[1]; row[8] = B[2];
      row[9] = C[0]; row[10] = C[1]; row[11] = C[2];
    } else {
      for (int c = 6; c < 12; c++) row[c] = 0.0;
    }
    h[j] = -(double)normvec[4 * i + 3];
  }
  return M;
}

// update_iterated_dyn_share_modified (esekfom.hpp:1620-1938).
//   state26 in: propagated state, out: posterior.  P in/out (23x23 row-major).  R = LASER_POINT_COV.
//   body[N*3]; knn callback invoked on search passes.  nbr/nbr_d2/nbr_cnt/sel/normvec: caller scratch that
//   persists to map_incremental (Nearest_Points cache).  stats[0]=passes run, [1]=search passes, [2]=last M,
//   [3]=converged count t.  trace (optional, may be NULL): per pass 26 doubles of the state after boxplus.
void orc_esikf_update(double* state26, double* P_io, double R, int maximum_iter, const double* limit23,
                      const float* body, int n, int extrinsic_est_en, orc_knn5_fn knn, void* map, float* world,
                      float* nbr, float* nbr_d2, int* nbr_cnt, unsigned char* sel, float* normvec, int* stats,
                      double* trace) {
  State x = load_state(state26);
  State x_prop = x;
  std::vector<double> P_prop(P_io, P_io + NS * NS);
  std::vector<double> P(NS * NS), L(NS * NS);
  std::vector<double> hx((size_t)std::max(n, 1) * 12), hv(std::max(n, 1));
  double K_h[NS], K_x[NS * NS];
  double dx_new[NS];
  for (int i = 0; i < NS; i++) dx_new[i] = 0;
  bool converge = true;
  int t = 0;
  int passes = 0, searches = 0, lastM = 0;
  for (int it = -1; it < maximum_iter; it++) {
    // ---- h_dyn_share(x_, dyn_share)
    double sx[26];
    store_state(x, sx);
    orc_transform(sx, body, n, world);
    if (converge) {
      knn(map, world, n, nbr, nbr_d2, nbr_cnt);
      searches++;
    }
    double tot;
    int M = orc_residual_pass(sx, body, world, n, nbr, nbr_d2, nbr_cnt, converge ? 1 : 0, sel, extrinsic_est_en,
                              normvec, hx.data(), hv.data(), &tot);
    passes++;
    if (M < 1) continue;  // valid=false -> continue (:1641-1644)
    lastM = M;
    double dx[NS];
    state_boxminus(x, x_prop, dx);  // :1655
    for (int i = 0; i < NS; i++) dx_new[i] = dx[i];
    P = P_prop;  // :1660
    // SO3 blocks :1665-1681
    const int so3_idx[2] = {3, 6};
    for (int b = 0; b < 2; b++) {
      int idx = so3_idx[b];
      double A[9], J[9];
      A_matrix(dx + idx, A);
      tr3(A, J);
      double tv[3];
      mv3(J, dx_new + idx, tv);
      dx_new[idx] = tv[0]; dx_new[idx + 1] = tv[1]; dx_new[idx + 2] = tv[2];
      left_rows<3>(P.data(), idx, J, P.data());
      right_cols<3>(P.data(), idx, J);
    }
    {  // S2 block :1683-1703
      const int idx = 21;
      double Nx[6], Mx[6], J[4];
      s2_Nx_yy(x.grav, Nx);
      s2_Mx(x_prop.grav, dx + idx, Mx);
      for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) {
          double s = 0;
          for (int k = 0; k < 3; k++) s += Nx[3 * i + k] * Mx[2 * k + j];
          J[2 * i + j] = s;
        }
      double t0 = J[0] * dx_new[idx] + J[1] * dx_new[idx + 1];
      double t1 = J[2] * dx_new[idx] + J[3] * dx_new[idx + 1];
      dx_new[idx] = t0; dx_new[idx + 1] = t1;
      left_rows<2>(P.data(), idx, J, P.data());
      right_cols<2>(P.data(), idx, J);
    }
    if (NS > M) {
      // :1720-1750  K = P Hc^T (Hc P Hc^T / R + I)^-1 / R
      std::vector<double> Hc((size_t)M * NS, 0.0);
      for (int r = 0; r < M; r++)
        for (int c = 0; c < 12; c++) Hc[(size_t)r * NS + c] = hx[(size_t)r * 12 + c];
      std::vector<double> PHt((size_t)NS * M), S((size_t)M * M), Sinv((size_t)M * M), K((size_t)NS * M);
      for (int i = 0; i < NS; i++)
        for (int r = 0; r < M; r++) {
          double s = 0;
          for (int k = 0; k < NS; k++) s += P[i * NS + k] * Hc[(size_t)r * NS + k];
          PHt[(size_t)i * M + r] = s;
        }
      for (int r = 0; r < M; r++)
        for (int c = 0; c < M; c++) {
          double s = 0;
          for (int k = 0; k < NS; k++) s += Hc[(size_t)r * NS + k] * PHt[(size_t)k * M + c];
          S[(size_t)r * M + c] = s / R + (r == c ? 1.0 : 0.0);
        }
      invert(S.data(), Sinv.data(), M);
      for (int i = 0; i < NS; i++)
        for (int c = 0; c < M; c++) {
          double s = 0;
          for (int k = 0; k < M; k++) s += PHt[(size_t)i * M + k] * Sinv[(size_t)k * M + c];
          K[(size_t)i * M + c] = s / R;
        }
      for (int i = 0; i < NS; i++) {
        double s = 0;
        for (int k = 0; k < M; k++) s += K[(size_t)i * M + k] * hv[k];
        K_h[i] = s;
        for (int j = 0; j < NS; j++) {
          double q = 0;
          for (int k = 0; k < M; k++) q += K[(size_t)i * M + k] * Hc[(size_t)k * NS + j];
          K_x[i * NS + j] = q;
        }
      }
    } else {
      // :1788-1815
      std::vector<double> PR(NS * NS), P_temp(NS * NS), P_inv(NS * NS);
      for (int i = 0; i < NS * NS; i++) PR[i] = P[i] / R;
      invert(PR.data(), P_temp.data(), NS);
      double HTH[144];
      for (int a = 0; a < 12; a++)
        for (int b = 0; b < 12; b++) {
          double s = 0;
          for (int r = 0; r < M; r++) s += hx[(size_t)r * 12 + a] * hx[(size_t)r * 12 + b];
          HTH[a * 12 + b] = s;
        }
      for (int a = 0; a < 12; a++)
        for (int b = 0; b < 12; b++) P_temp[a * NS + b] += HTH[a * 12 + b];
      invert(P_temp.data(), P_inv.data(), NS);
      // K_h = (P_inv[:, :12] * h_x^T) * h
      std::vector<double> T((size_t)NS * M);
      for (int i = 0; i < NS; i++)
        for (int r = 0; r < M; r++) {
          double s = 0;
          for (int k = 0; k < 12; k++) s += P_inv[i * NS + k] * hx[(size_t)r * 12 + k];
          T[(size_t)i * M + r] = s;
        }
      for (int i = 0; i < NS; i++) {
        double s = 0;
        for (int r = 0; r < M; r++) s += T[(size_t)i * M + r] * hv[r];
        K_h[i] = s;
      }
      for (int i = 0; i < NS * NS; i++) K_x[i] = 0;
      for (int i = 0; i < NS; i++)
        for (int b = 0; b < 12; b++) {
          double s = 0;
          for (int k = 0; k < 12; k++) s += P_inv[i * NS + k] * HTH[k * 12 + b];
          K_x[i * NS + b] = s;
        }
    }
    // :1821  dx_ = K_h + (K_x - I) * dx_new
    double dx_[NS];
    for (int i = 0; i < NS; i++) {
      double s = 0;
      for (int j = 0; j < NS; j++) s += (K_x[i * NS + j] - (i == j ? 1.0 : 0.0)) * dx_new[j];
      dx_[i] = K_h[i] + s;
    }
    state_boxplus(x, dx_);  // :1823
    if (trace) store_state(x, trace + (size_t)26 * (passes - 1));
    converge = true;
    for (int i = 0; i < NS; i++)
      if (std::fabs(dx_[i]) > limit23[i]) { converge = false; break; }
    if (converge) t++;
    if (!t && it == maximum_iter - 2) converge = true;  // :1835-1838
    if (t > 1 || it == maximum_iter - 1) {
      // :1841-1931
      L = P;
      for (int b = 0; b < 2; b++) {
        int idx = so3_idx[b];
        double A[9], J[9];
        A_matrix(dx_ + idx, A);
        tr3(A, J);
        left_rows<3>(L.data(), idx, J, P.data());
        for (int c = 0; c < 12; c++) {
          double tv[3];
          for (int i = 0; i < 3; i++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += J[3 * i + k] * K_x[(idx + k) * NS + c];
            tv[i] = s;
          }
          for (int i = 0; i < 3; i++) K_x[(idx + i) * NS + c] = tv[i];
        }
        right_cols<3>(L.data(), idx, J);
        right_cols<3>(P.data(), idx, J);
      }
      {
        const int idx = 21;
        double Nx[6], Mx[6], J[4];
        s2_Nx_yy(x.grav, Nx);
        s2_Mx(x_prop.grav, dx_ + idx, Mx);
        for (int i = 0; i < 2; i++)
          for (int j = 0; j < 2; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += Nx[3 * i + k] * Mx[2 * k + j];
            J[2 * i + j] = s;
          }
        left_rows<2>(L.data(), idx, J, P.data());
        for (int c = 0; c < 12; c++) {
          double t0 = J[0] * K_x[idx * NS + c] + J[1] * K_x[(idx + 1) * NS + c];
          double t1 = J[2] * K_x[idx * NS + c] + J[3] * K_x[(idx + 1) * NS + c];
          K_x[idx * NS + c] = t0;
          K_x[(idx + 1) * NS + c] = t1;
        }
        right_cols<2>(L.data(), idx, J);
        right_cols<2>(P.data(), idx, J);
      }
      // P_ = L_ - K_x[:, :12] * P_[:12, :]
      std::vector<double> Pn(NS * NS);
      for (int i = 0; i < NS; i++)
        for (int j = 0; j < NS; j++) {
          double s = 0;
          for (int k = 0; k < 12; k++) s += K_x[i * NS + k] * P[k * NS + j];
          Pn[i * NS + j] = L[i * NS + j] - s;
        }
      P = Pn;
      store_state(x, state26);
      for (int i = 0; i < NS * NS; i++) P_io[i] = P[i];
      if (stats) { stats[0] = passes; stats[1] = searches; stats[2] = lastM; stats[3] = t; }
      return;
    }
  }
  // loop ran out (every remaining pass invalid): x_ keeps its last value, P_ stays as last assigned.
  // If no pass was ever valid, P_ was never touched (still P_propagated).
  store_state(x, state26);
  if (lastM > 0)
    for (int i = 0; i < NS * NS; i++) P_io[i] = P[i];
  if (stats) { stats[0] = passes; stats[1] = searches; stats[2] = lastM; stats[3] = t; }
}

// map_incremental classification (laserMapping.cpp:1440-1496).  cls[i]: 0 = dropped, 1 = PointToAdd (downsample),
// 2 = PointNoNeedDownsample.  world[N*3] is recomputed from the posterior state (:1451).
void orc_map_incremental_classify(const double* state26, const float* body, int n, const float* nbr,
                                  const int* nbr_cnt, int flg_EKF_inited, double filter_size_map_min, float* world,
                                  unsigned char* cls) {
  State s = load_state(state26);
  const double fs = filter_size_map_min;
  for (int i = 0; i < n; i++) {
    float* pw = world + 3 * i;
    body_to_world(s, body + 3 * i, pw);
    if (nbr_cnt[i] > 0 && flg_EKF_inited) {
      const float* pn = nbr + 15 * i;
      float mid[3];
      for (int a = 0; a < 3; a++) mid[a] = (float)(std::floor(pw[a] / fs) * fs + 0.5 * fs);  // :1459-1461 (double)
      float dist = (pw[0] - mid[0]) * (pw[0] - mid[0]) + (pw[1] - mid[1]) * (pw[1] - mid[1]) +
                   (pw[2] - mid[2]) * (pw[2] - mid[2]);  // calc_dist, common_lib.h:496-499 (float)
      // :1465  fabs(float) compared with double 0.5*fs
      if (std::fabs(pn[0] - mid[0]) > 0.5 * fs && std::fabs(pn[1] - mid[1]) > 0.5 * fs &&
          std::fabs(pn[2] - mid[2]) > 0.5 * fs) {
        cls[i] = 2;
        continue;
      }
      bool need_add = true;
      for (int j = 0; j < NMATCH; j++) {
        if (nbr_cnt[i] < NMATCH) break;
        const float* q = pn + 3 * j;
        float dq = (q[0] - mid[0]) * (q[0] - mid[0]) + (q[1] - mid[1]) * (q[1] - mid[1]) +
                   (q[2] - mid[2]) * (q[2] - mid[2]);
        if (dq < dist) { need_add = false; break; }
      }
      cls[i] = need_add ? 1 : 0;
    } else {
      cls[i] = 1;
    }
  }
}

// lasermap_fov_segment (laserMapping.cpp:1136-1200).  local_map[6] (vertex_min xyz, vertex_max xyz) and
// *initialized are in/out.  boxes_out receives up to 3 boxes x 6 floats; returns the number of boxes.
int orc_fov_segment(const double* pos_lid, double cube_len, float det_range, float* local_map, int* initialized,
                    float* boxes_out) {
  const float MOV_THRESHOLD = 1.5f;
  if (!*initialized) {
    for (int i = 0; i < 3; i++) {
      local_map[i] = (float)(pos_lid[i] - cube_len / 2.0);
      local_map[3 + i] = (float)(pos_lid[i] + cube_len / 2.0);
    }
    *initialized = 1;
    return 0;
  }
  float dist_to_map_edge[3][2];
  bool need_move = false;
  for (int i = 0; i < 3; i++) {
    dist_to_map_edge[i][0] = (float)std::fabs(pos_lid[i] - local_map[i]);
    dist_to_map_edge[i][1] = (float)std::fabs(pos_lid[i] - local_map[3 + i]);
    if (dist_to_map_edge[i][0] <= MOV_THRESHOLD * det_range || dist_to_map_edge[i][1] <= MOV_THRESHOLD * det_range)
      need_move = true;
  }
  if (!need_move) return 0;
  float newmap[6], tmp[6];
  for (int i = 0; i < 6; i++) newmap[i] = local_map[i];
  float mov_dist =
      (float)std::max((cube_len - 2.0 * MOV_THRESHOLD * det_range) * 0.5 * 0.9, double(det_range * (MOV_THRESHOLD - 1)));
  int nb = 0;
  for (int i = 0; i < 3; i++) {
    for (int k = 0; k < 6; k++) tmp[k] = local_map[k];
    if (dist_to_map_edge[i][0] <= MOV_THRESHOLD * det_range) {
      newmap[3 + i] -= mov_dist;
      newmap[i] -= mov_dist;
      tmp[i] = local_map[3 + i] - mov_dist;
      for (int k = 0; k < 6; k++) boxes_out[6 * nb + k] = tmp[k];
      nb++;
    } else if (dist_to_map_edge[i][1] <= MOV_THRESHOLD * det_range) {
      newmap[3 + i] += mov_dist;
      newmap[i] += mov_dist;
      tmp[3 + i] = local_map[i] + mov_dist;
      for (int k = 0; k < 6; k++) boxes_out[6 * nb + k] = tmp[k];
      nb++;
    }
  }
  for (int i = 0; i < 6; i++) local_map[i] = newmap[i];
  return nb;
}

// exposed for unit tests of the manifold algebra
void orc_boxplus(double* state26, const double* d23) {
  State x = load_state(state26);
  state_boxplus(x, d23);
  store_state(x, state26);
}
void orc_boxminus(const double* a26, const double* b26, double* r23) {
  State a = load_state(a26), b = load_state(b26);
  state_boxminus(a, b, r23);
}
void orc_A_matrix(const double* v3, double* out9) { A_matrix(v3, out9); }
int orc_invert(const double* A, double* Ainv, int n) { return invert(A, Ainv, n) ? 1 : 0; }
void orc_s2_mats(const double* vec3, const double* delta2, double* Bx6, double* Nx6, double* Mx6) {
  s2_Bx(vec3, Bx6);
  s2_Nx_yy(vec3, Nx6);
  s2_Mx(vec3, delta2, Mx6);
}

}  // extern "C"
