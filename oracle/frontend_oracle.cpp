// ============================================================================================================
// TEST INFRASTRUCTURE ONLY — CPU ORACLE for the rows either side of the per-scan path.
// Never linked, imported or executed by the shipped GPU path.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use it.
//
// Dependency-free restatements (no Eigen / PCL / ROS in this image), each citing what it follows:
//
//   orc_undistort            ImuProcess::UndistortPcl, backward pass   src/IMU_Processing.hpp:241-243 (sort),
//                            :334-386 (per-point compensation); Exp(ang_vel, dt) include/math_tools.h:39-61;
//                            Pose6D layout msg/Pose6D.msg, set_pose6d include/common_lib.h:446-460
//   orc_voxel_grid           pcl::VoxelGrid<PointXYZINormal>::applyFilter as called at src/laserMapping.cpp:2322-2323
//                            (leaf set at :2135) and :640-643.  PCL is a third-party dependency that is NOT under
//                            /root/reference (README.md:16 "PCL >= 1.10" => 1.10.0 on Ubuntu 20.04): this restates the
//                            published algorithm of PCL 1.10 filters/impl/voxel_grid.hpp (getMinMax3D, float leaf
//                            index, std::sort by leaf index, CentroidPoint accumulation, division by the count).
//   orc_transform_cloud_rpy  transformPointCloud include/common_lib.h:711-734 with pcl::getTransformation
//                            (PCL 1.10 common/impl/eigen.hpp, float), used by recontructIKdTree laserMapping.cpp:636
//   orc_body_to_world        pointBodyToWorld / RGBpointBodyToWorld  src/laserMapping.cpp:1077-1110
//
// PARITY PIN STATUS: "parity unpinned" — the reference ships no tests, vectors or data for these functions and
// neither Eigen nor PCL can be compiled here.  The restatements are cross-validated against independent numpy
// implementations in tests/test_oracle_frontend.py.
// ============================================================================================================
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

namespace {

struct Q4 { double x, y, z, w; };
inline Q4 qconj(const Q4& a) { return Q4{-a.x, -a.y, -a.z, a.w}; }
// Eigen QuaternionBase::_transformVector:  uv = q.vec x v; uv += uv;  v + w*uv + q.vec x uv
inline void qrot(const Q4& q, const double v[3], double out[3]) {
  double uv[3] = {q.y * v[2] - q.z * v[1], q.z * v[0] - q.x * v[2], q.x * v[1] - q.y * v[0]};
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  const double c[3] = {q.y * uv[2] - q.z * uv[1], q.z * uv[0] - q.x * uv[2], q.x * uv[1] - q.y * uv[0]};
  out[0] = v[0] + q.w * uv[0] + c[0];
  out[1] = v[1] + q.w * uv[1] + c[1];
  out[2] = v[2] + q.w * uv[2] + c[2];
}

// Exp(ang_vel, dt), math_tools.h:39-61: Rodrigues, "Eye3 + sin(r)*K + (1-cos(r))*K*K" == (I + s*K) + ((1-c)*K)*K
inline void exp_so3(const double w[3], double dt, double R[9]) {
  const double n = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
  if (!(n > 0.0000001)) return;
  const double a[3] = {w[0] / n, w[1] / n, w[2] / n};
  const double K[9] = {0.0, -a[2], a[1], a[2], 0.0, -a[0], -a[1], a[0], 0.0};
  const double r = n * dt;
  const double s = std::sin(r), c1 = 1.0 - std::cos(r);
  double cK[9];
  for (int i = 0; i < 9; i++) cK[i] = c1 * K[i];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      const double kk = cK[i * 3 + 0] * K[0 * 3 + j] + cK[i * 3 + 1] * K[1 * 3 + j] + cK[i * 3 + 2] * K[2 * 3 + j];
      R[i * 3 + j] = (R[i * 3 + j] + s * K[i * 3 + j]) + kk;
    }
}
inline void mat3mul(const double A[9], const double B[9], double C[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
inline void mat3vec(const double A[9], const double v[3], double o[3]) {
  for (int i = 0; i < 3; i++) o[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}

constexpr int POSE_DOUBLES = 22;  // offset_time, acc[3], gyr[3], vel[3], pos[3], rot[9]   (msg/Pose6D.msg)

// one compensation of one point with segment (head, tail), IMU_Processing.hpp:353-378
inline void compensate(float p[3], double t, const double* head, const double* tail, const Q4& rot_end, const double pos_end[3],
                       const Q4& offR, const double offT[3]) {
  const double dt = t - head[0];
  const double* R_imu = head + 13;
  const double* vel = head + 7;
  const double* pos = head + 10;
  const double* acc = tail + 1;
  const double* gyr = tail + 4;
  double E[9], Ri[9];
  exp_so3(gyr, dt, E);
  mat3mul(R_imu, E, Ri);
  const double Pi[3] = {(double)p[0], (double)p[1], (double)p[2]};
  double T_ei[3];
  for (int k = 0; k < 3; k++) T_ei[k] = ((pos[k] + vel[k] * dt) + ((0.5 * acc[k]) * dt) * dt) - pos_end[k];
  double a[3], b[3], c[3], d[3];
  qrot(offR, Pi, a);
  for (int k = 0; k < 3; k++) a[k] += offT[k];
  mat3vec(Ri, a, b);
  for (int k = 0; k < 3; k++) b[k] += T_ei[k];
  qrot(qconj(rot_end), b, c);
  for (int k = 0; k < 3; k++) c[k] -= offT[k];
  qrot(qconj(offR), c, d);
  p[0] = (float)d[0]; p[1] = (float)d[1]; p[2] = (float)d[2];
}

}  // namespace

extern "C" {

// UndistortPcl backward pass, literally: sort by curvature (std::sort, IMU_Processing.hpp:243), then the double sweep
// (:334-386) including its quirks (points with t <= IMUpose[0].offset_time stay untouched; the FIRST sorted point is
// compensated once more by every earlier segment whose head time it exceeds, because the `break` at begin() leaves
// it_pcl on it).  poses = np x 22 doubles (Pose6D), state26 = imu_state after the forward propagation.
// out_xyz[n*3] and perm[n] are in SORTED order like pcl_out (perm[j] = input index of sorted point j).
void orc_undistort(const float* xyz, const float* curvature, int n, const double* poses, int np, const double* state26,
                   float* out_xyz, int* perm) {
  std::vector<int> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return curvature[a] < curvature[b]; });  // time_list
  for (int j = 0; j < n; j++) {
    perm[j] = idx[j];
    for (int k = 0; k < 3; k++) out_xyz[3 * j + k] = xyz[3 * idx[j] + k];
  }
  if (n == 0 || np < 2) return;
  const Q4 rot_end{state26[3], state26[4], state26[5], state26[6]};
  const Q4 offR{state26[7], state26[8], state26[9], state26[10]};
  const double* pos_end = state26;
  const double* offT = state26 + 11;
  int it = n - 1;
  for (int kp = np - 1; kp != 0; kp--) {
    const double* head = poses + (size_t)(kp - 1) * POSE_DOUBLES;
    const double* tail = poses + (size_t)kp * POSE_DOUBLES;
    for (; curvature[idx[it]] / double(1000) > head[0]; it--) {
      compensate(out_xyz + 3 * it, curvature[idx[it]] / double(1000), head, tail, rot_end, pos_end, offR, offT);
      if (it == 0) break;
    }
  }
}

// pcl::VoxelGrid<PointXYZINormal>::applyFilter (PCL 1.10), downsample_all_data = true, min_points_per_voxel = 0, no
// filter field; fields carried: x,y,z,intensity (pts4) and curvature.  order_mode 0: std::sort by leaf index like PCL
// (the order INSIDE a leaf is whatever introsort leaves — unspecified by PCL, deterministic for one libstdc++);
// 1: stable order (ascending input index inside a leaf) — the order the GPU kernel sums in, for bit-exact tests.
// Returns the number of output points; -1 when PCL's overflow guard triggers (output = input, laserMapping unchanged).
int orc_voxel_grid(const float* pts4, const float* curvature, int n, float leaf, int order_mode, float* out4, float* out_curv) {
  if (n == 0) return 0;
  const float inv = 1.0f / leaf;   // inverse_leaf_size_ = Array4f::Ones() / leaf_size_.array()
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = 0; i < n; i++) {   // getMinMax3D (dense clouds: no finiteness test)
    for (int k = 0; k < 3; k++) {
      mn[k] = std::min(mn[k], pts4[4 * i + k]);
      mx[k] = std::max(mx[k], pts4[4 * i + k]);
    }
  }
  const std::int64_t dx = static_cast<std::int64_t>((mx[0] - mn[0]) * inv) + 1;
  const std::int64_t dy = static_cast<std::int64_t>((mx[1] - mn[1]) * inv) + 1;
  const std::int64_t dz = static_cast<std::int64_t>((mx[2] - mn[2]) * inv) + 1;
  if ((dx * dy * dz) > static_cast<std::int64_t>(INT32_MAX)) {
    for (int i = 0; i < n; i++) {
      for (int k = 0; k < 4; k++) out4[4 * i + k] = pts4[4 * i + k];
      if (out_curv) out_curv[i] = curvature ? curvature[i] : 0.f;
    }
    return -1;
  }
  int min_b[3], max_b[3], div_b[3];
  for (int k = 0; k < 3; k++) {
    min_b[k] = static_cast<int>(std::floor(mn[k] * inv));
    max_b[k] = static_cast<int>(std::floor(mx[k] * inv));
    div_b[k] = max_b[k] - min_b[k] + 1;
  }
  const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  struct CPI { unsigned idx; unsigned cloud_point_index; };
  std::vector<CPI> iv;
  iv.reserve(n);
  for (int i = 0; i < n; i++) {
    const int ijk0 = static_cast<int>(std::floor(pts4[4 * i + 0] * inv) - static_cast<float>(min_b[0]));
    const int ijk1 = static_cast<int>(std::floor(pts4[4 * i + 1] * inv) - static_cast<float>(min_b[1]));
    const int ijk2 = static_cast<int>(std::floor(pts4[4 * i + 2] * inv) - static_cast<float>(min_b[2]));
    const int idx = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2];
    iv.push_back(CPI{static_cast<unsigned>(idx), static_cast<unsigned>(i)});
  }
  auto less = [](const CPI& a, const CPI& b) { return a.idx < b.idx; };
  if (order_mode == 0) std::sort(iv.begin(), iv.end(), less);
  else std::stable_sort(iv.begin(), iv.end(), less);
  int total = 0;
  size_t index = 0;
  while (index < iv.size()) {
    size_t i = index + 1;
    while (i < iv.size() && iv[i].idx == iv[index].idx) ++i;
    // CentroidPoint<PointXYZINormal>: AccumulatorXYZ (Vector3f sum), AccumulatorIntensity, AccumulatorCurvature
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f, sc = 0.f;
    for (size_t li = index; li < i; ++li) {
      const unsigned p = iv[li].cloud_point_index;
      sx += pts4[4 * p + 0]; sy += pts4[4 * p + 1]; sz += pts4[4 * p + 2];
      si += pts4[4 * p + 3];
      if (curvature) sc += curvature[p];
    }
    const float cnt = static_cast<float>(i - index);
    out4[4 * total + 0] = sx / cnt; out4[4 * total + 1] = sy / cnt; out4[4 * total + 2] = sz / cnt;
    out4[4 * total + 3] = si / cnt;
    if (out_curv) out_curv[total] = sc / cnt;
    ++total;
    index = i;
  }
  return total;
}

// transformPointCloud(cloudIn, PointTypePose*) common_lib.h:711-734: Affine3f from pcl::getTransformation(x,y,z,roll,
// pitch,yaw) (float), then per point "t00*x + t01*y + t02*z + t03" in float, intensity copied.  pose6 = x,y,z,roll,pitch,yaw.
void orc_rpy_matrix(const float* pose6, float* t12) {
  const float x = pose6[0], y = pose6[1], z = pose6[2], roll = pose6[3], pitch = pose6[4], yaw = pose6[5];
  const float A = std::cos(yaw), B = std::sin(yaw), Cc = std::cos(pitch), D = std::sin(pitch);
  const float E = std::cos(roll), F = std::sin(roll), DE = D * E, DF = D * F;
  t12[0] = A * Cc; t12[1] = A * DF - B * E; t12[2] = B * F + A * DE; t12[3] = x;
  t12[4] = B * Cc; t12[5] = A * E + B * DF; t12[6] = B * DE - A * F; t12[7] = y;
  t12[8] = -D;     t12[9] = Cc * F;         t12[10] = Cc * E;        t12[11] = z;
}
void orc_transform_cloud_rpy(const float* pts4, int n, const float* pose6, float* out4) {
  float t[12];
  orc_rpy_matrix(pose6, t);
  for (int i = 0; i < n; i++) {
    const float* p = pts4 + 4 * i;
    out4[4 * i + 0] = t[0] * p[0] + t[1] * p[1] + t[2] * p[2] + t[3];
    out4[4 * i + 1] = t[4] * p[0] + t[5] * p[1] + t[6] * p[2] + t[7];
    out4[4 * i + 2] = t[8] * p[0] + t[9] * p[1] + t[10] * p[2] + t[11];
    out4[4 * i + 3] = p[3];
  }
}

// pointBodyToWorld / RGBpointBodyToWorld (laserMapping.cpp:1077-1110): double math, stored to float, intensity copied.
void orc_body_to_world4(const double* state26, const float* pts4, int n, float* out4) {
  const Q4 rot{state26[3], state26[4], state26[5], state26[6]};
  const Q4 offR{state26[7], state26[8], state26[9], state26[10]};
  for (int i = 0; i < n; i++) {
    const double p[3] = {(double)pts4[4 * i], (double)pts4[4 * i + 1], (double)pts4[4 * i + 2]};
    double a[3], g[3];
    qrot(offR, p, a);
    for (int k = 0; k < 3; k++) a[k] += state26[11 + k];
    qrot(rot, a, g);
    for (int k = 0; k < 3; k++) out4[4 * i + k] = (float)(g[k] + state26[k]);
    out4[4 * i + 3] = pts4[4 * i + 3];
  }
}

}  // extern "C"
