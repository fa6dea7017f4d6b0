// Compiles the C++ facades the way src/laserMapping.cpp uses the reference headers (KD_TREE<PointType> global,
// h_share_model callback with state_ikfom / dyn_share_datastruct look-alikes) and, when a GPU is present, runs them.
// Built by tests/test_facade_cpp.py with:  g++ -Ioracle/shim -Iinclude tests/cpp/facade_smoke.cpp -Lbetter_fastlio2_b200 -lfastlio_b200
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include <fastlio_b200/ikd_tree_facade.hpp>
#include <fastlio_b200/lio_gpu_frontend.hpp>
#include <fastlio_b200/scan_frontend_facade.hpp>

typedef pcl::PointXYZINormal PointType;                       // common_lib.h:161
typedef std::vector<PointType, Eigen::aligned_allocator<PointType>> PointVector;  // common_lib.h:163

// minimal stand-ins with the member names of use-ikfom.hpp:21-30 / esekfom.hpp:79-89 (Eigen is absent in this image)
struct Vec3 { double v[3]; double operator[](int i) const { return v[i]; } };
struct Quat { double c[4]; const double* coeffs() const { return c; } };
struct state_ikfom { Vec3 pos; Quat rot; Quat offset_R_L_I; Vec3 offset_T_L_I, vel, bg, ba, grav; };
struct MatX { std::vector<double> a; int r = 0, c = 0; void resize(int R, int C) { r = R; c = C; a.assign((size_t)R * C, 0.0); } double* data() { return a.data(); } int rows() const { return r; } };
struct VecX { std::vector<double> a; void resize(int n) { a.assign(n, 0.0); } double* data() { return a.data(); } };
struct dyn_share_datastruct { bool valid = true, converge = true; MatX h_x; VecX h; };

struct Pose6D { double offset_time, acc[3], gyr[3], vel[3], pos[3], rot[9]; };   // msg/Pose6D.msg
struct PointCloudXYZI { PointVector points; };                                    // pcl::PointCloud<PointType> look-alike

KD_TREE<PointType> ikdtree;  // laserMapping.cpp:116
flb::LioGpu gpu;
void h_share_model(state_ikfom& s, dyn_share_datastruct& d) { gpu.h_share_model(s, d); }

int main() {
  if (flb_device_count() <= 0) { std::printf("NO_GPU compile-only ok\n"); return 0; }
  std::mt19937 rng(5);
  std::uniform_real_distribution<float> U(-10.f, 10.f);
  std::normal_distribution<float> N(0.f, 0.01f);
  PointVector cloud;
  for (int i = 0; i < 60000; ++i) { PointType p{}; p.x = U(rng); p.y = U(rng); p.z = N(rng); cloud.push_back(p); }           // ground
  for (int i = 0; i < 30000; ++i) { PointType p{}; p.x = 6.f + N(rng); p.y = U(rng); p.z = 0.3f * (U(rng) + 10.f); cloud.push_back(p); }  // wall
  for (size_t i = 0; i < cloud.size(); ++i) cloud[i].intensity = (float)(i % 251);   // whole records live in the map (ikd_Tree.h:64-86)
  ikdtree.set_capacity(1 << 20, 1 << 16);
  if (ikdtree.Root_Node != nullptr) return 2;
  ikdtree.set_downsample_param(0.2f);
  ikdtree.Build(cloud);
  if (ikdtree.Root_Node == nullptr || ikdtree.validnum() != (int)cloud.size() || ikdtree.size() != (int)cloud.size()) return 3;
  PointVector near; std::vector<float> d2;
  PointType q{}; q.x = 1.f; q.y = 2.f; q.z = 0.05f;
  ikdtree.Nearest_Search(q, 5, near, d2);
  if (near.size() != 5 || !(d2[0] <= d2[4]) || d2[4] > 0.2f) return 4;
  // brute-force check of the nearest distance
  float best = 1e30f;
  for (auto& p : cloud) { float d = (p.x - q.x) * (p.x - q.x) + (p.y - q.y) * (p.y - q.y) + (p.z - q.z) * (p.z - q.z); if (d < best) best = d; }
  if (best != d2[0]) return 5;
  // the neighbours come back with their intensity (Nearest_Points are whole PointType records, ikd_Tree.cpp:391-395)
  for (auto& nb : near) {
    bool ok = false;
    for (auto& p : cloud) if (p.x == nb.x && p.y == nb.y && p.z == nb.z) { ok = p.intensity == nb.intensity; break; }
    if (!ok) return 40;
  }
  {
    PointVector all;
    ikdtree.flatten(ikdtree.Root_Node, all, NOT_RECORD);
    double si = 0, so = 0;
    for (auto& p : cloud) si += p.intensity;
    for (auto& p : all) so += p.intensity;
    if (all.size() != cloud.size() || si != so) return 41;
  }
  PointVector add = cloud; add.resize(2000);
  for (auto& p : add) p.z += 0.05f;
  ikdtree.Add_Points(add, true);
  std::vector<BoxPointType> boxes(1);
  boxes[0] = BoxPointType{{-10.f, -10.f, -1.f}, {0.f, 10.f, 10.f}};
  const int before = ikdtree.validnum();
  const int nd = ikdtree.Delete_Point_Boxes(boxes);
  if (nd <= 0 || ikdtree.validnum() != before - nd) return 6;
  PointVector().swap(ikdtree.PCL_Storage);
  ikdtree.flatten(ikdtree.Root_Node, ikdtree.PCL_Storage, NOT_RECORD);  // laserMapping.cpp:2363-2364
  if ((int)ikdtree.PCL_Storage.size() != ikdtree.validnum()) return 7;

  // measurement callback on a scan = a noisy subsample of the remaining cloud seen from the origin
  if (!gpu.attach(ikdtree.handle(), false, 3, 0.2)) return 8;
  PointVector body;
  for (size_t i = 0; i < ikdtree.PCL_Storage.size(); i += 7) body.push_back(ikdtree.PCL_Storage[i]);
  gpu.begin_scan(&body[0].x, (int)body.size(), sizeof(PointType));
  state_ikfom s{};
  s.rot.c[3] = 1.0; s.offset_R_L_I.c[3] = 1.0; s.grav.v[2] = -9.809;
  s.pos.v[0] = 0.02; s.pos.v[2] = -0.01;  // small prior error
  dyn_share_datastruct d;
  h_share_model(s, d);
  if (!d.valid || d.h_x.rows() != gpu.effct_feat_num || gpu.effct_feat_num < 1000) return 9;
  const int M = d.h_x.rows();
  gpu.set_row_mode(flb::LioGpu::COMPRESSED_ROWS);
  dyn_share_datastruct d2s; d2s.converge = false;
  h_share_model(s, d2s);
  if (!d2s.valid || d2s.h_x.rows() != 24) return 10;
  // compressed rows reproduce H^T H and H^T h of the exact rows
  double maxerr = 0, scale = 0;
  for (int a = 0; a < 12; ++a)
    for (int b = 0; b < 12; ++b) {
      double e = 0, c = 0;
      for (int r = 0; r < M; ++r) e += d.h_x.a[(size_t)a * M + r] * d.h_x.a[(size_t)b * M + r];
      for (int r = 0; r < 24; ++r) c += d2s.h_x.a[(size_t)a * 24 + r] * d2s.h_x.a[(size_t)b * 24 + r];
      maxerr = std::fmax(maxerr, std::fabs(e - c)); scale = std::fmax(scale, std::fabs(e));
    }
  if (maxerr > 1e-9 * scale) { std::printf("HTH mismatch %g / %g\n", maxerr, scale); return 11; }
  const int added = gpu.map_incremental(s, true);

  // ---- front-end facades: UndistortPcl backward pass -> VoxelGrid -> h_share_model without a host round trip
  flb::ScanFrontEnd fe;
  if (!fe.attach(gpu.handle(), 1 << 17)) return 12;
  PointCloudXYZI lidar, pcl_out, feats_down_body, world;
  for (size_t i = 0; i < body.size(); ++i) {
    PointType p = body[i];
    p.intensity = (float)(i % 200);
    p.curvature = (float)((i * 37) % 1000) * 0.1f;   // 0..99.9 ms, shuffled
    lidar.points.push_back(p);
  }
  std::vector<Pose6D> IMUpose(3);
  for (int k = 0; k < 3; ++k) {   // a sensor at rest: the compensation must be the identity
    Pose6D q{};
    q.offset_time = 0.05 * k;
    for (int i = 0; i < 3; ++i) q.pos[i] = s.pos.v[i];
    q.rot[0] = q.rot[4] = q.rot[8] = 1.0;
    IMUpose[k] = q;
  }
  if (!fe.undistort(lidar, IMUpose, s, &pcl_out) || pcl_out.points.size() != lidar.points.size()) return 13;
  double sum_in = 0, sum_out = 0;
  for (size_t i = 0; i < lidar.points.size(); ++i) {
    if (i && pcl_out.points[i].curvature < pcl_out.points[i - 1].curvature) return 14;   // time order
    sum_in += (double)lidar.points[i].x + lidar.points[i].y + lidar.points[i].z + lidar.points[i].intensity;
    sum_out += (double)pcl_out.points[i].x + pcl_out.points[i].y + pcl_out.points[i].z + pcl_out.points[i].intensity;
  }
  if (std::fabs(sum_in - sum_out) > 1e-6 * std::fabs(sum_in)) { std::printf("undistort identity broken %g %g\n", sum_in, sum_out); return 15; }
  flb::VoxelGridGpu<PointType> downSizeFilterSurf(&fe);
  downSizeFilterSurf.setLeafSize(0.5f, 0.5f, 0.5f);       // laserMapping.cpp:2135
  downSizeFilterSurf.setInputCloud(&pcl_out);              // :2322
  downSizeFilterSurf.filter(feats_down_body);              // :2323
  const int feats_down_size = (int)feats_down_body.points.size();
  if (feats_down_size < 100 || feats_down_size >= (int)lidar.points.size()) return 16;
  gpu.set_row_mode(flb::LioGpu::EXACT_ROWS);
  dyn_share_datastruct d3;
  h_share_model(s, d3);                                    // runs on the filtered scan the front end left on the device
  if (!d3.valid || d3.h_x.rows() != gpu.effct_feat_num || gpu.effct_feat_num < 50) return 17;
  state_ikfom ident{};
  ident.rot.c[3] = 1.0; ident.offset_R_L_I.c[3] = 1.0;
  if (!fe.to_world(ident, false, world, 1 << 17) || (int)world.points.size() != feats_down_size) return 18;
  for (int i = 0; i < feats_down_size; ++i)
    if (world.points[i].x != feats_down_body.points[i].x || world.points[i].intensity != feats_down_body.points[i].intensity) return 19;
  std::printf("FRONTEND_OK raw=%d down=%d M=%d\n", (int)lidar.points.size(), feats_down_size, gpu.effct_feat_num);
  std::printf("FACADE_OK M=%d deleted=%d map=%d added=%d\n", M, nd, ikdtree.validnum(), added);
  return 0;
}
