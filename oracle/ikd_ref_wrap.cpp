// TEST INFRASTRUCTURE ONLY (oracle). Not part of the shipped product path.
//
// Thin extern "C" harness around the REFERENCE's own, unmodified ikd-Tree
// (/root/reference/include/ikd-Tree/ikd_Tree.{h,cpp}), which oracle/Makefile compiles in place into
// oracle/_ref/libikd_ref.so. Nothing of the reference is copied here: this file only calls its public API
// (ikd_Tree.h:225-249) the way src/laserMapping.cpp does (Build :2339, Nearest_Search :1909,
// Add_Points :1492-1493, Delete_Point_Boxes :1198, reconstruct :656, flatten :2364, size/validnum :2343-2344).
//
// It is the pin for the map/k-NN half of the oracle and the "reference" kind of CPU baseline in bench.py.
#include <ikd-Tree/ikd_Tree.h>
#include <omp.h>
#include <cstdint>

using PointT = pcl::PointXYZINormal;
using Tree = KD_TREE<PointT>;
using PV = Tree::PointVector;

static inline PointT mk(const float* p) {
  PointT q;
  std::memset(&q, 0, sizeof(q));
  q.x = p[0];
  q.y = p[1];
  q.z = p[2];
  return q;
}

extern "C" {

// KD_TREE embeds a 1e6-entry op log inline (ikd_Tree.h:18,45) -> must live on the heap.
void* ikdref_create(float delete_param, float balance_param, float box_length) {
  return new Tree(delete_param, balance_param, box_length);
}
void ikdref_destroy(void* h) { delete static_cast<Tree*>(h); }
void ikdref_set_downsample(void* h, float ds) { static_cast<Tree*>(h)->set_downsample_param(ds); }
int ikdref_has_root(void* h) { return static_cast<Tree*>(h)->Root_Node != nullptr; }

void ikdref_build(void* h, const float* xyz, int n) {
  PV v(n);
  for (int i = 0; i < n; i++) v[i] = mk(xyz + 3 * i);
  static_cast<Tree*>(h)->Build(v);
}
void ikdref_reconstruct(void* h, const float* xyz, int n) {
  PV v(n);
  for (int i = 0; i < n; i++) v[i] = mk(xyz + 3 * i);
  static_cast<Tree*>(h)->reconstruct(v);
}

// nq queries, k neighbours each. out_xyz[nq*k*3], out_d2[nq*k], out_cnt[nq]. Unfilled entries are NaN/INF.
// threads<=0 -> omp default. Mirrors the OpenMP loop of h_share_model (laserMapping.cpp:1883-1909).
void ikdref_nearest(void* h, const float* q, int nq, int k, float* out_xyz, float* out_d2, int* out_cnt, int threads) {
  Tree* t = static_cast<Tree*>(h);
  if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 256)
  for (int i = 0; i < nq; i++) {
    PV nn;
    std::vector<float> d2;
    t->Nearest_Search(mk(q + 3 * i), k, nn, d2);
    int c = (int)nn.size();
    out_cnt[i] = c;
    for (int j = 0; j < k; j++) {
      if (j < c) {
        out_xyz[(size_t)(i * (size_t)k + j) * 3 + 0] = nn[j].x;
        out_xyz[(size_t)(i * (size_t)k + j) * 3 + 1] = nn[j].y;
        out_xyz[(size_t)(i * (size_t)k + j) * 3 + 2] = nn[j].z;
        out_d2[i * (size_t)k + j] = d2[j];
      } else {
        out_xyz[(size_t)(i * (size_t)k + j) * 3 + 0] = NAN;
        out_xyz[(size_t)(i * (size_t)k + j) * 3 + 1] = NAN;
        out_xyz[(size_t)(i * (size_t)k + j) * 3 + 2] = NAN;
        out_d2[i * (size_t)k + j] = INFINITY;
      }
    }
  }
}

// same with the reference's max_dist argument (ikd_Tree.h:236, candidates with dist > max_dist^2 are not accepted)
void ikdref_nearest_md(void* h, const float* q, int nq, int k, double max_dist, float* out_xyz, float* out_d2, int* out_cnt) {
  Tree* t = static_cast<Tree*>(h);
#pragma omp parallel for schedule(dynamic, 256)
  for (int i = 0; i < nq; i++) {
    PV nn;
    std::vector<float> d2;
    t->Nearest_Search(mk(q + 3 * i), k, nn, d2, max_dist);
    int c = (int)nn.size();
    out_cnt[i] = c;
    for (int j = 0; j < k; j++) {
      const size_t o = (size_t)i * k + j;
      out_xyz[3 * o] = j < c ? nn[j].x : NAN;
      out_xyz[3 * o + 1] = j < c ? nn[j].y : NAN;
      out_xyz[3 * o + 2] = j < c ? nn[j].z : NAN;
      out_d2[o] = j < c ? d2[j] : INFINITY;
    }
  }
}

// orc_knn5_fn-compatible adapter (see lio_oracle.cpp): 5-NN with the thread count set by ikdref_set_threads.
static int g_knn_threads = 0;
void ikdref_set_threads(int t) { g_knn_threads = t; }
void ikdref_knn5(void* h, const float* q, int nq, float* out_xyz, float* out_d2, int* out_cnt) {
  ikdref_nearest(h, q, nq, 5, out_xyz, out_d2, out_cnt, g_knn_threads);
}

int ikdref_add_points(void* h, const float* xyz, int n, int downsample_on) {
  if (n <= 0) return 0;
  PV v(n);
  for (int i = 0; i < n; i++) v[i] = mk(xyz + 3 * i);
  return static_cast<Tree*>(h)->Add_Points(v, downsample_on != 0);
}

// boxes: nb * 6 floats (min xyz, max xyz), the BoxPointType of ikd_Tree.h:32-35.
int ikdref_delete_boxes(void* h, const float* boxes, int nb) {
  std::vector<BoxPointType> b(nb);
  for (int i = 0; i < nb; i++)
    for (int a = 0; a < 3; a++) {
      b[i].vertex_min[a] = boxes[6 * i + a];
      b[i].vertex_max[a] = boxes[6 * i + 3 + a];
    }
  return static_cast<Tree*>(h)->Delete_Point_Boxes(b);
}

void ikdref_delete_points(void* h, const float* xyz, int n) {
  PV v(n);
  for (int i = 0; i < n; i++) v[i] = mk(xyz + 3 * i);
  static_cast<Tree*>(h)->Delete_Points(v);
}

// ---- the same with PointType::intensity (x, y, z, intensity records): the reference tree stores whole points
static inline PointT mk4(const float* p) {
  PointT q = mk(p);
  q.intensity = p[3];
  return q;
}
void ikdref_build_i(void* h, const float* xyzi, int n) {
  PV v(n);
  for (int i = 0; i < n; i++) v[i] = mk4(xyzi + 4 * i);
  static_cast<Tree*>(h)->Build(v);
}
int ikdref_add_points_i(void* h, const float* xyzi, int n, int downsample_on) {
  if (n <= 0) return 0;
  PV v(n);
  for (int i = 0; i < n; i++) v[i] = mk4(xyzi + 4 * i);
  return static_cast<Tree*>(h)->Add_Points(v, downsample_on != 0);
}
int ikdref_flatten_i(void* h, float* out, int cap) {
  Tree* t = static_cast<Tree*>(h);
  PV st;
  if (t->Root_Node == nullptr) return 0;
  t->flatten(t->Root_Node, st, NOT_RECORD);
  int n = (int)st.size();
  if (out)
    for (int i = 0; i < n && i < cap; i++) {
      out[4 * i + 0] = st[i].x;
      out[4 * i + 1] = st[i].y;
      out[4 * i + 2] = st[i].z;
      out[4 * i + 3] = st[i].intensity;
    }
  return n;
}
// k-NN returning the neighbours' intensities as well: out_xyzi[nq*k*4]
void ikdref_nearest_i(void* h, const float* q, int nq, int k, float* out_xyzi, float* out_d2, int* out_cnt) {
  Tree* t = static_cast<Tree*>(h);
#pragma omp parallel for schedule(dynamic, 256)
  for (int i = 0; i < nq; i++) {
    PV nn;
    std::vector<float> d2;
    t->Nearest_Search(mk(q + 3 * i), k, nn, d2);
    int c = (int)nn.size();
    out_cnt[i] = c;
    for (int j = 0; j < k; j++) {
      const size_t o = (size_t)i * k + j;
      out_xyzi[4 * o] = j < c ? nn[j].x : NAN;
      out_xyzi[4 * o + 1] = j < c ? nn[j].y : NAN;
      out_xyzi[4 * o + 2] = j < c ? nn[j].z : NAN;
      out_xyzi[4 * o + 3] = j < c ? nn[j].intensity : NAN;
      out_d2[o] = j < c ? d2[j] : INFINITY;
    }
  }
}

int ikdref_size(void* h) { return static_cast<Tree*>(h)->size(); }
int ikdref_validnum(void* h) { return static_cast<Tree*>(h)->validnum(); }

// Returns number of valid points; writes up to cap of them (xyz) when out != nullptr.
int ikdref_flatten(void* h, float* out, int cap) {
  Tree* t = static_cast<Tree*>(h);
  PV st;
  if (t->Root_Node == nullptr) return 0;
  t->flatten(t->Root_Node, st, NOT_RECORD);
  int n = (int)st.size();
  if (out)
    for (int i = 0; i < n && i < cap; i++) {
      out[3 * i + 0] = st[i].x;
      out[3 * i + 1] = st[i].y;
      out[3 * i + 2] = st[i].z;
    }
  return n;
}

}  // extern "C"
