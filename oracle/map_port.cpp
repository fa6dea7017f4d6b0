// ============================================================================================================
// TEST INFRASTRUCTURE ONLY — CPU ORACLE ("port" kind).  Never used by the shipped GPU path.
//
// Restatement of the LOGICAL semantics of the reference ikd-Tree (include/ikd-Tree/ikd_Tree.cpp) as a plain point
// set, for use where oracle/_ref/libikd_ref.so (the reference compiled unmodified) is unavailable, and as a second
// opinion on it.  The k-d tree's balancing / lazy-delete / rebuild machinery does not change the set of valid
// points nor exact k-NN answers, so none of it is restated.
//
//   Nearest_Search  ikd_Tree.cpp:366-397,868-1013  exact k-NN over valid points, float sq-distances
//                   (calc_dist :1373-1378), ascending.  Equal-distance ties: (dist, x, y, z) order here; the
//                   reference's tie order depends on tree traversal (ikd_Tree.h:102-105) — documented deviation.
//   Add_Points      ikd_Tree.cpp:413-489  downsample_on: per point, box = its voxel (float arithmetic :424-432),
//                   winner = closest to the voxel centre among {new point, valid points in the box}; if the box held
//                   >1 point or the new point wins: box cleared, winner inserted.  downsample off: plain insert.
//   Delete_Point_Boxes ikd_Tree.cpp:535-556,647-710  half-open box test min <= p < max (:670); returns #deleted.
//   Build / reconstruct ikd_Tree.cpp:352-364,1393-1405  verbatim insert, no dedupe.
// ============================================================================================================
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <unordered_map>
#include <algorithm>
#include <omp.h>

namespace {

struct P3 { float x, y, z; };

struct PortMap {
  float ds = 0.2f;  // downsample_size
  std::vector<P3> pts;
  std::vector<unsigned char> alive;
  std::unordered_map<uint64_t, std::vector<int>> cells;  // voxel key -> indices (alive or dead)
  int nvalid = 0;
  int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  bool has_bounds = false;
};

inline int vox(float x, float ds) { return (int)std::floor(x / ds); }  // float division, as ikd_Tree.cpp:424
inline uint64_t keyof(int ix, int iy, int iz) {
  const uint64_t B = 1u << 20;
  return ((uint64_t)(ix + (int64_t)B) << 42) | ((uint64_t)(iy + (int64_t)B) << 21) | (uint64_t)(iz + (int64_t)B);
}
inline float d2f(const P3& a, const P3& b) {
  return (a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z);
}

void raw_insert(PortMap* m, P3 p) {
  int id = (int)m->pts.size();
  m->pts.push_back(p);
  m->alive.push_back(1);
  int c[3] = {vox(p.x, m->ds), vox(p.y, m->ds), vox(p.z, m->ds)};
  m->cells[keyof(c[0], c[1], c[2])].push_back(id);
  m->nvalid++;
  for (int a = 0; a < 3; a++) {
    if (!m->has_bounds) { m->lo[a] = m->hi[a] = c[a]; }
    else { m->lo[a] = std::min(m->lo[a], c[a]); m->hi[a] = std::max(m->hi[a], c[a]); }
  }
  m->has_bounds = true;
}

struct Cand { float d; P3 p; };
inline bool cand_less(const Cand& a, const Cand& b) {
  if (a.d != b.d) return a.d < b.d;
  if (a.p.x != b.p.x) return a.p.x < b.p.x;
  if (a.p.y != b.p.y) return a.p.y < b.p.y;
  return a.p.z < b.p.z;
}

void knn_one(const PortMap* m, P3 q, int k, std::vector<Cand>& best) {
  best.clear();
  if (m->nvalid == 0) return;
  const float ds = m->ds;
  int c[3] = {vox(q.x, ds), vox(q.y, ds), vox(q.z, ds)};
  auto push = [&](const P3& p) {
    Cand cd{d2f(q, p), p};
    if ((int)best.size() < k) {
      best.push_back(cd);
      std::sort(best.begin(), best.end(), cand_less);
    } else if (cand_less(cd, best.back())) {
      best.back() = cd;
      std::sort(best.begin(), best.end(), cand_less);
    }
  };
  // max ring needed to cover the occupied bounding box
  int maxring = 0;
  for (int a = 0; a < 3; a++) maxring = std::max(maxring, std::max(std::abs(c[a] - m->lo[a]), std::abs(c[a] - m->hi[a])));
  const int RING_CAP = 24;
  for (int r = 0; r <= std::min(maxring, RING_CAP); r++) {
    for (int dz = -r; dz <= r; dz++)
      for (int dy = -r; dy <= r; dy++)
        for (int dx = -r; dx <= r; dx++) {
          if (std::max(std::abs(dx), std::max(std::abs(dy), std::abs(dz))) != r) continue;
          auto it = m->cells.find(keyof(c[0] + dx, c[1] + dy, c[2] + dz));
          if (it == m->cells.end()) continue;
          for (int id : it->second)
            if (m->alive[id]) push(m->pts[id]);
        }
    if ((int)best.size() == k) {
      // everything within `cover` of q has been visited (with a safety margin for float voxel rounding)
      double cover = 1e30;
      const double qd[3] = {q.x, q.y, q.z};
      for (int a = 0; a < 3; a++) {
        double lo = (double)(c[a] - r) * (double)ds, hi = (double)(c[a] + r + 1) * (double)ds;
        cover = std::min(cover, std::min(qd[a] - lo, hi - qd[a]));
      }
      cover -= 1e-4 * (double)ds + 1e-5 * (std::fabs(qd[0]) + std::fabs(qd[1]) + std::fabs(qd[2])) * 1e-2;
      if (cover > 0 && (double)best.back().d < cover * cover) return;
    }
    if (r == maxring) return;  // whole map visited
  }
  // far query: brute force over everything
  best.clear();
  for (size_t id = 0; id < m->pts.size(); id++)
    if (m->alive[id]) push(m->pts[id]);
}

}  // namespace

extern "C" {

void* mapport_create(float box_length) {
  PortMap* m = new PortMap();
  m->ds = box_length;
  return m;
}
void mapport_destroy(void* h) { delete static_cast<PortMap*>(h); }
void mapport_set_downsample(void* h, float ds) { static_cast<PortMap*>(h)->ds = ds; }

static void clear_all(PortMap* m) {
  m->pts.clear(); m->alive.clear(); m->cells.clear(); m->nvalid = 0; m->has_bounds = false;
}
void mapport_build(void* h, const float* xyz, int n) {
  PortMap* m = static_cast<PortMap*>(h);
  clear_all(m);
  for (int i = 0; i < n; i++) raw_insert(m, P3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]});
}
void mapport_reconstruct(void* h, const float* xyz, int n) { mapport_build(h, xyz, n); }

void mapport_nearest(void* h, const float* q, int nq, int k, float* out_xyz, float* out_d2, int* out_cnt, int threads) {
  PortMap* m = static_cast<PortMap*>(h);
  if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel
  {
    std::vector<Cand> best;
#pragma omp for schedule(dynamic, 64)
    for (int i = 0; i < nq; i++) {
      knn_one(m, P3{q[3 * i], q[3 * i + 1], q[3 * i + 2]}, k, best);
      int c = (int)best.size();
      out_cnt[i] = c;
      for (int j = 0; j < k; j++) {
        size_t o = (size_t)i * k + j;
        if (j < c) {
          out_xyz[3 * o] = best[j].p.x; out_xyz[3 * o + 1] = best[j].p.y; out_xyz[3 * o + 2] = best[j].p.z;
          out_d2[o] = best[j].d;
        } else {
          out_xyz[3 * o] = out_xyz[3 * o + 1] = out_xyz[3 * o + 2] = NAN;
          out_d2[o] = INFINITY;
        }
      }
    }
  }
}
// orc_knn5_fn-compatible adapter (see lio_oracle.cpp)
void mapport_knn5(void* h, const float* q, int nq, float* out_xyz, float* out_d2, int* out_cnt) {
  mapport_nearest(h, q, nq, 5, out_xyz, out_d2, out_cnt, 0);
}

int mapport_add_points(void* h, const float* xyz, int n, int downsample_on) {
  PortMap* m = static_cast<PortMap*>(h);
  const float ds = m->ds;
  int counter = 0;
  for (int i = 0; i < n; i++) {
    P3 p{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    if (!downsample_on) { raw_insert(m, p); continue; }
    // ikd_Tree.cpp:424-432 (float), centre = min + (max-min)/2.0 (double, stored to float)
    float bmin[3], bmax[3];
    P3 mid;
    const float pc[3] = {p.x, p.y, p.z};
    float* midp[3] = {&mid.x, &mid.y, &mid.z};
    int c[3];
    for (int a = 0; a < 3; a++) {
      c[a] = (int)std::floor(pc[a] / ds);
      bmin[a] = std::floor(pc[a] / ds) * ds;
      bmax[a] = bmin[a] + ds;
      *midp[a] = (float)(bmin[a] + (bmax[a] - bmin[a]) / 2.0);
    }
    // Search_by_range: valid points with min <= p < max.  Candidates: the voxel itself and (for 1-ulp boundary
    // disagreements between floor(x/ds) and the box test) its 26 neighbours.
    std::vector<int> in_box;
    for (int dz = -1; dz <= 1; dz++)
      for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
          auto it = m->cells.find(keyof(c[0] + dx, c[1] + dy, c[2] + dz));
          if (it == m->cells.end()) continue;
          for (int id : it->second) {
            if (!m->alive[id]) continue;
            const P3& e = m->pts[id];
            if (e.x >= bmin[0] && e.x < bmax[0] && e.y >= bmin[1] && e.y < bmax[1] && e.z >= bmin[2] && e.z < bmax[2])
              in_box.push_back(id);
          }
        }
    float min_dist = d2f(p, mid);
    int winner = -1;  // -1 = the new point
    for (int id : in_box) {
      float t = d2f(m->pts[id], mid);
      if (t < min_dist) { min_dist = t; winner = id; }
    }
    P3 res = winner < 0 ? p : m->pts[winner];
    bool same = std::fabs(p.x - res.x) < 1e-6 && std::fabs(p.y - res.y) < 1e-6 && std::fabs(p.z - res.z) < 1e-6;
    if (in_box.size() > 1 || same) {
      for (int id : in_box) { m->alive[id] = 0; m->nvalid--; }
      raw_insert(m, res);
      counter++;
    }
  }
  return counter;
}

int mapport_delete_boxes(void* h, const float* boxes, int nb) {
  PortMap* m = static_cast<PortMap*>(h);
  int cnt = 0;
  for (int b = 0; b < nb; b++) {
    const float* mn = boxes + 6 * b;
    const float* mx = boxes + 6 * b + 3;
    for (size_t id = 0; id < m->pts.size(); id++) {
      if (!m->alive[id]) continue;
      const P3& e = m->pts[id];
      if (e.x >= mn[0] && e.x < mx[0] && e.y >= mn[1] && e.y < mx[1] && e.z >= mn[2] && e.z < mx[2]) {
        m->alive[id] = 0;
        m->nvalid--;
        cnt++;
      }
    }
  }
  return cnt;
}

int mapport_validnum(void* h) { return static_cast<PortMap*>(h)->nvalid; }
int mapport_size(void* h) { return static_cast<PortMap*>(h)->nvalid; }
int mapport_flatten(void* h, float* out, int cap) {
  PortMap* m = static_cast<PortMap*>(h);
  int n = 0;
  for (size_t id = 0; id < m->pts.size(); id++)
    if (m->alive[id]) {
      if (out && n < cap) { out[3 * n] = m->pts[id].x; out[3 * n + 1] = m->pts[id].y; out[3 * n + 2] = m->pts[id].z; }
      n++;
    }
  return n;
}

}  // extern "C"
