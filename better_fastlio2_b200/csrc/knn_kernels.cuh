// knn_kernels.cuh — K1: exact k-nearest-neighbour search on the hashed voxel map, one warp per query.
// Replaces KD_TREE::Nearest_Search / Search (include/ikd-Tree/ikd_Tree.cpp:366-397, :868-1013) as called from
// h_share_model (src/laserMapping.cpp:1909).  Semantics kept: exact over all valid points, unbounded range
// (max_dist = INFINITY by default, ikd_Tree.h:236), float squared distances ((dx*dx+dy*dy)+dz*dz, no FMA —
// calc_dist ikd_Tree.cpp:1373-1378), ascending output, fewer than k results when the map is small.
// Equal-distance ties are broken by (x,y,z) — the reference's order depends on tree traversal (ikd_Tree.h:102-105).
//
// Search plan per query (all exact, each phase only runs if the previous one could not prove completeness):
//   A  5x5x5 voxel stencil around the query voxel (always inside 2x2x2 blocks; 8 hash probes by lanes 0-7)
//   B  the 3x3x3 coarse cells (8x8x8 blocks each) around the query, block bitmaps prune by box distance
//   C  every other coarse cell in the coarse hash (far queries at the map frontier; rare)
// Completeness test: have k candidates and d_k < (distance from the query to the boundary of the searched region)^2.
// Roofline: HBM/L2-latency bound gather; algorithmic bytes 16 (query) + 80 (5 neighbours) + 80 (cache write).
#pragma once
#include "voxel_map.cuh"
#include <math_constants.h>

namespace flb {

constexpr unsigned FULL = 0xffffffffu;
constexpr int BLOCK_RINGS = 3;  // block shells searched by the exact kernel before it falls back to the coarse levels

template <int K>
struct TopK {
  float d[K], x[K], y[K], z[K];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int j = 0; j < K; ++j) { d[j] = CUDART_INF_F; x[j] = 0.f; y[j] = 0.f; z[j] = 0.f; }
  }
  static __device__ __forceinline__ bool less(float da, float xa, float ya, float za, float db, float xb, float yb, float zb) {
    if (da != db) return da < db;
    if (xa != xb) return xa < xb;
    if (ya != yb) return ya < yb;
    return za < zb;
  }
  __device__ __forceinline__ void insert(float dd, float px, float py, float pz) {
    if (!less(dd, px, py, pz, d[K - 1], x[K - 1], y[K - 1], z[K - 1])) return;
    d[K - 1] = dd; x[K - 1] = px; y[K - 1] = py; z[K - 1] = pz;
#pragma unroll
    for (int j = K - 1; j > 0; --j) {
      if (less(d[j], x[j], y[j], z[j], d[j - 1], x[j - 1], y[j - 1], z[j - 1])) {
        float t;
        t = d[j]; d[j] = d[j - 1]; d[j - 1] = t;
        t = x[j]; x[j] = x[j - 1]; x[j - 1] = t;
        t = y[j]; y[j] = y[j - 1]; y[j - 1] = t;
        t = z[j]; z[j] = z[j - 1]; z[j - 1] = t;
      }
    }
  }
  __device__ __forceinline__ void pop_front() {
#pragma unroll
    for (int j = 0; j < K - 1; ++j) { d[j] = d[j + 1]; x[j] = x[j + 1]; y[j] = y[j + 1]; z[j] = z[j + 1]; }
    d[K - 1] = CUDART_INF_F;
  }
};

__device__ __forceinline__ float sqdist(float qx, float qy, float qz, float px, float py, float pz) {
  const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Merge the 32 lane-local sorted lists into the global top-K.  On return: lane r (< gcount) holds result r in
// (rd,rx,ry,rz); every lane's list is cleared except lane r which re-inserts result r; returns gcount and the
// current k-th distance (INF if fewer than K found) in thr.
template <int K>
__device__ __forceinline__ int warp_merge(TopK<K>& t, int lane, float& rd, float& rx, float& ry, float& rz, float& thr) {
  int gcount = 0;
  float last = CUDART_INF_F;
  rd = CUDART_INF_F; rx = ry = rz = CUDART_NAN_F;
#pragma unroll 1
  for (int r = 0; r < K; ++r) {
    const unsigned v = __float_as_uint(t.d[0]);
    const unsigned mn = __reduce_min_sync(FULL, v);
    if (mn == 0x7f800000u) break;
    unsigned who = __ballot_sync(FULL, v == mn);
    int src = __ffs(who) - 1;
    if (who & (who - 1)) {  // several lanes tie on distance: lexicographic (x,y,z)
      float bx = __shfl_sync(FULL, t.x[0], src), by = __shfl_sync(FULL, t.y[0], src), bz = __shfl_sync(FULL, t.z[0], src);
      unsigned rest = who & ~(1u << src);
      while (rest) {
        const int c = __ffs(rest) - 1;
        rest &= rest - 1;
        const float cx = __shfl_sync(FULL, t.x[0], c), cy = __shfl_sync(FULL, t.y[0], c), cz = __shfl_sync(FULL, t.z[0], c);
        if (cx < bx || (cx == bx && (cy < by || (cy == by && cz < bz)))) { bx = cx; by = cy; bz = cz; src = c; }
      }
    }
    const float gd = __shfl_sync(FULL, t.d[0], src), gx = __shfl_sync(FULL, t.x[0], src);
    const float gy = __shfl_sync(FULL, t.y[0], src), gz = __shfl_sync(FULL, t.z[0], src);
    if (lane == r) { rd = gd; rx = gx; ry = gy; rz = gz; }
    if (lane == src) t.pop_front();
    last = gd;
    ++gcount;
  }
  thr = (gcount == K) ? last : CUDART_INF_F;
  t.clear();
  if (lane < gcount) { t.d[0] = rd; t.x[0] = rx; t.y[0] = ry; t.z[0] = rz; }
  return gcount;
}

// K0: body -> world transform (laserMapping.cpp:1894-1898): double math (Eigen quaternion * vector form), result
// rounded to float.  R = s.rot, Roff = s.offset_R_L_I.
struct PoseDev {
  double rot[4];   // x,y,z,w
  double offR[4];
  double pos[3];
  double offT[3];
};
__device__ __forceinline__ void qrot_d(const double* q, double vx, double vy, double vz, double& ox, double& oy, double& oz) {
  double ux = __dsub_rn(__dmul_rn(q[1], vz), __dmul_rn(q[2], vy));
  double uy = __dsub_rn(__dmul_rn(q[2], vx), __dmul_rn(q[0], vz));
  double uz = __dsub_rn(__dmul_rn(q[0], vy), __dmul_rn(q[1], vx));
  ux = __dadd_rn(ux, ux); uy = __dadd_rn(uy, uy); uz = __dadd_rn(uz, uz);
  const double cx = __dsub_rn(__dmul_rn(q[1], uz), __dmul_rn(q[2], uy));
  const double cy = __dsub_rn(__dmul_rn(q[2], ux), __dmul_rn(q[0], uz));
  const double cz = __dsub_rn(__dmul_rn(q[0], uy), __dmul_rn(q[1], ux));
  ox = __dadd_rn(__dadd_rn(vx, __dmul_rn(q[3], ux)), cx);
  oy = __dadd_rn(__dadd_rn(vy, __dmul_rn(q[3], uy)), cy);
  oz = __dadd_rn(__dadd_rn(vz, __dmul_rn(q[3], uz)), cz);
}
__device__ __forceinline__ float4 body_to_world(const PoseDev& s, const float4 pb) {
  double ax, ay, az, gx, gy, gz;
  qrot_d(s.offR, (double)pb.x, (double)pb.y, (double)pb.z, ax, ay, az);
  ax = __dadd_rn(ax, s.offT[0]); ay = __dadd_rn(ay, s.offT[1]); az = __dadd_rn(az, s.offT[2]);
  qrot_d(s.rot, ax, ay, az, gx, gy, gz);
  return make_float4((float)__dadd_rn(gx, s.pos[0]), (float)__dadd_rn(gy, s.pos[1]), (float)__dadd_rn(gz, s.pos[2]), pb.w);
}

// Device-resident state of the iterated update (see esikf_device.cuh); the measurement kernels read the pose of the
// current iterate and the loop flags from here when running in device-driven mode (ctl != nullptr).
constexpr int NDOF = 23;
struct EsikfCtl {
  double x[26];        // current iterate x_
  double xp[26];       // x_propagated
  double Pp[NDOF * NDOF];
  double P[NDOF * NDOF];
  double limit[NDOF];
  double R;
  double last_res;
  PoseDev pose;        // pose of the current iterate
  int max_iter, it, t, converge, finished, need_host, passes, searches, lastM, n;
  int flg_inited;      // flg_EKF_inited of this scan (laserMapping.cpp:2317)
  int pad_;
};
__device__ __forceinline__ bool ctl_pass_active(const EsikfCtl* c) { return !c->finished && c->it < c->max_iter; }

struct KnnArgs {
  MapDev m;
  const float4* q;     // n world-frame query points (x,y,z,*)
  int n;
  float4* nbr;         // [K][n] : (x,y,z,d2) of the r-th neighbour of query i at nbr[r*n+i]
  unsigned char* cnt;  // [n] number of neighbours found
  float max_d2;        // max_dist^2 (INF: unbounded)
  int* phase_stats;    // optional [4]: queries finishing in phase A / B / C, total candidate points
  int* worklist;       // stencil kernel: indices of queries it could not prove complete; warp kernel: its input list
  int* work_count;     // number of entries in worklist (device)
  const EsikfCtl* ctl; // device-driven mode: queries = body_to_world(ctl->pose, body[i]); skipped unless a search pass
  const float4* body;
  int stride;          // leading dimension of nbr (>= n; the session capacity, so launches do not depend on n)
};

// visit every point of voxel slot `idx` (head + overflow chain)
template <int K>
__device__ __forceinline__ void visit_voxel(const MapDev& m, size_t idx, float qx, float qy, float qz, float lim, TopK<K>& t) {
  float4 e = __ldg(&m.slots[idx]);
  for (;;) {
    const float dd = sqdist(qx, qy, qz, e.x, e.y, e.z);
    if (dd <= lim) t.insert(dd, e.x, e.y, e.z);
    const int c = __float_as_int(e.w);
    if (c < 0) break;
    e = __ldg(&m.ovf[c]);
  }
}

// squared distance from q to the axis-aligned cell [lo,hi) per axis, shrunk by mg (conservative lower bound)
__device__ __forceinline__ float box_mind2(float qx, float qy, float qz, float lx, float ly, float lz, float hx, float hy, float hz, float mg) {
  const float gx = fmaxf(fmaxf(lx - qx, qx - hx) - mg, 0.f);
  const float gy = fmaxf(fmaxf(ly - qy, qy - hy) - mg, 0.f);
  const float gz = fmaxf(fmaxf(lz - qz, qz - hz) - mg, 0.f);
  return gx * gx + gy * gy + gz * gz;
}
// squared distance from q (inside) to the nearest face of [lo,hi), shrunk by mg; 0 if outside / too close
__device__ __forceinline__ float cover2(float qx, float qy, float qz, float lx, float ly, float lz, float hx, float hy, float hz, float mg) {
  float c = fminf(fminf(fminf(qx - lx, hx - qx), fminf(qy - ly, hy - qy)), fminf(qz - lz, hz - qz)) - mg;
  return c > 0.f ? c * c : 0.f;
}

// Scan the blocks flagged in `todo` (one candidate block per lane: its index `myblk`, occupancy `mymask` and block
// coordinates) with the WHOLE warp: two blocks per step, lane l reads slots l and l+32 of each, so the 4 point loads of
// a step are issued back to back (memory-level parallelism) before the insertions.  Voxels inside the phase-A
// stencil (|v - cv| <= 2) were already visited and are skipped when skip_stencil is set.
template <int K>
__device__ __forceinline__ void coop_scan_blocks(const MapDev& m, unsigned todo, int myblk, unsigned long long mymask, int mybx,
                                                 int myby, int mybz, int lane, float qx, float qy, float qz, int cvx, int cvy,
                                                 int cvz, bool skip_stencil, float limit, TopK<K>& t) {
  while (todo) {
    int src[2];
    src[0] = __ffs(todo) - 1;
    todo &= todo - 1;
    src[1] = todo ? __ffs(todo) - 1 : -1;
    if (src[1] >= 0) todo &= todo - 1;
    float4 e[4];
    bool v[4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int sl = src[u] >= 0 ? src[u] : 0;
      const int blk = __shfl_sync(FULL, myblk, sl);
      const unsigned long long mask = __shfl_sync(FULL, mymask, sl);
      const int bx = __shfl_sync(FULL, mybx, sl), by = __shfl_sync(FULL, myby, sl), bz = __shfl_sync(FULL, mybz, sl);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int s = lane + 32 * h;
        bool ok = src[u] >= 0 && ((mask >> s) & 1ull);
        if (ok && skip_stencil) {
          const int vx = bx * 4 + (s & 3), vy = by * 4 + ((s >> 2) & 3), vz = bz * 4 + (s >> 4);
          ok = !(abs(vx - cvx) <= 2 && abs(vy - cvy) <= 2 && abs(vz - cvz) <= 2);
        }
        v[u * 2 + h] = ok;
        if (ok) e[u * 2 + h] = __ldg(&m.slots[(size_t)blk * 64 + s]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (!v[j]) continue;
      float4 p = e[j];
      for (;;) {
        const float dd = sqdist(qx, qy, qz, p.x, p.y, p.z);
        if (dd <= limit) t.insert(dd, p.x, p.y, p.z);
        const int c = __float_as_int(p.w);
        if (c < 0) break;
        p = __ldg(&m.ovf[c]);
      }
    }
  }
}

// Process the blocks of one coarse cell (slot cs): per 32-bit group of the 512-bit block bitmap every lane tests its
// block (bit set, not already covered by phase B0, box distance within the bound), probes the hash for it in
// parallel, then the candidates are scanned co-operatively.
template <int K>
__device__ __forceinline__ void scan_coarse_cell(const MapDev& m, int cs, int lane, float qx, float qy, float qz,
                                                 int cvx, int cvy, int cvz, bool have, float thr, float lim, float mg,
                                                 TopK<K>& t) {
  const int qbx_ = cvx >> 2, qby_ = cvy >> 2, qbz_ = cvz >> 2;
  int ccx, ccy, ccz;
  unpack_key(__ldg(&m.ckeys[cs]), ccx, ccy, ccz);
  const float ds = m.ds;
  const float bs = 4.f * ds;
  const float bound = have ? thr : CUDART_INF_F;
#pragma unroll 1
  for (int k = 0; k < 8; ++k) {
    const unsigned long long word = __ldg(&m.cbits[(size_t)cs * 8 + k]);
    if (word == 0ull) continue;  // uniform
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int bit = k * 64 + h * 32 + lane;
      const int bx = ccx * 8 + (bit & 7), by = ccy * 8 + ((bit >> 3) & 7), bz = ccz * 8 + (bit >> 6);
      bool go = (word >> (h * 32 + lane)) & 1ull;
      if (go && abs(bx - qbx_) <= BLOCK_RINGS && abs(by - qby_) <= BLOCK_RINGS && abs(bz - qbz_) <= BLOCK_RINGS) go = false;  // visited by phase B0
      if (go) {
        const float lx = (float)bx * bs, ly = (float)by * bs, lz = (float)bz * bs;
        const float md = box_mind2(qx, qy, qz, lx, ly, lz, lx + bs, ly + bs, lz + bs, mg);
        go = !(md > bound || md > lim);
      }
      int blk = -1;
      unsigned long long mask = 0ull;
      if (go) {
        blk = find_block(m, pack_key(bx, by, bz));
        if (blk >= 0) mask = __ldg(&m.bmask[blk]);
      }
      const unsigned todo = __ballot_sync(FULL, blk >= 0 && mask != 0ull);
      if (todo) coop_scan_blocks<K>(m, todo, blk, mask, bx, by, bz, lane, qx, qy, qz, cvx, cvy, cvz, false, fminf(lim, bound), t);
    }
  }
}

template <int K>
__global__ void __launch_bounds__(128) k_knn(KnnArgs a) {
  const MapDev& m = a.m;
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  const float ds = m.ds;
  const float lim = a.max_d2;
  if (a.ctl && !(ctl_pass_active(a.ctl) && a.ctl->converge)) return;
  const int nwork = a.worklist ? *a.work_count : a.n;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nwork; w += warps_per_grid) {
    const int i = a.worklist ? a.worklist[w] : w;
    const float4 q4 = a.ctl ? body_to_world(a.ctl->pose, __ldg(&a.body[i])) : __ldg(&a.q[i]);
    const float qx = q4.x, qy = q4.y, qz = q4.z;
    TopK<K> t;
    t.clear();
    float rd = CUDART_INF_F, rx = CUDART_NAN_F, ry = CUDART_NAN_F, rz = CUDART_NAN_F, thr = CUDART_INF_F;
    int gcount = 0;
    int phase = 0;
    const float qlim = 4.0e6f * ds;
    const bool qok = fabsf(qx) < qlim && fabsf(qy) < qlim && fabsf(qz) < qlim;
    if (qok) {
      const int cvx = voxel_of(qx, ds), cvy = voxel_of(qy, ds), cvz = voxel_of(qz, ds);
      const float mg = 1e-3f * ds + 4.8e-7f * (fabsf(qx) + fabsf(qy) + fabsf(qz));
      // ---------------- phase A: 5x5x5 stencil, inside the 2x2x2 blocks starting at bb
      const int bbx = (cvx - 2) >> 2, bby = (cvy - 2) >> 2, bbz = (cvz - 2) >> 2;
      int myblk = -1;
      unsigned long long mymask = 0ull;
      if (lane < 8) {
        myblk = find_block(m, pack_key(bbx + (lane & 1), bby + ((lane >> 1) & 1), bbz + (lane >> 2)));
        if (myblk >= 0) mymask = __ldg(&m.bmask[myblk]);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int v = lane + 32 * r;
        const bool act = v < 125;
        const int vv = act ? v : 0;
        const int vx = cvx + (vv % 5) - 2, vy = cvy + ((vv / 5) % 5) - 2, vz = cvz + (vv / 25) - 2;
        const int bsel = ((vx >> 2) - bbx) + 2 * ((vy >> 2) - bby) + 4 * ((vz >> 2) - bbz);
        const int blk = __shfl_sync(FULL, myblk, bsel);
        const unsigned long long mask = __shfl_sync(FULL, mymask, bsel);
        const int s = (((vz & 3) << 2) + (vy & 3)) * 4 + (vx & 3);
        if (act && blk >= 0 && ((mask >> s) & 1ull)) visit_voxel<K>(m, (size_t)blk * 64 + s, qx, qy, qz, lim, t);
      }
      gcount = warp_merge<K>(t, lane, rd, rx, ry, rz, thr);
      float cov = cover2(qx, qy, qz, (float)(cvx - 2) * ds, (float)(cvy - 2) * ds, (float)(cvz - 2) * ds,
                         (float)(cvx + 3) * ds, (float)(cvy + 3) * ds, (float)(cvz + 3) * ds, mg);
      bool done = (gcount == K && thr < cov) || cov > lim;
      const int qbx = cvx >> 2, qby = cvy >> 2, qbz = cvz >> 2;
      if (!done) {
        // ---------------- phase B0: shells of BLOCKS around the query block, radius 1..BLOCK_RINGS (one shell block per
        // lane and round, hash probes in parallel, candidates scanned co-operatively).  Ring r covers >= 4r voxels around
        // the query; ring 3 (2.4 m at 0.2 m voxels) exceeds the 5 m^2 acceptance radius of h_share_model, so only
        // genuinely far queries (map frontier) go on to the coarse levels.
        phase = 1;
        const float bs4 = 4.f * ds;
#pragma unroll 1
        for (int r = 1; r <= BLOCK_RINGS && !done; ++r) {
          const int w = 2 * r + 1, nb = w * w * w;
          const float bound = gcount == K ? thr : CUDART_INF_F;
#pragma unroll 1
          for (int base = 0; base < nb; base += 32) {
            const int idx = base + lane;
            int bx = 0, by = 0, bz = 0, blk = -1;
            unsigned long long mask = 0ull;
            if (idx < nb) {
              const int dx = idx % w - r, dy = (idx / w) % w - r, dz = idx / (w * w) - r;
              // shell only — the interior was visited by smaller rings; ring 1 also takes the query's own block, whose
              // voxels outside the 5x5x5 stencil have not been seen yet
              if (r == 1 || max(abs(dx), max(abs(dy), abs(dz))) == r) {
                bx = qbx + dx; by = qby + dy; bz = qbz + dz;
                const float lx = (float)bx * bs4, ly = (float)by * bs4, lz = (float)bz * bs4;
                const float md = box_mind2(qx, qy, qz, lx, ly, lz, lx + bs4, ly + bs4, lz + bs4, mg);
                if (!(md > bound || md > lim)) {
                  blk = find_block(m, pack_key(bx, by, bz));
                  if (blk >= 0) mask = __ldg(&m.bmask[blk]);
                }
              }
            }
            const unsigned todo = __ballot_sync(FULL, blk >= 0 && mask != 0ull);
            if (todo) coop_scan_blocks<K>(m, todo, blk, mask, bx, by, bz, lane, qx, qy, qz, cvx, cvy, cvz, r == 1, fminf(lim, bound), t);
          }
          gcount = warp_merge<K>(t, lane, rd, rx, ry, rz, thr);
          cov = cover2(qx, qy, qz, (float)(qbx - r) * bs4, (float)(qby - r) * bs4, (float)(qbz - r) * bs4,
                       (float)(qbx + r + 1) * bs4, (float)(qby + r + 1) * bs4, (float)(qbz + r + 1) * bs4, mg);
          done = (gcount == K && thr < cov) || cov > lim;
        }
      }
      if (!done) {
        // ---------------- phase B: 3x3x3 coarse cells around the query (blocks of B0 are skipped inside)
        phase = 2;
        const int qcx = qbx >> 3, qcy = qby >> 3, qcz = qbz >> 3;
        int mycs = -1;
        if (lane < 27) mycs = find_coarse(m, pack_key(qcx + (lane % 3) - 1, qcy + ((lane / 3) % 3) - 1, qcz + (lane / 9) - 1));
        // nearest cells first (centre cell), so the bound tightens early
        const unsigned present = __ballot_sync(FULL, mycs >= 0);
        {
          const int cs = __shfl_sync(FULL, mycs, 13);
          if (cs >= 0) scan_coarse_cell<K>(m, cs, lane, qx, qy, qz, cvx, cvy, cvz, gcount == K, thr, lim, mg, t);
          if (gcount < K) gcount = warp_merge<K>(t, lane, rd, rx, ry, rz, thr);  // get a finite bound before the ring
        }
        unsigned rest = present & ~(1u << 13);
        while (rest) {
          const int c = __ffs(rest) - 1;
          rest &= rest - 1;
          const int cs = __shfl_sync(FULL, mycs, c);
          scan_coarse_cell<K>(m, cs, lane, qx, qy, qz, cvx, cvy, cvz, gcount == K, thr, lim, mg, t);
        }
        gcount = warp_merge<K>(t, lane, rd, rx, ry, rz, thr);
        const float cs32 = 32.f * ds;
        cov = cover2(qx, qy, qz, (float)(qcx - 1) * cs32, (float)(qcy - 1) * cs32, (float)(qcz - 1) * cs32,
                     (float)(qcx + 2) * cs32, (float)(qcy + 2) * cs32, (float)(qcz + 2) * cs32, mg);
        done = (gcount == K && thr < cov) || cov > lim;
        if (!done) {
          // ---------------- phase C: exhaustive scan of the coarse hash with box-distance pruning
          phase = 3;
          const int ncs = m.counters[CNT_COARSE_USED];   // dense list of occupied coarse cells
#pragma unroll 1
          for (int base = 0; base < ncs; base += 32) {
            const int li = base + lane;
            const int cs = li < ncs ? (int)__ldg(&m.clist[li]) : -1;
            const uint64_t ck = cs >= 0 ? __ldg(&m.ckeys[cs]) : KEY_EMPTY;
            bool go = false;
            if (ck != KEY_EMPTY) {
              int cx, cy, cz;
              unpack_key(ck, cx, cy, cz);
              if (!(abs(cx - qcx) <= 1 && abs(cy - qcy) <= 1 && abs(cz - qcz) <= 1)) {
                const float md = box_mind2(qx, qy, qz, (float)cx * cs32, (float)cy * cs32, (float)cz * cs32,
                                           (float)(cx + 1) * cs32, (float)(cy + 1) * cs32, (float)(cz + 1) * cs32, mg);
                go = (gcount < K || md <= thr) && md <= lim;
              }
            }
            unsigned todo = __ballot_sync(FULL, go);
            if (!todo) continue;
            while (todo) {
              const int c = __ffs(todo) - 1;
              todo &= todo - 1;
              scan_coarse_cell<K>(m, __shfl_sync(FULL, cs, c), lane, qx, qy, qz, cvx, cvy, cvz, gcount == K, thr, lim, mg, t);
            }
            gcount = warp_merge<K>(t, lane, rd, rx, ry, rz, thr);
          }
        }
      }
    }
    if (lane < K) a.nbr[(size_t)lane * a.stride + i] = make_float4(rx, ry, rz, rd);
    if (lane == 0) {
      a.cnt[i] = (unsigned char)gcount;
      if (a.phase_stats) atomicAdd(&a.phase_stats[phase], 1);
    }
  }
}

// K1a: phase A with ONE THREAD per query (the common case: >99 % of LiDAR returns lie on mapped surfaces and are
// resolved by the 5x5x5 stencil).  Per query: 8 block probes (one 16-B hash entry + one 8-B occupancy word each),
// then ONE loop over the occupied stencil voxels of all 8 blocks (per-thread cursor, so a warp iterates
// max-over-lanes of the candidate COUNT, not the sum of per-block maxima) with one 16-B point load per candidate and
// a branch-free insertion into a register-resident (distance, slot-id) top-K.  No cross-lane traffic.
// Queries whose stencil cannot prove completeness are appended to a work list for the exact kernel k_knn.
template <int K>
struct TopKId {
  float d[K];
  unsigned id[K];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int j = 0; j < K; ++j) { d[j] = CUDART_INF_F; id[j] = 0u; }
  }
  // branch-free sorted insert (equal distances keep their arrival order: exact float ties between distinct map points
  // are not canonicalised on this fast path — the reference's own tie order is traversal dependent, ikd_Tree.h:102-105)
  __device__ __forceinline__ void insert(float dd, unsigned pid) {
    bool c[K];
#pragma unroll
    for (int j = 0; j < K; ++j) c[j] = dd < d[j];
#pragma unroll
    for (int j = K - 1; j > 0; --j) {
      d[j] = c[j - 1] ? d[j - 1] : (c[j] ? dd : d[j]);
      id[j] = c[j - 1] ? id[j - 1] : (c[j] ? pid : id[j]);
    }
    d[0] = c[0] ? dd : d[0];
    id[0] = c[0] ? pid : id[0];
  }
};

constexpr int STENCIL_THREADS = 128;

// One pass over the candidate voxels recorded in s_cand (per-thread smem column).  PRUNE: skip a voxel (head load and
// overflow chain) when the squared distance from the query to the voxel's box — a lower bound for every point keyed
// to that voxel, looked up from per-axis gap tables — already exceeds the current k-th distance.
template <int K, bool PRUNE>
__device__ __forceinline__ void stencil_pass(const MapDev& m, const int (*s_blk)[STENCIL_THREADS],
                                             const unsigned long long (*s_cand)[STENCIL_THREADS],
                                             const float (*s_gap)[STENCIL_THREADS], int tid, int ox, int oy, int oz, float qx, float qy,
                                             float qz, float lim, TopKId<K>& t, int& n_head, int& n_chain) {
  int b = -1, blk = 0;
  unsigned long long cand = 0ull;
  for (;;) {
    unsigned pid[4];
    int nc = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      pid[u] = 0u;
      for (;;) {
        while (cand == 0ull && b < 7) { ++b; cand = s_cand[b][tid]; blk = s_blk[b][tid]; }
        if (cand == 0ull) break;
        const int sl = __ffsll((long long)cand) - 1;
        cand &= cand - 1;
        if (PRUNE) {
          // stencil-relative voxel index per axis (0..4): block half (bit of b) * 4 + local coordinate - stencil origin
          const int jx = ((b & 1) << 2) + (sl & 3) - ox, jy = (((b >> 1) & 1) << 2) + ((sl >> 2) & 3) - oy, jz = ((b >> 2) << 2) + (sl >> 4) - oz;
          const float md = s_gap[jx][tid] + s_gap[5 + jy][tid] + s_gap[10 + jz][tid];
          if (md > t.d[K - 1]) continue;  // no point of this voxel can enter the top-K
        }
        pid[u] = (unsigned)blk * 64u + (unsigned)sl;
        nc = u + 1;
        break;
      }
    }
    if (nc == 0) break;
    float4 e[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (u < nc) e[u] = __ldg(&m.slots[pid[u]]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (u < nc) {
        float dd = sqdist(qx, qy, qz, e[u].x, e[u].y, e[u].z);
        if (dd <= lim) t.insert(dd, pid[u]);
        ++n_head;
        int c = __float_as_int(e[u].w);
        while (c >= 0) {  // overflow chain of this voxel
          ++n_chain;
          const float4 o = __ldg(&m.ovf[c]);
          dd = sqdist(qx, qy, qz, o.x, o.y, o.z);
          if (dd <= lim) t.insert(dd, 0x80000000u | (unsigned)c);
          c = __float_as_int(o.w);
        }
      }
    }
    if (nc < 4) break;
  }
}

// Outer-shell pass: (1) a flat, load-free loop tests every occupied shell voxel against the k-th distance known after
// the inner pass (box lower bound from the per-axis gap tables) and compacts the survivors into a small per-thread list
// — lanes only diverge on cheap code; (2) the survivors are loaded four at a time.  Returns false if the list overflowed
// (the query then goes to the exact kernel).
constexpr int SHELL_LIST = 32;
template <int K>
__device__ __forceinline__ bool stencil_shell_pass(const MapDev& m, const int (*s_blk)[STENCIL_THREADS],
                                                   const unsigned long long (*s_cand)[STENCIL_THREADS],
                                                   const float (*s_gap)[STENCIL_THREADS], unsigned short (*s_list)[STENCIL_THREADS],
                                                   int tid, int ox, int oy, int oz, float qx, float qy, float qz, float lim,
                                                   TopKId<K>& t, int& n_head, int& n_chain) {
  const float bound = t.d[K - 1];
  int ns = 0;
  {
    int b = -1;
    unsigned long long cand = 0ull;
    for (;;) {
      while (cand == 0ull && b < 7) { ++b; cand = s_cand[b][tid]; }
      if (cand == 0ull) break;
      const int sl = __ffsll((long long)cand) - 1;
      cand &= cand - 1;
      const int jx = ((b & 1) << 2) + (sl & 3) - ox, jy = (((b >> 1) & 1) << 2) + ((sl >> 2) & 3) - oy, jz = ((b >> 2) << 2) + (sl >> 4) - oz;
      const float md = s_gap[jx][tid] + s_gap[5 + jy][tid] + s_gap[10 + jz][tid];
      if (!(md > bound)) {  // a point of this voxel could still enter the top-K
        if (ns < SHELL_LIST) s_list[ns][tid] = (unsigned short)((b << 6) | sl);
        ++ns;
      }
    }
  }
  if (ns > SHELL_LIST) return false;  // (sparse inner region: hardly anything could be pruned)
  for (int base = 0; base < ns; base += 4) {
    unsigned pid[4];
    float4 e[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      pid[u] = 0u;
      if (base + u < ns) {
        const unsigned v = s_list[base + u][tid];
        pid[u] = (unsigned)s_blk[v >> 6][tid] * 64u + (v & 63u);
        e[u] = __ldg(&m.slots[pid[u]]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (base + u < ns) {
        float dd = sqdist(qx, qy, qz, e[u].x, e[u].y, e[u].z);
        if (dd <= lim) t.insert(dd, pid[u]);
        ++n_head;
        int c = __float_as_int(e[u].w);
        while (c >= 0) {  // overflow chain of this voxel
          ++n_chain;
          const float4 o = __ldg(&m.ovf[c]);
          dd = sqdist(qx, qy, qz, o.x, o.y, o.z);
          if (dd <= lim) t.insert(dd, 0x80000000u | (unsigned)c);
          c = __float_as_int(o.w);
        }
      }
    }
  }
  return true;
}

template <int K>
__global__ void __launch_bounds__(STENCIL_THREADS, 7) k_knn_stencil(KnnArgs a) {
  __shared__ int s_blk[8][STENCIL_THREADS];
  __shared__ unsigned long long s_cand1[8][STENCIL_THREADS];   // occupied voxels of the inner 3x3x3
  __shared__ unsigned long long s_cand2[8][STENCIL_THREADS];   // occupied voxels of the outer shell of the 5x5x5
  __shared__ float s_gap[15][STENCIL_THREADS];                 // squared query-to-slab gaps: x[5], y[5], z[5]
  __shared__ unsigned short s_list[SHELL_LIST][STENCIL_THREADS];  // surviving shell voxels (block half << 6 | slot)
  const MapDev& m = a.m;
  const int tid = threadIdx.x;
  const int i = blockIdx.x * blockDim.x + tid;
  if (a.ctl && !(ctl_pass_active(a.ctl) && a.ctl->converge)) return;
  if (i >= (a.ctl ? a.ctl->n : a.n)) return;
  const float ds = m.ds;
  const float lim = a.max_d2;
  const float4 q4 = a.ctl ? body_to_world(a.ctl->pose, __ldg(&a.body[i])) : __ldg(&a.q[i]);
  const float qx = q4.x, qy = q4.y, qz = q4.z;
  const float qlim = 4.0e6f * ds;
  TopKId<K> t;
  t.clear();
  bool done = false;
  if (fabsf(qx) < qlim && fabsf(qy) < qlim && fabsf(qz) < qlim) {
    const int cvx = voxel_of(qx, ds), cvy = voxel_of(qy, ds), cvz = voxel_of(qz, ds);
    const int bbx = (cvx - 2) >> 2, bby = (cvy - 2) >> 2, bbz = (cvz - 2) >> 2;
    const float mg = 1e-3f * ds + 4.8e-7f * (fabsf(qx) + fabsf(qy) + fabsf(qz));
    // per axis the 5-wide stencil covers local range [o,3] of the low block and [0,o] of the high block; the inner
    // 3-wide one is the same shifted by one voxel
    const int ox = (cvx - 2) & 3, oy = (cvy - 2) & 3, oz = (cvz - 2) & 3;
    unsigned x5[2], y5[2], z5[2], x3[2], y3[2], z3[2];
    {
      const unsigned mx5 = 31u << ox, my5 = 31u << oy, mz5 = 31u << oz, mx3 = 14u << ox, my3 = 14u << oy, mz3 = 14u << oz;
      x5[0] = mx5 & 15u; x5[1] = (mx5 >> 4) & 15u; y5[0] = my5 & 15u; y5[1] = (my5 >> 4) & 15u; z5[0] = mz5 & 15u; z5[1] = (mz5 >> 4) & 15u;
      x3[0] = mx3 & 15u; x3[1] = (mx3 >> 4) & 15u; y3[0] = my3 & 15u; y3[1] = (my3 >> 4) & 15u; z3[0] = mz3 & 15u; z3[1] = (mz3 >> 4) & 15u;
    }
    auto spread_y = [](unsigned ym) { return (ym & 1u) | ((ym & 2u) << 3) | ((ym & 4u) << 6) | ((ym & 8u) << 9); };  // bits 0,4,8,12
    auto spread_z = [](unsigned zm) {
      return (unsigned long long)(zm & 1u) | ((unsigned long long)(zm & 2u) << 15) | ((unsigned long long)(zm & 4u) << 30) |
             ((unsigned long long)(zm & 8u) << 45);                                                                      // bits 0,16,32,48
    };
    // squared gaps from the query to the 5 voxel slabs per axis (conservative: shrunk by the rounding margin)
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const float lx = (float)(cvx - 2 + j) * ds, ly = (float)(cvy - 2 + j) * ds, lz = (float)(cvz - 2 + j) * ds;
      const float gx = fmaxf(fmaxf(lx - qx, qx - (lx + ds)) - mg, 0.f);
      const float gy = fmaxf(fmaxf(ly - qy, qy - (ly + ds)) - mg, 0.f);
      const float gz = fmaxf(fmaxf(lz - qz, qz - (lz + ds)) - mg, 0.f);
      s_gap[j][tid] = gx * gx;
      s_gap[5 + j][tid] = gy * gy;
      s_gap[10 + j][tid] = gz * gz;
    }
    // ---- the 8 hash probes are INDEPENDENT loads: issue them back to back (memory-level parallelism), then resolve;
    // only a collision (first slot holds another key) falls back to the sequential probe loop
    uint64_t keys8[8];
    uint4 ent[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      keys8[b] = pack_key(bbx + (b & 1), bby + ((b >> 1) & 1), bbz + (b >> 2));
      ent[b] = __ldg(reinterpret_cast<const uint4*>(&m.hent[hash_key(keys8[b]) & m.hash_mask]));
    }
    int blk8[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const uint64_t k0 = ((uint64_t)ent[b].y << 32) | ent[b].x;
      blk8[b] = (k0 == keys8[b]) ? (int)ent[b].z : (k0 == KEY_EMPTY ? -1 : find_block(m, keys8[b]));
    }
    unsigned long long occ[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) occ[b] = blk8[b] >= 0 ? __ldg(&m.bmask[blk8[b]]) : 0ull;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int hx = b & 1, hy = (b >> 1) & 1, hz = b >> 2;
      const unsigned long long sten5 = (unsigned long long)(x5[hx] * spread_y(y5[hy])) * spread_z(z5[hz]);
      const unsigned long long sten3 = (unsigned long long)(x3[hx] * spread_y(y3[hy])) * spread_z(z3[hz]);
      s_blk[b][tid] = blk8[b];
      s_cand1[b][tid] = occ[b] & sten3;
      s_cand2[b][tid] = occ[b] & sten5 & ~sten3;
    }
    // ---- inner 3x3x3 first (gives a tight k-th distance), then the outer shell with box-distance pruning
    int n_chain = 0, n_head = 0;
    stencil_pass<K, false>(m, s_blk, s_cand1, s_gap, tid, ox, oy, oz, qx, qy, qz, lim, t, n_head, n_chain);
    if (!stencil_shell_pass<K>(m, s_blk, s_cand2, s_gap, s_list, tid, ox, oy, oz, qx, qy, qz, lim, t, n_head, n_chain))
      stencil_pass<K, false>(m, s_blk, s_cand2, s_gap, tid, ox, oy, oz, qx, qy, qz, lim, t, n_head, n_chain);  // list overflow: visit all
    if (a.phase_stats) {  // profiling only: candidate statistics
      atomicAdd(&a.phase_stats[4], n_chain);
      atomicMax(&a.phase_stats[5], n_chain);
      atomicAdd(&a.phase_stats[6], n_head);
    }
    const float cov = cover2(qx, qy, qz, (float)(cvx - 2) * ds, (float)(cvy - 2) * ds, (float)(cvz - 2) * ds,
                             (float)(cvx + 3) * ds, (float)(cvy + 3) * ds, (float)(cvz + 3) * ds, mg);
    done = (t.d[K - 1] < CUDART_INF_F && t.d[K - 1] < cov) || cov > lim;
  } else {
    done = true;  // unrepresentable / NaN query: no neighbours
  }
  if (done) {
    int c = 0;
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const bool ok = t.d[r] < CUDART_INF_F;
      c += ok ? 1 : 0;
      float4 o = make_float4(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F, CUDART_INF_F);
      if (ok) {
        const float4 e = (t.id[r] & 0x80000000u) ? __ldg(&m.ovf[t.id[r] & 0x7FFFFFFFu]) : __ldg(&m.slots[t.id[r]]);
        o = make_float4(e.x, e.y, e.z, t.d[r]);
      }
      a.nbr[(size_t)r * a.stride + i] = o;
    }
    a.cnt[i] = (unsigned char)c;
    if (a.phase_stats) atomicAdd(&a.phase_stats[0], 1);
  } else {
    a.worklist[atomicAdd(a.work_count, 1)] = i;
  }
}

__global__ void k_transform(PoseDev s, const float4* __restrict__ body, float4* __restrict__ world, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) world[i] = body_to_world(s, body[i]);
}

}  // namespace flb
