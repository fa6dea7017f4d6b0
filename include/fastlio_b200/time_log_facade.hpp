// time_log_facade.hpp — the reference's per-scan timing log wired to the counters of the B200 path
// (second half).  Keeps the row layout, the running averages and the CSV format of
// src/laserMapping.cpp:2429-2449 (per scan) and :2559-2574 ("fast_lio_time_log.csv"), so existing plotting scripts
// keep working:
//
//   flb::TimeLog tlog;                                            // next to the T1[] / s_plot*[] globals (:23)
//   ... per scan, after flb_scan_step(...) filled `flb_scan_result r`:
//   tlog.add(Measures.lidar_beg_time, t5 - t0, feats_undistort->points.size(), kdtree_size_st, r, preprocess_time);
//   ... at shutdown (:2559):  tlog.save(log_path + "fast_lio_time_log.csv");
//
// Column mapping: "incremental time" = device time of map_incremental's kernels + the fov delete (gpu_ms_total −
// update.gpu_ms; a device-driven step stamps both on the device: the update ends behind its last update kernel, the whole
// step behind the last insert kernel), "search time" = 0 exactly as in the reference (kdtree_search_time is reset every scan at :2249 and
// never accumulated), "delete size" = kdtree_delete_counter, "delete time" folded into the incremental column (the
// delete runs inside the same stream segment), tree sizes = ikdtree.size() before / after, "add point size" =
// add_point_size (:1494) = n_to_add + n_no_downsample.
#pragma once
#include <cstdio>
#include <string>
#include <vector>

#include "../fastlio_b200.h"

namespace flb {

class TimeLog {
 public:
  struct Row {
    double time_stamp, total_time;
    int scan_points;
    double incremental_time, search_time;
    int delete_size;
    double delete_time;
    int tree_size_st, tree_size_end, add_point_size;
    double preprocess_time;
  };

  // total_s: wall time of the whole scan on the host (t5 - t0, :2439); device times arrive in ms and are logged in seconds
  void add(double lidar_beg_time, double total_s, int scan_points, int tree_size_st, const flb_scan_result& r,
           double preprocess_s = 0.0) {
    Row w;
    w.time_stamp = lidar_beg_time;
    w.total_time = total_s;
    w.scan_points = scan_points;
    w.incremental_time = 1e-3 * (double)(r.gpu_ms_total - r.update.gpu_ms);
    w.search_time = 0.0;
    w.delete_size = r.n_deleted;
    w.delete_time = 0.0;
    w.tree_size_st = tree_size_st;
    w.tree_size_end = r.map_valid;
    w.add_point_size = r.n_to_add + r.n_no_downsample;
    w.preprocess_time = preprocess_s;
    rows_.push_back(w);
    // running averages of :2432-2437
    const double n = (double)rows_.size();
    aver_time_consu = aver_time_consu * (n - 1) / n + total_s / n;
    aver_time_icp = aver_time_icp * (n - 1) / n + 1e-3 * (double)r.update.gpu_ms / n;
    aver_time_incre = aver_time_incre * (n - 1) / n + w.incremental_time / n;
  }

  // fast_lio_time_log.csv, header and row format of :2564-2567
  bool save(const std::string& path) const {
    FILE* fp = std::fopen(path.c_str(), "w");
    if (!fp) return false;
    std::fprintf(fp, "time_stamp, total time, scan point size, incremental time, search time, delete size, delete time, tree size st, tree size end, add point size, preprocess time\n");
    for (const Row& w : rows_)
      std::fprintf(fp, "%0.8f,%0.8f,%d,%0.8f,%0.8f,%d,%0.8f,%d,%d,%d,%0.8f\n", w.time_stamp, w.total_time, w.scan_points,
                   w.incremental_time, w.search_time, w.delete_size, w.delete_time, w.tree_size_st, w.tree_size_end,
                   w.add_point_size, w.preprocess_time);
    std::fclose(fp);
    return true;
  }

  const std::vector<Row>& rows() const { return rows_; }
  double aver_time_consu = 0.0, aver_time_icp = 0.0, aver_time_incre = 0.0;

 private:
  std::vector<Row> rows_;
};

}  // namespace flb
