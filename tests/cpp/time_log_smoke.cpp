// Host-only check of include/fastlio_b200/time_log_facade.hpp: rows, running averages and the CSV text of
// src/laserMapping.cpp:2564-2567.  Built and run by tests/test_time_log.py.
#include <cmath>
#include <cstdio>

#include <fastlio_b200/time_log_facade.hpp>

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  flb::TimeLog t;
  for (int k = 0; k < 3; ++k) {
    flb_scan_result r{};
    r.update.gpu_ms = 0.25f + 0.01f * k;
    r.gpu_ms_total = 0.30f + 0.02f * k;
    r.n_deleted = 10 * k;
    r.map_valid = 1000 + 5 * k;
    r.n_to_add = 7 + k;
    r.n_no_downsample = 2;
    t.add(100.0 + 0.1 * k, 0.001 * (k + 1), 120000 + k, 990 + 5 * k, r, 0.0005);
  }
  if (t.rows().size() != 3) return 3;
  if (std::fabs(t.aver_time_consu - 0.002) > 1e-12) return 4;          // mean of 1, 2, 3 ms
  if (std::fabs(t.aver_time_icp - 0.00026) > 1e-9) return 5;            // mean of 0.25, 0.26, 0.27 ms
  if (!t.save(argv[1])) return 6;
  std::printf("TIME_LOG_OK\n");
  return 0;
}
