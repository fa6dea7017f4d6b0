"""CPU: the timing-log facade (include/fastlio_b200/time_log_facade.hpp) keeps the reference's CSV format
(src/laserMapping.cpp:2564-2567) when fed from flb_scan_result."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_time_log_csv_format(tmp_path):
    exe = str(tmp_path / "time_log_smoke")
    csv = str(tmp_path / "fast_lio_time_log.csv")
    subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "time_log_smoke.cpp"), "-o", exe], check=True, capture_output=True, text=True)
    out = subprocess.run([exe, csv], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "TIME_LOG_OK" in out.stdout, (out.returncode, out.stdout, out.stderr)
    lines = open(csv).read().splitlines()
    assert lines[0] == ("time_stamp, total time, scan point size, incremental time, search time, delete size, delete time, "
                        "tree size st, tree size end, add point size, preprocess time")
    assert len(lines) == 4
    f = lines[2].split(",")
    assert len(f) == 11
    assert f[0] == "100.10000000" and f[1] == "0.00200000" and f[2] == "120001"
    assert abs(float(f[3]) - 1e-3 * (0.32 - 0.26)) < 1e-7 and float(f[4]) == 0.0      # incremental = total - update; search = 0
    assert f[5] == "10" and f[7] == "995" and f[8] == "1005" and f[9] == "10" and f[10] == "0.00050000"
