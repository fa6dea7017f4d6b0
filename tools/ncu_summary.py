#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into small text summaries under profiles/ (tracked).

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r1_launches.txt
  python tools/ncu_summary.py full gpurun_out/prof_knn.ncu-rep profiles/r1_knn_full.txt
"""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u in ("nsecond", "ns") else (v * 1e3 if u in ("msecond", "ms") else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; source {src}\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':58s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:58]:58s} {v[0]:8d} {v[1]:12.1f} {v[1] / v[0]:10.2f} {100 * v[1] / tot:6.1f}%\n")
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units = r[0], r[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ; source {src}\n")
        for row in r[2:]:
            f.write(f"--- {row[hdr.index('Kernel Name')]}\n")
            for w in WANT:
                if w in hdr:
                    f.write(f"    {w:70s} {row[hdr.index(w)]} {units[hdr.index(w)]}\n")
    print(open(dst).read())


def traffic(src, dst):
    """dram bytes (read + write) of ONE 5-NN search pass = the first captured k_knn_stencil + the k_knn that follows it;
    stamped with the md5 of csrc/knn_kernels.cuh so that bench.py drops the number once the kernels change."""
    import hashlib
    import json
    import os
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units = r[0], r[1]
    ik, ir, iw, it = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")

    def tobytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    rows = r[2:]
    best = None
    for i, row in enumerate(rows):   # the heaviest stencil launch = a search pass; add the exact kernel right after it
        if "k_knn_stencil" in row[ik]:
            t = float(row[it].replace(",", ""))
            if best is None or t > best[0]:
                best = (t, i)
    if best is None:
        raise SystemExit("no k_knn_stencil launch in " + src)
    i = best[1]
    tot = tobytes(rows[i][ir], units[ir]) + tobytes(rows[i][iw], units[iw])
    names = [rows[i][ik]]
    if i + 1 < len(rows) and "k_knn<" in rows[i + 1][ik]:
        tot += tobytes(rows[i + 1][ir], units[ir]) + tobytes(rows[i + 1][iw], units[iw])
        names.append(rows[i + 1][ik])
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    md5 = hashlib.md5(open(os.path.join(root, "better_fastlio2_b200", "csrc", "knn_kernels.cuh"), "rb").read()).hexdigest()
    json.dump({"dram_bytes_per_launch": tot, "kernels": names, "source": os.path.basename(src), "knn_kernels_md5": md5,
               "what": "dram__bytes_read.sum + dram__bytes_write.sum of one 5-NN search pass (k_knn_stencil + k_knn), ncu --set full, cold cache"},
              open(dst, "w"), indent=1)
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
