#!/usr/bin/env python
"""A/B of the stencil 5-NN kernel against the bulk-copy (cp.async.bulk + mbarrier) staged variant (csrc/knn_tile.cuh) on the
cfg2 map: kernel-only time of one search pass over a full scan, CUDA events around `iters` back-to-back launches, for

  * the scan in its native order (ring-major, azimuth inside a ring: neighbouring threads are neighbouring returns), and
  * the same queries sorted by the Morton code of their 0.8 m map block (what a per-scan radix sort would give),

each at a prior pose (first search pass of a scan: queries ~5 cm / 0.5 deg off the surfaces) and at the true pose (second
search pass).  Also checks that both kernels return the same distances / counts / unresolved sets.

  python tools/knn_tile_ab.py [--iters 20] > profiles/r2_tma_ab.txt
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def morton_order(q, cell):
    k = np.floor(q / cell).astype(np.int64) + (1 << 20)
    def spread(v):
        v = v & 0x1FFFFF
        v = (v | (v << 32)) & 0x1F00000000FFFF
        v = (v | (v << 16)) & 0x1F0000FF0000FF
        v = (v | (v << 8)) & 0x100F00F00F00F00F
        v = (v | (v << 4)) & 0x10C30C30C30C30C3
        v = (v | (v << 2)) & 0x1249249249249249
        return v
    code = spread(k[:, 0]) | (spread(k[:, 1]) << 1) | (spread(k[:, 2]) << 2)
    return np.argsort(code, kind="stable")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    bench.protect_stdout()
    from better_fastlio2_b200 import capi, synth
    L = capi.lib()
    f = L.flb_debug_knn_bench
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
    work = bench.make_workload(bench.SEED, 8)
    tree = capi.KDTree(voxel_size=bench.DS, max_points=16 << 20, max_blocks=2 << 20)
    bench.build_map(tree, work["map"])
    out = [f"# stencil 5-NN kernel A/B on the cfg2 map ({tree.validnum()} points), {args.iters} launches per measurement, 1xB200",
           "# variant 0 = k_knn_stencil<5> (product), 1 = k_knn_tile<5> (cp.async.bulk staged buckets, csrc/knn_tile.cuh)",
           f"{'pose':8s} {'order':8s} {'variant':8s} {'ms/launch':>10s} {'unresolved':>11s} {'same result':>12s}"]
    for pose_name, states in (("prior", work["priors"]), ("truth", work["truths"])):
        q = synth.body_to_world_np(states[3], work["scans"][3]).astype(np.float32)
        for order_name, idx in (("scan", np.arange(len(q))), ("morton", morton_order(q, 4 * bench.DS))):
            qq = np.ascontiguousarray(q[idx])
            res = {}
            for variant in (0, 1):
                ms, unres = C.c_float(0), C.c_int(0)
                d2 = np.empty((len(qq), 5), np.float32)
                cnt = np.empty(len(qq), np.int32)
                rc = f(tree.h, qq.ctypes.data, len(qq), 12, variant, args.iters, C.byref(ms), C.byref(unres), d2.ctypes.data, cnt.ctypes.data)
                if rc:
                    raise SystemExit(L.flb_last_error().decode())
                res[variant] = (ms.value, unres.value, d2, cnt)
            same = bool(np.array_equal(res[0][2], res[1][2]) and np.array_equal(res[0][3], res[1][3]) and res[0][1] == res[1][1])
            for variant in (0, 1):
                out.append(f"{pose_name:8s} {order_name:8s} {variant:8d} {res[variant][0]:10.4f} {res[variant][1]:11d} {str(same):>12s}")
    bench.emit_text("\n".join(out)) if hasattr(bench, "emit_text") else os.write(bench._REAL_STDOUT or 1, ("\n".join(out) + "\n").encode())
    tree.close()


if __name__ == "__main__":
    main()
