#!/usr/bin/env python
"""Full-size closed-loop parity diagnosis (BASELINE cfg2): the GPU path and the CPU oracle (reference ikd-Tree compiled
unmodified + restated h_share_model / ESIKF) replay the SAME scans from the SAME initial map; per frame the posterior
pose difference, the pass statistics, the 5-NN difference of the last search pass and the map point-set difference are
written to a JSON file.  Test infrastructure (uses oracle/): run on the GPU box,
    python tools/fullsize_parity.py --frames 25 --out gpurun_out/parity.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def rowset(a):
    a = np.ascontiguousarray(a, np.float32)
    return a.view([("x", "f4"), ("y", "f4"), ("z", "f4")]).reshape(-1)


def set_diff(a, b):
    va, vb = rowset(a), rowset(b)
    only_a = np.setdiff1d(va, vb)
    only_b = np.setdiff1d(vb, va)
    return only_a, only_b


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=25)
    ap.add_argument("--scans", type=int, default=60, help="n_scans of make_workload")
    ap.add_argument("--seed", type=int, default=20)
    ap.add_argument("--flatten-every", type=int, default=5)
    ap.add_argument("--chain", action="store_true", help="chain the filter (prior = previous posterior + true motion)")
    ap.add_argument("--out", default="gpurun_out/parity.json")
    args = ap.parse_args()
    bench.protect_stdout()
    from better_fastlio2_b200 import capi, synth
    from oracle import pyoracle as po
    po.build()
    work = bench.make_workload(args.seed, args.scans)
    ncores = os.cpu_count() or 1
    tree = capi.KDTree(voxel_size=bench.DS, max_points=16 << 20, max_blocks=2 << 20)
    bench.build_map(tree, work["map"])
    ref = po.make_map(ds=bench.DS, threads=ncores)
    t0 = time.time()
    bench.build_map(ref, work["map"])
    bench.log(f"ref build {time.time() - t0:.1f}s valid {ref.validnum()} gpu {tree.validnum()}")
    out = {"build": {}, "frames": []}
    og, oc = set_diff(tree.flatten(), ref.flatten())
    out["build"] = {"gpu_valid": tree.validnum(), "ref_valid": ref.validnum(), "only_gpu": len(og), "only_ref": len(oc),
                    "only_gpu_pts": [list(map(float, p)) for p in og[:20].tolist()],
                    "only_ref_pts": [list(map(float, p)) for p in oc[:20].tolist()]}
    bench.log("build diff", out["build"]["only_gpu"], out["build"]["only_ref"])
    nmax = max(len(s) for s in work["scans"])
    ses = capi.Session(tree, max_scan_points=max(131072, nmax), max_iterations=bench.MAX_ITER, filter_size_map_min=bench.DS)
    fov_g = capi.make_fov(cube_len=1000.0, det_range=100.0)
    fov_c = po.FovSegment(cube_len=1000.0, det_range=100.0)
    pos_lid_c = np.zeros(3)
    P0 = work["P"]
    for k in range(args.frames):
        body = work["scans"][k]
        # the same scan through the HOST-driven engine first (update only: the map is not touched), as a third opinion
        ses.scan_upload(body)
        ses.set_update_engine(False)
        s_h, P_h, st_h = ses.update_iterated_dyn_share_modified(work["priors"][k], P0)
        ses.set_update_engine(True)
        s_g, P_g, r = ses.scan_step(fov_g, body, work["priors"][k], P0, True)
        nb = ses.neighbors()
        boxes = fov_c.step(pos_lid_c)
        if len(boxes):
            ref.Delete_Point_Boxes(boxes)
        s_c, P_c, sc, st, _ = po.esikf_update(work["priors"][k], P0, body, ref, max_iter=bench.MAX_ITER)
        pos_lid_c = s_c[0:3] + synth.quat_to_mat(s_c[3:7]) @ s_c[11:14]
        na, nn = po.map_incremental(s_c, body, sc, ref, True, bench.DS)
        n = len(body)
        d2g, d2c = nb["d2"][:n], sc.nbr_d2[:n]
        fin = np.isfinite(d2g) & np.isfinite(d2c)
        nbr_diff = int(((d2g != d2c) & ~(~np.isfinite(d2g) & ~np.isfinite(d2c))).any(axis=1).sum())
        world_diff = int((nb["world"][:n] != sc.world[:n]).any(axis=1).sum())
        sel_diff = int((nb["sel"][:n] != sc.sel[:n]).sum())
        cnt_diff = int((nb["cnt"][:n] != sc.nbr_cnt[:n]).sum())
        rec = {"k": k, "dpos": float(np.abs(s_g[:3] - s_c[:3]).max()), "dpos_l2": float(np.linalg.norm(s_g[:3] - s_c[:3])),
               "dq": float(np.abs(s_g[3:7] - s_c[3:7]).max()), "dstate": float(np.abs(s_g - s_c).max()),
               "dP": float(np.abs(P_g - P_c).max()),
               "dpos_host_engine_vs_cpu": float(np.abs(s_h[:3] - s_c[:3]).max()),
               "dpos_host_engine_vs_device_engine": float(np.abs(s_h[:3] - s_g[:3]).max()),
               "host_engine": {k2: st_h[k2] for k2 in ("passes", "search_passes", "effct_feat_num", "converged_count")},
               "gpu": {"passes": r.update.passes, "searches": r.update.search_passes, "M": r.update.effct_feat_num,
                       "t": r.update.converged_count,
                       "add": r.n_to_add, "no_ds": r.n_no_downsample, "valid": r.map_valid, "deleted": r.n_deleted},
               "cpu": {"stats": [int(x) for x in st], "add": na, "no_ds": nn, "valid": ref.validnum()},
               "nbr_rows_diff": nbr_diff, "world_rows_diff": world_diff, "sel_diff": sel_diff, "cnt_diff": cnt_diff,
               "err_truth_gpu": float(np.linalg.norm(s_g[:3] - work["truths"][k][:3])),
               "err_truth_cpu": float(np.linalg.norm(s_c[:3] - work["truths"][k][:3]))}
        if nbr_diff and world_diff == 0:
            bad = np.nonzero((d2g != d2c).any(axis=1))[0][:5]
            rec["nbr_examples"] = [{"i": int(i), "q": nb["world"][i].tolist(), "d2_gpu": d2g[i].tolist(), "d2_cpu": d2c[i].tolist(),
                                    "nbr_gpu": nb["nbr"][i].tolist(), "nbr_cpu": sc.nbr[i].tolist()} for i in bad]
        if (k + 1) % args.flatten_every == 0 or k == args.frames - 1:
            og, oc = set_diff(tree.flatten(), ref.flatten())
            rec["map_only_gpu"], rec["map_only_ref"] = len(og), len(oc)
            rec["map_only_gpu_pts"] = [list(map(float, p)) for p in og[:10].tolist()]
            rec["map_only_ref_pts"] = [list(map(float, p)) for p in oc[:10].tolist()]
        out["frames"].append(rec)
        bench.log(json.dumps({a: rec[a] for a in ("k", "dpos", "dq", "dpos_host_engine_vs_cpu", "dpos_host_engine_vs_device_engine", "nbr_rows_diff", "sel_diff")}),
                  rec["gpu"]["t"], rec["host_engine"]["converged_count"], rec["cpu"]["stats"][3],
                  rec.get("map_only_gpu"), rec.get("map_only_ref"), rec["gpu"]["valid"], rec["cpu"]["valid"])
    out["max_dpos"] = max(f["dpos"] for f in out["frames"])
    out["max_dq"] = max(f["dq"] for f in out["frames"])
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)
    bench.emit({"max_dpos": out["max_dpos"], "max_dq": out["max_dq"], "frames": args.frames})
    ses.close()
    tree.close()


if __name__ == "__main__":
    main()
