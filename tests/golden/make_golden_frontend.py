"""Regenerates tests/golden/frontend/frontend_mini.npz: small seeded inputs of the front-end rows and
what the CPU oracle (oracle/frontend_oracle.cpp) produces for them — UndistortPcl backward pass, pcl::VoxelGrid in both
in-leaf orders, the key-frame transform, pointBodyToWorld.  The reference has no vectors for these functions and needs
Eigen/PCL (absent here), so the fixture freezes the restatement ("parity unpinned"); it guards the oracle
against drift and lets the GPU box check the CUDA path against committed numbers.

    python tests/golden/make_golden_frontend.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from better_fastlio2_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def main():
    rng = np.random.default_rng(41)
    world = synth.city_world(half_extent=80, seed=41)
    st = synth.trajectory_state(3)
    body = synth.scan_from_pose(world, st, synth.lidar_dirs("vlp16", rng), rng, max_range=60.0)[::7]
    xyz, inten, cur = synth.raw_scan_with_times(body, rng)
    poses, end = synth.imu_pose_sequence(st, rng, n_imu=11)
    und, perm = po.undistort(xyz, cur, poses, end)
    p4 = np.column_stack([und, inten[perm]]).astype(np.float32)
    ds_pcl, dc_pcl, _ = po.voxel_grid(p4, 0.5, curvature=cur[perm], order="pcl")
    ds_stb, dc_stb, _ = po.voxel_grid(p4, 0.5, curvature=cur[perm], order="stable")
    pose6 = np.array([4.5, -1.25, 0.5, 0.02, -0.01, 0.7], np.float32)
    tr = po.transform_cloud_rpy(p4, pose6)
    w = po.body_to_world4(end, ds_stb)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "frontend", "frontend_mini.npz")
    np.savez_compressed(out, xyz=xyz, intensity=inten, curvature=cur, poses=poses, end=end, undistorted=und, perm=perm,
                        leaf=np.array(0.5, np.float32), down_pcl=ds_pcl, down_pcl_curv=dc_pcl, down_stable=ds_stb,
                        down_stable_curv=dc_stb, pose6=pose6, transformed=tr, world=w)
    print("frontend_mini: raw", xyz.shape, "down", ds_stb.shape, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
