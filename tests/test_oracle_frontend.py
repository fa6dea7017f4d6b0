"""CPU tests: the front-end oracle (oracle/frontend_oracle.cpp — UndistortPcl backward pass, pcl::VoxelGrid, the key-frame
transform, pointBodyToWorld) cross-validated against independent numpy/scipy implementations.  These rows are "parity
unpinned" (the reference ships no vectors and needs Eigen/PCL, absent here); this is the strongest pin available."""
import os

import numpy as np
import pytest
from scipy.spatial.transform import Rotation as Rot

from better_fastlio2_b200 import synth
from tests.helpers import small_scene


@pytest.fixture(scope="module")
def raw():
    sc = small_scene(seed=11, map_half=20.0, half_extent=80.0)
    rng = np.random.default_rng(5)
    xyz, inten, cur = synth.raw_scan_with_times(sc["body"], rng)
    poses, end = synth.imu_pose_sequence(sc["st_true"], rng)
    return dict(xyz=xyz, inten=inten, cur=cur, poses=poses, end=end, scene=sc)


def _np_undistort(xyz, cur, poses, end):
    """Vectorised restatement with scipy rotations (float64), no quirks: segment = last head earlier than the point."""
    t = cur.astype(np.float64) / 1000.0
    off = poses[:, 0]
    h = np.searchsorted(off[:-1], t, side="left") - 1   # last h in [0, np-2] with off[h] < t (offsets ascending)
    out = xyz.astype(np.float64).copy()
    ok = h >= 0
    hh = np.clip(h, 0, len(poses) - 2)
    dt = t - off[hh]
    Rh = poses[hh, 13:22].reshape(-1, 3, 3)
    gyr = poses[hh + 1, 4:7]
    acc = poses[hh + 1, 1:4]
    Ri = Rh @ Rot.from_rotvec(gyr * dt[:, None]).as_matrix()
    T = poses[hh, 10:13] + poses[hh, 7:10] * dt[:, None] + 0.5 * acc * dt[:, None] ** 2 - end[0:3]
    Roff = Rot.from_quat(end[7:11]).as_matrix()
    Rend = Rot.from_quat(end[3:7]).as_matrix()
    a = xyz.astype(np.float64) @ Roff.T + end[11:14]
    b = np.einsum("nij,nj->ni", Ri, a) + T
    c = b @ Rend - end[11:14]          # Rend^T * b
    d = c @ Roff                       # Roff^T * c
    out[ok] = d[ok]
    return out, h


def test_undistort_matches_numpy(raw, oracle):
    out, perm = oracle.undistort(raw["xyz"], raw["cur"], raw["poses"], raw["end"])
    assert sorted(perm.tolist()) == list(range(len(perm)))
    cs = raw["cur"][perm]
    assert (np.diff(cs) >= 0).all()                       # time order (IMU_Processing.hpp:243)
    ref, h = _np_undistort(raw["xyz"], raw["cur"], raw["poses"], raw["end"])
    ref = ref[perm]
    err = np.abs(out.astype(np.float64) - ref).max(1)
    # everything but the first sorted point (which the reference's sweep may compensate repeatedly) agrees to float rounding
    assert err[1:].max() < 2e-5, err[1:].max()
    untouched = raw["cur"][perm] <= 0
    assert untouched.any()
    assert np.array_equal(out[untouched], raw["xyz"][perm][untouched])   # t <= IMUpose[0].offset_time: not compensated


def test_undistort_first_point_quirk(oracle):
    """The first time-sorted point is compensated again by every earlier segment whose head time it exceeds
    (IMU_Processing.hpp:382-383 leaves the iterator on begin())."""
    rng = np.random.default_rng(3)
    st = synth.trajectory_state(0)
    poses, end = synth.imu_pose_sequence(st, rng, n_imu=6)
    xyz = rng.uniform(-20, 20, (50, 3)).astype(np.float32)
    cur = rng.uniform(45.0, 99.0, 50).astype(np.float32)      # every point lies in a late segment
    out, perm = oracle.undistort(xyz, cur, poses, end)
    ref, h = _np_undistort(xyz, cur, poses, end)
    assert h.min() >= 1
    ref = ref[perm]
    assert np.abs(out[1:] - ref[1:]).max() < 2e-5
    # replay the quirk for point 0 with the single-segment formula applied repeatedly
    p = xyz[perm[0]].copy()
    t = np.float32(cur[perm[0]])
    for g in range(h[perm[0]], -1, -1):
        if float(t) / 1000.0 > poses[g, 0]:
            sub = np.vstack([poses[g], poses[g + 1]])
            q, _ = _np_undistort(p[None, :], np.array([t], np.float32) - np.float32(0), np.vstack([sub]), end)
            # _np_undistort measures dt from sub[0]'s offset: same segment arithmetic
            p = q[0].astype(np.float32)
    assert np.abs(out[0] - p).max() < 5e-5
    assert np.abs(out[0] - ref[0]).max() > 1e-3       # and it differs from the single compensation


def _np_voxel_grid(p4, leaf):
    inv = np.float32(1.0) / np.float32(leaf)
    xyz = p4[:, :3].astype(np.float32)
    mn = np.floor(xyz.min(0) * inv).astype(np.int64)
    mx = np.floor(xyz.max(0) * inv).astype(np.int64)
    div = mx - mn + 1
    ijk = (np.floor(xyz * inv) - mn.astype(np.float32)).astype(np.int64)
    key = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    uk, inv_idx, cnt = np.unique(key, return_inverse=True, return_counts=True)
    s = np.zeros((len(uk), p4.shape[1]), np.float64)
    np.add.at(s, inv_idx, p4.astype(np.float64))
    return (s / cnt[:, None]), cnt


@pytest.mark.parametrize("leaf", [0.5, 0.2, 1.0])
def test_voxel_grid_matches_numpy(raw, oracle, leaf):
    p4 = np.column_stack([raw["xyz"], raw["inten"]]).astype(np.float32)
    out, oc, ovf = oracle.voxel_grid(p4, leaf, curvature=raw["cur"])
    assert not ovf
    ref, cnt = _np_voxel_grid(np.column_stack([p4, raw["cur"]]), leaf)
    assert len(out) == len(ref)                                  # same leaves, output ordered by leaf index
    assert np.abs(out[:, :3] - ref[:, :3]).max() < 1e-4
    assert np.abs(out[:, 3] - ref[:, 3]).max() < 1e-2            # intensities up to 255, float sums
    assert np.abs(oc - ref[:, 4]).max() < 1e-2
    out_s, _, _ = oracle.voxel_grid(p4, leaf, order="stable")
    assert len(out_s) == len(out)
    assert np.abs(out_s - out).max() < 1e-3                      # only the summation order inside a leaf differs
    assert cnt.max() > 1 and len(out) < len(p4)


def test_voxel_grid_edge_cases(oracle):
    empty, _, ovf = oracle.voxel_grid(np.zeros((0, 4), np.float32), 0.5)
    assert len(empty) == 0 and not ovf
    one = np.array([[1.0, -2.0, 3.0, 7.0]], np.float32)
    o, _, _ = oracle.voxel_grid(one, 0.5)
    assert np.array_equal(o, one)
    # PCL's guard: "Leaf size is too small for the input dataset" -> output = input
    far = np.array([[0, 0, 0, 1], [500, 500, 500, 2], [-100, 3, 9, 3]], np.float32)
    o, _, ovf = oracle.voxel_grid(far, 0.001)
    assert ovf and np.array_equal(o, far)
    # points exactly on leaf faces and negative coordinates
    g = np.array([[-0.5, 0.0, 0.5, 0], [-0.5000001, 0.0, 0.5, 0], [0.4999999, 0.0, 0.5, 0], [0.0, 0.0, 0.999, 0]], np.float32)
    o, _, _ = oracle.voxel_grid(g, 0.5)
    ref, _ = _np_voxel_grid(g, 0.5)
    assert len(o) == len(ref) and np.abs(o - ref).max() < 1e-6


def test_transform_cloud_rpy_matches_scipy(raw, oracle):
    p4 = np.column_stack([raw["xyz"], raw["inten"]]).astype(np.float32)[:5000]
    pose6 = np.array([12.5, -3.25, 1.5, 0.03, -0.02, 1.1], np.float32)    # x,y,z,roll,pitch,yaw
    out = oracle.transform_cloud_rpy(p4, pose6)
    R = Rot.from_euler("ZYX", [pose6[5], pose6[4], pose6[3]]).as_matrix()   # Rz(yaw) Ry(pitch) Rx(roll)
    assert np.abs(oracle.rpy_matrix(pose6)[:, :3] - R).max() < 1e-6
    ref = p4[:, :3].astype(np.float64) @ R.T + pose6[:3].astype(np.float64)
    assert np.abs(out[:, :3] - ref).max() < 2e-4
    assert np.array_equal(out[:, 3], p4[:, 3])


def test_body_to_world4_matches_scipy(raw, oracle):
    st = raw["scene"]["st_true"]
    p4 = np.column_stack([raw["xyz"], raw["inten"]]).astype(np.float32)[:5000]
    out = oracle.body_to_world4(st, p4)
    ref = synth.body_to_world_np(st, p4[:, :3])
    assert np.abs(out[:, :3] - ref).max() < 1e-4
    assert np.array_equal(out[:, 3], p4[:, 3])


GOLD_FE = os.path.join(os.path.dirname(__file__), "golden", "frontend", "frontend_mini.npz")


def test_oracle_reproduces_frontend_golden(oracle):
    """The committed fixture (tests/golden/make_golden_frontend.py) pins the front-end oracle against drift."""
    g = np.load(GOLD_FE)
    und, perm = oracle.undistort(g["xyz"], g["curvature"], g["poses"], g["end"])
    assert np.array_equal(perm, g["perm"]) and np.array_equal(und, g["undistorted"])
    p4 = np.column_stack([und, g["intensity"][perm]]).astype(np.float32)
    for order, key in (("pcl", "down_pcl"), ("stable", "down_stable")):
        d, dc, ovf = oracle.voxel_grid(p4, float(g["leaf"]), curvature=g["curvature"][perm], order=order)
        assert not ovf and np.array_equal(d, g[key]) and np.array_equal(dc, g[key + "_curv"])
    assert np.array_equal(oracle.transform_cloud_rpy(p4, g["pose6"]), g["transformed"])
    assert np.array_equal(oracle.body_to_world4(g["end"], g["down_stable"]), g["world"])


# ---------------------------------------------------------------------------------------------- property-based checks
from hypothesis import given, settings, strategies as st  # noqa: E402
from hypothesis.extra import numpy as hnp  # noqa: E402

_clouds = hnp.arrays(np.float32, st.tuples(st.integers(1, 200), st.just(4)),
                     elements=st.floats(-50, 50, width=32, allow_nan=False, allow_infinity=False))


@settings(max_examples=40, deadline=None)
@given(p4=_clouds, leaf=st.sampled_from([0.1, 0.25, 0.5, 2.0]), seed=st.integers(0, 1000))
def test_voxel_grid_invariants(p4, leaf, seed):
    from oracle import pyoracle as po
    out, _, ovf = po.voxel_grid(p4, leaf, order="stable")
    assert not ovf and 1 <= len(out) <= len(p4)
    # every centroid lies in the bounding box of the cloud; the leaves are a partition: a permutation of the input gives
    # the same number of leaves and (up to float summation order) the same centroids in the same leaf order
    tol = 2e-3   # float32 accumulation of up to 200 coordinates of magnitude 50 (PCL sums in float too)
    assert (out[:, :3] >= p4[:, :3].min(0) - tol).all() and (out[:, :3] <= p4[:, :3].max(0) + tol).all()
    perm = np.random.default_rng(seed).permutation(len(p4))
    out2, _, _ = po.voxel_grid(p4[perm], leaf, order="stable")
    assert len(out2) == len(out)
    assert np.abs(out2 - out).max() <= tol
    # PCL's own in-leaf order (std::sort) only changes rounding
    out3, _, _ = po.voxel_grid(p4, leaf, order="pcl")
    assert len(out3) == len(out) and np.abs(out3 - out).max() <= tol
    # total mass: leaf counts are implied by sum(points) = sum(count * centroid); check with counts from numpy
    ref, cnt = _np_voxel_grid(p4, leaf)
    assert len(ref) == len(out) and int(cnt.sum()) == len(p4)


@settings(max_examples=25, deadline=None)
@given(n=st.integers(1, 300), seed=st.integers(0, 10000))
def test_undistort_at_rest_is_identity_and_sorted(n, seed):
    """A sensor at rest (zero velocity / rates, poses equal to the end state) must leave every point where it is; the
    output is in time order whatever the input order."""
    from oracle import pyoracle as po
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-30, 30, (n, 3)).astype(np.float32)
    cur = rng.uniform(0, 100, n).astype(np.float32)
    st0 = synth.make_state(pos=(1.5, -2.0, 0.7))
    R = synth.quat_to_mat(st0[3:7]).reshape(-1)
    poses = np.array([np.concatenate([[0.02 * k], np.zeros(3), np.zeros(3), np.zeros(3), st0[0:3], R]) for k in range(6)])
    out, perm = po.undistort(xyz, cur, poses, st0)
    assert (np.diff(cur[perm]) >= 0).all()
    assert np.abs(out - xyz[perm]).max() <= 4e-6 * 30      # float rounding of the (identity) double transform chain
