"""GPU parity tests of the front-end rows through the C ABI against the CPU oracle
(oracle/frontend_oracle.cpp): UndistortPcl backward pass, pcl::VoxelGrid, recontructIKdTree's transform+filter+rebuild,
world re-projection for publishing, and the whole raw-scan -> posterior pipeline.

Tolerances (floating point rows):
  * undistortion: double math rounded to float; device sincos vs glibc sin/cos may differ in the last double ulp, so
    a coordinate may land on the neighbouring float: <= 2 float ulp, and >= 99.9 % of coordinates bit-equal;
  * voxel grid: bit-exact against the oracle summing in the same (stable) order; <= 1e-3 against the PCL std::sort
    order (float sums of <= a few dozen coordinates up to ~100 m);
  * pipeline: posterior pose within 1e-4 m / 1e-4 rad of the oracle pipeline (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

from better_fastlio2_b200 import capi, synth
from tests.helpers import small_scene, sort_rows

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def raw():
    sc = small_scene(seed=11, map_half=40.0, half_extent=100.0)
    rng = np.random.default_rng(5)
    xyz, inten, cur = synth.raw_scan_with_times(sc["body"], rng)
    poses, end = synth.imu_pose_sequence(sc["st_true"], rng)
    return dict(xyz=xyz, inten=inten, cur=cur, poses=poses, end=end, scene=sc, pts48=capi.pack_pointtype(xyz, inten, cur))


@pytest.fixture()
def rig(raw):
    tree = capi.KDTree(voxel_size=0.2, max_points=1 << 21, max_blocks=1 << 18)
    ses = capi.Session(tree, max_scan_points=1 << 17, max_iterations=3)
    fe = capi.FrontEnd(ses, max_raw_points=1 << 17)
    yield tree, ses, fe
    fe.close()
    ses.close()
    tree.close()


def _ulp_close(a, b, ulps=2):
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    tol = ulps * np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32))
    return np.abs(a - b) <= tol


def test_undistort_parity(raw, rig, oracle):
    tree, ses, fe = rig
    fe.upload(raw["pts48"])
    fe.undistort(raw["poses"], raw["end"])
    g_xyzi, g_cur, g_perm = fe.download_undistorted()
    o_xyz, o_perm = oracle.undistort(raw["xyz"], raw["cur"], raw["poses"], raw["end"])
    n = len(raw["xyz"])
    assert sorted(g_perm.tolist()) == list(range(n))
    assert np.array_equal(g_cur, raw["cur"][g_perm]) and (np.diff(g_cur) >= 0).all()     # time order, stable
    assert (np.diff(g_perm)[np.diff(g_cur) == 0] > 0).all()
    assert np.array_equal(g_xyzi[:, 3], raw["inten"][g_perm])
    g_by_in = np.empty((n, 3), np.float32)
    o_by_in = np.empty((n, 3), np.float32)
    g_by_in[g_perm] = g_xyzi[:, :3]
    o_by_in[o_perm] = o_xyz
    ok = _ulp_close(g_by_in, o_by_in)
    assert ok.all(), f"{(~ok).sum()} coordinates differ by more than 2 ulp; max {np.abs(g_by_in - o_by_in).max()}"
    assert (g_by_in == o_by_in).mean() > 0.999
    moved = np.abs(g_by_in - raw["xyz"]).max(1)
    assert moved.max() > 0.05          # the compensation is not a no-op on this trajectory (10 m/s)
    assert (moved[raw["cur"] <= 0] == 0).all()


def test_undistort_first_point_quirk(rig, oracle):
    tree, ses, fe = rig
    rng = np.random.default_rng(3)
    poses, end = synth.imu_pose_sequence(synth.trajectory_state(0), rng, n_imu=6)
    xyz = rng.uniform(-20, 20, (500, 3)).astype(np.float32)
    cur = rng.uniform(45.0, 99.0, 500).astype(np.float32)
    fe.upload(capi.pack_pointtype(xyz, None, cur))
    fe.undistort(poses, end)
    g, gc, gp = fe.download_undistorted()
    o, op = oracle.undistort(xyz, cur, poses, end)
    assert np.array_equal(gp, op)                      # distinct stamps: one possible order
    assert _ulp_close(g[:, :3], o, ulps=4).all()       # includes point 0, compensated by every earlier segment
    assert np.isfinite(g).all()


@pytest.mark.parametrize("leaf", [0.5, 0.2])
def test_voxel_filter_bit_exact(raw, rig, oracle, leaf):
    tree, ses, fe = rig
    fe.upload(raw["pts48"])
    n_out = fe.voxel_filter(leaf)                      # no undistortion: upload order is the summation order
    g, gc = fe.download_down()
    p4 = np.column_stack([raw["xyz"], raw["inten"]]).astype(np.float32)
    o, oc, ovf = oracle.voxel_grid(p4, leaf, curvature=raw["cur"], order="stable")
    assert not ovf and n_out == len(o) == len(g) == ses.n
    assert np.array_equal(g, o), np.abs(g - o).max()
    assert np.array_equal(gc, oc)
    o_pcl, _, _ = oracle.voxel_grid(p4, leaf, order="pcl")
    assert np.abs(g - o_pcl).max() < 1e-3              # PCL's own (unspecified) in-leaf order: rounding only


def test_undistort_then_filter_and_publish(raw, rig, oracle):
    tree, ses, fe = rig
    fe.upload(raw["pts48"])
    fe.undistort(raw["poses"], raw["end"])
    und, ucur, perm = fe.download_undistorted()
    n_out = fe.voxel_filter(0.5)
    g, gc = fe.download_down()
    o, oc, _ = oracle.voxel_grid(und, 0.5, curvature=ucur, order="stable")
    assert n_out == len(o)
    assert np.array_equal(g, o) and np.array_equal(gc, oc)
    # publish_frame_world: RGBpointBodyToWorld of feats_down_body / feats_undistort (laserMapping.cpp:1502-1540)
    st = raw["scene"]["st_true"]
    w0 = fe.points_to_world(0, st)
    w1 = fe.points_to_world(1, st)
    assert np.array_equal(w0, oracle.body_to_world4(st, g))
    assert np.array_equal(w1, oracle.body_to_world4(st, und))


def test_process_one_call_equals_three(raw, rig):
    tree, ses, fe = rig
    fe.upload(raw["pts48"])
    fe.undistort(raw["poses"], raw["end"])
    n3 = fe.voxel_filter(0.5)
    a, ac = fe.download_down()
    buf = np.ascontiguousarray(raw["pts48"])
    poses = np.ascontiguousarray(raw["poses"], np.float64)
    end = np.ascontiguousarray(raw["end"], np.float64)
    n1 = fe.process_ptr(buf.ctypes.data, len(buf), poses, end, 0.5)
    b, bc = fe.download_down()
    assert n1 == n3 and np.array_equal(a, b) and np.array_equal(ac, bc)
    # without IMU poses the scan is only filtered (upload order)
    n0 = fe.process_ptr(buf.ctypes.data, len(buf), np.zeros((0, 22)), end, 0.5)
    und, _, perm = fe.download_undistorted()
    assert np.array_equal(perm, np.arange(len(buf))) and np.array_equal(und[:, :3], raw["xyz"]) and n0 > 0


def test_raw_scan_pipeline_pose_parity(raw, rig, oracle):
    """meas.lidar -> UndistortPcl -> VoxelGrid -> update -> map_incremental on the GPU vs the same chain on the oracle."""
    tree, ses, fe = rig
    sc = raw["scene"]
    tree.Build(sc["map"])
    fe.upload(raw["pts48"])
    fe.undistort(raw["poses"], raw["end"])
    n_out = fe.voxel_filter(0.5)
    assert n_out > 1000
    s_gpu, P_gpu, r = ses.scan_step(None, None, sc["prior"], sc["P"])
    ref = oracle.make_map(ds=0.2)
    ref.Build(sc["map"])
    o_xyz, o_perm = oracle.undistort(raw["xyz"], raw["cur"], raw["poses"], raw["end"])
    o_ds, _, _ = oracle.voxel_grid(np.column_stack([o_xyz, raw["inten"][o_perm]]), 0.5, order="pcl")
    assert len(o_ds) == n_out
    s_cpu, P_cpu, *_ = oracle.esikf_update(sc["prior"], sc["P"], o_ds[:, :3], ref, max_iter=3)
    assert np.abs(s_gpu[:3] - s_cpu[:3]).max() <= 1e-4
    assert np.abs(s_gpu[3:7] - s_cpu[3:7]).max() <= 1e-4
    assert r.update.effct_feat_num > 500


def test_voxel_grid_filter_standalone_and_guards(raw, oracle):
    tree = capi.KDTree(voxel_size=0.2, max_points=1 << 16, max_blocks=1 << 14)
    p4 = np.column_stack([raw["xyz"], raw["inten"]]).astype(np.float32)
    g = capi.voxel_grid_filter(tree, raw["pts48"], 0.4)
    o, _, _ = oracle.voxel_grid(p4, 0.4, order="stable")
    assert np.array_equal(g, o)
    assert len(capi.voxel_grid_filter(tree, np.zeros((0, 12), np.float32), 0.4)) == 0
    one = capi.pack_pointtype(np.array([[1.0, -2.0, 3.0]], np.float32), [7.0], [0.0])
    assert np.array_equal(capi.voxel_grid_filter(tree, one, 0.5), np.array([[1.0, -2.0, 3.0, 7.0]], np.float32))
    # PCL's int32 overflow guard: the cloud is returned unchanged
    far = np.array([[0, 0, 0], [500, 500, 500], [-100, 3, 9]], np.float32)
    out = capi.voxel_grid_filter(tree, capi.pack_pointtype(far, [1, 2, 3]), 0.001)
    assert np.array_equal(out[:, :3], far) and np.array_equal(out[:, 3], np.array([1, 2, 3], np.float32))
    # leaf faces / negative coordinates
    gpts = np.array([[-0.5, 0.0, 0.5], [-0.5000001, 0.0, 0.5], [0.4999999, 0.0, 0.5], [0.0, 0.0, 0.999]], np.float32)
    og, _, _ = oracle.voxel_grid(np.column_stack([gpts, np.zeros(4, np.float32)]), 0.5, order="stable")
    assert np.array_equal(capi.voxel_grid_filter(tree, capi.pack_pointtype(gpts), 0.5), og)
    with pytest.raises(capi.FlbError):
        capi.voxel_grid_filter(tree, raw["pts48"], 0.0)
    tree.close()


def test_frontend_capacity_and_argument_errors(raw, rig):
    tree, ses, fe = rig
    small = capi.FrontEnd(ses, max_raw_points=100)
    with pytest.raises(capi.FlbError):
        small.upload(raw["pts48"])
    small.close()
    fe.upload(raw["pts48"])
    with pytest.raises(capi.FlbError):
        fe.undistort(np.zeros((300, 22)), raw["end"])          # more IMU poses than FLB_MAX_IMU_POSES
    with pytest.raises(capi.FlbError):
        fe.voxel_filter(-1.0)
    # a scan with a single IMU pose has no segment: nothing is compensated, only sorted by time
    fe.undistort(raw["poses"][:1], raw["end"])
    g, gc, gp = fe.download_undistorted()
    assert np.array_equal(g[:, :3], raw["xyz"][gp])
    # empty scan
    fe.upload(np.zeros((0, 12), np.float32))
    fe.undistort(raw["poses"], raw["end"])
    assert fe.voxel_filter(0.5) == 0


def test_reconstruct_keyframes(raw, oracle):
    """recontructIKdTree (laserMapping.cpp:632-664): subMap += transformPointCloud(kf, pose); VoxelGrid; reconstruct."""
    rng = np.random.default_rng(9)
    tree = capi.KDTree(voxel_size=0.2, max_points=1 << 20, max_blocks=1 << 17)
    tree.Build(raw["scene"]["map"][:5000])           # previous content must disappear
    clouds, clouds48, poses = [], [], []
    for k in range(5):
        idx = rng.choice(len(raw["xyz"]), 6000, replace=False)
        p4 = np.column_stack([raw["xyz"][idx], raw["inten"][idx]]).astype(np.float32)
        clouds.append(p4)
        clouds48.append(capi.pack_pointtype(p4[:, :3], p4[:, 3]))
        poses.append([3.0 * k, 0.2 * k, 0.1, 0.01 * k, -0.02, 0.3 * k])
    clouds48.insert(2, np.zeros((0, 12), np.float32))    # an empty key frame
    poses.insert(2, [0, 0, 0, 0, 0, 0])
    clouds.insert(2, np.zeros((0, 4), np.float32))
    leaf = 0.4
    feats = capi.reconstruct_keyframes(tree, clouds48, np.array(poses, np.float32), leaf)
    sub = np.concatenate([oracle.transform_cloud_rpy(c, np.array(p, np.float32)) for c, p in zip(clouds, poses)])
    o, _, _ = oracle.voxel_grid(sub, leaf, order="stable")
    assert np.array_equal(feats, o)
    assert tree.validnum() == len(o) == tree.size()
    assert np.array_equal(sort_rows(tree.flatten()), sort_rows(o[:, :3]))
    # empty sub-map: reconstruct deletes everything
    none = capi.reconstruct_keyframes(tree, [], np.zeros((0, 6), np.float32), leaf)
    assert len(none) == 0 and tree.validnum() == 0
    tree.close()


def test_frontend_golden_fixture(rig):
    """GPU path vs the committed fixture tests/golden/frontend/frontend_mini.npz (no oracle call needed on the box)."""
    tree, ses, fe = rig
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "frontend", "frontend_mini.npz"))
    fe.upload(capi.pack_pointtype(g["xyz"], g["intensity"], g["curvature"]))
    fe.undistort(g["poses"], g["end"])
    und, cur, perm = fe.download_undistorted()
    g_by_in = np.empty_like(g["xyz"])
    o_by_in = np.empty_like(g["xyz"])
    g_by_in[perm] = und[:, :3]
    o_by_in[g["perm"]] = g["undistorted"]
    assert _ulp_close(g_by_in, o_by_in).all()
    n = fe.voxel_filter(float(g["leaf"]))
    d, dc = fe.download_down()
    assert abs(n - len(g["down_stable"])) <= 2                    # a 1-ulp difference may move a point across a leaf face
    if n == len(g["down_stable"]):
        assert np.abs(d - g["down_stable"]).max() < 1e-3 and np.abs(d - g["down_pcl"]).max() < 1e-3
