"""Shared scene builders for the parity tests (seeded, small enough for the CPU oracle to finish in seconds)."""
import numpy as np

from better_fastlio2_b200 import synth


def sort_rows(a):
    a = np.asarray(a)
    if len(a) == 0:
        return a
    idx = np.lexsort((a[:, 2], a[:, 1], a[:, 0]))
    return a[idx]


def small_scene(seed=1, model="vlp16", map_half=50.0, ds=0.2, half_extent=120.0):
    rng = np.random.default_rng(seed)
    world = synth.city_world(half_extent=half_extent, seed=seed)
    st_true = synth.trajectory_state(0)
    body = synth.scan_from_pose(world, st_true, synth.lidar_dirs(model, rng), rng)
    mp = synth.sample_surface_map(world, (0, 0, 0), map_half, ds, rng)
    prior = synth.perturb_state(st_true, rng)
    return dict(rng=rng, world=world, st_true=st_true, body=body, map=mp, prior=prior, P=synth.default_cov(), ds=ds)


def knn_equal(d2_a, xyz_a, cnt_a, d2_b, xyz_b, cnt_b):
    """Exact k-NN agreement: counts and sorted float distances bit-equal; coordinates equal wherever a query has no
    duplicated distance (ties may legitimately be ordered differently, SURVEY.md §7 hard part 1)."""
    assert np.array_equal(cnt_a, cnt_b)
    fin = np.isfinite(d2_a)
    assert np.array_equal(fin, np.isfinite(d2_b))
    assert np.array_equal(d2_a[fin], d2_b[fin]), f"max |dd2| = {np.abs(d2_a[fin] - d2_b[fin]).max()}"
    k = d2_a.shape[1]
    tie = np.zeros(len(d2_a), bool)
    for j in range(k - 1):
        tie |= (d2_a[:, j] == d2_a[:, j + 1]) & np.isfinite(d2_a[:, j])
    ok = ~tie
    xa = np.where(np.isfinite(xyz_a[ok]), xyz_a[ok], 0)
    xb = np.where(np.isfinite(xyz_b[ok]), xyz_b[ok], 0)
    assert np.array_equal(xa, xb)
    return int(tie.sum())


def maps_match(a, b, tol=2e-6):
    """Same point set up to float rounding of individual coordinates: the inserted world points are float roundings
    of a double transform, so two engines whose states agree to ~1e-13 can differ by 1 ulp in a few coordinates."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    if len(a) != len(b):
        return False, f"sizes {len(a)} != {len(b)}"
    va = a.view([("", a.dtype)] * 3).ravel()
    vb = b.view([("", b.dtype)] * 3).ravel()
    only_a = a[~np.isin(va, vb)]
    only_b = b[~np.isin(vb, va)]
    if len(only_a) != len(only_b):
        return False, f"unmatched {len(only_a)} vs {len(only_b)}"
    if len(only_a) > max(20, len(a) // 1000):
        return False, f"too many rounding differences: {len(only_a)}"
    for p in only_a:
        d = np.abs(only_b - p).max(1)
        j = int(np.argmin(d))
        if d[j] > tol * max(1.0, float(np.abs(p).max())):
            return False, f"point {p} has no partner within tolerance (best {d[j]})"
        only_b = np.delete(only_b, j, 0)
    return True, f"{len(only_a)} coordinates differ by rounding"


# ---------------------------------------------------------------------------------------------- size-independent properties
def brute_knn_d2(points, q, k=5):
    """k smallest squared distances from q to points in float32, (dx*dx + dy*dy) + dz*dz like calc_dist (ikd_Tree.cpp:1373)."""
    p = np.asarray(points, np.float32)
    q = np.asarray(q, np.float32)
    d = ((p[:, 0] - q[0]) ** 2 + (p[:, 1] - q[1]) ** 2) + (p[:, 2] - q[2]) ** 2
    k = min(k, len(d))
    return np.sort(np.partition(d, k - 1)[:k])


def map_properties(tree, queries, rng, n_brute=32):
    """Properties every exact k-NN map must satisfy at ANY size (used on the CPU with the reference ikd-Tree at a small
    size and on the GPU at BASELINE's full size): sorted distances, coordinates reproduce distances, agreement with a
    brute-force scan of the map's own content, idempotence of the downsampled insert, box delete bookkeeping."""
    queries = np.ascontiguousarray(queries, np.float32)
    content = tree.flatten()
    assert len(content) == tree.validnum() and len(content) >= 5
    xyz, d2, cnt = tree.Nearest_Search(queries, 5)
    assert (cnt == 5).all()
    assert (np.diff(d2, axis=1) >= 0).all()                                     # ascending
    dd = ((xyz[:, :, 0] - queries[:, None, 0]) ** 2 + (xyz[:, :, 1] - queries[:, None, 1]) ** 2) + \
         (xyz[:, :, 2] - queries[:, None, 2]) ** 2
    assert np.array_equal(dd.astype(np.float32), d2)                            # the returned points ARE at those distances
    pick = rng.choice(len(queries), min(n_brute, len(queries)), replace=False)
    for i in pick:
        assert np.array_equal(brute_knn_d2(content, queries[i]), d2[i]), i      # exact: nothing nearer exists in the map
    # downsampled insert twice = once (each touched voxel already holds its winner)
    X = (queries[::3] + rng.normal(0, 0.03, (len(queries[::3]), 3))).astype(np.float32)
    tree.Add_Points(X, True)
    v1, c1 = tree.validnum(), sort_rows(tree.flatten())
    tree.Add_Points(X, True)
    assert tree.validnum() == v1
    c2 = sort_rows(tree.flatten())
    assert np.array_equal(c1, c2)
    # box delete: count, emptiness of the half-open box, search still exact afterwards
    ctr = np.median(queries, axis=0)
    box = np.array([ctr[0] - 6, ctr[1] - 6, ctr[2] - 3, ctr[0] + 6, ctr[1] + 6, ctr[2] + 30], np.float32)
    inside = ((c2 >= box[:3]) & (c2 < box[3:])).all(1)
    nd = tree.Delete_Point_Boxes(box[None, :])
    assert nd == int(inside.sum()) and nd > 0
    after = tree.flatten()
    assert len(after) == len(c2) - nd == tree.validnum()
    assert not ((after >= box[:3]) & (after < box[3:])).all(1).any()
    near = queries[((queries >= box[:3] - 1) & (queries < box[3:] + 1)).all(1)][:8]
    if len(near):
        _, d3, c3 = tree.Nearest_Search(near, 5)
        for j in range(len(near)):
            assert np.array_equal(brute_knn_d2(after, near[j]), d3[j][:c3[j]])
    return dict(map_points=len(content), queries=len(queries), deleted=nd)
