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
