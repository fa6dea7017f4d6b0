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
