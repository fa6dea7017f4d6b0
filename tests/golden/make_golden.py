"""Regenerates tests/golden/*.npz.  Run in the BUILD container (needs /root/reference to have been compiled into
oracle/_ref by `make -C oracle`): the fixtures freeze what the reference's own ikd-Tree (compiled unmodified) plus
the restated h_share_model / ESIKF produce on small seeded scenes, so that the GPU box — where /root/reference
does not exist — can check both the oracle and the CUDA path against them.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from better_fastlio2_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from tests.helpers import sort_rows  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make(name, seed, model, ext, map_half, stride):
    rng = np.random.default_rng(seed)
    world = synth.city_world(half_extent=100, seed=seed)
    st_true = synth.trajectory_state(0)
    body = synth.scan_from_pose(world, st_true, synth.lidar_dirs(model, rng), rng, max_range=60.0)[::stride]
    mp = synth.sample_surface_map(world, (0, 0, 0), map_half, 0.2, rng)
    prior = synth.perturb_state(st_true, rng)
    P = synth.default_cov()
    assert po.have_ref(), "build oracle/_ref first"
    ref = po.RefIkdTree(ds=0.2)
    ref.Build(mp)
    w0 = po.transform(prior, body)
    x0, d0, c0 = ref.Nearest_Search(w0, 5)
    sel = np.ones(len(body), np.uint8)
    M0, hx0, h0, nv0, tot0 = po.residual_pass(prior, body, w0, x0, d0, c0, True, sel, ext)
    st, Pp, sc, stats, trace = po.esikf_update(prior, P, body, ref, max_iter=3, extrinsic_est_en=ext, want_trace=True)
    wpost, cls = po.map_incremental_classify(st, body, sc.nbr, sc.nbr_cnt, True, 0.2)
    ref.Add_Points(wpost[cls == 1], True)
    ref.Add_Points(wpost[cls == 2], False)
    final = sort_rows(ref.flatten())
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz")
    np.savez_compressed(out, map=mp, body=body, prior=prior, P=P, ext=np.array(ext), nn_d2=d0, nn_cnt=c0, sel0=sel,
                        M0=np.array(M0), HTH0=hx0.T @ hx0, HTh0=hx0.T @ h0, tot0=np.array(tot0), post=st, P_post=Pp,
                        stats=stats, trace=trace, cls=cls, final_count=np.array(len(final)),
                        final_sha256=np.array(digest(final)), final_sorted=final, truth=st_true)
    print(name, "map", mp.shape, "body", body.shape, "M0", M0, "stats", stats, "final", len(final),
          os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    make("vlp16_mini", 11, "vlp16", False, 14.0, 6)
    make("vlp16_mini_extrinsic", 12, "vlp16", True, 14.0, 6)
