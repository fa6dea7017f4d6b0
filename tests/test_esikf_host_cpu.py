"""CPU: the PRODUCT's host update engine (csrc/esikf_host.hpp: the 23-DOF manifold algebra and the two 23x23 inverses of
update_iterated_dyn_share_modified, esekfom.hpp:1620-1938, consuming reduced normal equations) driven with measurement
passes computed by the oracle, against the oracle's own restatement of the whole update.  No GPU involved: this pins the
host-side half of the host-driven engine (flb_session_set_update_engine(0), boundaries B1/B2/B3) in the CPU suite."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.helpers import small_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def iu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("iu") / "libesikf_host_shim.so")
    subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                    "-I", os.path.join(ROOT, "better_fastlio2_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "esikf_host_shim.cpp"), "-o", so], check=True, capture_output=True, text=True)
    L = C.CDLL(so)
    L.iu_create.argtypes = [dp, dp, C.c_double, C.c_int, dp]
    L.iu_create.restype = C.c_void_p
    for f in ("iu_destroy", "iu_skip"):
        getattr(L, f).argtypes = [C.c_void_p]
        getattr(L, f).restype = None
    for f in ("iu_more", "iu_need_search", "iu_converged_count"):
        getattr(L, f).argtypes = [C.c_void_p]
        getattr(L, f).restype = C.c_int
    L.iu_current_state.argtypes = [C.c_void_p, dp]
    L.iu_step.argtypes = [C.c_void_p, dp, dp]
    L.iu_step_rows.argtypes = [C.c_void_p, dp, dp, C.c_int]
    L.iu_result.argtypes = [C.c_void_p, dp, dp]
    return L


def _run_host_engine(L, oracle, prior, P, body, mp, max_iter, ext):
    """What flb's host-driven engine does per scan, with the oracle standing in for the GPU measurement kernels."""
    lim = np.full(23, 0.001)
    h = L.iu_create(np.ascontiguousarray(prior, np.float64), np.ascontiguousarray(P, np.float64).reshape(-1), 0.001, max_iter, lim)
    n = len(body)
    sel = np.ones(n, np.uint8)                      # point_selected_surf := true per scan (laserMapping.cpp:2131)
    nbr = d2 = cnt = None
    passes = searches = 0
    st = np.zeros(26)
    while L.iu_more(h):
        L.iu_current_state(h, st)
        world = oracle.transform(st, body)
        search = bool(L.iu_need_search(h))
        if search:
            nbr, d2, cnt = mp.Nearest_Search(world, 5)
            searches += 1
        M, hx, hv, _, _ = oracle.residual_pass(st, body, world, nbr, d2, cnt, search, sel, ext)
        passes += 1
        if M < 1:
            L.iu_skip(h)
        elif M < 23:
            L.iu_step_rows(h, np.ascontiguousarray(hx).reshape(-1), np.ascontiguousarray(hv), M)
        else:
            L.iu_step(h, np.ascontiguousarray(hx.T @ hx).reshape(-1), np.ascontiguousarray(hx.T @ hv))
    out_s, out_P = np.zeros(26), np.zeros(23 * 23)
    L.iu_result(h, out_s, out_P)
    L.iu_destroy(h)
    return out_s, out_P.reshape(23, 23), passes, searches


@pytest.mark.parametrize("ext,max_iter", [(False, 3), (True, 3), (False, 4)])
def test_host_engine_matches_oracle_update(iu, oracle, ext, max_iter):
    sc = small_scene(seed=5, map_half=25.0, half_extent=80.0)
    body = sc["body"][::4]
    mp = oracle.make_map(ds=0.2)
    mp.Build(sc["map"])
    s_ref, P_ref, _, stats, _ = oracle.esikf_update(sc["prior"], sc["P"], body, mp, max_iter=max_iter, extrinsic_est_en=ext)
    s, P, passes, searches = _run_host_engine(iu, oracle, sc["prior"], sc["P"], body, mp, max_iter, ext)
    assert passes == int(stats[0]) and searches == int(stats[1])
    # same algebra, different summation orders / inverse routines: ~1e-12
    assert np.abs(s - s_ref).max() < 1e-9, np.abs(s - s_ref).max()
    assert np.abs(P - P_ref).max() < 1e-10, np.abs(P - P_ref).max()
    assert np.abs(s[:3] - sc["st_true"][:3]).max() < 0.1


def test_host_engine_underdetermined_branch(iu, oracle):
    """M < 23 takes the K = P H^T (H P H^T / R + I)^-1 / R branch (esekfom.hpp:1720-1750) with explicit rows."""
    sc = small_scene(seed=6, map_half=25.0, half_extent=80.0)
    mp = oracle.make_map(ds=0.2)
    mp.Build(sc["map"])
    world = oracle.transform(sc["st_true"], sc["body"])
    _, d2, cnt = mp.Nearest_Search(world, 5)
    good = np.where((cnt == 5) & (d2[:, 4] < 0.2))[0][:12]      # a dozen well-supported points: M <= 12 < 23
    body = sc["body"][good]
    s_ref, P_ref, _, stats, _ = oracle.esikf_update(sc["prior"], sc["P"], body, mp, max_iter=3)
    s, P, passes, searches = _run_host_engine(iu, oracle, sc["prior"], sc["P"], body, mp, 3, False)
    assert 0 < int(stats[2]) < 23
    assert passes == int(stats[0])
    assert np.abs(s - s_ref).max() < 1e-9 and np.abs(P - P_ref).max() < 1e-10
