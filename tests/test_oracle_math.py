"""CPU tests: the restated ESIKF / plane-fit / manifold maths of the oracle cross-validated against numpy / scipy
(the reference has no tests or golden vectors for these and needs Eigen, which is absent here: "parity unpinned")."""
import numpy as np
from scipy.spatial.transform import Rotation

from better_fastlio2_b200 import synth


def rand_state(rng):
    q = Rotation.random(random_state=rng.integers(1 << 30)).as_quat()
    q2 = Rotation.from_rotvec(rng.normal(0, 0.05, 3)).as_quat()
    g = rng.normal(0, 1, 3) + np.array([0, 0, -3.0])
    g = g / np.linalg.norm(g) * synth.G_LEN
    return synth.make_state(pos=rng.normal(0, 5, 3), rot=q, offR=q2, offT=rng.normal(0, 0.1, 3), vel=rng.normal(0, 1, 3),
                            bg=rng.normal(0, 0.01, 3), ba=rng.normal(0, 0.01, 3), grav=g)


def test_esti_plane_matches_lstsq(oracle):
    rng = np.random.default_rng(0)
    nok = 0
    for _ in range(300):
        n = rng.normal(0, 1, 3)
        n /= np.linalg.norm(n)
        c = rng.uniform(-40, 40, 3)
        u = np.cross(n, [1, 0, 0.3])
        u /= np.linalg.norm(u)
        v = np.cross(n, u)
        pts = (c + rng.uniform(-0.3, 0.3, (5, 1)) * u + rng.uniform(-0.3, 0.3, (5, 1)) * v + rng.normal(0, 0.01, (5, 1)) * n)
        pts = pts.astype(np.float32)
        ok, pabcd = oracle.esti_plane(pts, 0.1)
        A64 = pts.astype(np.float64)
        x = np.linalg.lstsq(A64, -np.ones(5), rcond=None)[0]
        nn = np.linalg.norm(x)
        ref = np.r_[x / nn, 1 / nn]
        # the reference solves in FLOAT (common_lib.h:520); points tens of metres from the origin with a 0.3 m spread
        # make A ill-conditioned, so the agreement with a float64 solve is bounded by eps32 * cond(A)
        tol = max(2e-4, 40 * 1.2e-7 * np.linalg.cond(A64))
        assert np.allclose(pabcd[:3], ref[:3], atol=tol), (pabcd, ref, tol)
        assert abs(pabcd[3] - ref[3]) < tol * max(1.0, np.linalg.norm(c)) * 2
        # and a float32 pivoted-QR least squares from scipy (same algorithm family) agrees much more tightly
        from scipy.linalg import qr, solve_triangular
        Q, Rm, piv = qr(pts, mode="economic", pivoting=True)
        y = solve_triangular(Rm, Q.T @ (-np.ones(5, np.float32)))
        xs = np.zeros(3, np.float32)
        xs[piv] = y
        xs = xs / np.linalg.norm(xs)
        assert np.allclose(pabcd[:3], xs, atol=max(5e-5, tol * 0.2)), (pabcd, xs)
        resid = np.abs(pts @ ref[:3] + ref[3]).max()
        if resid < 0.08:
            assert ok
            nok += 1
    assert nok > 250
    # outlier -> rejected
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0.5, 0.5, 0.9]], np.float32) + 3
    ok, _ = oracle.esti_plane(pts, 0.1)
    assert not ok


def test_boxplus_boxminus_roundtrip(oracle):
    rng = np.random.default_rng(1)
    L = oracle.lio()
    for _ in range(100):
        s = rand_state(rng)
        d = rng.normal(0, 0.02, 23)
        s2 = s.copy()
        L.orc_boxplus(s2, d)
        r = np.zeros(23)
        L.orc_boxminus(s2, s, r)
        assert np.allclose(r, d, atol=1e-9), np.abs(r - d).max()
        # rotation block agrees with scipy: q <- q * exp(d)
        q = (Rotation.from_quat(s[3:7]) * Rotation.from_rotvec(d[3:6])).as_quat()
        assert min(np.abs(q - s2[3:7]).max(), np.abs(q + s2[3:7]).max()) < 1e-12
        assert abs(np.linalg.norm(s2[23:26]) - synth.G_LEN) < 1e-9


def test_A_matrix_is_right_jacobian_transposed(oracle):
    # A_matrix(v) (mtkmath.hpp:235-247) = I + (1-cos)/|v|^2 [v]x + (1 - sin/|v|)/|v|^2 [v]x^2  (SO(3) left Jacobian)
    rng = np.random.default_rng(2)
    L = oracle.lio()
    for _ in range(50):
        v = rng.normal(0, 0.3, 3)
        A = np.zeros(9)
        L.orc_A_matrix(v, A)
        A = A.reshape(3, 3)
        th = np.linalg.norm(v)
        K = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
        ref = np.eye(3) + (1 - np.cos(th)) / th ** 2 * K + (1 - np.sin(th) / th) / th ** 2 * K @ K
        assert np.allclose(A, ref, atol=1e-13)
        # defining property of the left Jacobian: Exp(v + dv) ~= Exp(J dv) Exp(v)
        dv = rng.normal(0, 1e-6, 3)
        lhs = Rotation.from_rotvec(v + dv)
        rhs = Rotation.from_rotvec(ref @ dv) * Rotation.from_rotvec(v)
        assert (lhs * rhs.inv()).magnitude() < 1e-10
    A = np.zeros(9)
    L.orc_A_matrix(np.zeros(3), A)
    assert np.array_equal(A.reshape(3, 3), np.eye(3))


def test_invert(oracle):
    rng = np.random.default_rng(3)
    L = oracle.lio()
    for n in (1, 5, 23):
        A = rng.normal(0, 1, (n, n)) + n * np.eye(n)
        Ai = np.zeros((n, n))
        assert L.orc_invert(np.ascontiguousarray(A).reshape(-1), Ai.reshape(-1), n) == 1
        assert np.allclose(Ai, np.linalg.inv(A), rtol=1e-10, atol=1e-12)


def test_s2_matrices(oracle):
    L = oracle.lio()
    rng = np.random.default_rng(4)
    for _ in range(20):
        g = rng.normal(0, 1, 3)
        g = g / np.linalg.norm(g) * synth.G_LEN
        Bx, Nx, Mx = np.zeros(6), np.zeros(6), np.zeros(6)
        L.orc_s2_mats(g, np.zeros(2), Bx, Nx, Mx)
        Bx, Nx, Mx = Bx.reshape(3, 2), Nx.reshape(2, 3), Mx.reshape(3, 2)
        # Bx spans the tangent plane at g (orthogonal to g), columns orthonormal (S2.hpp:215-231)
        assert np.allclose(g @ Bx, 0, atol=1e-9)
        assert np.allclose(Bx.T @ Bx, np.eye(2), atol=1e-9)
        # Nx_yy * Mx(0) = identity on the tangent space (consistency of boxplus / boxminus Jacobians)
        assert np.allclose(Nx @ Mx, np.eye(2), atol=1e-9)


def test_first_pass_is_textbook_kalman_update(oracle):
    """With dx = 0 the first ESIKF pass reduces to x + K(-h), K = (H^T H + (P/R)^-1)^-1 H^T (esekfom.hpp:1788-1821)."""
    from tests.helpers import small_scene
    sc = small_scene(seed=2, map_half=30.0, half_extent=60.0)
    m = oracle.PortMap(ds=sc["ds"])
    m.Build(sc["map"])
    body = sc["body"][::8]
    st, P, scr, stats, trace = oracle.esikf_update(sc["prior"], sc["P"], body, m, max_iter=3, want_trace=True)
    world = oracle.transform(sc["prior"], body)
    x, d2, cnt = m.Nearest_Search(world, 5)
    sel = np.ones(len(body), np.uint8)
    M, hx, h, nv, tot = oracle.residual_pass(sc["prior"], body, world, x, d2, cnt, True, sel, False)
    assert M > 100
    H = np.zeros((M, 23))
    H[:, :12] = hx
    R = 0.001
    K = np.linalg.solve(H.T @ H + np.linalg.inv(sc["P"] / R), H.T)
    dx = K @ h
    s1 = sc["prior"].copy()
    oracle.lio().orc_boxplus(s1, dx)
    assert np.allclose(trace[0], s1, atol=1e-9), np.abs(trace[0] - s1).max()
    # posterior covariance is symmetric PSD-ish and shrinks the observed pose block
    assert np.allclose(P, P.T, atol=1e-10)
    assert np.all(np.diag(P)[:6] < np.diag(sc["P"])[:6])
    assert np.linalg.norm(st[:3] - sc["st_true"][:3]) < np.linalg.norm(sc["prior"][:3] - sc["st_true"][:3])


def test_fov_segment_moves_cube(oracle):
    f = oracle.FovSegment(cube_len=200.0, det_range=40.0)
    assert len(f.step(np.zeros(3))) == 0 and f.init[0] == 1
    assert np.allclose(f.local_map, [-100, -100, -100, 100, 100, 100])
    assert len(f.step(np.array([10.0, 0, 0]))) == 0          # far from every face (threshold 1.5*40 = 60)
    b = f.step(np.array([45.0, 0, 0]))                        # 55 m from the +x face -> move by mov_dist
    mov = max((200 - 2 * 1.5 * 40) * 0.5 * 0.9, 40 * 0.5)
    assert len(b) == 1
    assert np.allclose(b[0], [-100, -100, -100, -100 + mov, 100, 100])
    assert np.allclose(f.local_map, [-100 + mov, -100, -100, 100 + mov, 100, 100])
