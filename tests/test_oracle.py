"""CPU tests of the oracle itself (the checker must be right before it checks anything).

The reference has no tests or golden vectors for this path (parity
unpinned).  What can be pinned is pinned here:
  * the restated sampler against the REFERENCE's own sampler source compiled from
    /root/reference/src/sample_eigen.h (oracle/_ref, golden fixture sampler_ref.npz);
  * dual-number vs closed-form vs finite-difference Jacobians;
  * Ceres-LM invariants and an independent SciPy least-squares second opinion.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import binding
from workloads import synthetic
def _sample(I, gx, gy, y, x):
    out = np.zeros(3, dtype=np.float32)
    rows, cols = I.shape
    binding.lib().oracle_sample_linear(C.c_void_p(I.ctypes.data), C.c_void_p(gx.ctypes.data),
                                       C.c_void_p(gy.ctypes.data), rows, cols, C.c_float(y), C.c_float(x),
                                       C.c_void_p(out.ctypes.data))
    return out


def test_sampler_matches_reference_golden_bit_exact():
    """Golden vectors produced by the reference's own SampleLinear (sample_eigen.h:55-102)."""
    g = np.load(os.path.join(GOLDEN, "sampler_ref.npz"))
    I, gx, gy = g["I"], g["gx"], g["gy"]
    got = np.stack([_sample(I, gx, gy, y, x) for x, y in zip(g["xs"], g["ys"])])
    assert got.tobytes() == g["out"].tobytes()  # bit-exact, incl. borders, (-1,0) band, NaN/huge
    # Chain<float,2,Jet>::Rule (jet_extras.h:86-111): value = sample, v = gx*x.v + gy*y.v
    for i in range(g["xa"].shape[0]):
        s = _sample(I, gx, gy, np.float32(g["ya"][i]), np.float32(g["xa"][i]))
        assert float(s[0]) == g["ja"][i]
        np.testing.assert_allclose(float(s[1]) * g["xv"][i] + float(s[2]) * g["yv"][i], g["jv"][i], rtol=1e-14, atol=1e-14)


def test_sampler_matches_reference_binary_when_present():
    ref = binding.ref_lib()
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    rng = np.random.default_rng(5)
    rows, cols = 17, 29
    I = rng.uniform(0, 255, size=(rows, cols)).astype(np.float32)  # non-integer channel values too
    gx, gy = np.zeros_like(I), np.zeros_like(I)
    binding.lib().oracle_imgradient(C.c_void_p(I.ctypes.data), rows, cols, C.c_void_p(gx.ctypes.data), C.c_void_p(gy.ctypes.data))
    xs = rng.uniform(-2, cols + 1, size=3000).astype(np.float32)
    ys = rng.uniform(-2, rows + 1, size=3000).astype(np.float32)
    out = np.zeros(3, dtype=np.float32)
    for x, y in zip(xs, ys):
        ref.ref_sample_linear(C.c_void_p(I.ctypes.data), C.c_void_p(gx.ctypes.data), C.c_void_p(gy.ctypes.data),
                              rows, cols, C.c_float(y), C.c_float(x), C.c_void_p(out.ctypes.data))
        assert _sample(I, gx, gy, y, x).tobytes() == out.tobytes()


def test_imgradient_border_rule():
    """src/imgproc.cc:27-96: 0.5*central difference inside, zero on all four borders."""
    rng = np.random.default_rng(1)
    I = rng.integers(0, 256, size=(9, 13)).astype(np.float32)
    gx, gy = np.empty_like(I), np.empty_like(I)
    binding.lib().oracle_imgradient(C.c_void_p(I.ctypes.data), 9, 13, C.c_void_p(gx.ctypes.data), C.c_void_p(gy.ctypes.data))
    ex, ey = np.zeros_like(I), np.zeros_like(I)
    ex[1:-1, 1:-1] = 0.5 * (I[1:-1, 2:] - I[1:-1, :-2])
    ey[1:-1, 1:-1] = 0.5 * (I[2:, 1:-1] - I[:-2, 1:-1])
    assert np.array_equal(gx, ex) and np.array_equal(gy, ey)
    for g in (gx, gy):
        assert not g[0].any() and not g[-1].any() and not g[:, 0].any() and not g[:, -1].any()


def test_patch_weights():
    w = np.zeros(25)
    binding.lib().oracle_patch_weights(2, 0, C.c_void_p(w.ctypes.data))
    assert np.array_equal(w, np.ones(25))
    binding.lib().oracle_patch_weights(2, 1, C.c_void_p(w.ctypes.data))
    np.testing.assert_allclose(w, synthetic.patch_weights(2, True), rtol=1e-15)
    assert abs(w.sum() - 1.0) < 1e-14 and w[12] == w.max()


def test_pose_params_roundtrip():
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(3)
    for _ in range(200):
        rv = rng.normal(size=3)
        rv *= rng.choice([1e-9, 1e-3, 0.5, 3.0]) / np.linalg.norm(rv)  # |rv| < pi
        T = np.eye(4)
        T[:3, :3] = Rotation.from_rotvec(rv).as_matrix()
        T[:3, 3] = rng.normal(size=3)
        p = binding.pose_to_params(T)
        np.testing.assert_allclose(p[:3], rv, atol=1e-12)
        np.testing.assert_allclose(binding.params_to_pose(p), T, atol=1e-12)
    # angle-axis point rotation agrees with the matrix form
    aa, pt, out = rng.normal(size=3), rng.normal(size=3), np.zeros(3)
    binding.lib().oracle_angle_axis_rotate_point(C.c_void_p(aa.ctypes.data), C.c_void_p(pt.ctypes.data), C.c_void_p(out.ctypes.data))
    np.testing.assert_allclose(out, Rotation.from_rotvec(aa).as_matrix() @ pt, atol=1e-14)


def test_autodiff_vs_analytic_vs_fd(small_win):
    """Jet<double,9> path == closed form to 1e-9; both match central differences of the
    *smooth surrogate* (the Jacobian uses the interpolated gradient image, not the derivative
    of the interpolant — SURVEY App. A.3 — so FD is checked on the geometric chain only)."""
    w = small_win
    ow = binding.OracleWindow(w)
    rng = np.random.default_rng(0)
    for p in rng.choice(w.n_points, size=25, replace=False):
        for o in range(w.obs_offsets[p], w.obs_offsets[p + 1]):
            f = int(w.obs_frame[o])
            cam = w.cams_init[f] + (1e-3 if f == 0 else 0.0)  # also exercise non-tiny angle on frame 0
            r1, Jc1, Jp1 = ow.residual_block(f, cam, w.points_init[p], w.desc[p], mode=1)
            r0, Jc0, Jp0 = ow.residual_block(f, cam, w.points_init[p], w.desc[p], mode=0)
            r2, _, _ = ow.residual_block(f, cam, w.points_init[p], w.desc[p], mode=2)
            np.testing.assert_allclose(r1, r0, atol=1e-9)
            np.testing.assert_allclose(r2, r0, atol=1e-9)
            np.testing.assert_allclose(Jc1, Jc0, rtol=1e-8, atol=1e-8)
            np.testing.assert_allclose(Jp1, Jp0, rtol=1e-8, atol=1e-8)
    # tiny-angle branch (theta^2 <= eps): d/dw = -[X]x
    cam = np.array([1e-9, -2e-9, 5e-10, 0.01, 0.02, 0.03])
    r1, Jc1, Jp1 = ow.residual_block(1, cam, w.points_init[0], w.desc[0], mode=1)
    r0, Jc0, Jp0 = ow.residual_block(1, cam, w.points_init[0], w.desc[0], mode=0)
    np.testing.assert_allclose(Jc1, Jc0, rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(Jp1, Jp0, rtol=1e-8, atol=1e-8)


def test_geometric_chain_fd(small_win):
    """d(u,v)/d(cam,point) from the Jets vs central finite differences of the projection."""
    w = small_win
    ow = binding.OracleWindow(w)
    # Use a linear ramp image: I = 2x + 3y -> sample = 2u + 3v (exact for bilinear), gx = 2, gy = 3
    rows, cols = w.rows, w.cols
    yy, xx = np.mgrid[0:rows, 0:cols].astype(np.float32)
    ramp = (0.25 * xx + 0.5 * yy).astype(np.float32)[None, None].repeat(w.n_frames, axis=0)
    ow2 = binding.OracleWindow(w, planes=ramp)
    p, f = 40, 2
    cam, X = w.cams_init[f].copy(), w.points_init[p].copy()
    r, Jc, Jp = ow2.residual_block(f, cam, X, w.desc[p], mode=1)
    h = 1e-3  # large step: the fp32 sampler quantises values at ~8e-6
    for q in range(9):
        d = np.zeros(9); d[q] = h
        rp, _, _ = ow2.residual_block(f, cam + d[:6], X + d[6:], w.desc[p], mode=2)
        rm, _, _ = ow2.residual_block(f, cam - d[:6], X - d[6:], w.desc[p], mode=2)
        fd = (rp - rm) / (2 * h)
        an = Jc[:, q] if q < 6 else Jp[:, q - 6]
        np.testing.assert_allclose(an, fd, rtol=2e-3, atol=2e-2)  # fp32 sampler quantisation limits FD


def test_evaluate_blocks_consistent(small_ragged_win):
    w = small_ragged_win
    ow = binding.OracleWindow(w)
    e = ow.evaluate(w.cams_init, w.points_init, 1)
    assert abs(e["cost"] - ow.cost(w.cams_init, w.points_init)) <= 1e-9 * e["cost"]
    # rebuild blocks from per-observation Jacobians in numpy
    F, n = w.n_frames, w.n_points
    U = np.zeros((F, 6, 6)); gc = np.zeros((F, 6)); V = np.zeros((n, 3, 3)); gp = np.zeros((n, 3))
    a = w.huber
    for p in range(n):
        for o in range(w.obs_offsets[p], w.obs_offsets[p + 1]):
            f = int(w.obs_frame[o])
            r, Jc, Jp = ow.residual_block(f, w.cams_init[f], w.points_init[p], w.desc[p], mode=1)
            s = float(r @ r)
            assert abs(s - e["obs_sqnorm"][o]) <= 1e-12 * max(1.0, s)
            rho1 = 1.0 if s <= a * a else a / np.sqrt(s)
            V[p] += rho1 * Jp.T @ Jp; gp[p] += rho1 * Jp.T @ r
            if f != w.fixed_frame:
                U[f] += rho1 * Jc.T @ Jc; gc[f] += rho1 * Jc.T @ r
                np.testing.assert_allclose(e["W"][o], rho1 * Jc.T @ Jp, rtol=1e-10, atol=1e-9)
            else:
                assert not e["W"][o].any()
    np.testing.assert_allclose(e["U"], U, rtol=1e-10, atol=1e-8)
    np.testing.assert_allclose(e["gc"], gc, rtol=1e-10, atol=1e-8)
    np.testing.assert_allclose(e["V"], V, rtol=1e-10, atol=1e-8)
    np.testing.assert_allclose(e["gp"], gp, rtol=1e-10, atol=1e-8)
    assert not e["U"][w.fixed_frame].any()


def test_lm_invariants(small_win):
    w = small_win
    ow = binding.OracleWindow(w)
    cams, pts, s, tr = ow.solve(w.cams_init, w.points_init)
    assert s["termination_type"] == 0
    assert s["final_cost"] < s["initial_cost"]
    assert np.array_equal(cams[w.fixed_frame], w.cams_init[w.fixed_frame])  # gauge: first camera constant
    acc = [t for t in tr if t["step_is_successful"]]
    costs = [tr[0]["cost"]] + [t["cost"] for t in acc]
    assert all(b < a for a, b in zip(costs, costs[1:]))                      # monotone accepted costs
    assert abs(costs[-1] - s["final_cost"]) == 0.0
    assert abs(ow.cost(cams, pts) - s["final_cost"]) <= 1e-9 * s["final_cost"]
    assert s["num_successful_steps"] == len(acc)
    for prev, t in zip(tr, tr[1:]):
        if t["step_is_successful"]:
            assert t["relative_decrease"] > 1e-3
            assert t["trust_region_radius"] >= prev["trust_region_radius"] / 2 - 1e-9 or t["relative_decrease"] < 0.25
        elif t["step_is_valid"]:
            assert t["trust_region_radius"] < prev["trust_region_radius"]
    # the two Jacobian paths give the same solve
    cams0, pts0, s0, tr0 = ow.solve(w.cams_init, w.points_init, use_autodiff=0)
    assert len(tr0) == len(tr)
    np.testing.assert_allclose(cams0, cams, atol=1e-8)
    assert abs(s0["final_cost"] - s["final_cost"]) <= 1e-9 * s["final_cost"]
    # thread count does not change the answer beyond rounding
    cams1, _, s1, tr1 = binding.OracleWindow(w, num_threads=1).solve(w.cams_init, w.points_init)
    assert len(tr1) == len(tr)
    np.testing.assert_allclose(cams1, cams, atol=1e-9)


def test_scipy_second_opinion():
    """Independent optimiser (SciPy TRF) on a tiny window with the same robustified residuals
    (sqrt(rho'(s)) applied per block, Ceres corrector semantics) must not find a meaningfully
    lower cost than the oracle's LM, and must agree on the optimum it reaches from there."""
    from scipy.optimize import least_squares
    w = synthetic.small_window(seed=3, n_frames=4, grid=(5, 6), rows=96, cols=128)
    ow = binding.OracleWindow(w)
    cams, pts, s, tr = ow.solve(w.cams_init, w.points_init)

    F, n = w.n_frames, w.n_points

    def unpack(x):
        c = w.cams_init.copy()
        c[1:] = x[: 6 * (F - 1)].reshape(F - 1, 6)
        return c, x[6 * (F - 1):].reshape(n, 3)

    def cost(x):
        c, p = unpack(x)
        return ow.cost(c, p)

    x_or = np.concatenate([cams[1:].ravel(), pts.ravel()])
    assert abs(cost(x_or) - s["final_cost"]) <= 1e-9 * s["final_cost"]

    def fun(x):
        c, p = unpack(x)
        e = ow.evaluate(c, p, 0)
        r = e["residuals"].copy()
        a = w.huber
        ssq = e["obs_sqnorm"]
        rho = np.where(ssq <= a * a, ssq, 2 * a * np.sqrt(ssq) - a * a)
        # residual vector whose squared norm is sum rho(s): scale each block to norm sqrt(rho)
        scale = np.sqrt(rho / np.maximum(ssq, 1e-300))
        return (r * scale[:, None]).ravel()

    x0 = np.concatenate([w.cams_init[1:].ravel(), w.points_init.ravel()])
    res = least_squares(fun, x0, method="trf", x_scale="jac", max_nfev=60, xtol=1e-10, ftol=1e-10)
    # SciPy differentiates the piecewise, fp32-quantised objective numerically, so it may stop
    # early; what must hold: it does not find a meaningfully LOWER cost than the oracle's LM ...
    assert s["final_cost"] <= 1.02 * res.cost
    assert s["final_cost"] < 0.7 * s["initial_cost"]
    # ... and restarted from the oracle's optimum it cannot improve it by more than 1 %.
    res2 = least_squares(fun, x_or, method="trf", x_scale="jac", max_nfev=30, xtol=1e-10, ftol=1e-10)
    assert res2.cost >= 0.99 * s["final_cost"]


def test_camera_model_and_patch_weights_match_reference_binary_when_present():
    """The reference's own Calibration (src/calibration.h: project<T>, project(Vec3), pyrDown), ImageSize::pyrDown
    (src/types.h:70-73) and MakePatchWeights (src/photobundle.cc:617-646), compiled from where they lie into
    oracle/_ref/libref_calib.so, against the oracle's restatement AND the host side's own code (compat.h /
    photobundle.cc, through libpba_host.so): bit for bit."""
    ref = binding.ref_calib_lib()
    if ref is None:
        pytest.skip("oracle/_ref/libref_calib.so not built (no /root/reference on this box)")
    from photobundle_b200 import host_capi
    host = host_capi.lib()
    host.pbah_pyr_down.argtypes = [C.c_void_p, C.c_double, C.c_int32, C.c_int32, C.c_void_p]
    rng = np.random.default_rng(17)
    k4 = np.array([718.856, 718.856, 607.1928, 185.2157])                 # config/kitti_stereo.cfg's camera
    uv_ref, uv_ora, uv_vec, uv_host = (np.zeros(2) for _ in range(4))
    for _ in range(2000):
        X = rng.normal(size=3) * np.array([5.0, 2.0, 20.0])
        if rng.random() < 0.05:
            X[2] = rng.choice([1e-12, -1e-3, 1e3])
        ref.ref_project(C.c_void_p(k4.ctypes.data), C.c_void_p(X.ctypes.data), C.c_void_p(uv_ref.ctypes.data))
        binding.lib().oracle_project(C.c_void_p(k4.ctypes.data), C.c_void_p(X.ctypes.data), C.c_void_p(uv_ora.ctypes.data))
        assert uv_ref.tobytes() == uv_ora.tobytes()                       # the residual functor's projection
        ref.ref_project_vec3(C.c_void_p(k4.ctypes.data), C.c_void_p(X.ctypes.data), C.c_void_p(uv_vec.ctypes.data))
        host.pbah_project(C.c_void_p(k4.ctypes.data), C.c_void_p(X.ctypes.data), C.c_void_p(uv_host.ctypes.data))
        # addFrame's projection, normHomog(K X) = (1 / p_z) * p: the host divides instead, at most 1 ulp apart
        np.testing.assert_allclose(uv_host, uv_vec, rtol=4e-16, atol=0)
    o_ref, o_host = np.zeros(7), np.zeros(7)
    for rows, cols in ((376, 1241), (375, 1242), (188, 621), (1, 1), (47, 156)):
        ref.ref_pyr_down(C.c_void_p(k4.ctypes.data), 0.5371657, rows, cols, C.c_void_p(o_ref.ctypes.data))
        host.pbah_pyr_down(C.c_void_p(k4.ctypes.data), 0.5371657, rows, cols, C.c_void_p(o_host.ctypes.data))
        assert o_ref.tobytes() == o_host.tobytes()
        assert list(o_ref[5:]) == [(rows + 1) // 2, (cols + 1) // 2] and o_ref[0] == 0.5 * k4[0] and o_ref[4] == 2 * 0.5371657
    for radius in (0, 1, 2, 3, 4):
        for gauss in (0, 1):
            n = (2 * radius + 1) ** 2
            w_ref, w_ora, w_host = np.zeros(n), np.zeros(n), np.zeros(n)
            assert ref.ref_patch_weights(radius, gauss, C.c_void_p(w_ref.ctypes.data)) == n
            binding.lib().oracle_patch_weights(radius, gauss, C.c_void_p(w_ora.ctypes.data))
            assert host.pbah_patch_weights(radius, gauss, C.c_void_p(w_host.ctypes.data)) == n
            assert w_ref.tobytes() == w_ora.tobytes() == w_host.tobytes()


def test_oracle_is_deterministic_for_a_fixed_thread_count():
    """Per-thread partial sums are combined in thread order (no OpenMP reduction clauses on the path): two solves with
    the same thread count give the same bits, whatever the scheduling."""
    from workloads import synthetic
    win = synthetic.small_window(seed=11, ragged=True, n_frames=8, grid=(10, 14))
    for nt in (1, 3, 4):
        runs = []
        for _ in range(3):
            ow = binding.OracleWindow(win, num_threads=nt)
            cams, pts, summ, tr = ow.solve(win.cams_init, win.points_init)
            runs.append((cams.tobytes(), pts.tobytes(), summ["final_cost"], len(tr)))
        assert runs[0] == runs[1] == runs[2], nt


def test_imgradient_and_disparity_match_reference_binary_when_present():
    """The reference's own imgradient (src/imgproc.cc:26-106) and disparityToDepth (:274-322), compiled from where they
    lie into oracle/_ref/libref_imgproc.so.  The gradient the oracle restates - and K_A forms on the fly from uint8
    footprints - equals it bit for bit (uint8 and float sources, odd sizes).  disparityToDepth: the host side's version
    uses an exact reciprocal and one invalid mark where the reference's SSE body uses _mm_rcp_ps (~12 bits) and its scalar
    tail another mark: same valid set, values within the reciprocal's error, invalid pixels negative in both."""
    ref = binding.ref_imgproc_lib()
    if ref is None:
        pytest.skip("oracle/_ref/libref_imgproc.so not built (no /root/reference on this box)")
    rng = np.random.default_rng(23)
    for rows, cols in ((9, 13), (376, 1241), (31, 64)):
        I8 = rng.integers(0, 256, size=(rows, cols), dtype=np.uint8)
        If = I8.astype(np.float32)
        gx_r, gy_r, gx_o, gy_o = (np.full((rows, cols), 7.0, dtype=np.float32) for _ in range(4))
        ref.ref_imgradient_u8(C.c_void_p(I8.ctypes.data), rows, cols, C.c_void_p(gx_r.ctypes.data), C.c_void_p(gy_r.ctypes.data))
        binding.lib().oracle_imgradient(C.c_void_p(If.ctypes.data), rows, cols, C.c_void_p(gx_o.ctypes.data), C.c_void_p(gy_o.ctypes.data))
        assert gx_r.tobytes() == gx_o.tobytes() and gy_r.tobytes() == gy_o.tobytes()
        Fn = rng.uniform(0, 255, size=(rows, cols)).astype(np.float32)      # non-integer channel values (descriptor planes)
        ref.ref_imgradient_f32(C.c_void_p(Fn.ctypes.data), rows, cols, C.c_void_p(gx_r.ctypes.data), C.c_void_p(gy_r.ctypes.data))
        binding.lib().oracle_imgradient(C.c_void_p(Fn.ctypes.data), rows, cols, C.c_void_p(gx_o.ctypes.data), C.c_void_p(gy_o.ctypes.data))
        assert gx_r.tobytes() == gx_o.tobytes() and gy_r.tobytes() == gy_o.tobytes()
    from photobundle_b200 import host_capi
    for rows, cols in ((376, 1241), (16, 16), (5, 7)):
        d = rng.uniform(-1.0, 90.0, size=(rows, cols)).astype(np.float32)
        d[rng.random(d.shape) < 0.1] = 0.0
        z_ref = np.zeros_like(d)
        ref.ref_disparity_to_depth(C.c_void_p(d.ctypes.data), rows, cols, C.c_float(386.1), C.c_void_p(z_ref.ctypes.data))
        z = host_capi.disparity_to_depth(d, 386.1)
        valid = d > 0.01
        assert np.array_equal(z_ref > 0, valid) and np.array_equal(z > 0, valid)
        np.testing.assert_allclose(z[valid], z_ref[valid], rtol=4e-4)
        assert (z_ref[~valid] < 0).all() and (z[~valid] == np.float32(-0.1)).all()


def test_association_lookup_matches_reference_binary_when_present():
    """interp2 / interpolateFixedPatch<2> of the reference (src/photobundle.cc:258-310, compiled from where they lie into
    oracle/_ref/libref_calib.so): the 5x5 bilinear lookup ZnccPatch_::set starts from equals the host's (and through
    tests/test_host.py the device's) bit for bit, inside the image, on its last row / column and outside."""
    ref = binding.ref_calib_lib()
    if ref is None or not hasattr(ref, "ref_interp_patch5_u8"):
        pytest.skip("oracle/_ref/libref_calib.so not built (no /root/reference on this box)")
    from photobundle_b200 import host_capi
    host = host_capi.lib()
    ref.ref_interp_patch5_u8.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_void_p]
    host.pbah_interp_patch5_u8.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_void_p]
    rng = np.random.default_rng(29)
    rows, cols = 23, 37
    I = rng.integers(0, 256, size=(rows, cols), dtype=np.uint8)
    a, b = np.zeros(25, dtype=np.float32), np.zeros(25, dtype=np.float32)
    pts = [(rng.uniform(-4, cols + 3), rng.uniform(-4, rows + 3)) for _ in range(3000)]
    pts += [(cols - 3.0, 5.25), (7.5, rows - 3.0), (cols - 3.0, rows - 3.0), (cols - 1.0, rows - 1.0), (2.0, 2.0), (0.0, 0.0), (-0.5, 3.0)]
    for u, v in pts:
        ref.ref_interp_patch5_u8(C.c_void_p(I.ctypes.data), rows, cols, u, v, C.c_void_p(a.ctypes.data))
        host.pbah_interp_patch5_u8(C.c_void_p(I.ctypes.data), rows, cols, u, v, C.c_void_p(b.ctypes.data))
        assert a.tobytes() == b.tobytes(), (u, v, a, b)


def test_local_maximum_rule_matches_reference_binary_when_present():
    """IsLocalMax_ (src/imgproc.h:175-212, compiled from where it lies) against the restatement the device's
    pba_select_candidates is tested with (tests/ref_addframe.py::local_maxima): ties lose, masked pixels lose, radius 0
    accepts everything."""
    ref = binding.ref_imgproc_lib()
    if ref is None or not hasattr(ref, "ref_local_maxima"):
        pytest.skip("oracle/_ref/libref_imgproc.so not built (no /root/reference on this box)")
    from ref_addframe import local_maxima        # (tests/ is on sys.path: conftest lives there)
    rng = np.random.default_rng(31)
    for radius in (0, 1, 2):
        rows, cols = 41, 57
        S = rng.integers(0, 6, size=(rows, cols)).astype(np.float32)          # few levels: plenty of ties
        S[rng.random(S.shape) < 0.05] *= 7.5
        mask = (rng.random(S.shape) > 0.2).astype(np.uint8)
        out = np.zeros((rows, cols), dtype=np.uint8)
        ref.ref_local_maxima(C.c_void_p(S.ctypes.data), C.c_void_p(mask.ctypes.data), rows, cols, radius, C.c_void_p(out.ctypes.data))
        mine = local_maxima(S, mask.astype(bool), radius)
        inner = (slice(radius, rows - radius), slice(radius, cols - radius))
        assert np.array_equal(out[inner].astype(bool), mine[inner])
        assert out[inner].sum() > (20 if radius else 100)


def test_extract_patch_matches_reference_binary_when_present():
    """ExtractPatch (src/photobundle.cc:466-479, compiled from where it lies): the reference descriptor of a new point -
    integer pixel, clamped so that the whole patch stays inside the image - against the oracle's restatement (which the
    device's pba_extract_descriptors equals bit for bit, tests/test_descriptors.py)."""
    ref = binding.ref_calib_lib()
    if ref is None or not hasattr(ref, "ref_extract_patch_f32"):
        pytest.skip("oracle/_ref/libref_calib.so not built (no /root/reference on this box)")
    rng = np.random.default_rng(37)
    rows, cols = 19, 27
    plane = rng.uniform(0, 255, size=(rows, cols)).astype(np.float32)
    for radius in (1, 2, 3):
        P = (2 * radius + 1) ** 2
        xy = np.stack([rng.integers(-3, cols + 3, size=200), rng.integers(-3, rows + 3, size=200)], axis=1).astype(np.int32)
        xy = np.concatenate([xy, np.array([[0, 0], [cols - 1, rows - 1], [radius, radius], [cols - radius - 1, rows - radius - 1]], dtype=np.int32)])
        mine = np.zeros((xy.shape[0], P))
        binding.lib().oracle_extract_patches(C.c_void_p(plane.ctypes.data), 1, rows, cols, radius, xy.shape[0], C.c_void_p(xy.ctypes.data),
                                             C.c_void_p(mine.ctypes.data))
        theirs = np.zeros(P)
        for k, (x, y) in enumerate(xy):
            ref.ref_extract_patch_f32(C.c_void_p(plane.ctypes.data), rows, cols, int(x), int(y), radius, C.c_void_p(theirs.ctypes.data))
            assert theirs.tobytes() == mine[k].tobytes(), (radius, x, y)


def test_residual_functor_matches_reference_binary_when_present():
    """The reference's own residual functor - the body of DescriptorError::operator()<double> (src/photobundle.cc:696-727)
    over its own SampleWithDerivative and Calibration::project, compiled from where they lie into
    oracle/_ref/libref_functor.so (only ceres::AngleAxisRotatePoint is restated: Ceres is absent) - against the oracle's
    residual block on every observation of a ragged window, interior and border, 1 and 3 channels: bit for bit."""
    ref = binding.ref_functor_lib()
    if ref is None:
        pytest.skip("oracle/_ref/libref_functor.so not built (no /root/reference on this box)")
    for win, planes in ((synthetic.small_window(seed=11, ragged=True, n_frames=8, grid=(10, 14)), None),
                        (synthetic.small_window(seed=3, radius=1, ragged=True, n_frames=6), None)):
        ow = binding.OracleWindow(win, num_threads=1, planes=planes)
        F, Cn, rows, cols = ow.planes.shape
        k4 = np.array([win.fx, win.fy, win.cx, win.cy])
        rng = np.random.default_rng(41)
        n_checked = 0
        for p in range(win.n_points):
            for o in range(int(win.obs_offsets[p]), int(win.obs_offsets[p + 1])):
                f = int(win.obs_frame[o])
                cam = win.cams_init[f] + (rng.normal(size=6) * 1e-3 if o % 3 else 0.0)
                X = win.points_init[p] * (1.0 if o % 5 else 1.6)          # some projections land near / beyond the border
                r_o, _, _ = ow.residual_block(f, cam, X, ow.desc[p], mode=2)
                r_r = np.zeros(ow.CP)
                cam = np.ascontiguousarray(cam, dtype=np.float64); X = np.ascontiguousarray(X, dtype=np.float64)
                ok = ref.ref_residual_block(C.c_void_p(ow.planes[f].ctypes.data), C.c_void_p(ow.gx[f].ctypes.data), C.c_void_p(ow.gy[f].ctypes.data),
                                            Cn, rows, cols, C.c_void_p(k4.ctypes.data), win.radius, C.c_void_p(ow.desc[p].ctypes.data),
                                            C.c_void_p(ow.weights.ctypes.data), C.c_void_p(cam.ctypes.data), C.c_void_p(X.ctypes.data),
                                            C.c_void_p(r_r.ctypes.data))
                assert ok == 1 and r_r.tobytes() == r_o.tobytes(), (p, o, np.abs(r_r - r_o).max())
                n_checked += 1
        assert n_checked > 300


def test_residual_jacobian_matches_reference_functor_with_jets_when_present():
    """The same reference functor body instantiated with ceres::Jet<double, 9> as AutoDiffCostFunction<.., 6, 3> seeds it
    (reference sampler chain rule src/jet_extras.h; Jet algebra and AngleAxisRotatePoint restated from Ceres' published
    definitions) against the oracle's autodiff residual block: residuals and all nine Jacobian columns bit for bit."""
    ref = binding.ref_functor_lib()
    if ref is None:
        pytest.skip("oracle/_ref/libref_functor.so not built (no /root/reference on this box)")
    win = synthetic.small_window(seed=11, ragged=True, n_frames=8, grid=(10, 14))
    ow = binding.OracleWindow(win, num_threads=1)
    F, Cn, rows, cols = ow.planes.shape
    k4 = np.array([win.fx, win.fy, win.cx, win.cy])
    rng = np.random.default_rng(43)
    n_checked = 0
    for p in range(0, win.n_points, 2):
        for o in range(int(win.obs_offsets[p]), int(win.obs_offsets[p + 1])):
            f = int(win.obs_frame[o])
            cam = np.ascontiguousarray(win.cams_init[f] + rng.normal(size=6) * 1e-3, dtype=np.float64)
            X = np.ascontiguousarray(win.points_init[p] * (1.0 if o % 7 else 1.5), dtype=np.float64)
            r_o, Jc, Jp = ow.residual_block(f, cam, X, ow.desc[p], mode=1)
            r_r, J = np.zeros(ow.CP), np.zeros((ow.CP, 9))
            ok = ref.ref_residual_block_jet(C.c_void_p(ow.planes[f].ctypes.data), C.c_void_p(ow.gx[f].ctypes.data), C.c_void_p(ow.gy[f].ctypes.data),
                                            Cn, rows, cols, C.c_void_p(k4.ctypes.data), win.radius, C.c_void_p(ow.desc[p].ctypes.data),
                                            C.c_void_p(ow.weights.ctypes.data), C.c_void_p(cam.ctypes.data), C.c_void_p(X.ctypes.data),
                                            C.c_void_p(r_r.ctypes.data), C.c_void_p(J.ctypes.data))
            assert ok == 1 and r_r.tobytes() == r_o.tobytes()
            assert np.ascontiguousarray(J[:, :6]).tobytes() == Jc.tobytes() and np.ascontiguousarray(J[:, 6:]).tobytes() == Jp.tobytes(), \
                (p, o, np.abs(J[:, :6] - Jc).max(), np.abs(J[:, 6:] - Jp).max())
            n_checked += 1
    assert n_checked > 150
