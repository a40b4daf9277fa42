"""Multi-channel descriptors: DescriptorFrame::Create (src/photobundle.cc:220-248),
computeSaliencyMap (:212-220) and ExtractPatch (:466-479).
  * CPU: the oracle restatement against the OpenCV-generated golden vector (tests/golden/bitplanes_ref.npz,
    made by tests/golden/make_bitplanes_golden.py) and against the reference's definitions;
  * GPU: the device construction (k_prep.cu through the C ABI) against the oracle, bit for bit, and the
    8-channel BitPlanes window through K_A against the oracle evaluation."""
import ctypes as C
import dataclasses
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import binding as ob
from photobundle_b200 import capi
from workloads import synthetic
def _golden():
    return np.load(os.path.join(GOLDEN, "bitplanes_ref.npz"))


# ------------------------------------------------------------------------------------------ CPU
def test_bitplanes_oracle_matches_opencv_golden():
    g = _golden()
    img = g["image"]
    blur, census = np.zeros_like(img), np.zeros_like(img)
    ob.lib().oracle_bitplanes_stages(C.c_void_p(img.ctypes.data), img.shape[0], img.shape[1],
                                     C.c_void_p(blur.ctypes.data), C.c_void_p(census.ctypes.data))
    assert np.array_equal(blur, g["blur"])          # cv::GaussianBlur(uint8, 3x3, sigma 1): bit-exact
    assert np.array_equal(census, g["census"])      # census transform: bit-exact
    planes = ob.build_channels(img, "bitplanes")
    assert planes.shape == (8,) + img.shape
    # cv::GaussianBlur(float, 5x5, sigma 1.5): OpenCV's SIMD filter uses FMA, the restatement does not
    assert np.abs(planes - g["planes"]).max() <= 5e-7
    assert planes.min() >= 0.0 and planes.max() <= 1.0 + 1e-6


def test_census_definition_and_borders():
    g = _golden()
    blur, census = g["blur"], g["census"]
    assert not census[0].any() and not census[-1].any() and not census[:, 0].any() and not census[:, -1].any()
    rng = np.random.default_rng(0)
    for _ in range(200):     # bit k = neighbour k >= centre, neighbours row-major without the centre (imgproc.cc:140-152)
        y, x = int(rng.integers(1, blur.shape[0] - 1)), int(rng.integers(1, blur.shape[1] - 1))
        nb = [blur[y - 1, x - 1], blur[y - 1, x], blur[y - 1, x + 1], blur[y, x - 1], blur[y, x + 1], blur[y + 1, x - 1],
              blur[y + 1, x], blur[y + 1, x + 1]]
        assert census[y, x] == sum((1 << k) for k, v in enumerate(nb) if v >= blur[y, x])


def test_intensity_and_gradient_channels_oracle():
    img = _golden()["image"]
    p = ob.build_channels(img, "intensity_and_gradient")
    assert np.array_equal(p[0], img.astype(np.float32))
    gx, gy = np.zeros_like(p[0]), np.zeros_like(p[0])
    ob.lib().oracle_imgradient(C.c_void_p(p[0].ctypes.data), img.shape[0], img.shape[1], C.c_void_p(gx.ctypes.data), C.c_void_p(gy.ctypes.data))
    assert np.array_equal(p[1], gx) and np.array_equal(p[2], gy)     # imgradient(uint8) == imgradient(float(uint8))
    assert not p[1][0].any() and not p[1][:, 0].any() and not p[2][-1].any() and not p[2][:, -1].any()
    assert ob.build_channels(img, "intensity").shape == (1,) + img.shape


def test_saliency_and_extract_patch_oracle():
    img = _golden()["image"]
    planes = ob.build_channels(img, "intensity_and_gradient")
    sal = ob.saliency_map(planes)
    y, x = 7, 9
    want = 0.0
    for k in range(3):
        want += abs(0.5 * (planes[k, y, x + 1] - planes[k, y, x - 1])) + abs(0.5 * (planes[k, y + 1, x] - planes[k, y - 1, x]))
    assert abs(sal[y, x] - want) <= 1e-4 and sal[0, 5] == 0.0
    xy = np.array([[9, 7], [0, 0], [img.shape[1] - 1, img.shape[0] - 1]], dtype=np.int32)
    d = ob.extract_patches(planes, xy, 2)
    assert d.shape == (3, 75)
    for i, (px, py) in enumerate(xy):
        for k in range(3):
            assert np.array_equal(d[i, 25 * k:25 * (k + 1)], synthetic.extract_patch(planes[k], int(px), int(py), 2))


# ------------------------------------------------------------------------------------------ GPU
def _handle(rows, cols, n_channels, **kw):
    return capi.Handle(rows, cols, 400.0, 400.0, cols / 2.0, rows / 2.0, radius=2, n_channels=n_channels, huber=0.05,
                       max_frames=kw.get("max_frames", 2), max_points=64, max_observations=256)


@pytest.mark.gpu
@pytest.mark.parametrize("descriptor", ["intensity_and_gradient", "bitplanes"])
def test_device_channels_equal_oracle(descriptor):
    img = _golden()["image"]
    img2 = np.ascontiguousarray(img[::-1, ::-1])
    Cn = {"intensity_and_gradient": 3, "bitplanes": 8}[descriptor]
    h = _handle(img.shape[0], img.shape[1], Cn)
    h.set_frames_u8_descriptor(np.stack([img, img2]), descriptor)
    for f, im in enumerate((img, img2)):
        want = ob.build_channels(im, descriptor)
        for k in range(Cn):
            assert np.array_equal(h.get_channel_plane(f, k), want[k]), (descriptor, f, k)
    if descriptor == "bitplanes":
        assert np.abs(h.get_channel_plane(0, 3) - _golden()["planes"][3]).max() <= 5e-7     # and OpenCV's own output
    with pytest.raises(capi.PbaError):
        h.set_frames_u8_descriptor(np.stack([img]), "intensity")       # a 1-channel descriptor on a C-channel handle
    h.close()


@pytest.mark.gpu
@pytest.mark.parametrize("descriptor", ["intensity", "intensity_and_gradient", "bitplanes"])
def test_device_saliency_and_descriptors_equal_oracle(descriptor):
    img = _golden()["image"]
    Cn = {"intensity": 1, "intensity_and_gradient": 3, "bitplanes": 8}[descriptor]
    h = _handle(img.shape[0], img.shape[1], Cn)
    h.prepare_frame_u8(img, descriptor)
    planes = ob.build_channels(img, descriptor)
    assert np.array_equal(h.saliency_map(), ob.saliency_map(planes))
    rng = np.random.default_rng(3)
    xy = np.stack([rng.integers(-3, img.shape[1] + 3, 40), rng.integers(-3, img.shape[0] + 3, 40)], axis=1).astype(np.int32)
    assert np.array_equal(h.extract_descriptors(xy), ob.extract_patches(planes, xy, 2))
    assert h.extract_descriptors(np.zeros((0, 2), dtype=np.int32)).shape == (0, Cn * 25)
    h.close()


@pytest.mark.gpu
def test_k1_bitplanes_window_device_built(small_win):
    """8-channel BitPlanes window: planes built on the device from the uint8 frames, reference descriptors
    extracted on the device, evaluated by K_A; the oracle evaluates the same window on its own planes."""
    from test_gpu_parity import _first_px
    w = small_win
    planes = np.stack([ob.build_channels(w.images[f], "bitplanes") for f in range(w.n_frames)])
    h = capi.Handle(w.rows, w.cols, w.fx, w.fy, w.cx, w.cy, radius=w.radius, n_channels=8, huber=w.huber,
                    max_frames=w.n_frames, max_points=w.n_points, max_observations=w.n_obs)
    desc = np.zeros((w.n_points, 8 * 25))
    ref_frame = np.array([int(w.obs_frame[w.obs_offsets[p]]) for p in range(w.n_points)])
    px = np.array([_first_px(w, p) for p in range(w.n_points)], dtype=np.int32)
    for f in np.unique(ref_frame):
        h.prepare_frame_u8(w.images[f], "bitplanes")
        sel = np.nonzero(ref_frame == f)[0]
        desc[sel] = h.extract_descriptors(px[sel])
        assert np.array_equal(desc[sel], ob.extract_patches(planes[f], px[sel], w.radius))
    w8 = dataclasses.replace(w, desc=desc, n_channels=8)
    h.set_frames_u8_descriptor(w.images, "bitplanes")
    h.set_poses(w8.cams_init, w8.fixed_frame)
    h.set_points(w8.points_init, w8.desc, w8.obs_offsets, w8.obs_frame, w8.weights)
    ev = h.eval()
    ref = ob.OracleWindow(w8, planes=planes).evaluate(w8.cams_init, w8.points_init, 1)
    d = np.abs(ev["residuals"] - ref["residuals"])
    assert (d == 0).mean() >= 0.999 and d.max() <= 1e-4
    assert abs(ev["cost"] - ref["cost"]) <= 1e-7 * ref["cost"]
    for k in ("U", "gc", "V", "gp", "W"):
        assert np.abs(ev[k] - ref[k]).max() <= 1e-5 * np.abs(ref[k]).max(), k
    s = h.solve(max_num_iterations=5)
    assert s["final_cost"] < s["initial_cost"]
    h.close()
