"""Synthetic sliding-window generator (SURVEY.md §8d recipe).

The reference ships no image data (KITTI frames are not on disk) and no expected
outputs, so every parity test and the bench use this deterministic recipe: a
KITTI-sized pinhole camera looking at one textured plane, rendered exactly by
ray-plane intersection, with points lifted from frame pixels through noisy depth the
way PhotometricBundleAdjustment::addFrame does it (/root/reference/src/photobundle.cc:
548-573: X = T_w * (z * K^-1 * [x y 1]^T); descriptor = integer-pixel patch of the
reference frame, ExtractPatch :466-479).

Random numbers: numpy Generator(PCG64) seeded from SeedSequence(master_seed).spawn(4)
-> (texture, point grid jitter, depth noise, pose perturbation).  Master seed of the
bench/golden windows: 20161201.
"""
from __future__ import annotations

import dataclasses

import numpy as np

MASTER_SEED = 20161201

KITTI_ROWS, KITTI_COLS = 376, 1241
KITTI_FX = KITTI_FY = 718.856
KITTI_CX, KITTI_CY = 607.1928, 185.2157


def rodrigues(w: np.ndarray) -> np.ndarray:
    """ceres::AngleAxisToRotationMatrix (first-order form for tiny angles)."""
    w = np.asarray(w, dtype=np.float64)
    th2 = float(w @ w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)
    if th2 > np.finfo(np.float64).eps:
        th = np.sqrt(th2)
        k = K / th
        return np.eye(3) + np.sin(th) * k + (1.0 - np.cos(th)) * (k @ k)
    return np.eye(3) + K


@dataclasses.dataclass
class Window:
    """Everything optimize() holds when it builds the problem (photobundle.cc:764-806)."""

    images: np.ndarray        # [F, rows, cols] uint8
    fx: float
    fy: float
    cx: float
    cy: float
    radius: int
    huber: float
    weights: np.ndarray       # [P] float64 (MakePatchWeights)
    cams_gt: np.ndarray       # [F, 6] world->camera [w, t]
    cams_init: np.ndarray     # [F, 6]
    points_gt: np.ndarray     # [n, 3]
    points_init: np.ndarray   # [n, 3]
    desc: np.ndarray          # [n, C*P] float64
    obs_offsets: np.ndarray   # [n+1] int32
    obs_frame: np.ndarray     # [nnz] int32 window-local frame index
    fixed_frame: int = 0
    n_channels: int = 1

    @property
    def n_frames(self) -> int:
        return int(self.images.shape[0])

    @property
    def rows(self) -> int:
        return int(self.images.shape[1])

    @property
    def cols(self) -> int:
        return int(self.images.shape[2])

    @property
    def n_points(self) -> int:
        return int(self.points_init.shape[0])

    @property
    def n_obs(self) -> int:
        return int(self.obs_frame.shape[0])

    @property
    def patch_len(self) -> int:
        return (2 * self.radius + 1) ** 2

    @property
    def n_residuals(self) -> int:
        return self.n_obs * self.patch_len * self.n_channels

    def planes_f32(self) -> np.ndarray:
        """DescriptorFrame channels for the Intensity descriptor: uint8 -> float cast
        (photobundle.cc:231).  Shape [F, C=1, rows, cols]."""
        return np.ascontiguousarray(self.images.astype(np.float32)[:, None, :, :])


class PlaneScene:
    """One textured plane n.X = d; texture = sum of 32 cosines on plane coordinates."""

    def __init__(self, rng: np.random.Generator, fscale: float = 1.0):
        # fscale scales the spatial-frequency band with the focal length so that a small
        # test image sees the same texture period in pixels as the KITTI-sized one.
        n = np.array([0.10, -0.15, -1.0])
        self.n = n / np.linalg.norm(n)
        self.d = 15.0 * self.n[2]
        e1 = np.cross(self.n, np.array([0.0, 1.0, 0.0]))
        self.e1 = e1 / np.linalg.norm(e1)
        self.e2 = np.cross(self.n, self.e1)
        K = 32
        fmag = fscale * np.exp(rng.uniform(np.log(0.15), np.log(3.0), size=K))
        fdir = rng.uniform(0.0, 2.0 * np.pi, size=K)
        self.freq = np.stack([fmag * np.cos(fdir), fmag * np.sin(fdir)], axis=1)  # cycles/m
        self.phase = rng.uniform(0.0, 2.0 * np.pi, size=K)
        amp = 1.0 / fmag
        amp *= 40.0 / np.sqrt(0.5 * np.sum(amp * amp))  # std(T) = 40
        self.amp = amp

    def texture(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        t = np.full(a.shape, 127.5, dtype=np.float64)
        for k in range(self.amp.shape[0]):
            t += self.amp[k] * np.cos(2.0 * np.pi * (self.freq[k, 0] * a + self.freq[k, 1] * b) + self.phase[k])
        return t

    def intersect(self, cam: np.ndarray, K4, xs: np.ndarray, ys: np.ndarray):
        """Ray-plane intersection for pixels (xs, ys) of the camera with world->camera
        parameters cam=[w,t].  Returns (world points [...,3], depth z in that camera)."""
        fx, fy, cx, cy = K4
        R = rodrigues(cam[:3])
        C = -R.T @ cam[3:]
        dc = np.stack([(xs - cx) / fx, (ys - cy) / fy, np.ones_like(xs, dtype=np.float64)], axis=-1)
        dw = dc @ R  # R^T d for each ray (row-vector form)
        lam = (self.d - self.n @ C) / (dw @ self.n)
        Xw = C + lam[..., None] * dw
        return Xw, lam  # camera-frame z equals lam because dc_z = 1

    def render(self, cam: np.ndarray, K4, rows: int, cols: int) -> np.ndarray:
        ys, xs = np.meshgrid(np.arange(rows, dtype=np.float64), np.arange(cols, dtype=np.float64), indexing="ij")
        Xw, _ = self.intersect(cam, K4, xs, ys)
        a = Xw @ self.e1
        b = Xw @ self.e2
        t = self.texture(a, b)
        return np.clip(np.rint(t), 0, 255).astype(np.uint8)


def extract_patch(img: np.ndarray, x: int, y: int, radius: int) -> np.ndarray:
    """ExtractPatch, photobundle.cc:466-479 (integer pixel, clamped to [r, size-r-1])."""
    rows, cols = img.shape
    max_cols, max_rows = cols - radius - 1, rows - radius - 1
    out = np.empty((2 * radius + 1) ** 2, dtype=np.float64)
    i = 0
    for r in range(-radius, radius + 1):
        ri = max(radius, min(y + r, max_rows))
        for c in range(-radius, radius + 1):
            ci = max(radius, min(x + c, max_cols))
            out[i] = float(img[ri, ci])
            i += 1
    return out


def patch_weights(radius: int, gaussian: bool = False) -> np.ndarray:
    """MakePatchWeights, photobundle.cc:617-644 (s_x = s_y = a = 1)."""
    n = (2 * radius + 1) ** 2
    if not gaussian:
        return np.ones(n, dtype=np.float64)
    w = np.empty(n, dtype=np.float64)
    i = 0
    for r in range(-radius, radius + 1):
        for c in range(-radius, radius + 1):
            w[i] = np.exp(-0.5 * (r * r / 1.0 + c * c / 1.0))
            i += 1
    return w / w.sum()


def make_window(
    n_frames: int = 8,
    grid: tuple[int, int] = (50, 80),
    rows: int = KITTI_ROWS,
    cols: int = KITTI_COLS,
    intrinsics: tuple[float, float, float, float] | None = None,
    radius: int = 2,
    huber: float = 0.05,
    ragged: bool = False,
    margin: int = 48,
    seed: int = MASTER_SEED,
    gaussian_weights: bool = False,
    images: np.ndarray | None = None,
) -> Window:
    """cfg2/3: make_window() -> 8 frames x 4000 points x 5x5 (32 000 observations).
    cfg4 finest level: make_window(16, (100, 160)).
    `images` lets a caller reuse already rendered frames (rendering dominates the cost)."""
    if intrinsics is None:
        sx, sy = cols / KITTI_COLS, rows / KITTI_ROWS
        intrinsics = (KITTI_FX * sx, KITTI_FY * sy, KITTI_CX * sx, KITTI_CY * sy)
    K4 = tuple(float(v) for v in intrinsics)
    ss = np.random.SeedSequence(seed).spawn(4)
    rng_tex, rng_grid, rng_depth, rng_pose = (np.random.Generator(np.random.PCG64(s)) for s in ss)
    scene = PlaneScene(rng_tex, fscale=min(K4[0], K4[1]) / KITTI_FX)

    F = n_frames
    cams_gt = np.zeros((F, 6), dtype=np.float64)
    for i in range(F):
        w = np.array([0.0, 0.002 * i, 0.0])
        C = i * np.array([0.03, 0.0, 0.10])
        cams_gt[i, :3] = w
        cams_gt[i, 3:] = -rodrigues(w) @ C
    cams_init = cams_gt.copy()
    cams_init[1:, :3] += rng_pose.normal(0.0, 2e-3, size=(F - 1, 3))
    cams_init[1:, 3:] += rng_pose.normal(0.0, 0.02, size=(F - 1, 3))

    if images is None:
        images = np.stack([scene.render(cams_gt[i], K4, rows, cols) for i in range(F)])
    assert images.shape == (F, rows, cols) and images.dtype == np.uint8

    gr, gc = grid
    n = gr * gc
    jx = rng_grid.uniform(0.0, 1.0, size=(gr, gc))
    jy = rng_grid.uniform(0.0, 1.0, size=(gr, gc))
    jj, ii = np.meshgrid(np.arange(gc), np.arange(gr))
    px = np.floor(margin + (jj + jx) * (cols - 2 * margin) / gc).astype(np.int64).ravel()
    py = np.floor(margin + (ii + jy) * (rows - 2 * margin) / gr).astype(np.int64).ravel()
    depth_noise = rng_depth.normal(0.0, 1.0, size=n)

    if ragged:
        ref = rng_grid.integers(0, max(1, F - 2), size=n)
        run = rng_grid.integers(3, F + 1, size=n)
    else:
        ref = np.zeros(n, dtype=np.int64)
        run = np.full(n, F, dtype=np.int64)
    last = np.minimum(ref + run, F)  # exclusive

    points_gt = np.empty((n, 3))
    points_init = np.empty((n, 3))
    desc = np.empty((n, (2 * radius + 1) ** 2))
    fx, fy, cx, cy = K4
    for f in np.unique(ref):
        sel = np.nonzero(ref == f)[0]
        Xw, z = scene.intersect(cams_gt[f], K4, px[sel].astype(np.float64), py[sel].astype(np.float64))
        points_gt[sel] = Xw
        zn = z * (1.0 + 0.02 * depth_noise[sel])
        ray = np.stack([(px[sel] - cx) / fx, (py[sel] - cy) / fy, np.ones(sel.shape[0])], axis=-1)
        Xc = zn[:, None] * ray
        # lift with the frame's *initial* world pose, as addFrame does (photobundle.cc:560)
        R = rodrigues(cams_init[f, :3])
        points_init[sel] = (Xc - cams_init[f, 3:]) @ R  # R^T (Xc - t)
        for k in sel:
            desc[k] = extract_patch(images[f], int(px[k]), int(py[k]), radius)

    counts = (last - ref).astype(np.int32)
    obs_offsets = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(counts, out=obs_offsets[1:])
    obs_frame = np.concatenate([np.arange(ref[k], last[k], dtype=np.int32) for k in range(n)])

    return Window(
        images=images, fx=fx, fy=fy, cx=cx, cy=cy, radius=radius, huber=huber,
        weights=patch_weights(radius, gaussian_weights), cams_gt=cams_gt, cams_init=cams_init,
        points_gt=points_gt, points_init=points_init, desc=desc,
        obs_offsets=obs_offsets, obs_frame=obs_frame.astype(np.int32), fixed_frame=0,
    )


def small_window(seed: int = 7, ragged: bool = False, radius: int = 2, n_frames: int = 5,
                 grid: tuple[int, int] = (12, 16), rows: int = 120, cols: int = 160) -> Window:
    """A seconds-scale window for CPU tests and golden fixtures (same recipe, small image;
    intrinsics scaled with the image so the scene geometry is unchanged)."""
    return make_window(n_frames=n_frames, grid=grid, rows=rows, cols=cols, radius=radius,
                       intrinsics=(200.0, 200.0, 0.5 * cols - 0.3, 0.5 * rows + 0.2),
                       ragged=ragged, margin=16, seed=seed)


# ----------------------------------------------------------------------------------------------
# Synthetic *sequence* for the sliding-window entry point (addFrame): images, depth maps and the
# frame-to-frame pose initialisation the reference's apps/run_kitti.cc feeds it (run_kitti.cc:39-51).
# ----------------------------------------------------------------------------------------------
@dataclasses.dataclass
class Sequence:
    images: np.ndarray      # [N, rows, cols] uint8
    depths: np.ndarray      # [N, rows, cols] float32 (stereo-like noisy depth; <= 0 = invalid)
    K4: tuple               # fx, fy, cx, cy
    T_w_gt: np.ndarray      # [N, 4, 4] camera-to-world (the convention of Trajectory)
    T_rel_gt: np.ndarray    # [N, 4, 4] frame-to-frame poses whose chaining gives T_w_gt
    T_rel_init: np.ndarray  # [N, 4, 4] perturbed initialisation (what a VO front-end would give)


def cam_to_mat(cam: np.ndarray) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = rodrigues(cam[:3])
    T[:3, 3] = cam[3:]
    return T


def make_sequence(n_frames: int = 10, rows: int = 120, cols: int = 160, seed: int = 21,
                  intrinsics=None, rot_sigma: float = 1e-3, trans_sigma: float = 0.01,
                  depth_noise: float = 0.02) -> Sequence:
    if intrinsics is None:
        intrinsics = (200.0, 200.0, 0.5 * cols - 0.3, 0.5 * rows + 0.2)
    K4 = tuple(float(v) for v in intrinsics)
    ss = np.random.SeedSequence(seed).spawn(3)
    rng_tex, rng_depth, rng_pose = (np.random.Generator(np.random.PCG64(s)) for s in ss)
    scene = PlaneScene(rng_tex, fscale=min(K4[0], K4[1]) / KITTI_FX)
    cams = np.zeros((n_frames, 6))
    for i in range(n_frames):
        w = np.array([0.0, 0.002 * i, 0.0005 * i])
        C = i * np.array([0.03, 0.005, 0.10])
        cams[i, :3] = w
        cams[i, 3:] = -rodrigues(w) @ C
    images = np.stack([scene.render(cams[i], K4, rows, cols) for i in range(n_frames)])
    ys, xs = np.meshgrid(np.arange(rows, dtype=np.float64), np.arange(cols, dtype=np.float64), indexing="ij")
    depths = np.empty((n_frames, rows, cols), dtype=np.float32)
    for i in range(n_frames):
        _, z = scene.intersect(cams[i], K4, xs, ys)
        depths[i] = (z * (1.0 + depth_noise * rng_depth.normal(size=z.shape))).astype(np.float32)
    T_c = np.stack([cam_to_mat(c) for c in cams])              # world -> camera
    T_w = np.stack([np.linalg.inv(T) for T in T_c])            # camera -> world
    # convertPoseToLocal (src/pose_utils.cc:62-74): T_i = inv(T_w[i]) * T_w[i-1], T_0 = inv(T_w[0])
    T_rel = np.stack([np.linalg.inv(T_w[0])] + [np.linalg.inv(T_w[i]) @ T_w[i - 1] for i in range(1, n_frames)])
    T_init = T_rel.copy()
    for i in range(1, n_frames):
        d = np.concatenate([rng_pose.normal(0, rot_sigma, 3), rng_pose.normal(0, trans_sigma, 3)])
        T_init[i] = cam_to_mat(d) @ T_rel[i]
    return Sequence(images=images, depths=depths, K4=K4, T_w_gt=T_w, T_rel_gt=T_rel, T_rel_init=T_init)


# ----------------------------------------------------------------------------------------------
# Pyramid (BASELINE config 4).  The reference's pyramid class is unfinished, so
# the level semantics are defined here: same window and same world points at every level; level l
# uses frames reduced l times with the cv::pyrDown rule (what src/photobundle_pyramid.cc:46 intends),
# intrinsics per Calibration::pyrDown (src/calibration.h:72-78: K *= 0.5), and descriptors
# re-extracted per level as the bilinear 5x5 patch of the level-l reference frame at the point's
# level-0 pixel divided by 2^l (so that descriptor location and projection stay consistent).
# ----------------------------------------------------------------------------------------------
def _reflect101(i: np.ndarray, n: int) -> np.ndarray:
    i = np.asarray(i).copy()
    if n == 1:
        return np.zeros_like(i)
    for _ in range(4):
        i = np.where(i < 0, -i, i)
        i = np.where(i >= n, 2 * n - 2 - i, i)
    return i


def pyr_down_u8(img: np.ndarray) -> np.ndarray:
    """cv::pyrDown for CV_8U: separable [1 4 6 4 1], BORDER_REFLECT_101, (sum + 128) >> 8."""
    rows, cols = img.shape
    drows, dcols = (rows + 1) // 2, (cols + 1) // 2
    w = (1, 4, 6, 4, 1)
    src = img.astype(np.int32)
    h = sum(w[i] * src[:, _reflect101(2 * np.arange(dcols) + i - 2, cols)] for i in range(5))
    v = sum(w[j] * h[_reflect101(2 * np.arange(drows) + j - 2, rows), :] for j in range(5))
    return ((v + 128) >> 8).astype(np.uint8)


def _bilinear_patch(img: np.ndarray, x: float, y: float, radius: int) -> np.ndarray:
    rows, cols = img.shape
    out = np.empty((2 * radius + 1) ** 2)
    k = 0
    for dy in range(-radius, radius + 1):
        for dx in range(-radius, radius + 1):
            xx = min(max(x + dx, 0.0), cols - 1.0)
            yy = min(max(y + dy, 0.0), rows - 1.0)
            x0, y0 = min(int(xx), cols - 2), min(int(yy), rows - 2)
            ax, ay = xx - x0, yy - y0
            out[k] = ((1 - ay) * ((1 - ax) * img[y0, x0] + ax * img[y0, x0 + 1]) +
                      ay * ((1 - ax) * img[y0 + 1, x0] + ax * img[y0 + 1, x0 + 1]))
            k += 1
    return out.astype(np.float32).astype(np.float64)   # descriptors are float channel values widened to double


def pyramid_level(win: Window, level: int, images_l: np.ndarray, ref_px: np.ndarray, ref_frame: np.ndarray) -> Window:
    """The level-`level` window of `win`; images_l are win.images reduced `level` times."""
    s = float(2 ** level)
    img_f = images_l.astype(np.float64)
    desc = np.stack([_bilinear_patch(img_f[int(ref_frame[k])], ref_px[k, 0] / s, ref_px[k, 1] / s, win.radius)
                     for k in range(win.n_points)]) if level > 0 else win.desc
    return dataclasses.replace(win, images=images_l, fx=win.fx / s, fy=win.fy / s, cx=win.cx / s, cy=win.cy / s, desc=desc)


def reference_pixels(win: Window) -> tuple[np.ndarray, np.ndarray]:
    """(integer pixel of each point in its reference frame, reference frame index), recovered by
    projecting the initial point with the initial pose it was lifted with."""
    ref = win.obs_frame[win.obs_offsets[:-1]]
    px = np.empty((win.n_points, 2))
    for f in np.unique(ref):
        sel = np.nonzero(ref == f)[0]
        R = rodrigues(win.cams_init[f, :3])
        Xc = win.points_init[sel] @ R.T + win.cams_init[f, 3:]
        px[sel, 0] = np.rint(win.fx * Xc[:, 0] / Xc[:, 2] + win.cx)
        px[sel, 1] = np.rint(win.fy * Xc[:, 1] / Xc[:, 2] + win.cy)
    return px, ref
