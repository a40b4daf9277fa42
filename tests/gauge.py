"""Comparison of two solutions of the same window up to the monocular scale gauge.

A photometric window with one fixed camera still has one weakly determined direction: scaling the whole scene
about the fixed camera's centre C0 (X -> C0 + a (X - C0), C_f -> C0 + a (C_f - C0), rotations unchanged) moves
every projection only through the fixed camera's own residuals.  Two solvers that differ in the last bits can
end at different points along it; everything orthogonal to it is well determined.  `scale_gauge_diff` fits the
scale `a` between two solutions (least squares over the camera centres and points) and reports what is left."""
import numpy as np


def _rodrigues(aa):
    th = np.linalg.norm(aa)
    if th < 1e-14:
        return np.eye(3)
    k = aa / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def centres(cams):
    """cams [F, 6] = angle-axis + translation of the INVERSE pose (x_cam = R X + t): centre C = -R^T t."""
    return np.stack([-_rodrigues(c[:3]).T @ c[3:] for c in cams])


def scale_gauge_diff(cams_a, pts_a, cams_b, pts_b, fixed_frame):
    Ca, Cb = centres(cams_a), centres(cams_b)
    c0 = Cb[fixed_frame]
    va = np.concatenate([(Ca - c0).ravel(), (pts_a - c0).ravel()])
    vb = np.concatenate([(Cb - c0).ravel(), (pts_b - c0).ravel()])
    alpha = float(va @ vb / (va @ va))                     # a such that a * (A - c0) ~ (B - c0)
    size = max(1.0, float(np.abs(pts_b - c0).max()))
    return dict(alpha=alpha,
                rotations=float(np.abs(cams_a[:, :3] - cams_b[:, :3]).max()),
                centres=float(np.abs(alpha * (Ca - c0) - (Cb - c0)).max()),
                points=float(np.abs(alpha * (pts_a - c0) - (pts_b - c0)).max() / size))
