"""ctypes binding of the C++ host class PhotometricBundleAdjustment (libpba_host.so), the
drop-in for the reference's src/photobundle.h interface.  Used by tests only."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpba_host.so")


class Options(C.Structure):
    _fields_ = [("maxNumPoints", C.c_int32), ("slidingWindowSize", C.c_int32), ("patchRadius", C.c_int32),
                ("maskBlockRadius", C.c_int32), ("maxFrameDistance", C.c_int32), ("nonMaxSuppRadius", C.c_int32),
                ("doGaussianWeighting", C.c_int32), ("verbose", C.c_int32), ("device", C.c_int32), ("descriptorType", C.c_int32), ("gpuFrontEnd", C.c_int32),
                ("minScore", C.c_double), ("robustThreshold", C.c_double), ("minValidDepth", C.c_double),
                ("maxValidDepth", C.c_double), ("numPyramidLevels", C.c_int32), ("nGpus", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: run __graft_entry__.build()")
        from . import capi
        capi.lib()  # libpba_b200.so first (dependency)
        _lib = C.CDLL(LIB_PATH)
        _lib.pbah_last_error.restype = C.c_char_p
    return _lib


def default_options() -> "Options":
    """PhotometricBundleAdjustment::Options as its default constructor leaves it."""
    o = Options()
    lib().pbah_default_options(C.byref(o))
    return o


class BundleAdjuster:
    """Mirror of `PhotometricBundleAdjustment photoba(calib, size, options); photoba.addFrame(...)`."""

    def __init__(self, rows, cols, fx, fy, cx, cy, baseline=0.5, **opts):
        L = lib()
        o = Options()
        L.pbah_default_options(C.byref(o))
        for k, v in opts.items():
            setattr(o, k, v)
        self.opts = o
        self._h = C.c_void_p()
        if L.pbah_create(rows, cols, C.c_double(fx), C.c_double(fy), C.c_double(cx), C.c_double(cy),
                         C.c_double(baseline), C.byref(o), C.byref(self._h)) != 0:
            raise RuntimeError(L.pbah_last_error().decode())
        self.rows, self.cols = rows, cols

    def close(self):
        if self._h:
            lib().pbah_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_frame(self, image: np.ndarray, depth: np.ndarray, T: np.ndarray) -> bool:
        image = np.ascontiguousarray(image, dtype=np.uint8)
        depth = np.ascontiguousarray(depth, dtype=np.float32)
        Tc = np.asfortranarray(np.asarray(T, dtype=np.float64))
        ran = C.c_int32()
        if lib().pbah_add_frame(self._h, C.c_void_p(image.ctypes.data), C.c_void_p(depth.ctypes.data),
                                C.c_void_p(Tc.ctypes.data), C.byref(ran)) != 0:
            raise RuntimeError(lib().pbah_last_error().decode())
        return bool(ran.value)

    def result(self) -> dict:
        L = lib()
        n_poses, n_pts, n_it = C.c_int32(), C.c_int32(), C.c_int32()
        L.pbah_result_counts(self._h, C.byref(n_poses), C.byref(n_pts), C.byref(n_it))
        poses = np.zeros((max(1, n_poses.value), 16))
        refined = np.zeros((max(1, n_pts.value), 3))
        original = np.zeros((max(1, n_pts.value), 3))
        scalars = np.zeros(6)
        costs = np.zeros(max(1, n_it.value))
        msg = C.create_string_buffer(256)
        L.pbah_result_get(self._h, C.c_void_p(poses.ctypes.data), C.c_void_p(refined.ctypes.data),
                          C.c_void_p(original.ctypes.data), C.c_void_p(scalars.ctypes.data),
                          C.c_void_p(costs.ctypes.data), msg, 256)
        P = poses[: n_poses.value].reshape(-1, 4, 4).transpose(0, 2, 1)  # column-major -> [r, c]
        return dict(poses=P, refinedPoints=refined[: n_pts.value], originalPoints=original[: n_pts.value],
                    initialCost=scalars[0], finalCost=scalars[1], fixedCost=scalars[2],
                    numSuccessfulStep=int(scalars[3]), numResiduals=int(scalars[4]), totalTime=scalars[5],
                    message=msg.value.decode(), iterationCosts=costs[: n_it.value])

    def scene_points(self) -> list[dict]:
        L = lib()
        n = L.pbah_num_scene_points(self._h)
        out = []
        X = np.zeros(3)
        xy = np.zeros(2, dtype=np.int32)
        vis = np.zeros(64, dtype=np.uint32)
        desc = np.zeros(81 * 8)
        for i in range(n):
            nv = L.pbah_scene_point(self._h, i, C.c_void_p(X.ctypes.data), C.c_void_p(xy.ctypes.data),
                                    C.c_void_p(vis.ctypes.data), 64, C.c_void_p(desc.ctypes.data), desc.size)
            P = (2 * self.opts.patchRadius + 1) ** 2 * {0: 1, 1: 3, 2: 8}[int(self.opts.descriptorType)]
            out.append(dict(X=X.copy(), x=int(xy[0]), y=int(xy[1]), vis=vis[:nv].tolist(), desc=desc[:P].copy()))
        return out

    def write_poses(self, path: str):
        if lib().pbah_write_poses_kitti(self._h, path.encode()) != 0:
            raise RuntimeError("writePosesKittiFormat failed")


def disparity_to_depth(disparity: np.ndarray, Bf: float) -> np.ndarray:
    """disparityToDepth of the host shim (src/imgproc.cc:280-330 semantics)."""
    d = np.ascontiguousarray(disparity, dtype=np.float32)
    out = np.zeros_like(d)
    lib().pbah_disparity_to_depth(C.c_void_p(d.ctypes.data), d.shape[0], d.shape[1], C.c_float(Bf), C.c_void_p(out.ctypes.data))
    return out


def load_poses_kitti(path: str, cap: int = 100000) -> np.ndarray:
    buf = np.zeros((cap, 16))
    n = lib().pbah_load_poses_kitti(path.encode(), C.c_void_p(buf.ctypes.data), cap)
    if n < 0:
        raise RuntimeError(lib().pbah_last_error().decode())
    return buf[:n].reshape(-1, 4, 4).transpose(0, 2, 1)
