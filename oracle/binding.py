"""ctypes binding of the CPU oracle (oracle/libpba_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under photobundle_b200/ imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpba_oracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libref_sampler.so")


def build(force: bool = False) -> None:
    src = os.path.join(_HERE, "pba_oracle.cc")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "pba_oracle.h")))
    if force or stale or (os.path.isdir("/root/reference/src") and not os.path.exists(_REF_PATH)):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


class Problem(C.Structure):
    _fields_ = [
        ("rows", C.c_int32), ("cols", C.c_int32), ("n_frames", C.c_int32), ("n_channels", C.c_int32),
        ("radius", C.c_int32), ("n_points", C.c_int32), ("fixed_frame", C.c_int32), ("num_threads", C.c_int32),
        ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double), ("huber", C.c_double),
        ("planes", C.c_void_p), ("grad_x", C.c_void_p), ("grad_y", C.c_void_p), ("weights", C.c_void_p),
        ("desc", C.c_void_p), ("obs_offsets", C.c_void_p), ("obs_frame", C.c_void_p),
    ]


class Blocks(C.Structure):
    _fields_ = [
        ("cost", C.c_double), ("U", C.c_void_p), ("gc", C.c_void_p), ("V", C.c_void_p), ("gp", C.c_void_p),
        ("W", C.c_void_p), ("obs_sqnorm", C.c_void_p), ("residuals", C.c_void_p),
    ]


class IterationSummary(C.Structure):
    _fields_ = [
        ("iteration", C.c_int32), ("step_is_valid", C.c_int32), ("step_is_nonmonotonic", C.c_int32),
        ("step_is_successful", C.c_int32), ("cost", C.c_double), ("cost_change", C.c_double),
        ("gradient_max_norm", C.c_double), ("gradient_norm", C.c_double), ("step_norm", C.c_double),
        ("relative_decrease", C.c_double), ("trust_region_radius", C.c_double), ("eta", C.c_double),
        ("step_size", C.c_double), ("line_search_function_evaluations", C.c_int32),
        ("line_search_gradient_evaluations", C.c_int32), ("line_search_iterations", C.c_int32),
        ("linear_solver_iterations", C.c_int32), ("iteration_time_in_seconds", C.c_double),
        ("step_solver_time_in_seconds", C.c_double), ("cumulative_time_in_seconds", C.c_double),
    ]


class SolverOptions(C.Structure):
    _fields_ = [
        ("max_num_iterations", C.c_int32), ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
        ("initial_trust_region_radius", C.c_double), ("max_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double), ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
        ("max_num_consecutive_invalid_steps", C.c_int32), ("jacobi_scaling", C.c_int32),
        ("use_autodiff", C.c_int32),
    ]


class Summary(C.Structure):
    _fields_ = [
        ("initial_cost", C.c_double), ("final_cost", C.c_double), ("fixed_cost", C.c_double),
        ("num_successful_steps", C.c_int32), ("num_unsuccessful_steps", C.c_int32),
        ("num_residuals", C.c_int32), ("num_residual_blocks", C.c_int32), ("num_iterations", C.c_int32),
        ("termination_type", C.c_int32), ("num_jacobian_evals", C.c_int32), ("num_cost_evals", C.c_int32),
        ("total_time_in_seconds", C.c_double), ("jacobian_time_in_seconds", C.c_double),
        ("cost_time_in_seconds", C.c_double), ("linear_solver_time_in_seconds", C.c_double),
        ("message", C.c_char * 256),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_cost.restype = C.c_double
        _lib.oracle_solve.restype = C.c_int32
    return _lib


def ref_lib() -> C.CDLL | None:
    """The reference's own sampler compiled from /root/reference (oracle/_ref/), or None."""
    if not os.path.exists(_REF_PATH):
        return None
    r = C.CDLL(_REF_PATH)
    r.ref_sample_with_derivative_double.restype = C.c_double
    r.ref_sample_with_derivative_double.argtypes = [C.c_void_p] * 3 + [C.c_int32, C.c_int32, C.c_double, C.c_double]
    return r


_REF_CALIB_PATH = os.path.join(os.path.dirname(_REF_PATH), "libref_calib.so")


def ref_calib_lib() -> C.CDLL | None:
    """The reference's own camera model (src/calibration.h) and MakePatchWeights compiled from /root/reference, or None."""
    if not os.path.exists(_REF_CALIB_PATH):
        return None
    r = C.CDLL(_REF_CALIB_PATH)
    r.ref_pyr_down.argtypes = [C.c_void_p, C.c_double, C.c_int32, C.c_int32, C.c_void_p]
    r.ref_triangulate.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    return r


_REF_IMGPROC_PATH = os.path.join(os.path.dirname(_REF_PATH), "libref_imgproc.so")


def ref_imgproc_lib() -> C.CDLL | None:
    """The reference's own imgradient / disparityToDepth (src/imgproc.cc) compiled from /root/reference, or None."""
    if not os.path.exists(_REF_IMGPROC_PATH):
        return None
    r = C.CDLL(_REF_IMGPROC_PATH)
    r.ref_disparity_to_depth.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p]
    return r


_REF_FUNCTOR_PATH = os.path.join(os.path.dirname(_REF_PATH), "libref_functor.so")


def ref_functor_lib() -> C.CDLL | None:
    """The reference's residual functor body (src/photobundle.cc:696-727) over its own sampler and camera model, or None."""
    if not os.path.exists(_REF_FUNCTOR_PATH):
        return None
    r = C.CDLL(_REF_FUNCTOR_PATH)
    r.ref_residual_block.restype = C.c_int32
    r.ref_residual_block.argtypes = [C.c_void_p] * 3 + [C.c_int32] * 3 + [C.c_void_p, C.c_int32] + [C.c_void_p] * 5
    r.ref_residual_block_jet.restype = C.c_int32
    r.ref_residual_block_jet.argtypes = [C.c_void_p] * 3 + [C.c_int32] * 3 + [C.c_void_p, C.c_int32] + [C.c_void_p] * 6
    return r


def _p(a: np.ndarray) -> int:
    return a.ctypes.data


DESCRIPTOR_TYPES = {"intensity": 0, "intensity_and_gradient": 1, "bitplanes": 2}


def build_channels(image_u8: np.ndarray, descriptor: str) -> np.ndarray:
    """DescriptorFrame::Create restated: uint8 image -> fp32 channel planes [C, rows, cols]."""
    L = lib()
    img = np.ascontiguousarray(image_u8, dtype=np.uint8)
    t = DESCRIPTOR_TYPES[descriptor]
    Cn = int(L.oracle_descriptor_channels(t))
    out = np.zeros((Cn,) + img.shape, dtype=np.float32)
    L.oracle_build_channels(t, C.c_void_p(_p(img)), img.shape[0], img.shape[1], C.c_void_p(_p(out)))
    return out


def saliency_map(planes: np.ndarray) -> np.ndarray:
    planes = np.ascontiguousarray(planes, dtype=np.float32)
    out = np.zeros(planes.shape[1:], dtype=np.float32)
    lib().oracle_saliency_map(C.c_void_p(_p(planes)), planes.shape[0], planes.shape[1], planes.shape[2], C.c_void_p(_p(out)))
    return out


def extract_patches(planes: np.ndarray, xy: np.ndarray, radius: int) -> np.ndarray:
    planes = np.ascontiguousarray(planes, dtype=np.float32)
    xy = np.ascontiguousarray(xy, dtype=np.int32).reshape(-1, 2)
    P = (2 * radius + 1) ** 2
    out = np.zeros((xy.shape[0], planes.shape[0] * P), dtype=np.float64)
    lib().oracle_extract_patches(C.c_void_p(_p(planes)), planes.shape[0], planes.shape[1], planes.shape[2], int(radius),
                                 xy.shape[0], C.c_void_p(_p(xy)), C.c_void_p(_p(out)))
    return out


class OracleWindow:
    """Holds a synthetic.Window in the layout oracle_problem wants and keeps arrays alive."""

    def __init__(self, win, num_threads: int = 0, planes: np.ndarray | None = None):
        self.win = win
        self.planes = np.ascontiguousarray(win.planes_f32() if planes is None else planes, dtype=np.float32)
        F, Cn, rows, cols = self.planes.shape
        self.weights = np.ascontiguousarray(win.weights, dtype=np.float64)
        self.desc = np.ascontiguousarray(win.desc, dtype=np.float64)
        self.obs_offsets = np.ascontiguousarray(win.obs_offsets, dtype=np.int32)
        self.obs_frame = np.ascontiguousarray(win.obs_frame, dtype=np.int32)
        # gradients once (so repeated evaluate() calls do not redo them)
        self.gx = np.empty_like(self.planes)
        self.gy = np.empty_like(self.planes)
        L = lib()
        for f in range(F):
            for k in range(Cn):
                L.oracle_imgradient(C.c_void_p(_p(self.planes[f, k])), rows, cols,
                                    C.c_void_p(_p(self.gx[f, k])), C.c_void_p(_p(self.gy[f, k])))
        self.pb = Problem(
            rows=rows, cols=cols, n_frames=F, n_channels=Cn, radius=win.radius, n_points=win.n_points,
            fixed_frame=win.fixed_frame, num_threads=num_threads, fx=win.fx, fy=win.fy, cx=win.cx, cy=win.cy,
            huber=win.huber, planes=_p(self.planes), grad_x=_p(self.gx), grad_y=_p(self.gy),
            weights=_p(self.weights), desc=_p(self.desc), obs_offsets=_p(self.obs_offsets),
            obs_frame=_p(self.obs_frame))
        self.CP = Cn * win.patch_len

    def cost(self, cams, points) -> float:
        cams = np.ascontiguousarray(cams, dtype=np.float64)
        points = np.ascontiguousarray(points, dtype=np.float64)
        return float(lib().oracle_cost(C.byref(self.pb), C.c_void_p(_p(cams)), C.c_void_p(_p(points))))

    def residual_block(self, frame, cam6, xyz, desc, mode=1):
        """mode 1: Jet<9> autodiff; 0: analytic; 2: T=double (residual only)."""
        cam6 = np.ascontiguousarray(cam6, dtype=np.float64)
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        desc = np.ascontiguousarray(desc, dtype=np.float64)
        r = np.zeros(self.CP)
        Jc = np.zeros((self.CP, 6))
        Jp = np.zeros((self.CP, 3))
        lib().oracle_residual_block(C.byref(self.pb), int(frame), C.c_void_p(_p(cam6)), C.c_void_p(_p(xyz)),
                                    C.c_void_p(_p(desc)), int(mode), C.c_void_p(_p(r)),
                                    C.c_void_p(_p(Jc)), C.c_void_p(_p(Jp)))
        return r, Jc, Jp

    def evaluate(self, cams, points, use_autodiff: int = 1, want_residuals: bool = True) -> dict:
        cams = np.ascontiguousarray(cams, dtype=np.float64)
        points = np.ascontiguousarray(points, dtype=np.float64)
        F, n, nnz = self.pb.n_frames, self.pb.n_points, int(self.obs_offsets[-1])
        out = dict(U=np.zeros((F, 6, 6)), gc=np.zeros((F, 6)), V=np.zeros((n, 3, 3)), gp=np.zeros((n, 3)),
                   W=np.zeros((nnz, 6, 3)), obs_sqnorm=np.zeros(nnz))
        if want_residuals:
            out["residuals"] = np.zeros((nnz, self.CP))
        b = Blocks(U=_p(out["U"]), gc=_p(out["gc"]), V=_p(out["V"]), gp=_p(out["gp"]), W=_p(out["W"]),
                   obs_sqnorm=_p(out["obs_sqnorm"]),
                   residuals=_p(out["residuals"]) if want_residuals else None)
        lib().oracle_evaluate(C.byref(self.pb), C.c_void_p(_p(cams)), C.c_void_p(_p(points)), int(use_autodiff), C.byref(b))
        out["cost"] = float(b.cost)
        return out

    def solve(self, cams, points, **opt_overrides):
        """Returns (cams, points, summary dict, trace list of dicts)."""
        cams = np.array(cams, dtype=np.float64, order="C", copy=True)
        points = np.array(points, dtype=np.float64, order="C", copy=True)
        opt = SolverOptions()
        lib().oracle_default_options(C.byref(opt))
        for k, v in opt_overrides.items():
            setattr(opt, k, v)
        summ = Summary()
        trace = (IterationSummary * (opt.max_num_iterations + 2))()
        lib().oracle_solve(C.byref(self.pb), C.byref(opt), C.c_void_p(_p(cams)), C.c_void_p(_p(points)),
                           C.byref(summ), trace)
        sd = {f[0]: getattr(summ, f[0]) for f in Summary._fields_}
        sd["message"] = summ.message.decode()
        tr = [{f[0]: getattr(trace[i], f[0]) for f in IterationSummary._fields_} for i in range(summ.num_iterations)]
        return cams, points, sd, tr


def pose_to_params(T44: np.ndarray) -> np.ndarray:
    T = np.asfortranarray(np.asarray(T44, dtype=np.float64))
    p = np.zeros(6)
    lib().oracle_pose_to_params(C.c_void_p(T.ctypes.data), C.c_void_p(p.ctypes.data))
    return p


def params_to_pose(p6: np.ndarray) -> np.ndarray:
    p = np.ascontiguousarray(p6, dtype=np.float64)
    T = np.zeros((4, 4), order="F")
    lib().oracle_params_to_pose(C.c_void_p(p.ctypes.data), C.c_void_p(T.ctypes.data))
    return np.ascontiguousarray(T)
