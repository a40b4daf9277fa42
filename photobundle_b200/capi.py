"""ctypes binding of the C-ABI product library (include/pba_b200.h -> libpba_b200.so).

This is only the Python face of the boundary used by tests/ and bench.py; the host side
of the product is C++ (photobundle_b200/host) and CUDA (photobundle_b200/csrc).  It fails
loudly when the CUDA library is missing: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PBA_B200_LIB: developer override to A/B-test scratch builds of the same library (scripts/)
LIB_PATH = os.environ.get("PBA_B200_LIB") or os.path.join(_HERE, "libpba_b200.so")

PBA_MAX_FRAMES = 16
PBA_UNIQUE_ID_BYTES = 128

EXPORTED_SYMBOLS = [
    "pba_last_error", "pba_version", "pba_default_solver_options", "pba_create", "pba_destroy",
    "pba_set_frames_u8", "pba_set_frames_f32", "pba_set_frame_u8", "pba_set_frames_u8_pyr", "pba_pyrdown_u8", "pba_set_poses", "pba_set_points",
    "pba_eval", "pba_eval_timed", "pba_solve", "pba_save_state", "pba_restore_state", "pba_copy_state", "pba_host_alloc", "pba_host_free", "pba_graph_counters", "pba_begin_batch", "pba_end_batch", "pba_set_frame_u8_ex", "pba_get_results", "pba_get_poses", "pba_get_points", "pba_get_iterations",
    "pba_comm_unique_id", "pba_comm_init", "pba_comm_init_local", "pba_comm_speculates", "pba_comm_sharded", "pba_shard_range", "pba_comm_exchange_kind",
    "pba_descriptor_channels", "pba_set_frames_u8_descriptor", "pba_get_channel_plane", "pba_prepare_frame_u8",
    "pba_saliency_map", "pba_extract_descriptors", "pba_associate", "pba_select_candidates",
]


class PbaError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [
        ("rows", C.c_int32), ("cols", C.c_int32), ("n_channels", C.c_int32), ("patch_radius", C.c_int32),
        ("max_frames", C.c_int32), ("max_points", C.c_int32), ("max_observations", C.c_int32),
        ("device", C.c_int32), ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
        ("huber", C.c_double),
    ]


class SolverOptions(C.Structure):
    _fields_ = [
        ("max_num_iterations", C.c_int32), ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
        ("initial_trust_region_radius", C.c_double), ("max_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double), ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
        ("max_num_consecutive_invalid_steps", C.c_int32), ("jacobi_scaling", C.c_int32),
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


class Summary(C.Structure):
    _fields_ = [
        ("initial_cost", C.c_double), ("final_cost", C.c_double), ("fixed_cost", C.c_double),
        ("num_successful_steps", C.c_int32), ("num_unsuccessful_steps", C.c_int32),
        ("num_residuals", C.c_int32), ("num_residual_blocks", C.c_int32), ("num_iterations", C.c_int32),
        ("termination_type", C.c_int32), ("num_evaluations", C.c_int32), ("kernel_launches", C.c_int32),
        ("num_collectives", C.c_int32), ("total_time_in_seconds", C.c_double),
        ("device_time_in_seconds", C.c_double), ("kb_device_time_in_seconds", C.c_double), ("message", C.c_char * 256),
    ]


class EvalOut(C.Structure):
    _fields_ = [
        ("cost", C.c_double), ("U", C.c_void_p), ("gc", C.c_void_p), ("V", C.c_void_p), ("gp", C.c_void_p),
        ("W", C.c_void_p), ("obs_sqnorm", C.c_void_p), ("residuals", C.c_void_p), ("device_ms", C.c_double),
    ]


_lib = None


def lib() -> C.CDLL:
    """Load libpba_b200.so; raise (never fall back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PbaError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.pba_last_error.restype = C.c_char_p
        L.pba_version.restype = C.c_char_p
        for name in EXPORTED_SYMBOLS:
            getattr(L, name)  # AttributeError if the ABI and the header drift apart
        _lib = L
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise PbaError(f"{what} failed ({rc}): {lib().pba_last_error().decode()}")


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class Handle:
    """One window on one GPU (a thin, explicit wrapper: every method is one C call)."""

    def __init__(self, rows, cols, fx, fy, cx, cy, radius=2, n_channels=1, huber=0.05,
                 max_frames=8, max_points=4096, max_observations=None, device=-1):
        self.cfg = Config(rows=rows, cols=cols, n_channels=n_channels, patch_radius=radius,
                          max_frames=max_frames, max_points=max_points,
                          max_observations=max_observations or max_points * max_frames, device=device,
                          fx=fx, fy=fy, cx=cx, cy=cy, huber=huber)
        self._h = C.c_void_p()
        _check(lib().pba_create(C.byref(self.cfg), C.byref(self._h)), "pba_create")
        self.n_frames = 0
        self.n_points = 0
        self.n_obs = 0
        self.rows, self.cols = rows, cols
        self.P = (2 * radius + 1) ** 2
        self.CP = self.P * n_channels

    @classmethod
    def for_window(cls, win, device=-1, planes_f32: np.ndarray | None = None):
        """Create a handle sized for a synthetic.Window and upload it (uint8 Intensity frames
        unless explicit fp32 channel planes [F, C, rows, cols] are given)."""
        nch = 1 if planes_f32 is None else int(planes_f32.shape[1])
        h = cls(win.rows, win.cols, win.fx, win.fy, win.cx, win.cy, radius=win.radius, n_channels=nch,
                huber=win.huber, max_frames=win.n_frames, max_points=max(1, win.n_points),
                max_observations=max(1, win.n_obs), device=device)
        if planes_f32 is None:
            h.set_frames_u8(win.images)
        else:
            h.set_frames_f32(planes_f32)
        h.set_poses(win.cams_init, win.fixed_frame)
        h.set_points(win.points_init, win.desc, win.obs_offsets, win.obs_frame, win.weights)
        return h

    def close(self):
        if self._h:
            lib().pba_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_frames_u8(self, images: np.ndarray):
        images = np.ascontiguousarray(images, dtype=np.uint8)
        F = images.shape[0]
        arr = (C.c_void_p * F)(*[images[f].ctypes.data for f in range(F)])
        getattr(self, "_batch_keepalive", []).append(images)    # batch mode: the copy is still in flight when we return
        _check(lib().pba_set_frames_u8(self._h, F, arr), "pba_set_frames_u8")
        self.n_frames = F

    def set_frames_u8_pyr(self, images: np.ndarray, levels_down: int):
        """Upload LEVEL-0 frames and reduce them `levels_down` times on the device (cv::pyrDown rule)."""
        images = np.ascontiguousarray(images, dtype=np.uint8)
        F, r, c = images.shape
        arr = (C.c_void_p * F)(*[images[f].ctypes.data for f in range(F)])
        _check(lib().pba_set_frames_u8_pyr(self._h, F, arr, r, c, int(levels_down)), "pba_set_frames_u8_pyr")
        self.n_frames = F

    def set_frame_u8(self, slot: int, image: np.ndarray):
        image = np.ascontiguousarray(image, dtype=np.uint8)
        _check(lib().pba_set_frame_u8(self._h, slot, _ptr(image)), "pba_set_frame_u8")

    def set_frames_f32(self, planes: np.ndarray):
        planes = np.ascontiguousarray(planes, dtype=np.float32)  # [F, C, rows, cols]
        F, Cn = planes.shape[:2]
        arr = (C.c_void_p * (F * Cn))(*[planes[f, k].ctypes.data for f in range(F) for k in range(Cn)])
        _check(lib().pba_set_frames_f32(self._h, F, arr), "pba_set_frames_f32")
        self.n_frames = F

    DESCRIPTOR_TYPES = {"intensity": 0, "intensity_and_gradient": 1, "bitplanes": 2}

    def set_frames_u8_descriptor(self, images: np.ndarray, descriptor: str):
        """Upload uint8 frames and build the descriptor's channel planes on the device (DescriptorFrame::Create)."""
        images = np.ascontiguousarray(images, dtype=np.uint8)
        F = images.shape[0]
        arr = (C.c_void_p * F)(*[images[f].ctypes.data for f in range(F)])
        _check(lib().pba_set_frames_u8_descriptor(self._h, F, arr, self.DESCRIPTOR_TYPES[descriptor]), "pba_set_frames_u8_descriptor")
        self.n_frames = F

    def get_channel_plane(self, frame: int, channel: int) -> np.ndarray:
        out = np.zeros((self.rows, self.cols), dtype=np.float32)
        _check(lib().pba_get_channel_plane(self._h, int(frame), int(channel), _ptr(out)), "pba_get_channel_plane")
        return out

    def prepare_frame_u8(self, image: np.ndarray, descriptor: str):
        image = np.ascontiguousarray(image, dtype=np.uint8)
        assert image.shape == (self.rows, self.cols)
        _check(lib().pba_prepare_frame_u8(self._h, _ptr(image), self.DESCRIPTOR_TYPES[descriptor]), "pba_prepare_frame_u8")
        self._prepared_channels = int(lib().pba_descriptor_channels(self.DESCRIPTOR_TYPES[descriptor]))

    def saliency_map(self) -> np.ndarray:
        out = np.zeros((self.rows, self.cols), dtype=np.float32)
        _check(lib().pba_saliency_map(self._h, _ptr(out)), "pba_saliency_map")
        return out

    def extract_descriptors(self, xy: np.ndarray) -> np.ndarray:
        xy = np.ascontiguousarray(xy, dtype=np.int32).reshape(-1, 2)
        out = np.zeros((xy.shape[0], self._prepared_channels * self.P), dtype=np.float64)
        _check(lib().pba_extract_descriptors(self._h, xy.shape[0], _ptr(xy), _ptr(out)), "pba_extract_descriptors")
        return out

    def associate(self, xyz, ref_patch, ref_norm, T_c, K, border: int):
        """addFrame's data association on the frame of prepare_frame_u8: (score [n], row_col [n, 2]); score -2 = not tested."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        n = xyz.shape[0]
        ref_patch = np.ascontiguousarray(ref_patch, dtype=np.float32).reshape(n, 25)
        ref_norm = np.ascontiguousarray(ref_norm, dtype=np.float32).reshape(n)
        Tc = np.ascontiguousarray(np.asarray(T_c, dtype=np.float64).T)   # column-major 4x4
        Kr = np.ascontiguousarray(K, dtype=np.float64)                   # row-major 3x3
        score, rc = np.zeros(n, dtype=np.float32), np.zeros((n, 2), dtype=np.int32)
        _check(lib().pba_associate(self._h, n, _ptr(xyz), _ptr(ref_patch), _ptr(ref_norm), _ptr(Tc), _ptr(Kr), int(border),
                                   _ptr(score), _ptr(rc)), "pba_associate")
        return score, rc

    def select_candidates(self, depth, masked_rc, mask_radius: int, nms_radius: int, border: int, min_depth: float, max_depth: float,
                          capacity: int = 1 << 20):
        """New-point candidates of addFrame on the frame of prepare_frame_u8: (row_col [m, 2] in scan order, saliency [m])."""
        depth = None if depth is None else np.ascontiguousarray(depth, dtype=np.float32)   # None: no depth test on the device
        masked = np.ascontiguousarray(masked_rc, dtype=np.int32).reshape(-1, 2)
        rc, sal, n_out = np.zeros((capacity, 2), dtype=np.int32), np.zeros(capacity, dtype=np.float32), C.c_int32()
        _check(lib().pba_select_candidates(self._h, None if depth is None else _ptr(depth), masked.shape[0], _ptr(masked), int(mask_radius), int(nms_radius), int(border),
                                           C.c_double(min_depth), C.c_double(max_depth), capacity, _ptr(rc), _ptr(sal), C.byref(n_out)),
               "pba_select_candidates")
        assert n_out.value <= capacity
        return rc[:n_out.value], sal[:n_out.value]

    def set_poses(self, cams: np.ndarray, fixed_frame: int = 0):
        cams = np.ascontiguousarray(cams, dtype=np.float64)
        getattr(self, "_batch_keepalive", []).append(cams)
        _check(lib().pba_set_poses(self._h, cams.shape[0], _ptr(cams), int(fixed_frame)), "pba_set_poses")
        self.n_frames = cams.shape[0]

    def set_points(self, xyz, desc, obs_offsets, obs_frame, weights):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        desc = np.ascontiguousarray(desc, dtype=np.float64)
        obs_offsets = np.ascontiguousarray(obs_offsets, dtype=np.int32)
        obs_frame = np.ascontiguousarray(obs_frame, dtype=np.int32)
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        n = xyz.shape[0]
        assert desc.shape == (n, self.CP) and obs_offsets.shape == (n + 1,) and weights.shape == (self.P,)
        getattr(self, "_batch_keepalive", []).extend([xyz, desc, obs_offsets, obs_frame, weights])
        _check(lib().pba_set_points(self._h, n, _ptr(xyz), _ptr(desc), _ptr(obs_offsets), _ptr(obs_frame),
                                    _ptr(weights)), "pba_set_points")
        self.n_points = n
        self.n_obs = int(obs_offsets[-1])
        nr, rk = getattr(self, "n_ranks", 1), getattr(self, "rank", 0)
        if nr > 1 and not self.sharded():     # window too small to shard: every rank holds all of it
            nr, rk = 1, 0
        a, b = shard_range(obs_offsets, rk, nr)
        self.n_obs_local = int(obs_offsets[b] - obs_offsets[a])
        self.n_points_local = int(b - a)

    def eval(self, want_residuals: bool = True) -> dict:
        F, n, nnz = self.n_frames, self.n_points, self.n_obs
        out = dict(U=np.zeros((F, 6, 6)), gc=np.zeros((F, 6)), V=np.zeros((n, 3, 3)), gp=np.zeros((n, 3)),
                   W=np.zeros((nnz, 6, 3)), obs_sqnorm=np.zeros(nnz))
        if want_residuals:
            out["residuals"] = np.zeros((nnz, self.CP))
        eo = EvalOut(U=out["U"].ctypes.data, gc=out["gc"].ctypes.data, V=out["V"].ctypes.data,
                     gp=out["gp"].ctypes.data, W=out["W"].ctypes.data, obs_sqnorm=out["obs_sqnorm"].ctypes.data,
                     residuals=out["residuals"].ctypes.data if want_residuals else None)
        _check(lib().pba_eval(self._h, C.byref(eo)), "pba_eval")
        out["cost"] = float(eo.cost)
        out["device_ms"] = float(eo.device_ms)
        return out

    def eval_timed(self, iters: int) -> float:
        ms = C.c_double()
        _check(lib().pba_eval_timed(self._h, int(iters), C.byref(ms)), "pba_eval_timed")
        return float(ms.value)

    def solve(self, **opt_overrides):
        opt = SolverOptions()
        lib().pba_default_solver_options(C.byref(opt))
        for k, v in opt_overrides.items():
            setattr(opt, k, v)
        summ = Summary()
        _check(lib().pba_solve(self._h, C.byref(opt), C.byref(summ)), "pba_solve")
        sd = {f[0]: getattr(summ, f[0]) for f in Summary._fields_}
        sd["message"] = summ.message.decode()
        return sd

    def save_state(self):
        _check(lib().pba_save_state(self._h), "pba_save_state")

    def restore_state(self):
        _check(lib().pba_restore_state(self._h), "pba_restore_state")

    def begin_batch(self):
        """Uploads until the next solve() only enqueue their copies (the arrays passed must stay alive until then)."""
        _check(lib().pba_begin_batch(self._h), "pba_begin_batch")
        self._batch_keepalive = []

    def end_batch(self):
        _check(lib().pba_end_batch(self._h), "pba_end_batch")
        self._batch_keepalive = []

    def set_frame_u8_ex(self, slot: int, image: np.ndarray, levels_down: int = 0, descriptor_type: int = 0):
        img = np.ascontiguousarray(image, dtype=np.uint8)
        getattr(self, "_batch_keepalive", []).append(img)
        _check(lib().pba_set_frame_u8_ex(self._h, int(slot), _ptr(img), img.shape[0], img.shape[1], int(levels_down), int(descriptor_type)),
               "pba_set_frame_u8_ex")

    def get_results(self):
        cams, pts = np.zeros((self.n_frames, 6)), np.zeros((self.n_points, 3))
        _check(lib().pba_get_results(self._h, _ptr(cams), _ptr(pts)), "pba_get_results")
        return cams, pts

    def copy_state_from(self, src: "Handle"):
        """Poses and points of `src` (same window at another pyramid level) become this handle's, on the device."""
        _check(lib().pba_copy_state(self._h, src._h), "pba_copy_state")

    def comm_init(self, unique_id: bytes, rank: int, n_ranks: int):
        """One process per GPU: join the window's communicator (before set_points)."""
        assert len(unique_id) == PBA_UNIQUE_ID_BYTES
        buf = C.create_string_buffer(unique_id, PBA_UNIQUE_ID_BYTES)
        _check(lib().pba_comm_init(self._h, buf, int(rank), int(n_ranks)), "pba_comm_init")
        self.rank, self.n_ranks = rank, n_ranks

    def graph_counters(self):
        """(builds, in-place updates) of the handle's LM-loop graph."""
        b, u = C.c_int32(0), C.c_int32(0)
        _check(lib().pba_graph_counters(self._h, C.byref(b), C.byref(u)), "pba_graph_counters")
        return b.value, u.value

    @staticmethod
    def comm_init_local(handles):
        """Handles of THIS process, one per device, joined without NCCL / IPC; solve them with solve_all()."""
        arr = (C.c_void_p * len(handles))(*[h._h for h in handles])
        _check(lib().pba_comm_init_local(arr, len(handles)), "pba_comm_init_local")
        for r, h in enumerate(handles):
            h.rank, h.n_ranks = r, len(handles)

    @staticmethod
    def solve_all(handles, **kw):
        """pba_solve on every member concurrently, one thread per handle (ctypes releases the GIL)."""
        import threading
        out, err = [None] * len(handles), [None] * len(handles)

        def run(i):
            try:
                out[i] = handles[i].solve(**kw)
            except Exception as e:  # noqa: BLE001
                err[i] = e
        ts = [threading.Thread(target=run, args=(i,)) for i in range(len(handles))]
        for th in ts:
            th.start()
        for th in ts:
            th.join()
        for e in err:
            if e is not None:
                raise e
        return out

    def sharded(self) -> bool:
        return bool(lib().pba_comm_sharded(self._h))

    def speculates(self) -> bool:
        return bool(lib().pba_comm_speculates(self._h))

    def exchange_kind(self) -> str:
        return {0: "none", 1: "peer-memory", 2: "nccl"}[int(lib().pba_comm_exchange_kind(self._h))]

    def get_poses(self) -> np.ndarray:
        cams = np.zeros((self.n_frames, 6))
        _check(lib().pba_get_poses(self._h, _ptr(cams)), "pba_get_poses")
        return cams

    def get_points(self) -> np.ndarray:
        pts = np.zeros((self.n_points, 3))   # the FULL array on every rank
        _check(lib().pba_get_points(self._h, _ptr(pts)), "pba_get_points")
        return pts

    def get_iterations(self) -> list[dict]:
        n = C.c_int32()
        _check(lib().pba_get_iterations(self._h, None, 0, C.byref(n)), "pba_get_iterations")
        arr = (IterationSummary * max(1, n.value))()
        _check(lib().pba_get_iterations(self._h, arr, n.value, C.byref(n)), "pba_get_iterations")
        return [{f[0]: getattr(arr[i], f[0]) for f in IterationSummary._fields_} for i in range(n.value)]


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(PBA_UNIQUE_ID_BYTES)
    _check(lib().pba_comm_unique_id(buf), "pba_comm_unique_id")
    return buf.raw


def shard_range(obs_offsets: np.ndarray, rank: int, n_ranks: int) -> tuple[int, int]:
    off = np.ascontiguousarray(obs_offsets, dtype=np.int32)
    a, b = C.c_int32(), C.c_int32()
    _check(lib().pba_shard_range(off.shape[0] - 1, _ptr(off), rank, n_ranks, C.byref(a), C.byref(b)), "pba_shard_range")
    return a.value, b.value


def pyrdown_u8(img: np.ndarray, device: int = -1) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    r, c = img.shape
    out = np.zeros(((r + 1) // 2, (c + 1) // 2), dtype=np.uint8)
    _check(lib().pba_pyrdown_u8(_ptr(img), r, c, _ptr(out), device), "pba_pyrdown_u8")
    return out
