#!/usr/bin/env python
"""bench.py — headline benchmark of the photometric-BA inner loop on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Headline workload (BASELINE.json configs[1]/[2]): synthetic 8-frame x 4 000-point x
5x5-patch window at KITTI size (32 000 observations, 800 000 residuals).  One *step* = one full
Levenberg-Marquardt solve of the window from its initial poses/points through the C ABI (pba_solve):
per LM iteration K_A (back-substitution + residual/Jacobian/blocks) and K_B (decision + Schur
elimination + reduced camera solve).

    value      point-residual evaluations / second = numResiduals x K_A passes / device time, inputs
               resident in HBM (pba_restore_state between steps), device time from the CUDA events the
               library records on its own stream around the solve
    e2e        same metric through the same C ABI with HOST buffers: per step the frames, poses, points,
               descriptors are copied host->device from pinned memory, the window is solved, poses +
               points are read back; host wall clock
    roofline   the time-dominant kernel against the measured HBM peak (MEASURED_PEAKS.json), with the
               other kernel's and the whole solve's fractions beside it; algorithmic bytes per SURVEY §8d:
               K_A 656 B/observation, K_B 112 B/observation + 12 B/point
    cfg4       second workload (BASELINE.json configs[3]): 16 frames x 16 000 points x 3 pyramid levels,
               reported at every N; with N > 1 the points of the window are sharded over the GPUs
    cpu_baseline  the CPU oracle restating the reference's Ceres/autodiff path (Ceres itself is not
               installable here), all host cores, a bounded sample of the same window

`--impl reference` times that CPU path as the step itself (rank 0 only), same --steps / --warmup.
N > 1: one process per GPU (torchrun); see DESIGN.md §7 for what is sharded and exchanged.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_OBS_KA = 656          # 544 B read + 112 B written per observation
ALGO_BYTES_PER_OBS_KB = 112          # the Schur/solve pass re-reads the 112 B/observation ...
ALGO_BYTES_PER_POINT_KB = 12         # ... and writes 12 B/point
HBM_FALLBACK_GBS = 6650.0            # /opt/skills/guides/B200_PROFILING.md fallback

WORKLOAD = "8-frame x 4000-point x 5x5-patch synthetic window at KITTI size, full LM solve (BASELINE configs[2])"


def workload_config(n_obs: int, n_res: int) -> dict:
    """The `config` object: identical in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "n_observations": n_obs, "n_residuals": n_res,
            "cache_hygiene": "GPU arm: L2 flushed (256 MiB write) between timed steps, outside the timed intervals"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def static_traffic(name: str):
    """dram__bytes_read+write per launch from the committed ncu --set full capture of this round (profiles/)."""
    p = os.path.join(ROOT, "profiles", name)
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["dram_bytes_per_launch"])
        except Exception:
            return None
    return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        if self._run_nvml():
            return
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def _run_nvml(self):
        """NVML in-process (a query takes ~0.1 ms, so a 20 ms timed region gets real samples); False -> nvidia-smi."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            dev = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            mx = nv.nvmlDeviceGetMaxClockInfo(dev, nv.NVML_CLOCK_SM)
            bits = [("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap)]
        except Exception:
            return False
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(dev, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(dev)
                pw = nv.nvmlDeviceGetPowerUsage(dev) / 1000.0
                self.samples.append([str(self.gpu), str(sm), str(mx), str(pw), hex(r)] + ["Active" if r & b else "Not Active" for _, b in bits])
            except Exception:
                pass
            self.stop_flag.wait(0.002)
        return True

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        mx = [float(s[2]) for s in self.samples if s[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for s in self.samples:
            for n, v in zip(names, s[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def build_window():
    import numpy as np
    from workloads import synthetic
    cache = os.path.join("/tmp", "pba_cfg3_images_v1.npy")
    images = None
    if os.path.exists(cache):
        try:
            images = np.load(cache)
        except Exception:
            images = None
    win = synthetic.make_window(images=images)
    if images is None:
        try:
            np.save(cache, win.images)
        except Exception:
            pass
    return win


def build_cfg4_levels(levels: int = 3):
    """BASELINE configs[3]: 16 frames x 16 000 points, `levels` pyramid levels (level l = frames reduced l times with
    the cv::pyrDown rule, intrinsics halved, descriptors re-extracted; workloads/synthetic.py).  Cached in /tmp."""
    import pickle
    from workloads import synthetic
    cache = os.path.join("/tmp", f"pba_cfg4_levels{levels}_v1.pkl")
    if os.path.exists(cache):
        try:
            return pickle.load(open(cache, "rb"))
        except Exception:
            pass
    import numpy as np
    win = synthetic.make_window(n_frames=16, grid=(100, 160))
    px, ref = synthetic.reference_pixels(win)
    imgs = [win.images]
    for _ in range(levels - 1):
        imgs.append(np.stack([synthetic.pyr_down_u8(i) for i in imgs[-1]]))
    wins = [synthetic.pyramid_level(win, lv, imgs[lv], px, ref) for lv in range(levels)]
    try:
        tmp = cache + f".{os.getpid()}"
        pickle.dump(wins, open(tmp, "wb"))
        os.replace(tmp, cache)
    except Exception:
        pass
    return wins


def cpu_oracle_run(win, steps: int, warmup: int, threads: int = 0):
    """Times oracle_solve (Jet<double,9> autodiff structure) on the host cores."""
    from oracle import binding as ob
    # explicit thread count: torch.distributed.run exports OMP_NUM_THREADS=1, which would otherwise turn the
    # "all host cores" baseline into a single-threaded one whenever the bench runs under torchrun
    ow = ob.OracleWindow(win, num_threads=threads or (os.cpu_count() or 1))
    evals, iters, secs = 0, 0, 0.0
    summ = None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, _, summ, _ = ow.solve(win.cams_init, win.points_init)
        dt = time.perf_counter() - t0
        if i >= warmup:
            secs += dt
            evals += summ["num_jacobian_evals"] + summ["num_cost_evals"]
            iters += summ["num_iterations"] - 1
    return {"residual_evals": evals * win.n_residuals, "lm_iters": iters, "seconds": secs, "summary": summ}


def make_handle(capi, win, local_rank, rank, world, comm_id, levels_down=0, images0=None):
    h = capi.Handle(win.rows, win.cols, win.fx, win.fy, win.cx, win.cy, radius=win.radius, huber=win.huber,
                    max_frames=win.n_frames, max_points=win.n_points, max_observations=win.n_obs, device=local_rank)
    if world > 1:
        h.comm_init(comm_id, rank, world)
    if levels_down:
        h.set_frames_u8_pyr(images0, levels_down)      # level-0 frames in, reduced on the device
    else:
        h.set_frames_u8(win.images)
    h.set_poses(win.cams_init, win.fixed_frame)
    h.set_points(win.points_init, win.desc, win.obs_offsets, win.obs_frame, win.weights)
    return h


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cfg4", action="store_true", help="skip the second workload (developer runs)")
    args = ap.parse_args()
    steps, warmup = max(1, args.steps), max(0, args.warmup)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        win = build_window()
        r = cpu_oracle_run(win, steps, warmup)     # one step = one full CPU solve of the same window (~50 ms on 16 cores)
        val = r["residual_evals"] / r["seconds"]
        line = {
            "impl": "reference", "metric": "point_residual_evaluations_per_sec", "value": val, "unit": "residuals/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * r["seconds"] / steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64 (fp32 sampler)",
            "data": "synthetic", "lm_iters_per_sec": r["lm_iters"] / r["seconds"],
            "config": workload_config(win.n_obs, win.n_residuals),
            "what": "CPU oracle restating the reference's Ceres/autodiff path (Ceres 1.x, Eigen, Boost, OpenCV are not "
                    "installable in this image, so the reference binary cannot run); one step = one full LM solve",
            "cpu_baseline": {"value": val, "unit": "residuals/s", "cores": ncores, "kind": "port",
                             "sample": f"{steps} full LM solves of the bench window, OpenMP over {ncores} host threads; "
                                       f"final cost {r['summary']['final_cost']:.4f}"},
            "e2e": {"value": val, "unit": "residuals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import numpy as np
    import torch
    import __graft_entry__ as ge

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ge.build()
    from photobundle_b200 import capi

    def comm_id():
        if world == 1:
            return None
        ids = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        return ids[0]

    win = build_window()
    # N > 1: one window sharded by point block over the ranks (frames/poses replicated)
    h = make_handle(capi, win, local_rank, rank, world, comm_id())
    h.save_state()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def one_step():
        h.restore_state()
        flush.fill_(1)            # flush L2 outside the timed interval
        torch.cuda.synchronize()
        return h.solve()

    nwarm = max(3, warmup)
    for _ in range(nwarm):
        one_step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    dev_s, evals, iters, launches, xchg, decisions, kb_s, kb_n = 0.0, 0, 0, 0, 0, 0, 0.0, 0
    last = None
    for _ in range(steps):
        s = one_step()
        dev_s += s["device_time_in_seconds"]
        evals += s["num_evaluations"]
        iters += s["num_iterations"] - 1
        launches += s["kernel_launches"]
        xchg += s["num_collectives"]
        decisions += s["num_evaluations"]
        kb_s += s["kb_device_time_in_seconds"]
        kb_n += s["num_evaluations"] - 1      # the last K_B only takes the terminating decision
        last = s
    barrier()

    # K_A alone (config 2): average launch duration, CUDA events on the library's stream
    h.restore_state()
    ka_iters = 200
    h.eval_timed(20)
    ka_ms_cold = []
    for _ in range(10):           # cold-L2 launches: flush, then time ONE launch
        flush.fill_(2)
        torch.cuda.synchronize()
        ka_ms_cold.append(h.eval_timed(1))
    ka_ms_warm = h.eval_timed(ka_iters) / ka_iters
    launches += 20 + 10 + ka_iters
    ka_ms = statistics.median(ka_ms_cold)

    # batched windows (auxiliary): B independent windows solved concurrently, one handle + stream + host
    # thread each — a single 8x4k window is latency-bound and leaves most of the GPU idle
    batched = None
    if True:   # at N GPUs: B windows per GPU, no communication at all (what a multi-GPU box is good for with windows this small)
        import threading as _th
        B = 4
        hs = [capi.Handle.for_window(win, device=local_rank) for _ in range(B)]
        for hb in hs:
            hb.save_state()
            hb.solve()
        res = [None] * B
        def _work(i):
            ev = 0
            for _ in range(steps):
                hs[i].restore_state()
                ev += hs[i].solve()["num_evaluations"]
            res[i] = ev
        torch.cuda.synchronize()
        barrier()
        tb = time.perf_counter()
        ths = [_th.Thread(target=_work, args=(i,)) for i in range(B)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        torch.cuda.synchronize()
        tb = time.perf_counter() - tb
        ev_all = float(sum(res))
        if dist is not None:
            tt = torch.tensor([tb], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            te = torch.tensor([ev_all], device="cuda", dtype=torch.float64)
            dist.all_reduce(te, op=dist.ReduceOp.SUM)
            tb, ev_all = float(tt[0]), float(te[0])
        batched = {"windows_in_flight": B * world, "residuals_per_sec": ev_all * win.n_residuals / tb,
                   "solves_per_sec": B * world * steps / tb,
                   "what": f"{B} independent windows per GPU on {world} GPU(s), one handle + stream + host thread each, no communication",
                   "timing": "host wall clock around all threads, max over ranks"}
        launches += sum(res) * 2
        for hb in hs:
            hb.close()

    # e2e through the C ABI with host buffers (pinned), wall clock
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    images, cams0, pts0, desc = pin(win.images), pin(win.cams_init), pin(win.points_init), pin(win.desc)
    obs_off, obs_frame, weights = pin(win.obs_offsets), pin(win.obs_frame), pin(win.weights)
    h2d = images.nbytes + cams0.nbytes * 2 + pts0.nbytes * 2 + desc.nbytes // 2 + obs_off.nbytes + obs_frame.nbytes + weights.nbytes
    d2h = cams0.nbytes + pts0.nbytes

    def e2e_step():
        h.begin_batch()           # uploads are enqueued; the (pinned) host buffers stay alive until solve() returns
        h.set_frames_u8(images)
        h.set_poses(cams0, win.fixed_frame)
        h.set_points(pts0, desc, obs_off, obs_frame, weights)
        s = h.solve()
        h.get_results()           # poses + points, one synchronisation
        return s

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_evals, e2e_iters = 0, 0
    for _ in range(steps):
        s = e2e_step()
        e2e_evals += s["num_evaluations"]
        e2e_iters += s["num_iterations"] - 1
        launches += s["kernel_launches"]
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    # the same call sequence as the sliding-window host class issues it: the window stays resident in its ring slots and
    # only ONE frame is new per solve (src/photobundle.cc:608-612: one addFrame -> one optimize); reported beside the
    # headline e2e, which re-sends the whole window every step
    def e2e_ring_step(k):
        h.begin_batch()
        h.set_frame_u8_ex(k % win.n_frames, images[k % win.n_frames])
        h.set_poses(cams0, win.fixed_frame)
        h.set_points(pts0, desc, obs_off, obs_frame, weights)
        s = h.solve()
        h.get_results()
        return s
    for k in range(2):
        e2e_ring_step(k)
    barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        launches += e2e_ring_step(k)["kernel_launches"]
    torch.cuda.synchronize()
    e2e_ring_s = time.perf_counter() - t0
    barrier()
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    n_obs_local = h.n_obs_local if world > 1 else win.n_obs
    n_pts_local = getattr(h, "n_points_local", win.n_points) if world > 1 else win.n_points
    exchange_kind = h.exchange_kind()
    speculates = h.speculates()
    sharded = h.sharded()
    # the library does not shard a window that fits one K_A wave (pba_comm_sharded); the north-star's sharded layout is
    # measured beside it by forcing it
    forced = None
    if world > 1 and not sharded:
        os.environ["PBA_MGPU_REPLICATE"] = "0"
        h.set_poses(win.cams_init, win.fixed_frame)
        h.set_points(win.points_init, win.desc, win.obs_offsets, win.obs_frame, win.weights)
        h.save_state()
        f_dev, f_x = 0.0, 0
        for k in range(nwarm + steps):
            h.restore_state()
            flush.fill_(k & 0xff)
            torch.cuda.synchronize()
            barrier()
            s = h.solve()
            if k >= nwarm:
                f_dev += s["device_time_in_seconds"]
                f_x += s["num_collectives"]
        del os.environ["PBA_MGPU_REPLICATE"]
        if dist is not None:
            tt = torch.tensor([f_dev], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            f_dev = float(tt[0])
        forced = {"ms_per_step": 1e3 * f_dev / steps, "value": evals * win.n_residuals / f_dev if f_dev > 0 else None,
                  "exchanges_per_solve": f_x / steps, "what": "the same window FORCED into point shards over the ranks "
                  "(PBA_MGPU_REPLICATE=0): one in-kernel exchange per LM iteration over NVLink peer memory"}
    h.close()

    # ---- second workload: BASELINE configs[3], 16 frames x 16 000 points x 3 levels, coarse to fine ----------
    cfg4 = None
    if not args.no_cfg4:
        wins = build_cfg4_levels(3)
        hl = [make_handle(capi, wins[lv], local_rank, rank, world, comm_id(), levels_down=lv, images0=wins[0].images)
              for lv in range(3)]
        hl[2].save_state()
        c4_steps = max(2, min(steps, 5))
        def cfg4_step():
            hl[2].restore_state()
            flush.fill_(3)
            torch.cuda.synchronize()
            out = []
            for lv in (2, 1, 0):
                if lv < 2:
                    hl[lv].copy_state_from(hl[lv + 1])    # device-to-device hand-over, coarse -> fine
                out.append(hl[lv].solve())
            return out
        for _ in range(2):
            cfg4_step()
        barrier()
        c4_dev, c4_evals, c4_iters, c4_fine, c4_last, c4_kb, c4_fine_evals = 0.0, 0, 0, 0.0, None, 0.0, 0
        for _ in range(c4_steps):
            out = cfg4_step()
            c4_dev += sum(s["device_time_in_seconds"] for s in out)
            c4_fine += out[-1]["device_time_in_seconds"]
            c4_kb += out[-1]["kb_device_time_in_seconds"]
            c4_fine_evals += out[-1]["num_evaluations"]
            c4_evals += sum(s["num_evaluations"] for s in out)
            c4_iters += sum(s["num_iterations"] - 1 for s in out)
            launches += sum(s["kernel_launches"] for s in out)
            c4_last = out
        barrier()
        if dist is not None:
            t = torch.tensor([c4_dev, c4_fine], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            c4_dev, c4_fine = float(t[0]), float(t[1])
        cfg4 = {"workload": "16-frame x 16000-point x 5x5-patch window, 3-level pyramid coarse to fine (BASELINE configs[3]); "
                            "levels handed over on the device (pba_copy_state)",
                "n_observations_per_level": wins[0].n_obs, "n_residuals_per_level": wins[0].n_residuals,
                "ms_per_solve_all_levels": 1e3 * c4_dev / c4_steps, "ms_per_solve_finest_level": 1e3 * c4_fine / c4_steps,
                "residuals_per_sec": c4_evals * wins[0].n_residuals / c4_dev, "lm_iters_per_sec": c4_iters / c4_dev,
                "finest_level_us_per_iteration": {"k_b": 1e6 * c4_kb / max(1, c4_fine_evals - c4_steps),
                                                  "k_a_incl_launch_gaps": 1e6 * (c4_fine - c4_kb) / max(1, c4_fine_evals),
                                                  "note": "this rank's own stamps / events (not max over ranks)"},
                "lm_iterations_per_level": [s["num_iterations"] - 1 for s in c4_last],
                "final_cost_per_level": [s["final_cost"] for s in c4_last], "steps": c4_steps,
                "exchanges_per_solve_all_levels": sum(s["num_collectives"] for s in c4_last),
                "speculation": ("on" if hl[0].speculates() else "off (decision exchanged first: two exchanges per iteration)") if world > 1 else None,
                "timing": "sum of the three levels' CUDA-event intervals, max over ranks"}
        for hx in hl:
            hx.close()

    # max over ranks (device-timed)
    if dist is not None:
        t = torch.tensor([dev_s, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, e2e_s = float(t[0]), float(t[1])
        c = torch.tensor([launches], device="cuda", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        launches = float(c[0])   # evals / iters are global already: one window, sharded

    peak, peak_src = measured_peak()
    n_ka, n_kb = last["num_evaluations"], last["num_evaluations"] - 1   # per solve; the last K_B only takes the terminating decision
    ka_bytes = n_obs_local * ALGO_BYTES_PER_OBS_KA
    kb_bytes = n_obs_local * ALGO_BYTES_PER_OBS_KB + n_pts_local * ALGO_BYTES_PER_POINT_KB
    solve_us = 1e6 * dev_s / steps
    kb_us = 1e6 * kb_s / max(1, kb_n)                      # the kernel's own start/end stamps (globaltimer), averaged
    ka_loop_us = (solve_us - n_kb * kb_us) / max(1, n_ka)  # in-loop K_A (with back-substitution) + the two launch gaps of an iteration
    ka_frac = ka_bytes / (ka_ms * 1e-3) / 1e9 / peak
    kb_frac = kb_bytes / (kb_us * 1e-6) / 1e9 / peak
    whole_frac = (n_ka * ka_bytes + n_kb * kb_bytes) / (solve_us * 1e-6) / 1e9 / peak
    ka_dominant = n_ka * ka_loop_us >= n_kb * kb_us
    value = evals * win.n_residuals / dev_s
    line = {
        "metric": "point_residual_evaluations_per_sec", "value": value, "unit": "residuals/s",
        "n_gpus": world, "steps": steps, "warmup": nwarm, "ms_per_step": 1e3 * dev_s / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64 (fp32 sampler)", "data": "synthetic",
        "lm_iters_per_sec": iters / dev_s,
        "config": workload_config(win.n_obs, win.n_residuals),
        "run": {
            "lm_iterations_per_solve": last["num_iterations"] - 1, "ka_passes_per_solve": last["num_evaluations"],
            "final_cost": last["final_cost"], "initial_cost": last["initial_cost"], "termination": last["message"],
            "timing": "sum of per-step CUDA-event intervals recorded by the library on its own stream; "
                      "L2 flushed (256 MiB write) between steps outside the intervals",
            "parallelism": "1 window on 1 GPU" if world == 1 else
                           (f"1 window on {world} GPUs, NOT sharded: all of its points fit one wave of K_A on one GPU, so sharding cannot "
                            f"shorten an LM iteration and every exchange costs more than it saves; every rank solves the whole window "
                            f"(pba_comm_sharded = 0), the forced-sharding figure is under run.sharded_anyway") if not sharded else
                           f"1 window, points sharded over {world} GPUs (frames/poses replicated); per LM iteration the pose "
                           f"blocks + cost and the reduced camera system travel in ONE in-kernel exchange ({exchange_kind}); a second "
                           f"one only when neither speculated outcome of the decision held",
            "sharded_anyway": forced,
            "exchanges_per_solve": xchg / steps if world > 1 else 0,
            "speculation": ("on: both outcomes of the pending decision are eliminated while the evaluation sums travel" if speculates else
                            "off: shard too large, decision exchanged first (two exchanges per iteration)") if (world > 1 and exchange_kind == "peer-memory") else None,
            "mis_speculation_rate": (xchg - decisions) / max(1, decisions) if (world > 1 and exchange_kind == "peer-memory" and speculates) else None,
        },
        "k_a": {"us_per_launch_cold_l2": 1e3 * ka_ms, "us_per_launch_warm_l2": 1e3 * ka_ms_warm,
                "observations_per_launch": n_obs_local,
                "residuals_per_sec": 25 * n_obs_local / (ka_ms * 1e-3), "observations_per_sec": n_obs_local / (ka_ms * 1e-3),
                "what": "evaluation-only launch of k_step (BASELINE configs[1]); in the LM loop the same kernel also back-substitutes"},
        "k_b": {"us_per_launch": kb_us,
                "what": "decision + Schur elimination (DMMA) + reduced camera solve (k_schur_solve): the kernel's own start/end "
                        "stamps (globaltimer), averaged over the LM iterations of the timed solves",
                "algorithmic_bytes_per_launch": kb_bytes},
        "k_a_in_loop": {"us_per_iteration_incl_launch_gaps": ka_loop_us,
                        "what": "(solve time - K_B time) / K_A launches: the in-loop K_A (with back-substitution) plus the two "
                                "kernel-to-kernel gaps of an LM iteration"},
        "roofline": {"bound": "hbm",
                     "kernel": "k_step<2,u8,1> (K_A)" if ka_dominant else "k_schur_solve (K_B)",
                     "achieved": (ka_frac if ka_dominant else kb_frac) * peak, "peak": peak, "unit": "GB/s",
                     "frac": ka_frac if ka_dominant else kb_frac,
                     "traffic": static_traffic("k_a_ncu_summary.json" if ka_dominant else "k_b_ncu_summary.json"),
                     "traffic_source": "static: dram__bytes_read+write per launch of this kernel from this round's ncu --set full "
                                       "capture (profiles/k_a_ncu_summary.json, k_b_ncu_summary.json); not re-measured in this run",
                     "algorithmic_bytes_per_launch": ka_bytes if ka_dominant else kb_bytes, "peak_source": peak_src,
                     "k_a": {"frac": ka_frac, "share_of_solve_incl_launch_gaps": n_ka * ka_loop_us / solve_us, "traffic": static_traffic("k_a_ncu_summary.json")},
                     "k_b": {"frac": kb_frac, "share_of_solve": n_kb * kb_us / solve_us, "traffic": static_traffic("k_b_ncu_summary.json")},
                     "whole_solve": {"frac": whole_frac, "algorithmic_bytes": n_ka * ka_bytes + n_kb * kb_bytes},
                     "tensor_pipe": "fp64 DMMA (mma.sync.m8n8k4.f64) carries the Schur products Zt Zt^T and the trailing updates of the "
                                    "reduced factorisation inside K_B; utilisation in profiles/ (ncu sm__pipe_tensor_subpipe_dmma_cycles_active)",
                     "note": "K_A duration = median of 10 single launches after an L2 flush, CUDA events on the launching stream"},
        "e2e": {"value": e2e_evals * win.n_residuals / e2e_s, "unit": "residuals/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s / steps, "lm_iters_per_sec": e2e_iters / e2e_s,
                "resident_window": {"ms_per_step": 1e3 * e2e_ring_s / steps, "value": e2e_evals * win.n_residuals / e2e_ring_s,
                                    "h2d_bytes_per_step": int(h2d - images.nbytes + images[0].nbytes),
                                    "what": "as the sliding-window host class calls it: frames stay in their device ring slots, one "
                                            "new frame per solve (pba_set_frame_u8_ex) + poses, points, descriptors; same solve, same read-back"}},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }
    if batched is not None:
        line["batched"] = batched
    if cfg4 is not None:
        line["cfg4"] = cfg4

    # CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload
    if world == 1 and rank == 0:
        r = cpu_oracle_run(win, 3, 1)
        r4 = cpu_oracle_run(win, 2, 1, threads=4)      # the reference caps Ceres at 4 threads (src/photobundle.cc:823-829)
        line["cpu_baseline"] = {
            "value": r["residual_evals"] / r["seconds"], "unit": "residuals/s", "cores": ncores, "kind": "port",
            "lm_iters_per_sec": r["lm_iters"] / r["seconds"],
            "at_reference_thread_cap": {"cores": 4, "value": r4["residual_evals"] / r4["seconds"], "lm_iters_per_sec": r4["lm_iters"] / r4["seconds"]},
            "sample": f"3 full LM solves of the same window by the CPU oracle (reference's Ceres/autodiff structure), "
                      f"{ncores} OpenMP threads; final cost {r['summary']['final_cost']:.4f}"}
        line["parity"] = {"gpu_final_cost": last["final_cost"], "cpu_final_cost": r["summary"]["final_cost"],
                          "rel_cost_diff": abs(last["final_cost"] - r["summary"]["final_cost"]) / r["summary"]["final_cost"],
                          "note": "oracle = restated reference path; Ceres semantics themselves are unpinned (DESIGN.md §2)"}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
