#!/usr/bin/env python
"""bench.py — headline benchmark of the photometric-BA inner loop on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]/[2]): synthetic 8-frame x 4 000-point x
5x5-patch window at KITTI size (32 000 observations, 800 000 residuals).  One *step* = one
full Levenberg-Marquardt solve of the window from its initial poses/points through the
C ABI (pba_solve): K1 residual+Jacobian passes + Schur + reduced solve per iteration.

    value      point-residual evaluations / second = numResiduals x K1 passes / device time,
               inputs resident in HBM (pba_restore_state between steps), device time from the
               CUDA events the library records on its own stream around the solve
    e2e        same metric through the same C ABI with HOST buffers: per step the frames,
               poses, points, descriptors are copied host->device from pinned memory, the
               window is solved, poses+points are read back; host wall clock
    roofline   K1 (the dominant kernel): algorithmic bytes 656 B/observation x
               32 000 / the live CUDA-event launch duration; peak = MEASURED_PEAKS.json hbm_gbs
    cpu_baseline  the CPU oracle restating the reference's Ceres/autodiff path (Ceres itself is
               not installable here), all host cores, one bounded sample of the same window

`--impl reference` times that CPU path as the step itself (rank 0 only).
N > 1: one process per GPU (torchrun); see DESIGN.md §multi-GPU for what is sharded.
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

ALGO_BYTES_PER_OBS_K1 = 656          # 544 B read + 112 B written per observation
ALGO_BYTES_PER_OBS_LM_ITER = 916     # one Jacobian pass + one cost pass (reference structure)
HBM_FALLBACK_GBS = 6650.0            # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def k1_traffic_from_profile():
    """dram__bytes_read+write per K1 launch from the committed ncu capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "k1_ncu_summary.json")
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
    from photobundle_b200 import synthetic
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
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
        k = min(steps, 5)  # bounded: each step is one full CPU solve (~0.2-1 s)
        r = cpu_oracle_run(win, k, min(warmup, 1))
        val = r["residual_evals"] / r["seconds"]
        line = {
            "impl": "reference", "metric": "point_residual_evaluations_per_sec", "value": val, "unit": "residuals/s",
            "n_gpus": args.gpus, "steps": k, "warmup": min(warmup, 1), "ms_per_step": 1e3 * r["seconds"] / k,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64 (fp32 sampler)",
            "data": "synthetic", "lm_iters_per_sec": r["lm_iters"] / r["seconds"],
            "config": {"workload": "8-frame x 4000-point x 5x5 window, full LM solve (BASELINE configs[2])",
                       "what": "CPU oracle restating the reference's Ceres/autodiff path (Ceres 1.x, Eigen, Boost, OpenCV "
                               "are not installable in this image, so the reference binary cannot run)",
                       "n_observations": win.n_obs, "n_residuals": win.n_residuals},
            "cpu_baseline": {"value": val, "unit": "residuals/s", "cores": ncores, "kind": "port",
                             "sample": f"{k} full LM solves of the bench window, OpenMP over {ncores} host threads"},
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

    win = build_window()
    if world > 1:
        # one window sharded by point block over the ranks (frames/poses replicated); the exchange per LM
        # iteration is the all-reduce of the pose blocks + cost and of the reduced camera system
        ids = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        h = capi.Handle(win.rows, win.cols, win.fx, win.fy, win.cx, win.cy, radius=win.radius, huber=win.huber,
                        max_frames=win.n_frames, max_points=win.n_points, max_observations=win.n_obs, device=local_rank)
        h.comm_init(ids[0], rank, world)
        h.set_frames_u8(win.images)
        h.set_poses(win.cams_init, win.fixed_frame)
        h.set_points(win.points_init, win.desc, win.obs_offsets, win.obs_frame, win.weights)
    else:
        h = capi.Handle.for_window(win, device=local_rank)
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

    for _ in range(max(3, warmup)):
        one_step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    dev_s, evals, iters, launches = 0.0, 0, 0, 0
    last = None
    for _ in range(steps):
        s = one_step()
        dev_s += s["device_time_in_seconds"]
        evals += s["num_evaluations"]
        iters += s["num_iterations"] - 1
        launches += s["kernel_launches"]
        last = s
    barrier()

    # K1 alone (config 2): average launch duration, CUDA events on the library's stream
    h.restore_state()
    k1_iters = 200
    h.eval_timed(20)
    k1_ms_cold = []
    for _ in range(10):           # cold-L2 launches: flush, then time ONE launch
        flush.fill_(2)
        torch.cuda.synchronize()
        k1_ms_cold.append(h.eval_timed(1))
    k1_ms_warm = h.eval_timed(k1_iters) / k1_iters
    launches += 20 + 10 + k1_iters
    k1_ms = statistics.median(k1_ms_cold)

    # batched windows (auxiliary): B independent windows solved concurrently, one handle + stream + host
    # thread each — a single 8x4k window is latency-bound and leaves most of the GPU idle
    batched = None
    if world == 1:
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
        tb = time.perf_counter()
        ths = [_th.Thread(target=_work, args=(i,)) for i in range(B)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        torch.cuda.synchronize()
        tb = time.perf_counter() - tb
        batched = {"windows_in_flight": B, "residuals_per_sec": sum(res) * win.n_residuals / tb,
                   "solves_per_sec": B * steps / tb, "timing": "host wall clock around all threads"}
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
        h.set_frames_u8(images)
        h.set_poses(cams0, win.fixed_frame)
        h.set_points(pts0, desc, obs_off, obs_frame, weights)
        s = h.solve()
        h.get_poses()
        h.get_points()
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
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    # max over ranks (device-timed)
    if dist is not None:
        t = torch.tensor([dev_s, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s, e2e_s = float(t[0]), float(t[1])
        c = torch.tensor([launches], device="cuda", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        launches = float(c[0])   # evals / iters are global already: one window, sharded

    peak, peak_src = measured_peak()
    n_obs_local = h.n_obs_local if world > 1 else win.n_obs
    achieved = n_obs_local * ALGO_BYTES_PER_OBS_K1 / (k1_ms * 1e-3) / 1e9
    value = evals * win.n_residuals / dev_s
    line = {
        "metric": "point_residual_evaluations_per_sec", "value": value, "unit": "residuals/s",
        "n_gpus": world, "steps": steps, "warmup": max(3, warmup), "ms_per_step": 1e3 * dev_s / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64 (fp32 sampler)", "data": "synthetic",
        "lm_iters_per_sec": iters / dev_s,
        "config": {
            "workload": "8-frame x 4000-point x 5x5 window, full LM solve incl. Schur + reduced-pose solve "
                        "(BASELINE configs[2]; K1-only numbers of configs[1] under `k1`)",
            "n_observations": win.n_obs, "n_residuals": win.n_residuals,
            "lm_iterations_per_solve": last["num_iterations"] - 1, "k1_passes_per_solve": last["num_evaluations"],
            "final_cost": last["final_cost"], "initial_cost": last["initial_cost"], "termination": last["message"],
            "timing": "sum of per-step CUDA-event intervals recorded by the library on its own stream; "
                      "L2 flushed (256 MiB write) between steps outside the intervals",
            "parallelism": "1 window on 1 GPU" if world == 1 else
                           f"1 window, points sharded over {world} GPUs (frames/poses replicated); per LM iteration the pose "
                           f"blocks+cost and the reduced camera system are exchanged by {h.exchange_kind()} "
                           f"({last['num_collectives']} NCCL collectives/solve)",
        },
        "k1": {"us_per_launch_cold_l2": 1e3 * k1_ms, "us_per_launch_warm_l2": 1e3 * k1_ms_warm,
               "observations_per_launch": n_obs_local,
               "residuals_per_sec": 25 * n_obs_local / (k1_ms * 1e-3), "observations_per_sec": n_obs_local / (k1_ms * 1e-3)},
        "k_b": {"us_per_launch_incl_launch_gaps": (1e6 * dev_s / steps - last["num_evaluations"] * 1e3 * k1_ms_warm) / max(1, last["num_evaluations"]),
                "what": "decision + Schur elimination + reduced solve (k_schur_solve); latency-bound single-CTA tail, see DESIGN.md §4; "
                        "derived as (solve time - K_A launches x warm K_A time) / K_B launches (one decision per evaluation)",
                "algorithmic_bytes_per_launch": n_obs_local * 112 + (win.n_points // world) * 12},
        "roofline": {"bound": "hbm", "kernel": "k_step<2,u8,1> (K_A)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": k1_traffic_from_profile(),
                     "algorithmic_bytes_per_launch": n_obs_local * ALGO_BYTES_PER_OBS_K1, "peak_source": peak_src,
                     "note": "duration = median of 10 single launches after an L2 flush, CUDA events on the launching stream"},
        "e2e": {"value": e2e_evals * win.n_residuals / e2e_s, "unit": "residuals/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s / steps, "lm_iters_per_sec": e2e_iters / e2e_s},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }
    if batched is not None:
        line["batched"] = batched

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
                          "rel_cost_diff": abs(last["final_cost"] - r["summary"]["final_cost"]) / r["summary"]["final_cost"]}
    if rank == 0:
        print(json.dumps(line))
    h.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
