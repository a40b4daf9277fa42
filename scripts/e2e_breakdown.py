"""Scratch: wall-clock breakdown of the e2e path (C ABI with host buffers), cfg3 window."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from photobundle_b200 import capi
from workloads import synthetic
w = synthetic.make_window()
h = capi.Handle.for_window(w)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
images, cams0, pts0, desc = pin(w.images), pin(w.cams_init), pin(w.points_init), pin(w.desc)
obs_off, obs_frame, weights = pin(w.obs_offsets), pin(w.obs_frame), pin(w.weights)
names = ["set_frames_u8", "set_poses", "set_points", "solve", "get_poses", "get_points"]
acc = dict.fromkeys(names, 0.0)
N = 30
for it in range(N + 3):
    t = [time.perf_counter()]
    h.set_frames_u8(images); t.append(time.perf_counter())
    h.set_poses(cams0, w.fixed_frame); t.append(time.perf_counter())
    h.set_points(pts0, desc, obs_off, obs_frame, weights); t.append(time.perf_counter())
    s = h.solve(); t.append(time.perf_counter())
    h.get_poses(); t.append(time.perf_counter())
    h.get_points(); t.append(time.perf_counter())
    if it >= 3:
        for k, n in enumerate(names):
            acc[n] += t[k + 1] - t[k]
tot = sum(acc.values())
for n in names:
    print(f"{n:14s} {1e6 * acc[n] / N:8.1f} us")
print(f"{'total':14s} {1e6 * tot / N:8.1f} us   (solve device time {1e6 * s['device_time_in_seconds']:.1f} us)")

# the batched path bench.py's e2e uses: uploads only enqueue, one synchronisation in get_results
accb = dict.fromkeys(["begin+set_frames", "set_poses", "set_points", "solve (host wall)", "get_results"], 0.0)
dev = 0.0
for it in range(N + 3):
    t = [time.perf_counter()]
    h.begin_batch(); h.set_frames_u8(images); t.append(time.perf_counter())
    h.set_poses(cams0, w.fixed_frame); t.append(time.perf_counter())
    h.set_points(pts0, desc, obs_off, obs_frame, weights); t.append(time.perf_counter())
    s = h.solve(); t.append(time.perf_counter())
    h.get_results(); t.append(time.perf_counter())
    if it >= 3:
        dev += s["device_time_in_seconds"]
        for k, n in enumerate(accb):
            accb[n] += t[k + 1] - t[k]
print("batched:")
for n in accb:
    print(f"{n:18s} {1e6 * accb[n] / N:8.1f} us")
print(f"{'total':18s} {1e6 * sum(accb.values()) / N:8.1f} us   (solve device time {1e6 * dev / N:.1f} us)")

# cost of a solve whose sizes differ from the previous one (the sliding window changes its point count every frame):
# host-side overhead of pba_solve = total_time - device_time, same sizes vs alternating sizes
import dataclasses
def sub(win, n):
    o = int(win.obs_offsets[n])
    return dataclasses.replace(win, points_init=win.points_init[:n], points_gt=win.points_gt[:n], desc=win.desc[:n],
                               obs_offsets=win.obs_offsets[:n + 1], obs_frame=win.obs_frame[:o])
wa, wb = w, sub(w, w.n_points - 40)
for label, seq_ in (("same sizes", [wa, wa]), ("alternating sizes", [wa, wb])):
    over = []
    for it in range(12):
        ww = seq_[it % 2]
        h.set_poses(ww.cams_init, ww.fixed_frame)
        h.set_points(ww.points_init, ww.desc, ww.obs_offsets, ww.obs_frame, ww.weights)
        s = h.solve()
        if it >= 2:
            over.append(1e6 * (s["total_time_in_seconds"] - s["device_time_in_seconds"]))
    print(f"pba_solve host overhead, {label}: median {np.median(over):.1f} us (min {min(over):.1f}, max {max(over):.1f})")
