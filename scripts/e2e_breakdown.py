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
