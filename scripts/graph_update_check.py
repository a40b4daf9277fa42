"""Solves of alternating sizes on one handle: the instantiated LM-loop graph must be updated in place (PBA_DEBUG_GRAPH=1
prints the counters at pba_destroy) and every solve must equal the one a fresh handle gives for the same window."""
import dataclasses
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from photobundle_b200 import capi  # noqa: E402
from workloads import synthetic  # noqa: E402


def sub(win, n):
    o = int(win.obs_offsets[n])
    return dataclasses.replace(win, points_init=win.points_init[:n], points_gt=win.points_gt[:n], desc=win.desc[:n],
                               obs_offsets=win.obs_offsets[:n + 1], obs_frame=win.obs_frame[:o])


w = synthetic.make_window()
wins = [w, sub(w, w.n_points - 40), sub(w, 2000), sub(w, 150), w]
ref = []
for ww in wins:
    h = capi.Handle.for_window(ww)
    s = h.solve()
    ref.append((s["final_cost"], s["num_iterations"], h.get_poses()))
    h.close()
h = capi.Handle.for_window(w)
for rnd in range(3):
    for ww, (c, n, poses) in zip(wins, ref):
        h.set_poses(ww.cams_init, ww.fixed_frame)
        h.set_points(ww.points_init, ww.desc, ww.obs_offsets, ww.obs_frame, ww.weights)
        s = h.solve()
        assert s["num_iterations"] == n and abs(s["final_cost"] - c) <= 1e-9 * c, (ww.n_points, s["final_cost"], c)
        np.testing.assert_allclose(h.get_poses(), poses, atol=1e-8)
        over = 1e6 * (s["total_time_in_seconds"] - s["device_time_in_seconds"])
        if rnd == 2:
            print(f"n_points {ww.n_points:5d}: {n} iterations, host overhead of pba_solve {over:.1f} us")
h.close()
print("OK")
