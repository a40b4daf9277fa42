"""How far apart are the CUDA solve and the oracle on the two chaotic windows once the monocular scale gauge
(a similarity about the fixed camera's centre) is taken out?  Used to set the tolerances of tests/test_gpu_parity.py."""
import dataclasses
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import binding as ob  # noqa: E402
from photobundle_b200 import capi  # noqa: E402
from workloads import synthetic  # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gauge import scale_gauge_diff  # noqa: E402

for name, win in (("ragged", synthetic.small_window(seed=11, ragged=True, n_frames=8, grid=(10, 14))),
                  ("no-loss", dataclasses.replace(synthetic.small_window(), huber=0.0)),
                  ("dense", synthetic.small_window())):
    ow = ob.OracleWindow(win, num_threads=1)
    ocams, opts, osum, otr = ow.solve(win.cams_init, win.points_init)
    h = capi.Handle.for_window(win)
    s = h.solve()
    cams, pts = h.get_poses(), h.get_points()
    h.close()
    d = scale_gauge_diff(cams, pts, ocams, opts, win.fixed_frame)
    print(name, "iters", s["num_iterations"], osum["num_iterations"], "raw rot", np.abs(cams - ocams)[:, :3].max(), "raw t", np.abs(cams - ocams)[:, 3:].max(),
          "raw pts", np.abs(pts - opts).max(), "| alpha-1", d["alpha"] - 1, "centres", d["centres"], "points", d["points"], "rel cost", abs(s["final_cost"] - osum["final_cost"]) / osum["final_cost"])
