"""K_A (evaluation-only) and full-solve device time of the cfg3 window for patch radius 1 / 2 / 3 (the reference's own
KITTI configuration uses patchRadius = 1, BASELINE's workloads 2)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from photobundle_b200 import capi
from workloads import synthetic
for r in (1, 2, 3):
    w = synthetic.make_window(radius=r)
    h = capi.Handle.for_window(w)
    h.save_state()
    h.eval_timed(20)
    ka = h.eval_timed(200) / 200
    ts = []
    for _ in range(8):
        h.restore_state()
        s = h.solve()
        ts.append(s["device_time_in_seconds"])
    print(f"radius {r}: {w.n_residuals} residuals, K_A {1e3 * ka:.1f} us per launch, solve {1e3 * np.median(ts):.3f} ms, "
          f"{s['num_iterations']} iterations, K_B {1e6 * s['kb_device_time_in_seconds'] / max(1, s['num_evaluations'] - 1):.1f} us")
    h.close()
