"""ncu target: a few K1 launches + one full solve on the cfg2/3 window."""
import sys
import numpy as np
sys.path.insert(0, ".")
from photobundle_b200 import capi
from workloads import synthetic
import os
img = "/tmp/cfg3_images.npy"
w = synthetic.make_window(images=np.load(img) if os.path.exists(img) else None)
h = capi.Handle.for_window(w)
for _ in range(4):
    h.eval(want_residuals=False)
if len(sys.argv) > 1 and sys.argv[1] == "solve":
    s = h.solve()
    print(s["num_iterations"], s["device_time_in_seconds"])
h.close()
