"""Target for PBA_DEBUG_TIMELINE / ncu: the finest level of BASELINE configs[3] (16 frames x 16 000 points)."""
import sys
sys.path.insert(0, ".")
from photobundle_b200 import capi
from workloads import synthetic
w = synthetic.make_window(n_frames=16, grid=(100, 160))
h = capi.Handle.for_window(w)
s = h.solve()
print(s["num_iterations"], s["device_time_in_seconds"], s["kb_device_time_in_seconds"])
h.close()
