"""Per-warp phase timeline of K_A (needs a build with -DK_STEP_TRACE)."""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, ".")
from photobundle_b200 import capi, synthetic
import os
img = "/tmp/cfg3_images.npy"
w = synthetic.make_window(images=np.load(img) if os.path.exists(img) else None)
h = capi.Handle.for_window(w)
for _ in range(3): h.eval(want_residuals=False)
s = h.solve()   # last K_A launches include the back-substitution
buf = np.zeros(4000 * 16, dtype=np.int64)
capi.lib().pba_debug_kstep_trace(C.c_void_p(buf.ctypes.data), buf.size)
t = buf.reshape(4000, 16)
names = ["kernel entry", "after prologue sync", "after loads+backsub", "after geometry", "after staging", "q0 sampled", "q0 reduced", "q0 huber", "q0 emitted", "q1 sampled", "q1 reduced", "q1 huber", "q1 emitted", "obs loop done", "after end sync", "exit"]
t0 = t[:, 0:1]
d = np.diff(t, axis=1)
print("phase durations in cycles (median / p90 over 4000 warps):")
for i in range(15):
    print(f"  {names[i]:>22s} -> {names[i+1]:<22s} {np.median(d[:, i]):8.0f} {np.percentile(d[:, i], 90):8.0f}")
print("warp total (entry->exit): median", np.median(t[:, 15] - t[:, 0]), "p90", np.percentile(t[:, 15] - t[:, 0], 90))
starts = t[:, 0] - t[:, 0].min()
print("kernel span:", (t[:, 15].max() - t[:, 0].min()), "cycles; warp start offsets p50/p90/max:", np.median(starts), np.percentile(starts, 90), starts.max())
