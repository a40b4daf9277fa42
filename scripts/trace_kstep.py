"""Per-warp phase timeline of K_A (needs a build with -DK_STEP_TRACE; scripts: make OUT=... EXTRA=-DK_STEP_TRACE)."""
import ctypes as C, sys
import numpy as np
sys.path.insert(0, ".")
from photobundle_b200 import capi
from workloads import synthetic
w = synthetic.make_window()
h = capi.Handle.for_window(w)
for _ in range(3): h.eval(want_residuals=False)
s = h.solve()   # the last K_A launches include the back-substitution
buf = np.zeros(4000 * 16, dtype=np.int64)
capi.lib().pba_debug_kstep_trace(C.c_void_p(buf.ctypes.data), buf.size)
t = buf.reshape(4000, 16)
order = [(0, "kernel entry"), (7, "state arrived (done check)"), (5, "pose constants done (warps 0-2) / skipped"), (8, "sstep, zeroing, weights done"), (12, "first point requested, at prologue barrier"), (1, "after prologue barrier"), (6, "staged inputs arrived"), (2, "back-substitution done"), (3, "geometry + staging done"),
         (4, "stage loop entered"), (9, "last group sampled"), (10, "last group reduced"), (11, "last group parked"),
         (13, "corrector + expansion done (obs loop done)"), (14, "after end sync"), (15, "exit")]
print("phase durations in cycles (median / p90 over 4000 warps):")
for (a, na), (b, nb) in zip(order, order[1:]):
    d = t[:, b] - t[:, a]
    print(f"  {na:>45s} -> {nb:<45s} {np.median(d):8.0f} {np.percentile(d, 90):8.0f}")
for wsel, name in ((0, "warp 0 of each CTA (computes pose constants)"), (5, "warp 5 of each CTA")):
    sel = t[wsel::14]
    print(name, "medians:", [int(np.median(sel[:, b] - sel[:, a])) for (a, _), (b, _) in zip(order[:6], order[1:7])])
print("warp total (prologue sync -> exit): median", np.median(t[:, 15] - t[:, 1]), "p90", np.percentile(t[:, 15] - t[:, 1], 90))
