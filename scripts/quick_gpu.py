"""Scratch GPU check: K1 / solve timings for cfg3 (+ parity spot check)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle import binding as ob
from photobundle_b200 import capi, synthetic

w = synthetic.make_window()
ow = ob.OracleWindow(w)
h = capi.Handle.for_window(w)
ev = h.eval()
ref = ow.evaluate(w.cams_init, w.points_init, 1)
d = np.abs(ev["residuals"] - ref["residuals"])
print("residual exact frac", (d == 0).mean(), "max", d.max(), "cost rel", abs(ev["cost"] - ref["cost"]) / ref["cost"])
for k in ("U", "gc", "V", "gp", "W"):
    print(k, np.abs(ev[k] - ref[k]).max() / np.abs(ref[k]).max())
for it in (1, 10, 100, 100):
    ms = h.eval_timed(it)
    print(f"K1 cfg2: {it} launches {ms:.4f} ms -> {ms/it*1e3:.2f} us/launch, {w.n_obs*656/(ms/it*1e-3)/1e9:.1f} GB/s algorithmic")
for rep in range(3):
    h.set_poses(w.cams_init, 0); h.set_points(w.points_init, w.desc, w.obs_offsets, w.obs_frame, w.weights)
    t = time.time(); s = h.solve(); dt = time.time() - t
    print(f"solve cfg3: iters {s['num_iterations']} evals {s['num_evaluations']} device {s['device_time_in_seconds']*1e3:.3f} ms wall {dt*1e3:.3f} ms launches {s['kernel_launches']} final {s['final_cost']:.4f} {s['message']}")
h.close()
