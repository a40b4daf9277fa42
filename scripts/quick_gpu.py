"""Scratch GPU check: K1 / solve timings for cfg3 (+ parity spot check)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle import binding as ob
from photobundle_b200 import capi
from workloads import synthetic
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

# ragged trace / pose diff
wr = synthetic.small_window(seed=11, ragged=True, n_frames=8, grid=(10, 14))
owr = ob.OracleWindow(wr)
oc, op_, osum, otr = owr.solve(wr.cams_init, wr.points_init)
hr = capi.Handle.for_window(wr)
sr = hr.solve(); cr = hr.get_poses(); pr = hr.get_points(); trr = hr.get_iterations(); hr.close()
print("ragged: gpu iters", sr["num_iterations"], sr["final_cost"], "oracle", osum["num_iterations"], osum["final_cost"])
print(" decisions equal:", [t["step_is_successful"] for t in trr] == [t["step_is_successful"] for t in otr])
print(" max dpose", np.abs(cr - oc).max(0), "max dpts", np.abs(pr - op_).max())
for a, b in list(zip(trr, otr))[-6:]:
    print("  G %d %d %.10f r=%.5e | O %d %d %.10f r=%.5e" % (a["iteration"], a["step_is_successful"], a["cost"], a["trust_region_radius"], b["iteration"], b["step_is_successful"], b["cost"], b["trust_region_radius"]))
