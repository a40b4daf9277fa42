"""Per-frame cost of the sliding-window pipeline (addFrame + optimize) at KITTI size, host vs device front end."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from photobundle_b200 import host_capi
from workloads import synthetic
seq = synthetic.make_sequence(n_frames=12, rows=376, cols=1241, intrinsics=(718.856, 718.856, 607.1928, 185.2157))
rows, cols = seq.images.shape[1:]
for gpu in (0, 1):
    ba = host_capi.BundleAdjuster(rows, cols, *seq.K4, slidingWindowSize=8, maxNumPoints=4096, verbose=0, minScore=0.65, gpuFrontEnd=gpu)
    t = []
    for i in range(12):
        t0 = time.perf_counter()
        ran = ba.add_frame(seq.images[i], seq.depths[i], seq.T_rel_init[i])
        t.append((time.perf_counter() - t0, ran))
    res = ba.result()
    no_opt = [a for a, r in t[1:] if not r]
    opt = [a for a, r in t if r]
    print(f"gpuFrontEnd={gpu}: addFrame without solve {1e3*np.median(no_opt):.2f} ms, with solve {1e3*np.median(opt):.2f} ms "
          f"(solver time {1e3*res['totalTime']:.2f} ms, {res['numResiduals']} residuals, final cost {res['finalCost']:.3f}), points {len(ba.scene_points())}")
    ba.close()
