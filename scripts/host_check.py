import sys
import numpy as np
sys.path.insert(0, ".")
from photobundle_b200 import host_capi
from workloads import synthetic
seq = synthetic.make_sequence(n_frames=10)
rows, cols = seq.images.shape[1:]
n = seq.images.shape[0]
ba = host_capi.BundleAdjuster(rows, cols, *seq.K4, slidingWindowSize=5, maxNumPoints=2048, verbose=0, minScore=0.65)
for i in range(n):
    ran = ba.add_frame(seq.images[i], seq.depths[i], seq.T_rel_init[i])
    if ran:
        r = ba.result(); print(i, "cost %.3f -> %.3f" % (r["initialCost"], r["finalCost"]), r["numResiduals"], r["numSuccessfulStep"], r["message"][:40], len(r["refinedPoints"]))
res = ba.result()
T0 = [np.linalg.inv(seq.T_rel_init[0])]
for i in range(1, n): T0.append(T0[-1] @ np.linalg.inv(seq.T_rel_init[i]))
T0 = np.stack(T0)
def rot_err(A, B): return np.array([np.degrees(np.arccos(np.clip((np.trace(a[:3,:3].T @ b[:3,:3]) - 1) / 2, -1, 1))) for a, b in zip(A, B)])
print("rot err init", np.round(rot_err(T0, seq.T_w_gt), 4)); print("rot err ref ", np.round(rot_err(res["poses"], seq.T_w_gt), 4))
print("trans err init", np.round(np.linalg.norm(T0[:, :3, 3] - seq.T_w_gt[:, :3, 3], axis=1), 4))
print("trans err ref ", np.round(np.linalg.norm(res["poses"][:, :3, 3] - seq.T_w_gt[:, :3, 3], axis=1), 4))
# scale-aligned: relative step lengths
print("step len gt  ", np.round(np.linalg.norm(np.diff(seq.T_w_gt[:, :3, 3], axis=0), axis=1), 4))
print("step len init", np.round(np.linalg.norm(np.diff(T0[:, :3, 3], axis=0), axis=1), 4))
print("step len ref ", np.round(np.linalg.norm(np.diff(res["poses"][:, :3, 3], axis=0), axis=1), 4))
