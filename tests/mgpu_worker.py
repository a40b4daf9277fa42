"""Worker for the multi-GPU parity test: launched with torchrun (one process per GPU).  Every
rank joins the window's communicator, uploads the FULL window (the library keeps its shard),
solves, and rank 0 prints the result as JSON."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from photobundle_b200 import capi  # noqa: E402
from workloads import synthetic  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    kind = sys.argv[1] if len(sys.argv) > 1 else "small"
    if kind == "small":
        win = synthetic.small_window(seed=7)
    elif kind == "one_point":   # fewer points than ranks: one rank's shard is empty
        import dataclasses
        w0 = synthetic.small_window(seed=7)
        o = int(w0.obs_offsets[1])
        win = dataclasses.replace(w0, points_init=w0.points_init[:1], points_gt=w0.points_gt[:1], desc=w0.desc[:1],
                                  obs_offsets=w0.obs_offsets[:2], obs_frame=w0.obs_frame[:o])
    elif kind == "cfg4":      # finest level of BASELINE configs[3]: 16 frames x 16 000 points
        win = synthetic.make_window(n_frames=16, grid=(100, 160))
    else:
        win = synthetic.make_window()
    ids = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    h = capi.Handle(win.rows, win.cols, win.fx, win.fy, win.cx, win.cy, radius=win.radius, huber=win.huber,
                    max_frames=win.n_frames, max_points=win.n_points, max_observations=win.n_obs, device=local)
    h.comm_init(ids[0], rank, world)
    h.set_frames_u8(win.images)
    h.set_poses(win.cams_init, win.fixed_frame)
    h.set_points(win.points_init, win.desc, win.obs_offsets, win.obs_frame, win.weights)
    h.save_state()
    s = h.solve()
    cams, pts = h.get_poses(), h.get_points()
    second, second_ms = None, None
    if kind not in ("small", "one_point"):      # solve the same window again on the same handle (warm: graph instantiated, L2 primed)
        h.restore_state()
        s2 = h.solve()
        second, second_ms = s2["final_cost"], 1e3 * s2["device_time_in_seconds"]
    tr = h.get_iterations()
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(cams=cams.tolist(), cost=s["final_cost"], iters=s["num_iterations"]))
    if rank == 0:
        print("MGPU_RESULT " + json.dumps(dict(
            world=world, second_final_cost=second, second_device_ms=second_ms, final_cost=s["final_cost"], initial_cost=s["initial_cost"], iters=s["num_iterations"],
            collectives=s["num_collectives"], exchange=h.exchange_kind(), sharded=h.sharded(), launches=s["kernel_launches"], device_ms=1e3 * s["device_time_in_seconds"],
            cams=cams.tolist(), pts_head=pts[:5].tolist(), pts_tail=pts[-5:].tolist(), n_pts=int(pts.shape[0]),
            accepts=[t["step_is_successful"] for t in tr],
            ranks_agree=all(np.array_equal(np.array(g["cams"]), cams) and g["cost"] == s["final_cost"] for g in gathered))))
    h.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
