"""Multi-GPU: a window shards by contiguous point block, frames/poses are replicated, the pose
blocks + cost and the reduced camera system are all-reduced each LM iteration.
  * CPU (gloo, world_size 2): the sharding rule partitions the points, and summing per-shard
    pose blocks / costs over ranks reproduces the single-process result (oracle arithmetic).
  * GPU (needs >= 2 devices): the N-GPU solve equals the 1-GPU solve."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from photobundle_b200 import capi
from workloads import synthetic
def test_shard_range_partitions_and_balances(small_ragged_win):
    off = small_ragged_win.obs_offsets
    n = off.shape[0] - 1
    for R in (1, 2, 3, 4, 8):
        ranges = [capi.shard_range(off, r, R) for r in range(R)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        loads = [off[b] - off[a] for a, b in ranges]
        assert max(loads) - min(loads) <= 2 * int(np.diff(off).max()) + 1
    assert capi.shard_range(np.zeros(1, dtype=np.int32), 0, 2) == (0, 0)   # empty window


def _gloo_worker(rank, world, port, q):
    import dataclasses
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import binding as ob
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    win = synthetic.small_window(seed=11, ragged=True, n_frames=8, grid=(10, 14))
    # the opaque communicator id travels the same way bench.py sends it (broadcast_object_list)
    ids = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    assert ids[0] == bytes(range(128))
    a, b = capi.shard_range(win.obs_offsets, rank, world)
    o0, o1 = int(win.obs_offsets[a]), int(win.obs_offsets[b])
    shard = dataclasses.replace(win, points_init=win.points_init[a:b], points_gt=win.points_gt[a:b], desc=win.desc[a:b],
                                obs_offsets=(win.obs_offsets[a:b + 1] - o0).astype(np.int32), obs_frame=win.obs_frame[o0:o1])
    e = ob.OracleWindow(shard, num_threads=1).evaluate(win.cams_init, shard.points_init, 1, want_residuals=False)
    buf = torch.from_numpy(np.concatenate([e["U"].ravel(), e["gc"].ravel(), [e["cost"]]]))
    dist.all_reduce(buf)                      # the per-iteration exchange: pose blocks + cost
    if rank == 0:
        full = ob.OracleWindow(win, num_threads=1).evaluate(win.cams_init, win.points_init, 1, want_residuals=False)
        ref = np.concatenate([full["U"].ravel(), full["gc"].ravel(), [full["cost"]]])
        q.put(float(np.abs(buf.numpy() - ref).max() / np.abs(ref).max()))
    dist.destroy_process_group()


def test_sharded_blocks_sum_to_global_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) <= 1e-13


def _run_workers(n, kind, exchange=None, replicate=False):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + n), os.path.join(ROOT, "tests", "mgpu_worker.py"), kind]
    env = dict(os.environ)
    env.pop("PBA_MGPU_EXCHANGE", None)
    # the windows of these tests fit one K_A wave, which the library would not shard (pba_comm_sharded): force it
    env["PBA_MGPU_REPLICATE"] = "1" if replicate else "0"
    if exchange:
        env["PBA_MGPU_EXCHANGE"] = exchange
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("MGPU_RESULT ")][-1]
    return json.loads(line[len("MGPU_RESULT "):])


@pytest.mark.gpu
def test_two_gpus_equal_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    one = _run_workers(1, "small")
    assert one["collectives"] == 0 and one["exchange"] == "none"
    # default: in-kernel exchange over NVLink peer memory (no collective calls); fallback: NCCL all-reduce
    for exchange in (None, "nccl"):
        two = _run_workers(2, "small", exchange)
        if exchange == "nccl":
            assert two["exchange"] == "nccl" and two["collectives"] > 0
        else:
            assert two["exchange"] in ("peer-memory", "nccl")   # nccl only where peer mappings are unavailable
            if two["exchange"] == "peer-memory":
                # one in-kernel exchange per LM decision, a second one only when the speculated outcome missed
                assert two["iters"] <= two["collectives"] <= 2 * two["iters"] + 1
        assert two["ranks_agree"]
        assert two["accepts"] == one["accepts"]
        assert abs(two["final_cost"] - one["final_cost"]) <= 1e-9 * one["final_cost"]
        np.testing.assert_allclose(np.array(two["cams"]), np.array(one["cams"]), atol=1e-8)
        np.testing.assert_allclose(np.array(two["pts_tail"]), np.array(one["pts_tail"]), atol=1e-6)
        assert two["n_pts"] == one["n_pts"]


@pytest.mark.gpu
def test_two_gpus_small_window_is_not_sharded():
    """Default policy: a window that fits one wave of K_A is solved whole by every rank - no exchange at all, and the
    result is the single-GPU one (to rounding: fp64 atomics are unordered)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    one = _run_workers(1, "small")
    env_backup = os.environ.pop("PBA_MGPU_REPLICATE", None)
    try:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
               "--master-port", "29733", os.path.join(ROOT, "tests", "mgpu_worker.py"), "small"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    finally:
        if env_backup is not None:
            os.environ["PBA_MGPU_REPLICATE"] = env_backup
    assert r.returncode == 0, r.stderr[-2000:]
    two = json.loads([l for l in r.stdout.splitlines() if l.startswith("MGPU_RESULT ")][-1][len("MGPU_RESULT "):])
    assert two["sharded"] is False and two["collectives"] == 0
    assert two["accepts"] == one["accepts"] and two["n_pts"] == one["n_pts"]
    assert abs(two["final_cost"] - one["final_cost"]) <= 1e-9 * one["final_cost"]
    np.testing.assert_allclose(np.array(two["cams"]), np.array(one["cams"]), atol=1e-8)
    np.testing.assert_allclose(np.array(two["pts_tail"]), np.array(one["pts_tail"]), atol=1e-6)


@pytest.mark.gpu
def test_two_gpus_empty_shard():
    """Fewer points than ranks: the rank with the empty shard still takes part in every exchange (zero contribution)
    instead of leaving its peers to time out."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    one, two = _run_workers(1, "one_point"), _run_workers(2, "one_point")
    assert two["ranks_agree"] and two["n_pts"] == one["n_pts"] == 1
    assert two["accepts"] == one["accepts"]
    assert abs(two["final_cost"] - one["final_cost"]) <= 1e-9 * max(one["final_cost"], 1e-30)
    np.testing.assert_allclose(np.array(two["cams"]), np.array(one["cams"]), atol=1e-8)


@pytest.mark.gpu
def test_two_gpus_bench_window_repeated_solves():
    """cfg3 window on 2 GPUs, solved twice on the same handle (epochs keep counting across solves)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    one, two = _run_workers(1, "cfg3"), _run_workers(2, "cfg3")
    assert two["ranks_agree"] and two["accepts"] == one["accepts"]
    assert abs(two["final_cost"] - one["final_cost"]) <= 1e-9 * one["final_cost"]
    np.testing.assert_allclose(np.array(two["cams"]), np.array(one["cams"]), atol=1e-8)
    assert abs(two["second_final_cost"] - two["final_cost"]) <= 1e-12 * two["final_cost"]


@pytest.mark.gpu
def test_local_communicator_two_devices_one_process(small_win, monkeypatch):
    """pba_comm_init_local: two handles of ONE process on two devices (no NCCL, no IPC), pba_solve from one thread per
    handle, equals the 1-GPU solve; the members agree; a second solve on the same handles works."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    w = small_win
    one = capi.Handle.for_window(w)
    s1 = one.solve()
    c1, p1 = one.get_poses(), one.get_points()
    acc1 = [t["step_is_successful"] for t in one.get_iterations()]
    monkeypatch.setenv("PBA_MGPU_REPLICATE", "0")            # a window this small would not be sharded by default
    hs = [capi.Handle(w.rows, w.cols, w.fx, w.fy, w.cx, w.cy, radius=w.radius, huber=w.huber, max_frames=w.n_frames,
                      max_points=w.n_points, max_observations=w.n_obs, device=d) for d in (0, 1)]
    with pytest.raises(capi.PbaError):
        capi.Handle.comm_init_local([hs[0], hs[0]])          # the same device twice
    capi.Handle.comm_init_local(hs)
    assert [h.exchange_kind() for h in hs] == ["peer-memory", "peer-memory"]
    for h in hs:
        h.set_frames_u8(w.images)
        h.set_poses(w.cams_init, w.fixed_frame)
        h.set_points(w.points_init, w.desc, w.obs_offsets, w.obs_frame, w.weights)
        h.save_state()
    assert all(h.sharded() for h in hs)
    for _ in range(2):
        ss = capi.Handle.solve_all(hs)
        assert ss[0]["final_cost"] == ss[1]["final_cost"] and ss[0]["num_iterations"] == ss[1]["num_iterations"]
        assert [t["step_is_successful"] for t in hs[0].get_iterations()] == acc1
        assert abs(ss[0]["final_cost"] - s1["final_cost"]) <= 1e-9 * s1["final_cost"]
        assert ss[0]["num_iterations"] <= ss[0]["num_collectives"] <= 2 * ss[0]["num_iterations"] + 1
        np.testing.assert_array_equal(hs[0].get_poses(), hs[1].get_poses())
        np.testing.assert_allclose(hs[0].get_poses(), c1, atol=1e-8)
        for h in hs:                                          # any member gathers the shards of all
            np.testing.assert_allclose(h.get_points(), p1, atol=1e-6)
        for h in hs:
            h.restore_state()
    # a member solved alone cannot make progress: it fails after the host barrier's time-out instead of hanging
    # (not exercised here: the time-out is 60 s)
    for h in hs:
        h.close()
    one.close()
