"""-m gpu (needs >= 2 GPUs, skipped otherwise): bands + halo exchange through the C ABI — the NVLink peer-memory kernel
(vrs_peer_connect, replayed inside the frame's CUDA graph) and NCCL send/recv (vrs_comm_init) — must give the same bits
as one GPU rendering the whole frame."""
import os
import socket
import sys

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def worker(rank, world, port, flags, frames, out_dir, mode):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import vrs_pkg
    V = vrs_pkg.load()
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    W, H = 480, 270
    band = V.band_for_rank(H, rank, world)

    def make(band_, device):
        R = V.Renderer(W, H, spatial_iterations=2, band=band_, halo_rows=32, device=device)
        R.loadVDB(common.asset("smoke"))
        gi = R.gridInfo()
        lo, hi = list(gi.world_bbox_min), list(gi.world_bbox_max)
        lights = V.generate_point_lights(lo, hi, False, 64)
        R.createRestirLights(lights)
        R.m_restirUniforms.initialLightSampleCount = 16
        R.m_restirUniforms.spatialNeighbors = 5
        R.m_restirUniforms.flags = flags
        ctr = [(a + b) * 0.5 for a, b in zip(lo, hi)]
        return R, ctr

    R, ctr = make(band, rank)
    if mode == "nccl":
        uid = [V.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        R.commInit(uid[0], rank, world)
    else:
        blobs = [None] * world
        dist.all_gather_object(blobs, R.peerExport())
        R.peerConnect(rank, world, blobs)
    full = make(None, 0)[0] if rank == 0 else None
    for r_ in [R] + ([full] if full else []):
        r_.CameraManip.setLookat(common.orbit_eye(ctr, 4.2, 0.2, 10.0), ctr)
        r_.createRestirUniformBuffer()
    ok = True
    for f in range(frames):
        eye = common.orbit_eye(ctr, 4.2, 0.2, 10.0 + 1.0 * f)
        R.CameraManip.setLookat(eye, ctr)
        R.renderFrame(clock=f)
        R.synchronize()
        mine = torch.from_numpy(R.readFrame())
        parts = [torch.zeros((V.band_for_rank(H, r, world)[1] - V.band_for_rank(H, r, world)[0], W, 4)) for r in range(world)] if rank == 0 else None
        dist.gather(mine, parts, dst=0)
        if rank == 0:
            full.CameraManip.setLookat(eye, ctr)
            full.renderFrame(clock=f)
            ref = full.readFrame()
            got = torch.cat(parts, 0).numpy()
            ok = ok and bool((got.view(np.uint32) == ref.view(np.uint32)).all())
    if rank == 0:
        open(os.path.join(out_dir, "result"), "w").write("ok" if ok else "mismatch")
    dist.barrier()
    R.destroy()
    dist.destroy_process_group()


def bench_worker(rank, world, port, name, frames, out_dir, mode):
    """The bench workload at its benchmarked size on the bench's own 6 deg/frame orbit, bands + halo sized like bench.py."""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import bench
    import vrs_pkg
    V = vrs_pkg.load()
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    wl = bench.WORKLOADS[name]
    W, H = wl["W"], wl["H"]
    path = bench.asset_path(V, wl["asset"]) if rank == 0 else None
    dist.barrier()
    path = bench.asset_path(V, wl["asset"])

    def make(band_, device, halo):
        R = V.Renderer(W, H, spatial_iterations=wl["iters"], band=band_, halo_rows=halo, device=device)
        R.loadVDB(path)
        lights, ctr, diag = bench.build_scene_inputs(V, wl, R)
        R.createRestirLights(lights)
        u = R.m_restirUniforms
        u.initialLightSampleCount, u.spatialNeighbors, u.flags = wl["M"], wl["k"], wl["flags"]
        return R, ctr, diag

    full, ctr, diag = make(None, rank, 32) if rank == 0 else (None, None, None)
    probe, ctr, diag = make((0, 64), rank, 0)
    gi = probe.gridInfo()
    reach = bench.temporal_halo_rows(V, wl, list(gi.world_bbox_min), list(gi.world_bbox_max), ctr, diag, frames=frames + 1)
    halo = reach if mode == "nccl" else 32          # peer memory reads reprojected pixels in place from the adjacent band
    probe.destroy()
    edges = [0, int(H * 0.46), H] if world == 2 else [round(i * H / world) for i in range(world + 1)]     # uneven, like the cost-balanced split
    band = (edges[rank], edges[rank + 1])
    R, _, _ = make(band, rank, halo)
    if mode == "nccl":
        uid = [V.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        R.commInit(uid[0], rank, world)
    else:
        blobs = [None] * world
        dist.all_gather_object(blobs, R.peerExport())
        R.peerConnect(rank, world, blobs)
    radius = bench.ORBIT_RADIUS * diag
    for r_ in [R] + ([full] if full else []):
        r_.CameraManip.setLookat(bench.orbit_eye(ctr, radius, 0.0, 0.0), ctr)
        r_.createRestirUniformBuffer()
    ok = True
    for f in range(frames):
        eye = bench.orbit_eye(ctr, radius, 0.0, bench.ORBIT_DEG * f)
        R.CameraManip.setLookat(eye, ctr)
        R.renderFrame(clock=f)
        R.synchronize()                                  # raises VRS_ERR_COMM when a halo wait timed out
        rows_max = max(edges[r + 1] - edges[r] for r in range(world))          # gloo gather wants equal shapes: pad the bands
        mine = torch.zeros((rows_max, W, 4))
        mine[:band[1] - band[0]] = torch.from_numpy(R.readFrame())
        parts = [torch.zeros((rows_max, W, 4)) for r in range(world)] if rank == 0 else None
        dist.gather(mine, parts, dst=0)
        if rank == 0:
            full.CameraManip.setLookat(eye, ctr)
            full.renderFrame(clock=f)
            ref = full.readFrame()
            got = torch.cat([parts[r][:edges[r + 1] - edges[r]] for r in range(world)], 0).numpy()
            ok = ok and bool((got.view(np.uint32) == ref.view(np.uint32)).all())
    ooh = torch.tensor([float(R.counters().temporal_out_of_halo)])
    dist.all_reduce(ooh)
    if rank == 0:
        open(os.path.join(out_dir, "result"), "w").write(("ok" if ok else "mismatch") + " ooh=%d halo=%d" % (int(ooh[0]), halo))
    dist.barrier()
    R.destroy()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,name,frames", [("peer", "bunny_4k_full", 4), ("peer", "smoke_1080p_temporal", 6), ("nccl", "smoke_1080p_full", 4)])
def test_bench_workload_bands_equal_single_gpu(mode, name, frames, tmp_path):
    """the 2-GPU run of the bench workloads on the bench orbit (6 deg/frame) must equal one GPU bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(bench_worker, args=(world, free_port(), name, frames, str(tmp_path), mode), nprocs=world, join=True)
    res = open(tmp_path / "result").read()
    assert res.startswith("ok ooh=0 "), res


@pytest.mark.parametrize("mode,flags", [("peer", 1 | 2 | 4), ("peer", 1 | 2), ("nccl", 1 | 2 | 4)])
def test_bands_equal_single_gpu(mode, flags, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(worker, args=(world, free_port(), flags, 14, str(tmp_path), mode), nprocs=world, join=True)
    assert open(tmp_path / "result").read() == "ok"
