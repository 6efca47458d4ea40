"""not gpu: the multi-GPU decomposition (screen-space bands, grid replicated, halo rows exchanged) exercised with
world_size 2 and 3 over the gloo backend on CPU.  Each rank runs the oracle on its band only and exchanges halo rows
with dist.send / dist.recv on the schedule libvrs uses (vrs_render_frame); the assembled frames must be bit-identical
to a single-rank run — the property the NCCL path is tested for on the GPU box (tests/test_gpu_multi.py)."""
import os
import socket
import sys

import numpy as np
import pytest

import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def halo_exchange(dist, torch, planes, rank, world, band, halo, H, max_rows):
    """Python mirror of vrs::comm_exchange_halo (csrc/vrs_comm.cpp): first/last own rows -> neighbours' halo rows, at most
    `max_rows` of them (spatial exchanges move ceil(spatialRadius) rows, the temporal one the whole halo)."""
    y0, y1 = band
    ops = []
    for p in planes:
        t = torch.from_numpy(p)
        n = min(halo, y1 - y0, max_rows)
        if rank > 0:
            ops.append(dist.P2POp(dist.isend, t[y0:y0 + n].contiguous(), rank - 1))
            lo = max(0, y0 - min(halo, max_rows))
            ops.append(dist.P2POp(dist.irecv, t[lo:y0], rank - 1))
        if rank < world - 1:
            ops.append(dist.P2POp(dist.isend, t[y1 - n:y1].contiguous(), rank + 1))
            hi = min(H, y1 + min(halo, max_rows))
            ops.append(dist.P2POp(dist.irecv, t[y1:hi], rank + 1))
    for r in dist.batch_isend_irecv(ops):
        r.wait()


def worker(rank, world, port, flags, frames, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle as O
    import vrs_pkg
    V = vrs_pkg.load()
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    W, H, halo, radius = 96, 44 * world, 40, 12.0      # bands must be at least as tall as the halo (vrs_comm_init enforces it)
    lights = V.generate_point_lights([-4.2, -0.2, -1.4], [-1.5, 5.2, 1.6], False, 16)
    scene = common.oracle_scene(O, "smoke", lights)
    lo, hi = scene.world_bbox()
    ctr = [(a + b) * 0.5 for a, b in zip(lo, hi)]
    band = V.band_for_rank(H, rank, world)
    OR = O.OracleRenderer(scene, W, H, spatial_iterations=2)
    full = O.OracleRenderer(scene, W, H, spatial_iterations=2) if rank == 0 else None
    prev = None
    ok = True
    for f in range(frames):
        cam = O.Camera(common.orbit_eye(ctr, 4.5, 0.3, 20.0 + 1.5 * f), ctr)
        gu = O.global_uniforms(cam, W, H)
        ru = O.restir_uniforms(cam, prev, W, H, len(lights), M=8, flags=flags, k=5, radius=radius)
        pc = O.PushConstant(0, 0, 0, f, 1 if f < 2 else 0)
        img = OR.render(gu, ru, pc, f, band[0], band[1],
                        exchange=lambda planes, kind: halo_exchange(dist, torch, planes, rank, world, band, halo, H,
                                                                    halo if kind == "temporal" else int(np.ceil(radius))))
        mine = torch.from_numpy(np.ascontiguousarray(img[band[0]:band[1]]))
        parts = [torch.zeros((V.band_for_rank(H, r, world)[1] - V.band_for_rank(H, r, world)[0], W, 4)) for r in range(world)] if rank == 0 else None
        dist.gather(mine, parts, dst=0)
        if rank == 0:
            ref = full.render(gu, ru, pc, f).copy()
            got = torch.cat(parts, 0).numpy()
            ok = ok and bool((got.view(np.uint32) == ref.view(np.uint32)).all())
        prev = cam
    if rank == 0:
        open(os.path.join(out_dir, "result"), "w").write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,flags", [(2, 1 | 2 | 4), (3, 1 | 2 | 4), (2, 1 | 2)])
def test_bands_with_halo_exchange_equal_single_rank(world, flags, tmp_path, O):
    import torch.multiprocessing as mp
    port = free_port()
    mp.spawn(worker, args=(world, port, flags, 3, str(tmp_path)), nprocs=world, join=True)
    assert open(tmp_path / "result").read() == "ok"
