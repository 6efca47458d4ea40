"""-m gpu, ONE device: the multi-GPU band schedule (halo pushes + flag waits inside the frame graphs, vrs_peer_connect_local)
with every band a context of this process on the same GPU, driven by vrs_render_frame_group.  The halo exchange is the
same kernels and the same phase schedule as across GPUs (there the stores travel over NVLink), so a 1-GPU box proves: bands == single frame bit for bit on the bench's
own 6 deg/frame orbit, the halo sizing rule, the out-of-halo counter and the exchange time-out path."""
import numpy as np
import pytest

import bench
import common

pytestmark = pytest.mark.gpu


def make(V, wl, path, band, halo, lights=None):
    R = V.Renderer(wl["W"], wl["H"], spatial_iterations=wl["iters"], band=band, halo_rows=halo)
    R.loadVDB(path)
    if lights is None:
        lights, ctr, diag = bench.build_scene_inputs(V, wl, R)
        R._scene = (lights, ctr, diag)
    R.createRestirLights(lights)
    u = R.m_restirUniforms
    u.initialLightSampleCount, u.spatialNeighbors, u.flags = wl["M"], wl["k"], wl["flags"]
    return R


def run_bands(V, name, nbands, frames, halo=None, W=None, H=None, M=None, edges=None, check_counter=True, orbit_deg=None):
    wl = dict(bench.WORKLOADS[name])
    if W:
        wl["W"], wl["H"] = W, H
    if M:
        wl["M"] = M
    path = bench.asset_path(V, wl["asset"])
    full = make(V, wl, path, None, 32)
    lights, ctr, diag = full._scene
    gi = full.gridInfo()
    reach = bench.temporal_halo_rows(V, wl, list(gi.world_bbox_min), list(gi.world_bbox_max), ctr, diag, frames=frames + 1)
    if halo is None:
        halo = 32          # spatial reuse only (radius 30): temporal reprojections that leave a band are read in place from the adjacent band
    if edges is None:
        edges = [round(i * wl["H"] / nbands) for i in range(nbands + 1)]
    bands = [make(V, wl, path, (edges[i], edges[i + 1]), halo, lights) for i in range(nbands)]
    for i, b in enumerate(bands):
        b.peerConnectLocal(bands[i - 1] if i > 0 else None, bands[i + 1] if i + 1 < nbands else None)
    radius = bench.ORBIT_RADIUS * diag
    for r_ in [full] + bands:
        r_.CameraManip.setLookat(bench.orbit_eye(ctr, radius, 0.0, 0.0), ctr)
        r_.createRestirUniformBuffer()
    bad = []
    for f in range(frames):
        eye = bench.orbit_eye(ctr, radius, 0.0, (bench.ORBIT_DEG if orbit_deg is None else orbit_deg) * f)
        for b in bands:
            b.CameraManip.setLookat(eye, ctr)
        V.render_frame_group(bands, f)               # one host thread: phases interleaved across the bands (vrs_render_frame_group)
        for b in bands:
            b.synchronize()
        full.CameraManip.setLookat(eye, ctr)
        full.renderFrame(clock=f)
        ref = full.readFrame()
        got = np.concatenate([b.readFrame() for b in bands], 0)
        if not (common.u32(got) == common.u32(ref)).all():
            bad.append((f, int((common.u32(got) != common.u32(ref)).any(-1).sum())))
        rr = full.readReservoirs()
        gr = {k: np.concatenate([b.readReservoirs()[k] for b in bands], 0) for k in ("info", "weight")}
        hit = full.readGBuffer()["worldPos"][..., 3] > 0.5
        for k in ("info", "weight"):
            if not (common.u32(gr[k])[hit] == common.u32(rr[k])[hit]).all():
                bad.append((f, k))
    ooh = sum(b.counters().temporal_out_of_halo for b in bands)
    for r_ in bands + [full]:
        r_.destroy()
    if check_counter:
        assert ooh == 0, "temporal reprojection left the halo rows (%d pixels) although the halo was sized from the orbit" % ooh
    return bad, ooh, reach


def test_two_and_three_bands_equal_one_frame_smoke_1080p(V):
    """configs[1] at full size on the bench orbit, 2 and 3 bands (uneven edges like the cost-balanced split)."""
    bad, _, reach = run_bands(V, "smoke_1080p_temporal", 2, 5)
    assert not bad, bad
    assert reach > 32                                # the 6 deg/frame orbit reprojects further than the 32-row halo: those pixels come from the neighbour's band
    bad, _, _ = run_bands(V, "smoke_1080p_full", 3, 4, edges=[0, 420, 640, 1080])
    assert not bad, bad


def test_two_bands_equal_one_frame_bunny_4k(V):
    """configs[3] (bench default) at 3840x2160 with 10k lights, full spatiotemporal, 2 bands, 3 frames of the bench orbit."""
    bad, _, reach = run_bands(V, "bunny_4k_full", 2, 3)
    assert not bad, bad
    assert reach >= 64


def test_small_halo_is_detected_not_silent(V):
    """A band thinner than the orbit's reprojection distance: some previous-frame pixels lie two bands away, where nobody
    can supply them; those merges are dropped and the counter must say so."""
    bad, ooh, _ = run_bands(V, "smoke_1080p_temporal", 3, 4, halo=8, check_counter=False, edges=[0, 300, 308, 1080], orbit_deg=25.0)
    assert ooh > 0, (ooh, bad)


def test_halo_wait_timeout_is_an_error(V):
    """A neighbour that never publishes its rows: the wait kernel gives up after ~4 s and vrs_synchronize returns VRS_ERR_COMM."""
    wl = dict(bench.WORKLOADS["smoke_1080p_full"], W=256, H=128, M=4)
    path = bench.asset_path(V, wl["asset"])
    a = make(V, wl, path, (0, 64), 32)
    lights, ctr, diag = a._scene
    b = make(V, wl, path, (64, 128), 32, lights)
    a.peerConnectLocal(None, b)
    b.peerConnectLocal(a, None)
    a.CameraManip.setLookat(bench.orbit_eye(ctr, 1.25 * diag, 0.0, 0.0), ctr)
    a.createRestirUniformBuffer()
    a.renderFrame(clock=0)                           # b never renders (vrs_render_frame on one band only)
    with pytest.raises(V.VrsError) as e:
        a.synchronize()
    assert e.value.status == 6                       # VRS_ERR_COMM
    assert a.counters().comm_timeouts >= 1
    a.destroy(); b.destroy()
