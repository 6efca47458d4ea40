"""-m gpu: the EXACT bench workloads at their benchmarked size against the oracle — same asset, lights, M, k, iterations,
6 deg/frame orbit at 1.25 bbox half-diagonals as bench.py (the scene comes from bench.py's own builders), culling on, CUDA
graphs on, the launcher's large-frame kernel choices (k_ris_thread / L2-table k_ris_coop, multi-block compaction).
Bar (BASELINE.json north_star): integer state bit-exact, reservoir weights <= 1 ulp, images relMSE <= 1e-4."""
import ctypes as C

import numpy as np
import pytest

import bench
import common

pytestmark = pytest.mark.gpu


def run_workload(V, name, frames, same_lights=True):
    O, wl, scene, ctr, diag = bench.oracle_setup(name)
    W, H = wl["W"], wl["H"]
    R = V.Renderer(W, H, spatial_iterations=wl["iters"])
    R.loadVDB(bench.asset_path(V, wl["asset"]))
    lights, ctr_p, diag_p = bench.build_scene_inputs(V, bench.WORKLOADS[name], R)
    assert ctr_p == pytest.approx(ctr) and diag_p == pytest.approx(diag)
    if same_lights:          # both arms of bench.py build the same scene, bit for bit
        assert lights.shape == scene.lights.shape and (common.u32(lights) == common.u32(scene.lights)).all(), \
            "product and oracle arms of the bench build different lights (first rows %s vs %s)" % (lights[0], scene.lights[0])
    else:
        O, wl, scene, ctr, diag = bench.oracle_setup(name, lights_override=lights)
    R.createRestirLights(lights)
    u = R.m_restirUniforms
    u.initialLightSampleCount, u.spatialNeighbors, u.flags = wl["M"], wl["k"], wl["flags"]
    radius = bench.ORBIT_RADIUS * diag_p
    R.CameraManip.setLookat(bench.orbit_eye(ctr_p, radius, 0.0, 0.0), ctr_p)
    R.createRestirUniformBuffer()
    OR = O.OracleRenderer(scene, W, H, spatial_iterations=wl["iters"])
    for f in range(frames):
        R.CameraManip.setLookat(bench.orbit_eye(ctr_p, radius, 0.0, bench.ORBIT_DEG * f), ctr_p)
        R.renderFrame(clock=f)
        gu, ru, pc = common.oracle_uniforms(O, R)
        pc.initialize = R._last_initialize
        ref = OR.render(gu, ru, pc, f)
        # --- images
        img = R.readFrame()
        rel = common.rel_mse(img, ref)
        assert rel <= 1e-4, "frame %d: relMSE %g" % (f, rel)
        nbad = int((common.u32(img) != common.u32(ref)).any(-1).sum())
        assert nbad == 0, "frame %d: %d pixels of the accumulation image are not bit-exact" % (f, nbad)
        del img
        # --- G-buffer (4 planes) and reservoirs, one plane at a time (133 MB each at 4K)
        g, go = R.readGBuffer(), OR.gbuffer()
        for plane in ("worldPos", "albedo", "normal", "matProps"):
            assert (common.u32(g[plane]) == common.u32(go[plane])).all(), "frame %d: G-buffer %s differs" % (f, plane)
        hit = go["worldPos"][..., 3] > 0.5
        del g
        r, ro = R.readReservoirs(), OR.reservoirs()
        assert (common.u32(r["info"]) == common.u32(ro["info"])).all(), "frame %d: reservoir M / lightIndex / kind / sampleSeed differ" % f
        wd = np.abs(common.u32(r["weight"]).astype(np.int64) - common.u32(ro["weight"]).astype(np.int64))
        assert wd.max() <= 1, "frame %d: reservoir weights differ by %d ulp" % (f, wd.max())
        del r, wd
        cnt = R.counters()
        assert cnt.hits == int(hit.sum()), "frame %d: hit count %d vs oracle %d" % (f, cnt.hits, int(hit.sum()))
    M = common.u32(OR.reservoirs()["info"])[..., 0]
    frac = float(hit.mean())
    R.destroy()
    return M, frac


def test_smoke_1080p_temporal_full_size(V):
    """BASELINE configs[1] as benchmarked: 1920x1080, 64 lights, M=32, visibility + temporal, 4 frames of the orbit."""
    M, frac = run_workload(V, "smoke_1080p_temporal", 4)
    assert M.max() > 32 and 0.05 < frac < 0.5


def test_bunny_4k_full_full_size(V):
    """BASELINE configs[3] as benchmarked (bench.py default): 3840x2160, 10k lights, M=32, full spatiotemporal k=5 x2, 3 frames."""
    M, frac = run_workload(V, "bunny_4k_full", 3)
    assert M.max() > 32 and 0.05 < frac < 0.6


def test_explosion_1080p_full_full_size(V):
    """BASELINE configs[2] shape as benchmarked: emissive-voxel lights, full spatiotemporal, 3 frames."""
    M, frac = run_workload(V, "explosion_1080p_full", 3)
    assert M.max() > 32
