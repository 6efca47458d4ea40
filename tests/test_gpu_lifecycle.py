"""-m gpu: lifecycle rows of the drop-in boundary and the defects ADVICE.md round 1 listed:
vrs_resize, a light-table change in the middle of a sequence (stale reservoir indices), a sparse grid whose window has
more than 2^32 voxels (hit voxel no longer packed into 32 bits), the work counters."""
import ctypes as C
import os

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu
FLAGS = 1 | 2 | 4


def setup(V, R, n_lights, M=8, k=3):
    gi = R.gridInfo()
    lo, hi = list(gi.world_bbox_min), list(gi.world_bbox_max)
    lights = V.generate_point_lights(lo, hi, False, n_lights)
    R.createRestirLights(lights)
    u = R.m_restirUniforms
    u.initialLightSampleCount, u.spatialNeighbors, u.flags = M, k, FLAGS
    ctr = [(a + b) * 0.5 for a, b in zip(lo, hi)]
    diag = float(np.sqrt(sum(((b - a) * 0.5) ** 2 for a, b in zip(lo, hi))))
    return ctr, diag


def frames(R, ctr, diag, n, first=0):
    for f in range(first, first + n):
        R.CameraManip.setLookat(common.orbit_eye(ctr, 1.4 * diag, 0.1 * diag, 20.0 + 3.0 * f), ctr)
        R.renderFrame(clock=f)


def test_resize_equals_fresh_context(V):
    A = V.Renderer(200, 120, spatial_iterations=1)
    A.loadVDB(common.asset("smoke"))
    ctr, diag = setup(V, A, 16)
    A.CameraManip.setLookat(common.orbit_eye(ctr, 1.4 * diag, 0.1 * diag, 20.0), ctr)
    A.createRestirUniformBuffer()
    frames(A, ctr, diag, 3)
    for (w, h) in [(333, 187), (64, 48), (200, 120)]:
        A.resize(w, h)
        B = V.Renderer(w, h, spatial_iterations=1)
        B.loadVDB(common.asset("smoke"))
        setup(V, B, 16)
        for R in (A, B):
            R.m_pcRestirPost.frame, R.m_pcRestirPost.initialize, R._ref_cam = 0, 1, None
            R.CameraManip.setLookat(common.orbit_eye(ctr, 1.4 * diag, 0.1 * diag, 20.0), ctr)
            R.createRestirUniformBuffer()
            frames(R, ctr, diag, 3)
        a, b = A.readFrame(), B.readFrame()
        assert a.shape == (h, w, 4) and (common.u32(a) == common.u32(b)).all()
        ra, rb = A.readReservoirs(), B.readReservoirs()
        assert (common.u32(ra["info"]) == common.u32(rb["info"])).all() and (common.u32(ra["weight"]) == common.u32(rb["weight"])).all()
        B.destroy()
    with pytest.raises(V.VrsError):
        A.resize(0, 10)
    A.destroy()


def test_new_lights_drop_the_temporal_history(V):
    """vrs_set_lights with FEWER lights after frames with many: the stored reservoirs hold indices into the old table; the next
    frame must not read them (out of bounds) nor merge them, and equals a fresh context's first frame with the new lights."""
    A = V.Renderer(160, 96, spatial_iterations=1)
    A.loadVDB(common.asset("smoke"))
    ctr, diag = setup(V, A, 500)
    A.CameraManip.setLookat(common.orbit_eye(ctr, 1.4 * diag, 0.1 * diag, 20.0), ctr)
    A.createRestirUniformBuffer()
    frames(A, ctr, diag, 4)
    setup(V, A, 3)                                       # 500 -> 3 lights
    B = V.Renderer(160, 96, spatial_iterations=1)
    B.loadVDB(common.asset("smoke"))
    setup(V, B, 3)
    B.CameraManip.setLookat(common.orbit_eye(ctr, 1.4 * diag, 0.1 * diag, 20.0 + 3.0 * 3), ctr)
    B.createRestirUniformBuffer()
    frames(A, ctr, diag, 1, first=4)
    frames(B, ctr, diag, 1, first=4)
    ra, rb = A.readReservoirs(), B.readReservoirs()
    assert (common.u32(ra["info"]) == common.u32(rb["info"])).all() and (common.u32(ra["weight"]) == common.u32(rb["weight"])).all()
    assert common.u32(ra["info"])[..., 1].max() < 3
    frames(A, ctr, diag, 2, first=5)                     # and temporal reuse resumes on the new table
    assert common.u32(A.readReservoirs()["info"])[..., 0].max() > 8
    A.destroy(); B.destroy()


def test_counters(V):
    R = V.Renderer(128, 72, spatial_iterations=0)
    R.loadVDB(common.asset("smoke"))
    ctr, diag = setup(V, R, 4)
    R.m_restirUniforms.flags = 1
    R.CameraManip.setLookat(common.orbit_eye(ctr, 1.4 * diag, 0.0, 0.0), ctr)
    R.createRestirUniformBuffer()
    R.renderFrame(clock=0)
    c = R.counters()
    hit = int((R.readGBuffer()["worldPos"][..., 3] > 0.5).sum())
    assert c.hits == hit and c.candidates >= hit and c.shadow_rays <= hit and c.temporal_out_of_halo == 0 and c.comm_timeouts == 0
    assert c.temporal_reach_rows == 0                                    # no temporal pass ran yet
    R.m_restirUniforms.flags = 1 | 2
    for f in range(1, 4):                                                # orbit: the reprojection moves rows, the counter records how far
        R.CameraManip.setLookat(common.orbit_eye(ctr, 1.4 * diag, 0.0, 12.0 * f), ctr)
        R.renderFrame(clock=f)
    assert 0 < R.counters().temporal_reach_rows < 72
    R.setKernelTiming(True)
    R.renderFrame(clock=1)
    kt = R.kernelTimes()
    names = [n for n, _ in kt]
    assert "k_primary" in names and "k_ris" in names and "k_shade" in names and all(ms >= 0 for _, ms in kt)
    R.setKernelTiming(False)
    R.destroy()


def test_window_above_2_pow_32_voxels(V, O, tmp_path):
    """Two far-apart bricks: the leaf-aligned window is 2048 x 2048 x 1024 = 2^32 voxels.  The hit voxel travels between the
    kernels as (cell, offset), so the G-buffer of every hit must still describe the voxel the hit position lies in."""
    import vdb_write as W
    rng = np.random.default_rng(5)
    g = W.Grid("density", background=0.0, half=False, compression=W.COMPRESS_ACTIVE_MASK, voxel_size=1.0)
    vals = {}
    for o in [(0, 0, 0), (2040, 2040, 1016), (2032, 2040, 1016), (2040, 2032, 1016)]:
        v = rng.uniform(0.5, 3.0, 512).astype(np.float32)
        g.set_leaf(o, v, np.ones(512, bool))
        vals[o] = v
    path = str(tmp_path / "sparse.vdb")
    W.write_vdb(path, [g], version=224)
    R = V.Renderer(160, 120, spatial_iterations=0, density_scale=40.0)
    R.loadVDB(path)
    gi = R.gridInfo()
    assert (gi.bbox_max[0] - gi.bbox_min[0] + 1) * (gi.bbox_max[1] - gi.bbox_min[1] + 1) * (gi.bbox_max[2] - gi.bbox_min[2] + 1) >= 2 ** 32
    A, B = 0.05, np.array([-2.5, 0.5, 0.0])
    ctr = A * np.array([2040.0, 2040.0, 1020.0]) + B
    R.createRestirLights(V.generate_point_lights(list(ctr - 1.0), list(ctr + 1.0), True, 4))
    u = R.m_restirUniforms
    u.initialLightSampleCount, u.flags = 4, 1
    R.CameraManip.setLookat(tuple(ctr + np.array([1.5, 1.2, 1.0])), tuple(ctr))
    R.createRestirUniformBuffer()
    R.renderFrame(clock=0)
    gb = R.readGBuffer()
    hit = gb["worldPos"][..., 3] > 0.5
    assert hit.sum() > 200
    P = gb["worldPos"][hit][:, :3].astype(np.float64)
    ijk = np.floor((P - B) / A + 0.5).astype(np.int64)
    out4 = np.zeros(4, np.float32)
    ok = 0
    for (i, j, k), alb in zip(ijk, gb["albedo"][hit]):
        o = (int(i) & ~7, int(j) & ~7, int(k) & ~7)
        cands = []
        for di, dj, dk in [(0, 0, 0), (1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1)]:   # position rounding at a voxel face
            ii, jj, kk = int(i) + di, int(j) + dj, int(k) + dk
            oo = (ii & ~7, jj & ~7, kk & ~7)
            if oo in vals:
                O.lib().orc_voxel_albedo(float(vals[oo][((ii & 7) << 6) | ((jj & 7) << 3) | (kk & 7)]), out4.ctypes.data_as(C.c_void_p))
                cands.append(out4.copy())
        ok += any((common.u32(c) == common.u32(alb)).all() for c in cands)
    assert ok == int(hit.sum()), "%d of %d hits carry the material of another voxel" % (int(hit.sum()) - ok, int(hit.sum()))
    R.destroy()
