"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Bar: RNG streams, voxel indices and chosen light indices bit-exact; images relMSE <= 1e-4 (BASELINE.json north_star)."""
import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


def run_frames(V, O, name, W, H, n_lights, flags, frames, M=32, k=5, iterations=2, orbit_step=6.0, white=False, trace=True, eye=1.6):
    R, OR, ctr, diag = common.setup_pair(V, O, name, W, H, n_lights, white=white, iterations=iterations, trace=trace)
    R.m_restirUniforms.initialLightSampleCount = M
    R.m_restirUniforms.spatialNeighbors = k
    R.m_restirUniforms.flags = flags
    R.CameraManip.setLookat(common.orbit_eye(ctr, eye * diag, 0.2 * diag, 30.0), ctr)
    R.createRestirUniformBuffer()
    out = []
    for f in range(frames):
        R.CameraManip.setLookat(common.orbit_eye(ctr, eye * diag, 0.2 * diag, 30.0 + orbit_step * f), ctr)
        R.renderFrame(clock=f)
        gu, ru, pc = common.oracle_uniforms(O, R)
        pc.initialize = R._last_initialize      # main.cpp:441-443 flips it after the submit; replay the value this frame used
        img_o = OR.render(gu, ru, pc, f).copy()
        img_p = R.readFrame()
        out.append((img_p, img_o, R.readGBuffer(), {k2: v.copy() for k2, v in OR.gbuffer().items()}, R.readReservoirs(),
                    {k2: v.copy() for k2, v in OR.reservoirs().items()}, R.readTrace() if trace else None, OR.f.trace.copy()))
    R.destroy()
    return out


def check(frames, exact_images=True):
    for i, (img_p, img_o, g_p, g_o, r_p, r_o, t_p, t_o) in enumerate(frames):
        assert t_p is None or (t_p == t_o).all(), "frame %d: trace (voxel code / collisions / cells / final RNG state) differs at %d px" % (i, (t_p != t_o).any(-1).sum())
        for plane in ("worldPos", "albedo", "normal", "matProps"):
            assert (common.u32(g_p[plane]) == common.u32(g_o[plane])).all(), "frame %d: G-buffer %s differs" % (i, plane)
        assert (common.u32(r_p["info"]) == common.u32(r_o["info"])).all(), "frame %d: reservoir M/lightIndex/kind/seed differ" % i
        wdiff = np.abs(common.u32(r_p["weight"]).astype(np.int64) - common.u32(r_o["weight"]).astype(np.int64))
        assert wdiff.max() <= 1, "frame %d: reservoir weights differ by more than 1 ulp (max %d)" % (i, wdiff.max())
        assert common.rel_mse(img_p, img_o) <= 1e-4, "frame %d: relMSE %g" % (i, common.rel_mse(img_p, img_o))
        if exact_images:
            assert (common.u32(img_p) == common.u32(img_o)).all(), "frame %d: accumulation image not bit-exact" % i


def test_config1_cube_initial_only(V, O):
    """BASELINE.json configs[0]: cube.vdb 256x256, 1 point light, RIS M=32, no reuse, 1 frame."""
    frames = run_frames(V, O, "cube", 256, 256, 1, V.VISIBILITY_REUSE_FLAG, 1, white=True)
    check(frames)
    assert frames[0][2]["worldPos"][..., 3].mean() > 0.05      # the cube is in view


def test_smoke_initial_visibility(V, O):
    check(run_frames(V, O, "smoke", 320, 180, 64, V.VISIBILITY_REUSE_FLAG, 2))


def test_smoke_temporal_orbit(V, O):
    """configs[1] at reduced size: smoke.vdb, 64 lights, RIS M=32 + temporal reuse, orbiting camera."""
    frames = run_frames(V, O, "smoke", 320, 180, 64, V.VISIBILITY_REUSE_FLAG | V.TEMPORAL_REUSE_FLAG, 6)
    check(frames)
    M = common.u32(frames[-1][4]["info"])[..., 0]
    assert M.max() > 32, "temporal reuse never merged a previous reservoir"


def test_smoke_full_spatiotemporal(V, O):
    """configs[2] shape: full spatiotemporal ReSTIR, k=5 neighbours, 2 spatial iterations."""
    flags = V.VISIBILITY_REUSE_FLAG | V.TEMPORAL_REUSE_FLAG | V.SPATIAL_REUSE_FLAG
    frames = run_frames(V, O, "smoke", 320, 180, 64, flags, 4, orbit_step=2.0)
    check(frames)


def test_smoke_unbiased_flags(V, O):
    flags = V.FINAL_VISIBILITY_FLAG | V.FINALIZE_W_FLAG | V.TEMPORAL_REUSE_FLAG | V.SPATIAL_REUSE_FLAG
    check(run_frames(V, O, "smoke", 256, 144, 64, flags, 3, orbit_step=0.0))


def test_device_grid_lookup_matches_host_and_oracle(V, O):
    import grid_py
    import vdb_py
    R = V.Renderer(64, 64)
    R.loadVDB(common.asset("smoke"))
    g = grid_py.read_vrsg(common.asset("smoke"))
    raw, vmin, vdim = grid_py.dense_raw(g)
    dens = vdb_py.density_from_raw(raw, g.level_set, g.background)
    rng = np.random.default_rng(7)
    ijk = np.stack([rng.integers(vmin[a] - 20, vmin[a] + vdim[a] + 20, 200000) for a in range(3)], 1).astype(np.int32)
    got = R.gridSampleDevice(ijk)
    loc = ijk - np.array(vmin)
    inside = ((loc >= 0) & (loc < np.array(vdim))).all(1)
    exp = np.zeros(len(ijk), np.float32)
    exp[inside] = dens[loc[inside, 2], loc[inside, 1], loc[inside, 0]]
    assert (got.view(np.uint32) == exp.view(np.uint32)).all()
    R.destroy()


@pytest.mark.parametrize("name,eye", [("smoke", 1.6), ("smoke", 3.0), ("smoke", 0.45), ("cube", 1.2)])
def test_coverage_culling_changes_nothing(V, O, name, eye):
    """Without the trace buffer the product culls primary rays whose 8x8-pixel tile no occupied cell projects into
    (k_cover); every pixel, reservoir and the image must still equal the oracle's, which marches every ray."""
    flags = V.VISIBILITY_REUSE_FLAG | V.TEMPORAL_REUSE_FLAG | V.SPATIAL_REUSE_FLAG
    check(run_frames(V, O, name, 331, 187, 16, flags, 4, M=8, k=3, iterations=1, orbit_step=17.0, trace=False, eye=eye))
