"""Unbiasedness against a brute-force path-traced estimate of the same integrand (BASELINE.json north_star:
"stay unbiased against a 4096-spp path-traced reference").

The unbiased configuration is FINALIZE_W + FINAL_VISIBILITY (+ temporal + spatial reuse with the 1/Z normalisation);
the reference's literal selection-time `w` is biased and the test documents by how much.  Firefly clamp off."""
import numpy as np
import pytest

import common


def scene(O, V, W, H, n_lights=16):
    lights = V.generate_point_lights([-4.2, -0.2, -1.4], [-1.5, 5.2, 1.6], False, n_lights)
    sc = common.oracle_scene(O, "smoke", lights)
    lo, hi = sc.world_bbox()
    ctr = [(a + b) * 0.5 for a, b in zip(lo, hi)]
    cam = O.Camera((ctr[0] + 3.0, ctr[1] + 0.4, ctr[2] + 3.0), ctr)
    return sc, cam, lights, ctr


def oracle_mean(O, sc, cam, W, H, n_lights, flags, frames, M=8, iters=1):
    gu = O.global_uniforms(cam, W, H)
    ru = O.restir_uniforms(cam, cam, W, H, n_lights, M=M, flags=flags, k=3, radius=8.0, firefly=1e30)
    pc = O.PushConstant(0, 0, 0, 0, 1)
    R = O.OracleRenderer(sc, W, H, spatial_iterations=iters)
    acc = np.zeros((H, W, 3), np.float64)
    for c in range(frames):
        acc += R.render(gu, ru, pc, c)[..., :3]
    return acc / frames, gu, ru


def test_oracle_restir_is_unbiased_and_reference_w_is_not(O, V):
    W = H = 64
    sc, cam, lights, _ = scene(O, V, W, H)
    unbiased = O.FLAG_FINALIZE_W | O.FLAG_FINAL_VISIBILITY | O.FLAG_TEMPORAL | O.FLAG_SPATIAL
    img, gu, ru = oracle_mean(O, sc, cam, W, H, len(lights), unbiased, 240)
    ru.flags = O.FLAG_VISIBILITY
    pt = 0.5 * (O.path_trace(sc, gu, ru, 4096, 3).astype(np.float64) + O.path_trace(sc, gu, ru, 4096, 5))
    ratio = img.mean() / pt.mean()
    assert abs(ratio - 1.0) < 0.03, ratio
    # initial-only, both ways
    a, _, _ = oracle_mean(O, sc, cam, W, H, len(lights), O.FLAG_FINALIZE_W | O.FLAG_FINAL_VISIBILITY, 160)
    assert abs(a.mean() / pt.mean() - 1.0) < 0.03
    b, _, _ = oracle_mean(O, sc, cam, W, H, len(lights), O.FLAG_VISIBILITY, 160)
    assert b.mean() / pt.mean() > 1.15, "reference selection-time w (reservoir.glsl:51) should show its bias here"


@pytest.mark.gpu
def test_gpu_accumulation_matches_path_traced(O, V):
    """The CUDA path's own 256-frame accumulation (restir_post.frag running mean) against the 4096-spp estimate."""
    W, H = 96, 64
    R, OR, ctr, diag = common.setup_pair(V, O, "smoke", W, H, 16, iterations=1, trace=False)
    u = R.m_restirUniforms
    u.initialLightSampleCount, u.spatialNeighbors, u.spatialRadius, u.fireflyClampThreshold = 8, 3, 8.0, 1e30
    u.flags = V.FINALIZE_W_FLAG | V.FINAL_VISIBILITY_FLAG | V.TEMPORAL_REUSE_FLAG | V.SPATIAL_REUSE_FLAG
    R.CameraManip.setLookat((ctr[0] + 3.0, ctr[1] + 0.4, ctr[2] + 3.0), ctr)
    R.createRestirUniformBuffer()
    R.m_pcRestirPost.initialize = 0                 # accumulate from frame 1 (frame 0 stores)
    for f in range(256):
        R.renderFrame(clock=f)
    img = R.readFrame()[..., :3].astype(np.float64)
    gu, ru, pc = common.oracle_uniforms(O, R)
    ru.flags = O.FLAG_VISIBILITY
    pt = O.path_trace(OR.scene, gu, ru, 4096, 11).astype(np.float64)
    assert abs(img.mean() / pt.mean() - 1.0) < 0.03, img.mean() / pt.mean()
    # per-pixel agreement where there is signal: coarse 8x8 block means within 15 %
    bi = img.reshape(H // 8, 8, W // 8, 8, 3).mean((1, 3, 4)); bp = pt.reshape(H // 8, 8, W // 8, 8, 3).mean((1, 3, 4))
    m = bp > 0.25 * bp.max()
    assert np.abs(bi[m] / bp[m] - 1.0).max() < 0.15
    R.destroy()
