"""not gpu: the screen-space coverage mask of k_cover ("Coverage culling"), restated in numpy with the same
fp32 formulas, against the oracle's own primary hits: no pixel whose ray finds a collision may lie in an unmarked tile,
from far, grazing and close-up cameras.  (The product-side equality of culled and unculled frames is a GPU test,
tests/test_gpu_parity.py::test_coverage_culling_changes_nothing.)"""
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

TILE = 8


def cover_mask(scene, gu, W, H):
    """k_cover + its host set-up (vrs_api.cu make_params): returns (mask[tiles_y, tiles_x] or None when the mask is unusable)."""
    f32 = np.float32
    s = scene.c
    vi = np.array(gu.viewInverse[:], np.float64).reshape(4, 4).T      # column-major -> math layout
    pi = np.array(gu.projInverse[:], np.float64).reshape(4, 4).T
    M = (np.linalg.inv(pi) @ np.linalg.inv(vi)).astype(f32)
    cm = scene.cellmax                                                # [cz][cy][cx]
    tiles_x, tiles_y = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    mask = np.zeros((tiles_y, tiles_x), bool)
    cz, cy, cx = np.nonzero(cm > 0)
    A, B = f32(s.A), [f32(s.B[a]) for a in range(3)]
    vmin = [int(s.vmin[a]) for a in range(3)]
    for c in zip(cx.tolist(), cy.tolist(), cz.tolist()):
        xs, ys = [], []
        for k in range(8):
            w = [(f32(vmin[a] + 8 * (c[a] + ((k >> a) & 1))) - f32(0.5)) * A + B[a] for a in range(3)]
            q = [f32(f32(f32(M[r, 0] * w[0]) + f32(M[r, 1] * w[1])) + f32(M[r, 2] * w[2])) + M[r, 3] for r in range(4)]
            if not q[3] > f32(1e-6):
                return None
            xs.append((q[0] / q[3] + f32(1)) * f32(0.5) * f32(W))
            ys.append((q[1] / q[3] + f32(1)) * f32(0.5) * f32(H))
        x0, x1, y0, y1 = math.floor(min(xs)) - 2, math.ceil(max(xs)) + 2, math.floor(min(ys)) - 2, math.ceil(max(ys)) + 2
        if x1 < 0 or y1 < 0 or x0 > W - 1 or y0 > H - 1:
            continue
        tx0, ty0, tx1, ty1 = max(x0, 0) // TILE, max(y0, 0) // TILE, min(x1, W - 1) // TILE, min(y1, H - 1) // TILE
        if (tx1 - tx0 + 1) * (ty1 - ty0 + 1) > 4096:
            return None
        mask[ty0:ty1 + 1, tx0:tx1 + 1] = True
    return mask


@pytest.mark.parametrize("asset,eyes", [("smoke", [(1.25, 0.0, 0.0), (1.25, 0.0, 137.0), (3.0, 0.4, 60.0), (0.8, 0.9, 250.0), (0.45, 0.0, 20.0)]),
                                        ("cube", [(1.4, 0.3, 45.0), (2.5, -0.5, 200.0)])])
def test_no_hit_pixel_outside_the_coverage_mask(O, asset, eyes):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import common
    lights = np.ones((1, 8), np.float32)
    scene = common.oracle_scene(O, asset, lights)
    lo, hi = scene.world_bbox()
    ctr = [(a + b) * 0.5 for a, b in zip(lo, hi)]
    diag = math.sqrt(sum(((b - a) * 0.5) ** 2 for a, b in zip(lo, hi)))
    W, H = 200, 120
    culled_somewhere = False
    for scale, height, ang in eyes:
        cam = O.Camera(common.orbit_eye(ctr, scale * diag, height * diag, ang), ctr)
        gu = O.global_uniforms(cam, W, H)
        ru = O.restir_uniforms(cam, None, W, H, 1, M=1, flags=0)
        OR = O.OracleRenderer(scene, W, H, spatial_iterations=0)
        OR.render(gu, ru, O.PushConstant(0, 0, 0, 0, 1), 7)
        hit = OR.gbuffer()["worldPos"][..., 3] > 0.5
        mask = cover_mask(scene, gu, W, H)
        if mask is None:
            continue                                                     # the product switches culling off for such a frame
        per_px = np.repeat(np.repeat(mask, TILE, 0), TILE, 1)[:H, :W]
        assert not (hit & ~per_px).any(), "%s eye %s: %d hit pixels in unmarked tiles" % (asset, (scale, height, ang), int((hit & ~per_px).sum()))
        culled_somewhere |= bool((~per_px).any())
    assert culled_somewhere, "the mask never excluded a tile: the test would not notice a mask that is always full"
