"""Shared scene construction for the parity tests: the same inputs go to the CUDA path (through the C ABI)
and to the CPU oracle."""
import ctypes as C
import math
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = os.path.join(ROOT, "assets")


def asset(name):
    return os.path.join(ASSETS, name + ".vrsg")


def oracle_scene(O, name, lights, density_scale=10.0):
    import grid_py
    import vdb_py
    g = grid_py.read_vrsg(asset(name))
    raw, vmin, _ = grid_py.dense_raw(g)
    dens = vdb_py.density_from_raw(raw, g.level_set, g.background)
    bg = float(vdb_py.density_from_raw(np.array([g.background], np.float32), g.level_set, g.background)[0])
    return O.OracleScene(dens, vmin, g.voxel_size, g.translation, lights, bg_density=bg, density_scale=density_scale)


def orbit_eye(center, radius, height, angle_deg):
    a = math.radians(angle_deg)
    return (center[0] + radius * math.cos(a), center[1] + height, center[2] + radius * math.sin(a))


def setup_pair(V, O, name, W, H, n_lights, white=False, density_scale=10.0, iterations=2, trace=True, lights_scale=1.0):
    """Returns (product Renderer, OracleRenderer, centre of the grid's world bbox, bbox half-diagonal)."""
    R = V.Renderer(W, H, spatial_iterations=iterations, density_scale=density_scale, enable_trace=trace)
    R.loadVDB(asset(name))
    gi = R.gridInfo()
    lo, hi = list(gi.world_bbox_min), list(gi.world_bbox_max)
    ctr = [(a + b) * 0.5 for a, b in zip(lo, hi)]
    ext = [(b - a) * 0.5 * lights_scale for a, b in zip(lo, hi)]
    lights = V.generate_point_lights([c - e for c, e in zip(ctr, ext)], [c + e for c, e in zip(ctr, ext)], white, n_lights)
    R.createRestirLights(lights)
    scene = oracle_scene(O, name, lights, density_scale)
    OR = O.OracleRenderer(scene, W, H, spatial_iterations=iterations)
    diag = math.sqrt(sum(e * e for e in [(b - a) * 0.5 for a, b in zip(lo, hi)]))
    return R, OR, ctr, diag


def oracle_uniforms(O, R):
    """Bit-copy the product's uniform structs into the oracle's mirrors (identical layouts)."""
    gu = O.GlobalUniforms.from_buffer_copy(bytes(R.m_globalUniforms))
    ru = O.RestirUniforms.from_buffer_copy(bytes(R.m_restirUniforms))
    pc = O.PushConstant.from_buffer_copy(bytes(R.m_pcRestirPost))
    return gu, ru, pc


def rel_mse(a, b):
    a = a[..., :3].astype(np.float64); b = b[..., :3].astype(np.float64)
    return float(np.mean((a - b) ** 2) / max(np.mean(b ** 2), 1e-30))


def u32(a):
    return np.ascontiguousarray(a).view(np.uint32)
