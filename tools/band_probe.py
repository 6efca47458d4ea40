#!/usr/bin/env python3
"""Render one band of the default bench frame on one GPU (no exchange) — used under ncu to see the per-kernel latency
floor that limits strong scaling: python tools/band_probe.py <y0> <y1> [frames]"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import vrs_pkg  # noqa: E402

V = vrs_pkg.load()
y0, y1 = int(sys.argv[1]), int(sys.argv[2])
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 20
wl = bench.WORKLOADS["smoke_1080p_temporal"]
R = V.Renderer(wl["W"], wl["H"], spatial_iterations=0, band=(y0, y1), halo_rows=32)
R.loadVDB(bench.asset_path(V, wl["asset"]))
lights, ctr, diag = bench.build_scene_inputs(V, wl, R)
R.createRestirLights(lights)
u = R.m_restirUniforms
u.initialLightSampleCount, u.spatialNeighbors, u.flags = wl["M"], wl["k"], wl["flags"]
R.CameraManip.setLookat(bench.orbit_eye(ctr, 1.25 * diag, 0.0, 0.0), ctr)
R.createRestirUniformBuffer()
tot = 0.0
for f in range(frames):
    R.CameraManip.setLookat(bench.orbit_eye(ctr, 1.25 * diag, 0.0, 6.0 * f), ctr)
    R.renderFrame(clock=f)
    t = R.timings()
    if f >= 5:
        tot += t.frame_ms
print("band", y0, y1, "mean frame ms", tot / max(1, frames - 5), "last", t.initial_ms, t.shade_ms)
