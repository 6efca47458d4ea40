#!/usr/bin/env python3
"""A few frames of a bench workload, nothing else — the command ncu wraps (tools/prof.sh).  usage: ncu_frames.py <workload> [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import vrs_pkg  # noqa: E402

V = vrs_pkg.load()
name = sys.argv[1]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 5
wl = bench.WORKLOADS[name]
R = V.Renderer(wl["W"], wl["H"], spatial_iterations=wl["iters"])
R.loadVDB(bench.asset_path(V, wl["asset"]))
lights, ctr, diag = bench.build_scene_inputs(V, wl, R)
R.createRestirLights(lights)
u = R.m_restirUniforms
u.initialLightSampleCount, u.spatialNeighbors, u.flags = wl["M"], wl["k"], wl["flags"]
R.CameraManip.setLookat(bench.orbit_eye(ctr, bench.ORBIT_RADIUS * diag, 0.0, 0.0), ctr)
R.createRestirUniformBuffer()
for f in range(frames):
    R.CameraManip.setLookat(bench.orbit_eye(ctr, bench.ORBIT_RADIUS * diag, 0.0, bench.ORBIT_DEG * f), ctr)
    R.renderFrame(clock=f)
R.synchronize()
print("rendered", frames, "frames of", name, "hits", R.counters().hits)
