#!/usr/bin/env python3
"""Render one band of a bench workload on one GPU (no exchange) and print the LIVE per-kernel event times (no profiler, warm
caches) — the latency floor that limits strong scaling: python tools/band_probe.py <workload> <y0> <y1> [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import vrs_pkg  # noqa: E402

V = vrs_pkg.load()
name, y0, y1 = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 24
wl = bench.WORKLOADS[name]
R = V.Renderer(wl["W"], wl["H"], spatial_iterations=wl["iters"], band=(y0, y1), halo_rows=32)
R.loadVDB(bench.asset_path(V, wl["asset"]))
lights, ctr, diag = bench.build_scene_inputs(V, wl, R)
R.createRestirLights(lights)
u = R.m_restirUniforms
u.initialLightSampleCount, u.spatialNeighbors, u.flags = wl["M"], wl["k"], wl["flags"]
R.CameraManip.setLookat(bench.orbit_eye(ctr, 1.25 * diag, 0.0, 0.0), ctr)
R.createRestirUniformBuffer()
R.setPassTiming(True)
tot, n = 0.0, 0
for f in range(frames):                       # graph replay: frame time as the bench sees it
    R.CameraManip.setLookat(bench.orbit_eye(ctr, 1.25 * diag, 0.0, 6.0 * f), ctr)
    R.renderFrame(clock=f)
    t = R.timings()
    if f >= 8:
        tot += t.frame_ms; n += 1
print("band %d..%d of %s: frame %.4f ms (graph replay, no overlap), hits %d" % (y0, y1, name, tot / n, R.counters().hits))
R.setPassTiming(False)
import time  # noqa: E402
for rep in range(2):
    R.synchronize(); t0 = time.perf_counter()
    for f in range(frames, frames + 200):
        R.CameraManip.setLookat(bench.orbit_eye(ctr, 1.25 * diag, 0.0, 6.0 * f), ctr)
        R.renderFrame(clock=f)
    R.synchronize()
    print("  200 frames back to back (frames in flight%s): %.4f ms per frame (wall, includes the host-side uniform producers)" % (" OFF" if os.environ.get("VRS_PIPELINE") == "0" else "", 1e3 * (time.perf_counter() - t0) / 200))
frames += 200
R.setKernelTiming(True)
acc = {}
for f in range(frames, frames + 12):
    R.CameraManip.setLookat(bench.orbit_eye(ctr, 1.25 * diag, 0.0, 6.0 * f), ctr)
    R.renderFrame(clock=f)
    if f >= frames + 2:
        for k, ms in R.kernelTimes():
            acc[k] = acc.get(k, 0.0) + ms / 10.0
print("  per kernel (eager, events): " + "  ".join("%s %.1f us" % (k, 1e3 * v) for k, v in acc.items()) + "  | sum %.1f us" % (1e3 * sum(acc.values())))
