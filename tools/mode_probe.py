#!/usr/bin/env python3
"""Frame time of one workload under the launch modes of the runtime (graph replay with / without overlap, eager, eager with
per-kernel events), each over whole orbits with the uniforms precomputed: python tools/mode_probe.py <workload> [orbits]"""
import ctypes as C
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import vrs_pkg  # noqa: E402

V = vrs_pkg.load()
name = sys.argv[1]
orbits = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = bench.WORKLOADS[name]
R = V.Renderer(wl["W"], wl["H"], spatial_iterations=wl["iters"])
R.loadVDB(bench.asset_path(V, wl["asset"]))
lights, ctr, diag = bench.build_scene_inputs(V, wl, R)
R.createRestirLights(lights)
u = R.m_restirUniforms
u.initialLightSampleCount, u.spatialNeighbors, u.flags = wl["M"], wl["k"], wl["flags"]
R.CameraManip.setLookat(bench.orbit_eye(ctr, 1.25 * diag, 0.0, 0.0), ctr)
R.createRestirUniformBuffer()
L = V.lib()
inputs = []
for f in range(60 * orbits * 6 + 40):
    R.CameraManip.setLookat(bench.orbit_eye(ctr, 1.25 * diag, 0.0, 6.0 * f), ctr)
    R.updateUniformBuffer(); R.updateRestirUniformBuffer(); R.updateFrame()
    inputs.append((V.GlobalUniforms.from_buffer_copy(R.m_globalUniforms), V.RestirUniforms.from_buffer_copy(R.m_restirUniforms), V.PushConstantRestir.from_buffer_copy(R.m_pcRestirPost)))
fno = [0]


def run(n):
    for _ in range(n):
        gu, ru, pc = inputs[fno[0]]
        s = L.vrs_render_frame(R._ctx, C.byref(gu), C.byref(ru), C.byref(pc), fno[0])
        assert s == 0, L.vrs_last_error(R._ctx)
        fno[0] += 1


def clock():
    try:
        return subprocess.check_output(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"], text=True).strip()
    except Exception:
        return "?"


def timed(label, n):
    run(12); R.synchronize()
    t0 = time.perf_counter(); run(n); R.synchronize(); dt = time.perf_counter() - t0
    print("%-46s %.4f ms/frame   [%s]" % (label, 1e3 * dt / n, clock()))


n = 60 * orbits
timed("graphs, frames in flight (VRS_PIPELINE=%s)" % os.environ.get("VRS_PIPELINE", "4"), n)
R.setPassTiming(True); timed("graphs, per-pass events (no overlap)", n); R.setPassTiming(False)
R.setKernelTiming(True); timed("eager + per-kernel events", n)
acc = {}
for i in range(12):
    run(1)
    for k, ms in R.kernelTimes():
        acc[k] = acc.get(k, 0.0) + ms / 12
print("   kernels: " + "  ".join("%s %.1f" % (k, 1e3 * v) for k, v in acc.items()) + " | sum %.1f us" % (1e3 * sum(acc.values())))
R.setKernelTiming(False)
timed("graphs, frames in flight (again)", n)
