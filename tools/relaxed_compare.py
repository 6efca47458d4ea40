#!/usr/bin/env python3
"""What bit-exact fp32 costs: the same frames through libvrs.so (IEEE division / square root, no FMA contraction — bit-exact
against the oracle) and through libvrs_relaxed.so (-fmad=true -prec-div=false -prec-sqrt=false), each in its own process.
Prints frame time of both, relMSE of the accumulated image and how many hit pixels ended with another light index.
usage: python tools/relaxed_compare.py <workload> [frames]      (child mode: --child <out.npz>)"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(name, frames, out):
    import bench
    import vrs_pkg
    V = vrs_pkg.load()
    wl = bench.WORKLOADS[name]
    R = V.Renderer(wl["W"], wl["H"], spatial_iterations=wl["iters"])
    R.loadVDB(bench.asset_path(V, wl["asset"]))
    lights, ctr, diag = bench.build_scene_inputs(V, wl, R)
    R.createRestirLights(lights)
    u = R.m_restirUniforms
    u.initialLightSampleCount, u.spatialNeighbors, u.flags = wl["M"], wl["k"], wl["flags"]
    R.CameraManip.setLookat(bench.orbit_eye(ctr, bench.ORBIT_RADIUS * diag, 0.0, 0.0), ctr)
    R.createRestirUniformBuffer()
    t0 = None
    for f in range(frames + 24):
        if f == 24:
            R.synchronize(); t0 = time.perf_counter()
        R.CameraManip.setLookat(bench.orbit_eye(ctr, bench.ORBIT_RADIUS * diag, 0.0, bench.ORBIT_DEG * f), ctr)
        R.renderFrame(clock=f)
    R.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / frames
    r = R.readReservoirs()
    np.savez(out, ms=ms, image=R.readFrame(), info=r["info"].view(np.uint32), hit=R.readGBuffer()["worldPos"][..., 3] > 0.5)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[3], int(sys.argv[4]), sys.argv[2])
        sys.exit(0)
    name = sys.argv[1]
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    res = {}
    with tempfile.TemporaryDirectory() as d:
        for tag, lib in (("exact", "libvrs.so"), ("relaxed", "libvrs_relaxed.so")):
            out = os.path.join(d, tag + ".npz")
            subprocess.check_call([sys.executable, __file__, "--child", out, name, str(frames)], env=dict(os.environ, VRS_LIB=lib))
            res[tag] = dict(np.load(out))
    a, b = res["exact"], res["relaxed"]
    ia, ib = a["image"][..., :3].astype(np.float64), b["image"][..., :3].astype(np.float64)
    rel = float(np.mean((ia - ib) ** 2) / max(np.mean(ia ** 2), 1e-30))
    both = a["hit"] & b["hit"]
    mism = float((a["info"][..., 1][both] != b["info"][..., 1][both]).mean())
    print("%s, %d frames of the orbit: exact %.4f ms/frame, relaxed %.4f ms/frame (%.1f %% faster); hit pixels %d vs %d; relMSE of the last "
          "frame %.3g; light index differs on %.3f %% of the common hit pixels (after %d frames of temporal + spatial reuse: one early "
          "difference propagates)" % (name, frames, float(a["ms"]), float(b["ms"]), 100.0 * (float(a["ms"]) / float(b["ms"]) - 1.0),
                                      int(a["hit"].sum()), int(b["hit"].sum()), rel, 100.0 * mism, frames + 24))
