#!/usr/bin/env python3
"""Freeze the oracle's frame-level behaviour: SHA-256 digests of every plane the oracle produces (G-buffer, reservoirs,
RNG / voxel / cell trace, accumulated image) for a few small configurations -> tests/golden/oracle_frame_digests.json.
The volumetric half of the oracle (tracking, gradient normals, the spatial neighbour loop) has no reference implementation
to be pinned to; these digests make every change of that specification a deliberate, reviewed regeneration
of this file instead of a silent drift.  tests/test_oracle_pinning.py recomputes and compares them on every run.
usage: python tools/gen_frame_digests.py [--check]"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden", "oracle_frame_digests.json")

# name: (asset, W, H, lights, white, M, k, iterations, flags, frames, orbit step in degrees)
CASES = {
    "cube_initial_visibility": ("cube", 96, 64, 1, True, 32, 0, 0, 1, 1, 0.0),
    "smoke_temporal_orbit": ("smoke", 128, 72, 64, False, 32, 0, 0, 1 | 2, 3, 6.0),
    "smoke_full_spatiotemporal": ("smoke", 128, 72, 64, False, 16, 5, 2, 1 | 2 | 4, 3, 2.0),
    "smoke_unbiased_flags": ("smoke", 96, 54, 16, False, 8, 3, 1, 2 | 4 | 16 | 32, 2, 0.0),
}


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_case(O, common, spec):
    asset, W, H, n_lights, white, M, k, iters, flags, frames, step = spec
    import grid_py
    g = grid_py.read_vrsg(common.asset(asset))
    probe = common.oracle_scene(O, asset, np.ones((1, 8), np.float32))
    lo, hi = probe.world_bbox()
    lights = O.generate_point_lights(lo, hi, white, n_lights)
    scene = common.oracle_scene(O, asset, lights)
    ctr = [(a + b) * 0.5 for a, b in zip(lo, hi)]
    diag = float(np.sqrt(sum(((b - a) * 0.5) ** 2 for a, b in zip(lo, hi))))
    OR = O.OracleRenderer(scene, W, H, spatial_iterations=iters)
    prev = None
    out = {}
    for f in range(frames):
        cam = O.Camera(common.orbit_eye(ctr, 1.6 * diag, 0.2 * diag, 30.0 + step * f), ctr)
        gu = O.global_uniforms(cam, W, H)
        ru = O.restir_uniforms(cam, prev, W, H, n_lights, M=M, flags=flags, k=max(k, 1) if flags & 4 else 5)
        pc = O.PushConstant(0, 0, 0, f, 1 if f == 0 else 0)
        img = OR.render(gu, ru, pc, f)
        prev = cam
        d = {"image": digest(img), "trace": digest(OR.f.trace)}
        for name, plane in OR.gbuffer().items():
            d["g_" + name] = digest(plane)
        for name, plane in OR.reservoirs().items():
            d["r_" + name] = digest(plane)
        d["hit_pixels"] = int((OR.gbuffer()["worldPos"][..., 3] > 0.5).sum())
        out["frame%d" % f] = d
    del g
    return out


def compute():
    import common
    import oracle as O
    O.build()
    return {name: run_case(O, common, spec) for name, spec in CASES.items()}


def main():
    got = compute()
    if "--check" in sys.argv:
        want = json.load(open(OUT))
        bad = [(c, f, k) for c in want for f in want[c] for k in want[c][f] if got.get(c, {}).get(f, {}).get(k) != want[c][f][k]]
        print("digests match" if not bad else "MISMATCH: %s" % bad[:10])
        sys.exit(1 if bad else 0)
    json.dump(got, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT, {c: got[c]["frame0"]["hit_pixels"] for c in got})


if __name__ == "__main__":
    main()
