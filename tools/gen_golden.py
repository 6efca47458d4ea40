#!/usr/bin/env python3
"""Generate tests/golden/ref_vectors.json by EXECUTING THE REFERENCE'S OWN SOURCES (oracle/_ref/libvrs_ref.so, built by
oracle/ref/build_ref.py from /root/reference).  Run in the container that has /root/reference; the JSON travels.
All fp32 values are stored as their uint32 bit patterns so comparisons are exact."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ref_vectors.json")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32).tolist()


def p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefGInfo(C.Structure):
    _fields_ = [("camPos", C.c_float * 3), ("worldPos", C.c_float * 3), ("normal", C.c_float * 3), ("albedo", C.c_float * 4),
                ("emissive", C.c_float * 3), ("albedoLum", C.c_float), ("roughness", C.c_float), ("metallic", C.c_float),
                ("sampleSeed", C.c_uint32)]


class RefRes(C.Structure):
    _fields_ = [("lightPos", C.c_float * 3), ("M", C.c_uint32), ("lightIndex", C.c_uint32), ("lightKind", C.c_int32),
                ("sampleSeed", C.c_uint32), ("pHat", C.c_float), ("sumWeights", C.c_float), ("w", C.c_float)]


def ginfo_from16(g16, L):
    g = RefGInfo()
    g.camPos[:] = g16[0:3].tolist(); g.worldPos[:] = g16[3:6].tolist(); g.normal[:] = g16[6:9].tolist()
    g.albedo[:] = g16[9:13].tolist(); g.emissive[:] = [0, 0, 0]
    g.albedoLum, g.roughness, g.metallic, g.sampleSeed = float(g16[13]), float(g16[14]), float(g16[15]), 0
    return g


def res_from8(r8):
    r = RefRes()
    f = np.array(r8, np.uint32).view(np.float32)
    r.lightPos[:] = [0, 0, 0]
    r.M, r.lightIndex, r.lightKind, r.sampleSeed = int(r8[0]), int(r8[1]), int(np.int32(r8[2])), int(r8[3])
    r.pHat, r.sumWeights, r.w = float(f[4]), float(f[5]), float(f[6])
    return r


def res_to8(r):
    f = np.array([r.pHat, r.sumWeights, r.w], np.float32).view(np.uint32)
    return [int(r.M), int(r.lightIndex), int(np.uint32(np.int32(r.lightKind))), int(r.sampleSeed), int(f[0]), int(f[1]), int(f[2]), 0]


def random_ginfo(rng, L):
    g = np.zeros(16, np.float32)
    g[0:3] = rng.uniform(-6, 6, 3)                      # camPos
    g[3:6] = rng.uniform(-2, 2, 3)                      # worldPos
    n = rng.normal(size=3); n /= np.linalg.norm(n)
    g[6:9] = n
    g[9:13] = rng.uniform(0, 1, 4)
    g[13] = L.ref_luminance_common(float(g[9]), float(g[10]), float(g[11]))
    g[14] = rng.uniform(0.02, 1.0)
    g[15] = rng.uniform(0, 1) if rng.uniform() < 0.5 else 0.0001
    return g


def main():
    L = O.ref()
    if L is None:
        raise SystemExit("oracle/_ref/libvrs_ref.so missing: run oracle/ref/build_ref.py where /root/reference exists")
    rng = np.random.default_rng(20251017)
    V = {"generated_by": "tools/gen_golden.py from oracle/_ref/libvrs_ref.so (reference sources, g++ -O2 -ffp-contract=off)"}

    # ---- RNG
    o2 = (C.c_uint32 * 2)()
    pts = [(0, 0), (51, 69), (115140, 64740), (3839, 2159)] + [tuple(int(v) for v in rng.integers(0, 2 ** 32, 2)) for _ in range(60)]
    V["pcg2d"] = []
    for x, y in pts:
        L.ref_pcg2d(x, y, o2)
        V["pcg2d"].append([x, y, o2[0], o2[1]])
    V["lcg_rnd"] = []
    for s0 in [507651704, 941672005, 2909577513] + [int(v) for v in rng.integers(0, 2 ** 32, 29)]:
        s = C.c_uint32(s0)
        a = L.ref_lcg(C.byref(s)); b = L.ref_lcg(C.byref(s)); r = L.ref_rnd(C.byref(s))
        V["lcg_rnd"].append([s0, a, b, bits([r])[0], s.value])
    V["luminance"] = []
    for _ in range(32):
        c = rng.uniform(0, 4, 3).astype(np.float32)
        V["luminance"].append(bits(c) + bits([L.ref_luminance_common(*map(float, c)), L.ref_luminance_utils(*map(float, c))]))

    # ---- Disney BRDF
    V["brdf"] = []
    out3 = np.zeros(3, np.float32)
    for _ in range(256):
        cosv = rng.uniform(-0.2, 1.0, 4).astype(np.float32)
        alb = rng.uniform(0, 1, 3).astype(np.float32)
        lum = np.float32(L.ref_luminance_common(*map(float, alb)))
        rough, metal = np.float32(rng.uniform(0, 1)), np.float32(rng.uniform(0, 1))
        fl = L.ref_disneyBrdfLuminance(*map(float, cosv), float(lum), float(rough), float(metal))
        L.ref_disneyBrdfColor(*map(float, cosv), p(alb), float(rough), float(metal), p(out3))
        V["brdf"].append({"in": bits(list(cosv) + list(alb) + [lum, rough, metal]), "lum": bits([fl])[0], "color": bits(out3)})

    # ---- lights / alias table / p-hat
    nl = 64
    lights = O.generate_point_lights((-3, -3, -3), (3, 3, 3), False, nl)   # oracle output, checked against ref below
    ref_l = np.zeros((nl, 8), np.float32)
    L.ref_generatePointLights(p(np.array([-3, -3, -3], np.float32)), p(np.array([3, 3, 3], np.float32)), 0, nl, p(ref_l))
    V["generate_point_lights"] = {"min": [-3, -3, -3], "max": [3, 3, 3], "white": 0, "n": nl, "out": bits(ref_l.ravel())}
    white = np.zeros((5, 8), np.float32)
    L.ref_generatePointLights(p(np.array([-10, -10, -10], np.float32)), p(np.array([10, 10, 10], np.float32)), 1, 5, p(white))
    V["generate_point_lights_white"] = {"min": [-10, -10, -10], "max": [10, 10, 10], "white": 1, "n": 5, "out": bits(white.ravel())}
    lights = ref_l
    V["alias_tables"] = []
    for pdf in [np.array([1, 2, 3, 4], np.float32), np.array([1, 1, 1, 0.27], np.float32), lights[:, 7].copy(),
                rng.uniform(0, 1, 1000).astype(np.float32), np.array([5.0], np.float32), np.array([0, 1, 0, 3, 0], np.float32)]:
        t = np.zeros(len(pdf), dtype=[("alias", "<i4"), ("prob", "<f4"), ("pdf", "<f4"), ("aliasPdf", "<f4")])
        L.ref_createAliasTable(p(pdf), len(pdf), p(t))
        V["alias_tables"].append({"pdf": bits(pdf), "alias": t["alias"].tolist(), "prob": bits(t["prob"]), "pdf_out": bits(t["pdf"]),
                                  "aliasPdf": bits(t["aliasPdf"])})
    table = np.zeros(nl, dtype=[("alias", "<i4"), ("prob", "<f4"), ("pdf", "<f4"), ("aliasPdf", "<f4")])
    L.ref_createAliasTable(p(lights[:, 7].copy()), nl, p(table))
    L.ref_set_scene(p(lights), nl, None, p(table), nl)
    V["scene_lights"] = bits(lights.ravel())
    V["alias_sample"] = []
    idx, pr = C.c_uint32(), C.c_float()
    for _ in range(128):
        r1, r2 = (np.float32(int(v)) / np.float32(16777216.0) for v in rng.integers(0, 2 ** 24, 2))
        L.ref_aliasTableSample(float(r1), float(r2), C.byref(idx), C.byref(pr))
        V["alias_sample"].append(bits([r1, r2]) + [idx.value, bits([pr.value])[0]])

    V["phat"] = []
    for _ in range(256):
        g16 = random_ginfo(rng, L)
        li = int(rng.integers(0, nl))
        g = ginfo_from16(g16, L)
        ph = L.ref_evaluatePHat(li, 0, C.byref(g))
        L.ref_evaluatePHatFull(li, 0, C.byref(g), p(out3))
        V["phat"].append({"g": bits(g16), "light": li, "phat": bits([ph])[0], "full": bits(out3)})

    # ---- initial RIS loop (restir.rgen:203-227)
    V["initial_ris"] = []
    for _ in range(64):
        g16 = random_ginfo(rng, L)
        g = ginfo_from16(g16, L)
        s0 = int(rng.integers(0, 2 ** 32))
        seed = C.c_uint32(s0)
        r = RefRes()
        count = int(rng.choice([1, 4, 32, 64]))
        L.ref_initial_ris(C.byref(g), count, C.byref(seed), C.byref(r))
        V["initial_ris"].append({"g": bits(g16), "count": count, "seed": s0, "res": res_to8(r), "seed_out": seed.value})

    # ---- combineReservoirs (reservoir.glsl:56-90) on reservoirs produced by the RIS loop
    V["combine"] = []
    for _ in range(128):
        ga, gb = random_ginfo(rng, L), random_ginfo(rng, L)
        if rng.uniform() < 0.7:                       # similar geometry, as temporal / spatial reuse sees it
            gb[3:6] = ga[3:6] + rng.normal(scale=0.03, size=3).astype(np.float32)
            gb[6:9] = ga[6:9]
        ra, rb = RefRes(), RefRes()
        sa, sb = C.c_uint32(int(rng.integers(0, 2 ** 32))), C.c_uint32(int(rng.integers(0, 2 ** 32)))
        L.ref_initial_ris(C.byref(ginfo_from16(ga, L)), 32, C.byref(sa), C.byref(ra))
        L.ref_initial_ris(C.byref(ginfo_from16(gb, L)), 32, C.byref(sb), C.byref(rb))
        if rng.uniform() < 0.2:
            ra.w = 0.0
        a8, b8 = res_to8(ra), res_to8(rb)
        s0 = int(rng.integers(0, 2 ** 32))
        seed = C.c_uint32(s0)
        rg = res_from8(a8)
        L.ref_combineReservoirs_geom(C.byref(rg), C.byref(res_from8(b8)), C.byref(ginfo_from16(ga, L)), C.byref(ginfo_from16(gb, L)), C.byref(seed))
        seed2 = C.c_uint32(s0)
        rp = res_from8(a8)
        ph = L.ref_evaluatePHat(rb.lightIndex, 0, C.byref(ginfo_from16(ga, L)))
        L.ref_combineReservoirs_plain(C.byref(rp), C.byref(res_from8(b8)), ph, C.byref(seed2))
        V["combine"].append({"ga": bits(ga), "gb": bits(gb), "a": a8, "b": b8, "seed": s0, "geom": res_to8(rg), "geom_seed": seed.value,
                             "plain_phat": bits([ph])[0], "plain": res_to8(rp), "plain_seed": seed2.value})

    # ---- final shade (restir_post.frag:78-102)
    V["post"] = []
    for _ in range(128):
        g16 = random_ginfo(rng, L)
        if rng.uniform() < 0.2:
            g16[12] = 0.9                              # emissive override branch
        r = RefRes()
        sa = C.c_uint32(int(rng.integers(0, 2 ** 32)))
        L.ref_initial_ris(C.byref(ginfo_from16(g16, L)), 32, C.byref(sa), C.byref(r))
        if rng.uniform() < 0.3:
            r.w = r.w * 50.0                           # firefly branch
        thr = 2.0
        L.ref_post_shade(C.byref(r), C.byref(ginfo_from16(g16, L)), thr, p(out3))
        old = rng.uniform(0, 1, 3).astype(np.float32)
        acc = np.zeros(3, np.float32)
        frame = int(rng.integers(0, 40))
        L.ref_post_accumulate(p(old), p(out3), frame, 0, p(acc))
        V["post"].append({"g": bits(g16), "res": res_to8(r), "thr": bits([thr])[0], "color": bits(out3), "old": bits(old), "frame": frame,
                          "accum": bits(acc)})

    # ---- camera (nvmath)
    V["camera"] = []
    m = np.zeros(16, np.float32)
    for _ in range(16):
        eye = rng.uniform(-8, 8, 3).astype(np.float32); ctr = rng.uniform(-1, 1, 3).astype(np.float32)
        up = np.array([0, 1, 0], np.float32)
        fov, aspect = float(np.float32(rng.uniform(20, 90))), float(np.float32(rng.choice([16 / 9, 1.0, 4 / 3])))
        rec = {"eye": bits(eye), "center": bits(ctr), "up": bits(up), "fov": bits([fov])[0], "aspect": bits([aspect])[0]}
        L.ref_look_at(p(eye), p(ctr), p(up), p(m)); view = m.copy(); rec["view"] = bits(view)
        L.ref_perspectiveVK(fov, aspect, 0.1, 1000.0, p(m)); proj = m.copy(); rec["proj"] = bits(proj)
        L.ref_matmul(p(proj), p(view), p(m)); rec["projview"] = bits(m)
        L.ref_invert(p(view), p(m)); rec["view_inv"] = bits(m)
        L.ref_invert(p(proj), p(m)); rec["proj_inv"] = bits(m)
        V["camera"].append(rec)

    # ---- voxel material and struct layout
    V["voxel_albedo"] = []
    o4 = np.zeros(4, np.float32)
    for v in [0.0, 1e-5, 0.0015, 0.05, 0.2861328125, 1.0, 5.71484375] + rng.uniform(0, 6, 25).tolist():
        L.ref_voxel_albedo(float(np.float32(v)), p(o4))
        V["voxel_albedo"].append(bits([v]) + bits(o4))
    lay = (C.c_int * 16)()
    L.ref_struct_layout(lay)
    V["struct_layout"] = list(lay)

    with open(OUT, "w") as f:
        json.dump(V, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
