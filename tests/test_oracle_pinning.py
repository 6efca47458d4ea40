"""not gpu: pins the oracle (oracle/liboracle.so) to the reference.

1. against tests/golden/ref_vectors.json — outputs of the reference's OWN sources (tools/gen_golden.py), bit-exact;
2. against oracle/_ref/libvrs_ref.so live on fresh random inputs, when that library exists (it is built only where
   /root/reference is mounted; the committed vectors cover every other box);
3. the known-answer vectors of SURVEY.md §8c.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "ref_vectors.json")))


def f32(bits_list):
    return np.array(bits_list, np.uint32).view(np.float32)


def b(x):
    return int(np.array([x], np.float32).view(np.uint32)[0])


def p(a):
    return a.ctypes.data_as(C.c_void_p)


class Scene:
    """Oracle-side scene holding only lights + alias table (enough for the reference-derived math)."""

    def __init__(self, O, lights):
        self.lights = np.ascontiguousarray(lights, np.float32).reshape(-1, 8)
        self.table = O.create_alias_table(self.lights[:, 7].copy())
        s = O.Scene()
        s.lights, s.nlights = self.lights.ctypes.data, len(self.lights)
        s.table, s.ntable = self.table.ctypes.data, len(self.table)
        self.c = s


@pytest.fixture(scope="module")
def scene(O):
    return Scene(O, f32(GOLD["scene_lights"]))


def test_survey_known_answers(O):
    L = O.lib()
    o2 = (C.c_uint32 * 2)()
    for (x, y), (ex, ey, es) in {(0, 0): (417608103, 90043601, 507651704), (51, 69): (2248375744, 2988263557, 941672005),
                                 (115140, 64740): (3891299610, 3313245199, 2909577513), (3839, 2159): (2062298494, 492618734, 2554917228)}.items():
        L.orc_pcg2d(x, y, o2)
        assert (o2[0], o2[1], (o2[0] + o2[1]) & 0xFFFFFFFF) == (ex, ey, es)
    for s0, a, bb, rbits, s3 in [(507651704, 14071671, 962154, 0x3cc15820, 67504833), (941672005, 13992672, 4970687, 0x3e981e24, 2035027730),
                                 (2909577513, 2783348, 15288643, 0x3c577180, 4228079046)]:
        s = C.c_uint32(s0)
        assert L.orc_lcg(C.byref(s)) == a and L.orc_lcg(C.byref(s)) == bb
        assert b(L.orc_rnd(C.byref(s))) == rbits and s.value == s3
    # per-pixel seed = pcg2d(pixel * K).x + .y with K = clock*8 + pass + 1 (pixel (17,23), K=3 -> (51,69))
    assert L.orc_pixel_seed(17, 23, 0, 2) == 941672005
    t = O.create_alias_table([1, 2, 3, 4])
    assert t["alias"].tolist() == [2, 3, 3, 3]
    assert np.allclose(t["prob"], [0.4, 0.8, 0.6, 1.0], atol=1e-6) and np.allclose(t["aliasPdf"], [0.3, 0.4, 0.4, 0.4], atol=1e-6)
    t = O.create_alias_table([1, 1, 1, 0.27])
    assert t["alias"].tolist() == [1, 2, 2, 0] and abs(t["prob"][3] - 0.33027524) < 1e-7


def test_golden_rng_luminance(O):
    L = O.lib()
    o2 = (C.c_uint32 * 2)()
    for x, y, ex, ey in GOLD["pcg2d"]:
        L.orc_pcg2d(x, y, o2)
        assert (o2[0], o2[1]) == (ex, ey)
    for s0, a, bb, rb, s3 in GOLD["lcg_rnd"]:
        s = C.c_uint32(s0)
        assert L.orc_lcg(C.byref(s)) == a and L.orc_lcg(C.byref(s)) == bb and b(L.orc_rnd(C.byref(s))) == rb and s.value == s3
    for rec in GOLD["luminance"]:
        c = f32(rec[:3])
        assert b(L.orc_luminance_common(*map(float, c))) == rec[3] and b(L.orc_luminance_utils(*map(float, c))) == rec[4]


def test_golden_brdf(O):
    L = O.lib()
    out = np.zeros(3, np.float32)
    for rec in GOLD["brdf"]:
        x = f32(rec["in"])
        alb = x[4:7].copy()
        assert b(L.orc_disney_brdf_luminance(*map(float, x[0:4]), float(x[7]), float(x[8]), float(x[9]))) == rec["lum"]
        L.orc_disney_brdf_color(*map(float, x[0:4]), p(alb), float(x[8]), float(x[9]), p(out))
        assert out.view(np.uint32).tolist() == rec["color"]


def test_golden_alias_tables_and_lights(O):
    L = O.lib()
    for rec in GOLD["alias_tables"]:
        t = O.create_alias_table(f32(rec["pdf"]))
        assert t["alias"].tolist() == rec["alias"]
        for k, name in (("prob", "prob"), ("pdf_out", "pdf"), ("aliasPdf", "aliasPdf")):
            assert t[name].view(np.uint32).tolist() == rec[k], name
    for key in ("generate_point_lights", "generate_point_lights_white"):
        g = GOLD[key]
        got = O.generate_point_lights(g["min"], g["max"], g["white"], g["n"])
        assert got.ravel().view(np.uint32).tolist() == g["out"]
    lights = f32(GOLD["scene_lights"]).reshape(-1, 8)
    table = O.create_alias_table(lights[:, 7].copy())
    idx, pr = C.c_uint32(), C.c_float()
    for r1b, r2b, ei, epb in GOLD["alias_sample"]:
        r1, r2 = f32([r1b, r2b])
        L.orc_alias_table_sample(p(table), len(table), float(r1), float(r2), C.byref(idx), C.byref(pr))
        assert idx.value == ei and b(pr.value) == epb


def test_golden_phat(O, scene):
    L = O.lib()
    out = np.zeros(3, np.float32)
    for rec in GOLD["phat"]:
        g = f32(rec["g"]).copy()
        assert b(L.orc_evaluate_phat(p(scene.lights), rec["light"], p(g))) == rec["phat"]
        L.orc_evaluate_phat_full(p(scene.lights), rec["light"], p(g), p(out))
        assert out.view(np.uint32).tolist() == rec["full"]


def test_golden_initial_ris(O, scene):
    L = O.lib()
    for rec in GOLD["initial_ris"]:
        g = f32(rec["g"]).copy()
        seed = C.c_uint32(rec["seed"])
        r8 = np.zeros(8, np.uint32)
        L.orc_initial_ris(C.byref(scene.c), p(g), rec["count"], C.byref(seed), p(r8))
        assert r8.tolist() == rec["res"], "reservoir after the RIS loop (M, lightIndex, kind, sampleSeed, pHat, sumW, w)"
        assert seed.value == rec["seed_out"], "RNG stream position"


def test_golden_combine(O, scene):
    L = O.lib()
    for rec in GOLD["combine"]:
        ga, gb = f32(rec["ga"]).copy(), f32(rec["gb"]).copy()
        a = np.array(rec["a"], np.uint32); other = np.array(rec["b"], np.uint32)
        seed = C.c_uint32(rec["seed"])
        L.orc_combine_geom(C.byref(scene.c), p(a), p(other), p(ga), p(gb), C.byref(seed))
        assert a.tolist() == rec["geom"] and seed.value == rec["geom_seed"]
        a = np.array(rec["a"], np.uint32)
        seed = C.c_uint32(rec["seed"])
        L.orc_combine_plain(p(a), p(other), float(f32([rec["plain_phat"]])[0]), C.byref(seed))
        assert a.tolist() == rec["plain"] and seed.value == rec["plain_seed"]


def test_golden_post_shade(O, scene):
    L = O.lib()
    out = np.zeros(3, np.float32)
    for rec in GOLD["post"]:
        g = f32(rec["g"]).copy()
        r8 = np.array(rec["res"], np.uint32)
        L.orc_post_shade(C.byref(scene.c), p(r8), p(g), float(f32([rec["thr"]])[0]), p(out))
        assert out.view(np.uint32).tolist() == rec["color"]
        # running mean of restir_post.frag:94-102 = mix(old, new, 1/frame)
        old = f32(rec["old"]); frame = rec["frame"]
        if frame < 1:
            exp = out
        else:
            w = np.float32(1.0) / np.float32(frame)
            exp = old * (np.float32(1.0) - w) + out * w
        assert exp.astype(np.float32).view(np.uint32).tolist() == rec["accum"]


def test_golden_camera_and_material(O):
    L = O.lib()
    for rec in GOLD["camera"]:
        eye, ctr, up = f32(rec["eye"]), f32(rec["center"]), f32(rec["up"])
        fov, aspect = float(f32([rec["fov"]])[0]), float(f32([rec["aspect"]])[0])
        view = O.look_at(eye, ctr, up); proj = O.perspectiveVK(fov, aspect, 0.1, 1000.0)
        assert view.view(np.uint32).tolist() == rec["view"] and proj.view(np.uint32).tolist() == rec["proj"]
        assert O.matmul(proj, view).view(np.uint32).tolist() == rec["projview"]
        assert O.invert(view).view(np.uint32).tolist() == rec["view_inv"] and O.invert(proj).view(np.uint32).tolist() == rec["proj_inv"]
    o4 = np.zeros(4, np.float32)
    for rec in GOLD["voxel_albedo"]:
        L.orc_voxel_albedo(float(f32([rec[0]])[0]), p(o4))
        assert o4.view(np.uint32).tolist() == rec[1:]
    assert GOLD["struct_layout"][:9] == [320, 20, 40, 48, 64, 128, 192, 256, 264]
    assert C.sizeof(O.RestirUniforms) == 320 and O.RestirUniforms.flags.offset == 256 and O.RestirUniforms.screenSize.offset == 40


def test_live_reference_fuzz(O, scene):
    """Fresh random inputs through the reference's own sources (skipped where oracle/_ref was never built)."""
    R = O.ref()
    if R is None:
        pytest.skip("oracle/_ref/libvrs_ref.so not built on this box (needs /root/reference)")
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_golden", os.path.join(os.path.dirname(HERE), "tools", "gen_golden.py"))
    gg = importlib.util.module_from_spec(spec); spec.loader.exec_module(gg)
    L = O.lib()
    rng = np.random.default_rng(99)
    R.ref_set_scene(p(scene.lights), len(scene.lights), None, p(scene.table), len(scene.table))
    for _ in range(300):
        g16 = gg.random_ginfo(rng, R)
        g = gg.ginfo_from16(g16, R)
        li = int(rng.integers(0, len(scene.lights)))
        assert b(L.orc_evaluate_phat(p(scene.lights), li, p(g16))) == b(R.ref_evaluatePHat(li, 0, C.byref(g)))
        s0 = int(rng.integers(0, 2 ** 32))
        sa, sb = C.c_uint32(s0), C.c_uint32(s0)
        r8 = np.zeros(8, np.uint32); rr = gg.RefRes()
        L.orc_initial_ris(C.byref(scene.c), p(g16), 32, C.byref(sa), p(r8))
        R.ref_initial_ris(C.byref(g), 32, C.byref(sb), C.byref(rr))
        assert r8.tolist() == gg.res_to8(rr) and sa.value == sb.value


def test_neglog1m_accuracy(O):
    """-ln(1-u) from + - * / only, for every kind of u = k / 2^24 the RNG can produce."""
    L = O.lib()
    ks = np.unique(np.concatenate([np.arange(0, 4096), np.arange(2 ** 24 - 4096, 2 ** 24), np.random.default_rng(3).integers(0, 2 ** 24, 20000)]))
    u = ks.astype(np.float32) / np.float32(16777216.0)
    got = np.array([L.orc_neglog1m(float(x)) for x in u], np.float64)
    exp = -np.log1p(-u.astype(np.float64))
    assert got[0] == 0.0
    rel = np.abs(got[1:] - exp[1:]) / exp[1:]
    assert rel.max() < 4e-7, rel.max()


def test_oracle_frames_match_committed_digests(O):
    """The oracle's full-frame outputs (G-buffer, reservoirs, RNG / voxel / cell trace, accumulated image) for four small
    configurations must equal the SHA-256 digests committed in tests/golden/oracle_frame_digests.json
    (tools/gen_frame_digests.py): a change of the volumetric specification is a deliberate regeneration, never a drift."""
    import json
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import gen_frame_digests as G
    want = json.load(open(G.OUT))
    got = G.compute()
    assert set(got) == set(want)
    for case in want:
        for frame in want[case]:
            for key, val in want[case][frame].items():
                assert got[case][frame][key] == val, (case, frame, key)
        assert want[case]["frame0"]["hit_pixels"] > 100          # the volume is in view: the digests cover real work
