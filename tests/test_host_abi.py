"""not gpu: the C-ABI library loads and exports every symbol include/vrs.h declares; host-side helpers match the
reference (golden vectors + live reference where available); the VDB reader reproduces the facts measured from the
shipped assets; error behaviour without a device."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import common

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = json.load(open(os.path.join(HERE, "golden", "ref_vectors.json")))
REF_ASSETS = "/root/reference/assets"


def f32(bits_list):
    return np.array(bits_list, np.uint32).view(np.float32)


def test_header_symbols_exported(V):
    hdr = open(os.path.join(ROOT, "include", "vrs.h")).read()
    declared = sorted(set(re.findall(r"\b(vrs_[a-zA-Z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 35
    L = V.lib()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(V.EXPORTS) == declared
    assert L.vrs_abi_version() == 2


def test_struct_layouts_match_reference(V):
    lay = GOLD["struct_layout"]   # sizeof / offsetof measured on the reference's host_device.h
    U = V.RestirUniforms
    assert C.sizeof(U) == lay[0] == 320
    assert (U.spatialNeighbors.offset, U.screenSize.offset, U.currCamPos.offset, U.currFrameProjectionViewMatrix.offset, U.prevCamPos.offset,
            U.prevFrameProjectionViewMatrix.offset, U.flags.offset, U.gamma.offset) == tuple(lay[1:9])
    assert U.initialLightSampleCount.offset == lay[14] and U.temporalSampleCountMultiplier.offset == lay[15]
    assert C.sizeof(V.GlobalUniforms) == lay[9] and C.sizeof(V.PointLight) == lay[10] and C.sizeof(V.AliasTableCell) == lay[12]
    assert C.sizeof(V.PushConstantRestir) == lay[13]


def test_default_uniforms_follow_renderer(V):
    u = V.RestirUniforms()
    V.lib().vrs_default_restir_uniforms(C.byref(u), 1280, 720)     # Renderer.cpp:2341-2358
    assert (u.flags, u.spatialNeighbors, u.spatialRadius, u.initialLightSampleCount) == (7, 4, 30.0, 64)
    assert (u.environmentalPower, u.fireflyClampThreshold, u.temporalSampleCountMultiplier, u.gamma, u.debugMode) == (1.0, 2.0, 20, 4.0, 0)
    assert tuple(u.screenSize) == (1280, 720)
    c = V.Config()
    V.lib().vrs_default_config(C.byref(c), 64, 32)
    assert abs(c.world_scale - 0.05) < 1e-9 and [round(x, 3) for x in c.world_translate] == [-2.5, 0.5, 0.0]   # Renderer.cpp:1420-1423
    assert abs(c.roughness - 0.9) < 1e-7 and abs(c.metallic - 0.0001) < 1e-10                                  # Renderer.cpp:1498-1500


def test_alias_table_and_lights_match_reference(V, O):
    for rec in GOLD["alias_tables"]:
        t = V.create_alias_table(f32(rec["pdf"]))
        assert t["alias"].tolist() == rec["alias"]
        assert t["prob"].view(np.uint32).tolist() == rec["prob"] and t["pdf"].view(np.uint32).tolist() == rec["pdf_out"]
        assert t["aliasPdf"].view(np.uint32).tolist() == rec["aliasPdf"]
    for key in ("generate_point_lights", "generate_point_lights_white"):
        g = GOLD[key]
        got = V.generate_point_lights(g["min"], g["max"], bool(g["white"]), g["n"])
        assert got.ravel().view(np.uint32).tolist() == g["out"]
    rng = np.random.default_rng(5)
    for n in (1, 2, 17, 1000, 10000, 100000):           # 100 000 = configs[4]
        pdf = rng.uniform(0, 3, n).astype(np.float32)
        a, b = V.create_alias_table(pdf), O.create_alias_table(pdf)
        assert a.tobytes() == b.tobytes()
        assert (a["prob"] <= 1.0).all() and (a["alias"] >= 0).all() and (a["alias"] < n).all()
        # the table must reproduce the pdf: P(i) = (prob_i + sum_{j: alias_j = i} (1 - prob_j)) / n
        mass = a["prob"].astype(np.float64).copy()
        np.add.at(mass, a["alias"], 1.0 - a["prob"].astype(np.float64))
        assert np.allclose(mass / n, pdf / pdf.sum(), atol=2e-6)


def test_configs4_light_set_and_alias_table_match_the_reference_sources(V, O):
    """BASELINE configs[4]: 100 000 lights.  The light set bench.py builds for it and its alias table, product vs oracle bit for
    bit, and — when the reference's own restir_utils.cpp is compiled here (oracle/_ref) — against that too."""
    import ctypes as C
    lo, hi = [-3.2, 0.4, -1.9], [2.7, 6.1, 2.2]
    n = 100000
    a, b = V.generate_point_lights(lo, hi, False, n), O.generate_point_lights(lo, hi, False, n)
    assert a.shape == (n, 8) and a.tobytes() == b.tobytes()
    ta, tb = V.create_alias_table(a[:, 7]), O.create_alias_table(a[:, 7])
    assert ta.tobytes() == tb.tobytes()
    R = O.ref()
    if R is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    names = [s for s in ("ref_createAliasTable", "ref_create_alias_table") if hasattr(R, s)]
    if not names:
        pytest.skip("reference alias-table entry point not exported by this _ref build")
    out = np.zeros(n, dtype=ta.dtype)
    pdf = np.ascontiguousarray(a[:, 7])
    getattr(R, names[0])(pdf.ctypes.data_as(C.c_void_p), C.c_int(n), out.ctypes.data_as(C.c_void_p))
    assert out.tobytes() == ta.tobytes()


def test_camera_matches_reference(V):
    for rec in GOLD["camera"]:
        eye, ctr, up = f32(rec["eye"]), f32(rec["center"]), f32(rec["up"])
        fov, aspect = float(f32([rec["fov"]])[0]), float(f32([rec["aspect"]])[0])
        view, proj = V.look_at(eye, ctr, up), V.perspectiveVK(fov, aspect, 0.1, 1000.0)
        assert view.view(np.uint32).tolist() == rec["view"] and proj.view(np.uint32).tolist() == rec["proj"]
        assert V.mat4_mul(proj, view).view(np.uint32).tolist() == rec["projview"]
        assert V.invert(view).view(np.uint32).tolist() == rec["view_inv"] and V.invert(proj).view(np.uint32).tolist() == rec["proj_inv"]


def test_no_device_fails_loudly(V):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    with pytest.raises(V.VrsError) as e:
        V.Renderer(64, 64)
    assert e.value.status == 7 and "no CPU fallback" in str(e.value)     # VRS_ERR_NO_DEVICE


def test_convert_errors(V, tmp_path):
    with pytest.raises(V.VrsError) as e:
        V.convert_vdb("/nonexistent/file.vdb", str(tmp_path / "x.vrsg"))
    assert e.value.status == 3                                            # VRS_ERR_IO
    bad = tmp_path / "bad.vdb"
    bad.write_bytes(b"not a vdb file at all" * 10)
    with pytest.raises(V.VrsError) as e:
        V.convert_vdb(str(bad), str(tmp_path / "x.vrsg"))
    assert e.value.status == 4                                            # VRS_ERR_FORMAT
    trunc = tmp_path / "trunc.vrsg"
    trunc.write_bytes(open(common.asset("cube"), "rb").read()[:5000])
    import grid_py
    with pytest.raises(Exception):
        grid_py.read_vrsg(str(trunc))


SMOKE_KAT = [((56, 112, 56), 0.05035400390625, True), ((60, 40, 60), 2.845703125, True), ((50, 100, 50), 0.2861328125, True),
             ((64, 150, 64), 0.2166748046875, True), ((62, 31, 59), 5.71484375, True), ((0, 0, 0), 0.0, False), ((1, 2, 1), 0.0, False),
             ((111, 223, 112), 0.0, False), ((200, 200, 200), 0.0, False)]
CUBE_KAT = [((0, 0, 0), -0.1500244140625, False), ((-1, -1, -1), -0.1500244140625, False), ((100, 0, 0), -0.1500244140625, False),
            ((97, 3, -5), -0.1500244140625, False), ((110, 110, 110), -1.1920928955078125e-07, True), ((-112, -112, -112), 0.1500244140625, False),
            ((500, 0, 0), 0.1500244140625, False), ((108, 0, 0), -0.0999755859375, True), ((112, 0, 0), 0.0999755859375, True)]


@pytest.mark.parametrize("name,kat,leaves,active,bbox", [
    ("smoke", SMOKE_KAT, 3117, 1049275, ((1, 2, 1), (111, 223, 112))),
    ("cube", CUBE_KAT, 6812, 1452218, ((-112, -112, -112), (112, 112, 112)))])
def test_snapshot_matches_appendix_a(name, kat, leaves, active, bbox):
    """assets/*.vrsg were produced by the product's C++ .vdb reader (vrs_convert_vdb); read them back with the oracle's
    independent numpy reader and check the facts SURVEY.md Appendix A measured from the files."""
    import grid_py
    g = grid_py.read_vrsg(common.asset(name))
    assert len(g.leaf_origin) == leaves and int(g.leaf_mask.sum()) == active
    lo, hi = grid_py.active_bbox(g)
    assert (tuple(int(v) for v in lo), tuple(int(v) for v in hi)) == bbox
    raw, vmin, vdim = grid_py.dense_raw(g)
    for (i, j, k), val, _act in kat:
        x, y, z = i - vmin[0], j - vmin[1], k - vmin[2]
        inside = 0 <= x < vdim[0] and 0 <= y < vdim[1] and 0 <= z < vdim[2]
        got = raw[z, y, x] if inside else np.float32(g.background)
        assert float(got) == val, (name, (i, j, k))
    s = sum(float(v[m].astype(np.float64).sum()) for v, m in zip(g.leaf_value, g.leaf_mask))
    assert abs(s - {"smoke": 171499.69552135468, "cube": 2864.861377477646}[name]) < 1e-6


@pytest.mark.parametrize("name", ["smoke", "cube"])
def test_cpp_vdb_reader_against_numpy_reader(V, tmp_path, name):
    """Two independent readers of the same .vdb (C++ in the product, numpy in oracle/) must agree on every voxel."""
    path = os.path.join(REF_ASSETS, name + ".vdb")
    if not os.path.exists(path):
        pytest.skip("reference assets not mounted on this box")
    import grid_py
    import vdb_py
    out = str(tmp_path / (name + ".vrsg"))
    V.convert_vdb(path, out)
    assert open(out, "rb").read() == open(common.asset(name), "rb").read(), "committed snapshot is stale"
    a, vmin_a, vdim_a = grid_py.dense_raw(grid_py.read_vrsg(out))
    g = vdb_py.read_vdb(path)
    b, vmin_b, vdim_b = g.dense_raw()
    assert vmin_a == vmin_b and vdim_a == vdim_b and a.tobytes() == b.tobytes()
    assert g.active_voxel_count() == g.meta["file_voxel_count"]
    assert g.topology_end == g.block_pos and g.buffers_end == g.end_pos


def test_band_split(V):
    for h, n in ((1080, 8), (2160, 8), (1080, 3), (7, 4)):
        bands = [V.band_for_rank(h, r, n) for r in range(n)]
        assert bands[0][0] == 0 and bands[-1][1] == h
        assert all(bands[i][1] == bands[i + 1][0] for i in range(n - 1))


def test_bench_row_split_balances_cost_and_keeps_halo():
    """bench.py's band splitter: equal shares of the per-row cost, every band at least as tall as the halo."""
    import os
    import sys
    import numpy as np
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    H = 1080
    y = np.arange(H)
    cost = 1.0 + 50.0 * np.exp(-((y - 600.0) / 120.0) ** 2)          # a bump of expensive rows
    for world in (2, 4, 8):
        bands = bench.split_rows(cost.copy(), world)
        assert bands[0][0] == 0 and bands[-1][1] == H and all(b[1] == n[0] for b, n in zip(bands, bands[1:]))
        assert all(b[1] - b[0] >= 32 for b in bands)
        shares = np.array([cost[b[0]:b[1]].sum() for b in bands])
        assert shares.max() / shares.mean() < 1.15
    flat = bench.split_rows(np.zeros(H) + 1e-9, 8)                    # degenerate cost: still a valid partition
    assert flat[0][0] == 0 and flat[-1][1] == H and all(b[1] - b[0] >= 32 for b in flat)
    # several camera positions: the bump moves; the split minimises the per-position maximum, never worse than the average split
    costs = np.stack([1.0 + 50.0 * np.exp(-((y - c) / 120.0) ** 2) for c in (450.0, 600.0, 750.0)])
    for world in (2, 4, 8):
        avg = bench.split_rows(costs.sum(0), world, 48)
        opt = bench.split_rows(costs, world, 48)
        worst = lambda bands: sum(max(c[b[0]:b[1]].sum() for b in bands) for c in costs)
        assert opt[0][0] == 0 and opt[-1][1] == H and all(b[1] == n[0] for b, n in zip(opt, opt[1:])) and all(b[1] - b[0] >= 48 for b in opt)
        assert worst(opt) <= worst(avg) * (1.0 + 1e-9)


def test_bench_issue_roofline_helper():
    """bench.py's extra "issue_roofline" object: warp instructions per frame (committed ncu launch list of the workload) x
    live frames/s over 148 SMs x 4 schedulers x SM clock; None whenever its inputs do not apply, never an exception."""
    import bench
    tr = bench.load_traffic(bench.DEFAULT_WORKLOAD)
    assert tr and "k_ris" in tr
    r = bench.issue_roofline(tr, 450.0, {"sm_mhz": 1965.0}, 1)
    want = sum(k["warp_inst_M"] * k.get("launches_per_frame", 1) for k in tr.values()) * 1e6 * 450.0 / (148 * 4 * 1965.0e6)
    assert r is not None and abs(r["frac"] - want) < 1e-3 and 0.0 < r["frac"] < 1.0
    assert bench.issue_roofline(tr, 450.0, None, 1) is None
    assert bench.issue_roofline({}, 450.0, {"sm_mhz": 1965.0}, 1) is None
    assert bench.issue_roofline(tr, 450.0, {"sm_mhz": 1965.0}, 8) is None
    assert bench.issue_roofline({"k": {}}, 450.0, {"sm_mhz": 1965.0}, 1) is None
    assert bench.load_traffic("no_such_workload") == {}


def test_bench_halo_rows_cover_the_orbit(V):
    """bench.py sizes the band halo from the camera orbit: the fixed 32 rows of round 1 are too few for 6 deg/frame."""
    import bench
    lo, hi = [-4.0, 0.5, -2.0], [1.0, 6.0, 3.0]
    ctr = [(a + b) * 0.5 for a, b in zip(lo, hi)]
    diag = float(np.sqrt(sum(((b - a) * 0.5) ** 2 for a, b in zip(lo, hi))))
    h1080 = bench.temporal_halo_rows(V, dict(W=1920, H=1080), lo, hi, ctr, diag)
    h4k = bench.temporal_halo_rows(V, dict(W=3840, H=2160), lo, hi, ctr, diag)
    assert h1080 % 8 == 0 and 32 < h1080 < 200 and h1080 < h4k < 400
    assert abs(h4k - 2 * h1080) <= 16                                    # reprojection distance scales with the resolution


def test_bench_config_is_identical_for_both_arms():
    import bench
    wl = dict(bench.WORKLOADS[bench.DEFAULT_WORKLOAD])
    assert bench.DEFAULT_WORKLOAD == "bunny_4k_full" and wl["W"] == 3840 and wl["iters"] == 2 and wl["flags"] == 7
    c = bench.config_of(wl, bench.DEFAULT_WORKLOAD)
    assert "partition" not in c and c["name"] == "bunny_4k_full" and c["resolution"] == [3840, 2160]
