"""not gpu: generated .vdb fixtures (oracle/vdb_write.py) drive the product's C++ reader (vrs_convert_vdb) and the oracle's
numpy reader through the on-disk variants the two shipped assets never use: ZIP blocks, fp32 values, node-mask metadata
codes 0..6, internal / root tiles, negative coordinates, several grids per file, file versions 222-224, and malformed input."""
import os

import numpy as np
import pytest

import common


def spec_value(g, i, j, k):
    o = (i & ~7, j & ~7, k & ~7)
    if o in g.leaves:
        v, m = g.leaves[o]
        off = ((i & 7) << 6) | ((j & 7) << 3) | (k & 7)
        return float(v[off])
    for (to, lg, tv, _a) in g.tiles:
        s = 1 << lg
        if to[0] <= i < to[0] + s and to[1] <= j < to[1] + s and to[2] <= k < to[2] + s:
            return float(tv)
    return float(g.background)


def build_cases():
    import vdb_write as W
    rng = np.random.default_rng(11)
    cases = []
    for half, comp, version in [(False, W.COMPRESS_ZIP | W.COMPRESS_ACTIVE_MASK, 224), (True, W.COMPRESS_ACTIVE_MASK, 222),
                                (False, 0, 223), (True, W.COMPRESS_ZIP, 224)]:
        g = W.Grid("density", background=0.0, half=half, compression=comp, voxel_size=0.25, translation=(1.0, -2.0, 3.5))
        for o in [(-16, 0, 8), (-8, 0, 8), (0, 0, 0), (8, 8, 8), (120, 120, 120), (128, 0, 0), (-4096, 8, 16), (4096, -8, 0)]:
            m = rng.uniform(size=512) < 0.4
            v = np.where(m, rng.uniform(0.1, 3.0, 512), 0.0)
            g.set_leaf(o, v, m)
        g.add_tile((16, 0, 8), 3, 0.75, True)         # active 8^3 tile
        g.add_tile((256, 128, 0), 7, 1.5, True)       # active 128^3 tile
        cases.append((g, version))
    # level set with +-background inactive values (codes 1 and 3), inactive interior tiles
    bg = np.float32(0.15)
    g = W.Grid("ls", background=bg, half=False, compression=W.COMPRESS_ACTIVE_MASK | W.COMPRESS_ZIP, voxel_size=0.05, grid_class="level set")
    idx = np.arange(512); x = idx >> 6
    m = (x >= 3) & (x <= 5)
    g.set_leaf((0, 0, 0), np.where(m, (x - 4) * 0.05, np.where(x < 3, -bg, bg)), m)              # code 3
    g.set_leaf((-8, 0, 0), np.where(m, -0.01, -bg), m)                                             # code 1
    g.set_leaf((8, 0, 0), np.where(m, 0.02, bg), m)                                                # code 0
    g.add_tile((-16, 0, 0), 3, -bg, False)                                                         # inactive interior tile
    cases.append((g, 224))
    # explicit inactive values: codes 2, 4, 5, 6
    g = W.Grid("odd", background=0.5, half=False, compression=W.COMPRESS_ACTIVE_MASK, voxel_size=1.0)
    m = rng.uniform(size=512) < 0.3
    act = rng.uniform(1, 2, 512)
    g.set_leaf((0, 0, 0), np.where(m, act, 7.25), m)                                               # code 2
    g.set_leaf((8, 0, 0), np.where(m, act, np.where(idx % 2 == 0, 0.5, 9.0)), m)                   # code 4
    g.set_leaf((16, 0, 0), np.where(m, act, np.where(idx % 3 == 0, 3.0, 4.0)), m)                  # code 5
    g.set_leaf((24, 0, 0), np.where(m, act, (idx % 5).astype(np.float32)), m)                      # code 6
    g.add_tile((0, 4096, 0), 12, 2.5, False)                                                       # inactive root tile with a value
    cases.append((g, 224))
    return cases


@pytest.mark.parametrize("case", range(6))
def test_both_readers_reproduce_the_written_grid(V, tmp_path, case):
    import grid_py
    import vdb_py
    import vdb_write as W
    g, version = build_cases()[case]
    path = str(tmp_path / "fixture.vdb")
    W.write_vdb(path, [g], version=version)
    # oracle reader
    pg = vdb_py.read_vdb(path)
    assert pg.topology_end == pg.block_pos and pg.buffers_end == pg.end_pos
    dense_py, vmin, vdim = pg.dense_raw()
    # product reader
    out = str(tmp_path / "fixture.vrsg")
    V.convert_vdb(path, out)
    sg = grid_py.read_vrsg(out)
    dense_cc, vmin2, vdim2 = grid_py.dense_raw(sg)
    assert vmin == vmin2 and vdim == vdim2
    assert dense_py.tobytes() == dense_cc.tobytes()
    assert abs(sg.voxel_size - g.voxel_size) < 1e-12 and bool(sg.level_set) == (g.grid_class == "level set")
    if g.translation is not None:
        assert np.allclose(sg.translation, g.translation)
    # ground truth at random points inside the window and at every leaf corner
    rng = np.random.default_rng(case)
    pts = [tuple(int(rng.integers(vmin[a], vmin[a] + vdim[a])) for a in range(3)) for _ in range(min(4000, int(np.prod(vdim))))]
    pts += [o for o in g.leaves] + [(o[0] + 7, o[1] + 7, o[2] + 7) for o in g.leaves]
    for (i, j, k) in pts:
        x, y, z = i - vmin[0], j - vmin[1], k - vmin[2]
        if 0 <= x < vdim[0] and 0 <= y < vdim[1] and 0 <= z < vdim[2]:
            assert float(dense_py[z, y, x]) == spec_value(g, i, j, k), (i, j, k)
    active = sum(int(m.sum()) for _v, m in g.leaves.values()) + sum((1 << lg) ** 3 for (_o, lg, _v, a) in g.tiles if a)
    assert pg.active_voxel_count() == active and int(sg.leaf_mask.sum()) == sum(int(m.sum()) for _v, m in g.leaves.values())


@pytest.mark.parametrize("mode", ["openvdb", "nosplit", "blocks", "zlib", "memcpy", "raw"])
@pytest.mark.parametrize("half", [False, True])
def test_blosc_compressed_buffers(V, tmp_path, mode, half):
    """COMPRESS_BLOSC files (what OpenVDB >= 3 writes by default, and what the missing bunny_cloud / explosion / fire assets
    use): the product's own chunk decoder (LZ4 + unshuffle + split streams, no libblosc) and the numpy reader (LZ4 through
    pyarrow) must both reproduce the written grid, for every chunk layout a Blosc 1.x writer can produce."""
    import grid_py
    import vdb_py
    import vdb_write as W
    rng = np.random.default_rng(5)
    g = W.Grid("density", background=0.0, half=half, compression=W.COMPRESS_BLOSC | W.COMPRESS_ACTIVE_MASK, voxel_size=0.1)
    for n, o in enumerate([(0, 0, 0), (8, 0, 0), (-8, 16, 24), (128, 0, 8), (0, 4096, 0)]):
        m = rng.uniform(size=512) < (0.02 if n == 1 else 0.9 if n == 0 else 0.5)      # 10 active voxels: below BLOSC_MINIMUM_BYTES
        smooth = 1.0 + 0.01 * np.arange(512)                                           # compressible byte planes
        v = np.where(m, smooth if n % 2 == 0 else rng.uniform(0.1, 3.0, 512), 0.0)
        g.set_leaf(o, v, m)
    g.add_tile((256, 128, 0), 7, 1.5, True)
    W.BLOSC_MODE["mode"] = mode
    try:
        path = str(tmp_path / "blosc.vdb")
        W.write_vdb(path, [g], version=224)
    finally:
        W.BLOSC_MODE["mode"] = "openvdb"
    pg = vdb_py.read_vdb(path)
    assert pg.topology_end == pg.block_pos and pg.buffers_end == pg.end_pos
    dense_py, vmin, vdim = pg.dense_raw()
    out = str(tmp_path / "blosc.vrsg")
    V.convert_vdb(path, out)
    dense_cc, vmin2, vdim2 = grid_py.dense_raw(grid_py.read_vrsg(out))
    assert vmin == vmin2 and vdim == vdim2 and dense_py.tobytes() == dense_cc.tobytes()
    for o, (v, m) in g.leaves.items():
        for off in (0, 77, 300, 511):
            i, j, k = o[0] + (off >> 6), o[1] + ((off >> 3) & 7), o[2] + (off & 7)
            want = np.float32(np.float16(v[off])) if half else np.float32(v[off])
            assert dense_py[k - vmin[2], j - vmin[1], i - vmin[0]] == want


def test_blosc_chunk_errors(V, tmp_path):
    """Chunks this reader cannot decode fail loudly (VRS_ERR_FORMAT), they are never read as garbage."""
    import vdb_write as W
    g = W.Grid("density", background=0.0, compression=W.COMPRESS_BLOSC | W.COMPRESS_ACTIVE_MASK)
    g.set_leaf((0, 0, 0), 1.0 + 0.01 * np.arange(512), np.ones(512, bool))
    path = str(tmp_path / "b.vdb")
    W.write_vdb(path, [g], version=224)
    data = bytearray(open(path, "rb").read())
    head = bytes([2, 1, 0x1 | (1 << 5), 4])                   # the chunk header written by _blosc_chunk("openvdb")
    at = data.find(head)
    assert at > 0
    for patch, what in [((at + 2, 0x1 | (4 << 5)), "Zstd codec"), ((at + 2, 0x4 | (1 << 5)), "bit shuffle"), ((at + 4, data[at + 4] ^ 1), "size mismatch")]:
        bad = bytearray(data)
        bad[patch[0]] = patch[1]
        p2 = str(tmp_path / "bad.vdb")
        open(p2, "wb").write(bad)
        with pytest.raises(Exception) as e:
            V.convert_vdb(p2, str(tmp_path / "bad.vrsg"))
        assert "Blosc" in str(e.value), what
    trunc = str(tmp_path / "trunc.vdb")
    open(trunc, "wb").write(data[:at + 40])
    with pytest.raises(Exception):
        V.convert_vdb(trunc, str(tmp_path / "trunc.vrsg"))


def test_transform_maps(V, tmp_path):
    """Every way OpenVDB serialises a uniform scale (+ translation): Uniform/ScaleMap, Uniform/ScaleTranslateMap and the
    general AffineMap (what Houdini exports); non-uniform scale, rotation and frustum maps fail loudly in both readers."""
    import grid_py
    import vdb_py
    import vdb_write as W

    def grid(**kw):
        g = W.Grid("density", background=0.0, **kw)
        g.set_leaf((8, -8, 0), np.linspace(0.5, 1.5, 512), np.ones(512, bool))
        return g

    ok = [dict(voxel_size=0.25), dict(voxel_size=0.25, map_type="ScaleMap"), dict(voxel_size=0.125, translation=(1.0, -2.0, 3.5)),
          dict(voxel_size=0.125, translation=(1.0, -2.0, 3.5), map_type="ScaleTranslateMap"),
          dict(voxel_size=0.2, translation=(-4.0, 0.5, 9.0), map_type="AffineMap"), dict(voxel_size=0.2, map_type="AffineMap")]
    for n, kw in enumerate(ok):
        path, out = str(tmp_path / ("m%d.vdb" % n)), str(tmp_path / ("m%d.vrsg" % n))
        W.write_vdb(path, [grid(**kw)])
        pg = vdb_py.read_vdb(path)
        V.convert_vdb(path, out)
        sg = grid_py.read_vrsg(out)
        want_t = kw.get("translation") or (0.0, 0.0, 0.0)
        for g_ in (pg, sg):
            assert abs(g_.voxel_size - kw["voxel_size"]) < 1e-12 and np.allclose(g_.translation, want_t), (kw, g_.voxel_size, g_.translation)
    rot = [[0.0, 0.2, 0, 0], [-0.2, 0.0, 0, 0], [0, 0, 0.2, 0], [0, 0, 0, 1.0]]
    bad = [dict(voxel_size=(0.25, 0.5, 0.25), map_type="ScaleMap"), dict(voxel_size=(0.25, 0.25, 0.3), translation=(0.0, 0.0, 1.0), map_type="ScaleTranslateMap"),
           dict(voxel_size=0.2, map_type="AffineMap", affine=rot), dict(voxel_size=0.2, map_type="NonlinearFrustumMap")]
    for n, kw in enumerate(bad):
        path = str(tmp_path / ("b%d.vdb" % n))
        W.write_vdb(path, [grid(**kw)])
        with pytest.raises(V.VrsError) as e:
            V.convert_vdb(path, str(tmp_path / "b.vrsg"))
        assert e.value.status in (4, 5) and ("not supported" in str(e.value)), str(e.value)
        with pytest.raises(NotImplementedError):
            vdb_py.read_vdb(path)


def test_grid_selection_by_name_and_errors(V, tmp_path):
    import grid_py
    import vdb_write as W
    a = W.Grid("density", background=0.0)
    a.set_leaf((0, 0, 0), np.full(512, 1.0), np.ones(512, bool))
    b = W.Grid("temperature", background=0.0)
    b.set_leaf((8, 0, 0), np.full(512, 300.0), np.ones(512, bool))
    path = str(tmp_path / "two.vdb")
    W.write_vdb(path, [a, b])
    V.convert_vdb(path, str(tmp_path / "first.vrsg"))                       # first float grid
    V.convert_vdb(path, str(tmp_path / "temp.vrsg"), grid_name="temperature")
    assert grid_py.read_vrsg(str(tmp_path / "first.vrsg")).leaf_origin.tolist() == [[0, 0, 0]]
    g2 = grid_py.read_vrsg(str(tmp_path / "temp.vrsg"))
    assert g2.leaf_origin.tolist() == [[8, 0, 0]] and float(g2.leaf_value[0, 0]) == 300.0
    with pytest.raises(V.VrsError) as e:
        V.convert_vdb(path, str(tmp_path / "x.vrsg"), grid_name="velocity")
    assert e.value.status == 4
    # truncated file, blosc flag, old file version
    raw = open(path, "rb").read()
    (tmp_path / "trunc.vdb").write_bytes(raw[:len(raw) // 2])
    with pytest.raises(V.VrsError) as e:
        V.convert_vdb(str(tmp_path / "trunc.vdb"), str(tmp_path / "x.vrsg"))
    assert e.value.status == 4
    c = W.Grid("density", compression=4 | 2)                                   # COMPRESS_BLOSC: decoded (test_blosc_compressed_buffers)
    c.set_leaf((0, 0, 0), np.full(512, 1.0), np.ones(512, bool))
    W.write_vdb(str(tmp_path / "blosc.vdb"), [c])
    V.convert_vdb(str(tmp_path / "blosc.vdb"), str(tmp_path / "x.vrsg"))
    W.write_vdb(str(tmp_path / "old.vdb"), [a], version=220)
    with pytest.raises(V.VrsError) as e:
        V.convert_vdb(str(tmp_path / "old.vdb"), str(tmp_path / "x.vrsg"))
    assert e.value.status == 4 and "222" in str(e.value)


def test_procedural_standins_are_deterministic(V, tmp_path):
    import grid_py
    a, b = str(tmp_path / "a.vrsg"), str(tmp_path / "b.vrsg")
    V.write_procedural_vrsg("torus_knot_helix", 64, a)
    V.write_procedural_vrsg("torus_knot_helix", 64, b)
    assert open(a, "rb").read() == open(b, "rb").read()
    for kind in ("bunny_cloud", "explosion", "fire", "fire_torus"):
        V.write_procedural_vrsg(kind, 64, a)
        g = grid_py.read_vrsg(a)
        assert len(g.leaf_origin) > 10 and 0.0 < float(g.leaf_value.max()) < 10.0 and not g.level_set
        assert (g.leaf_value[~g.leaf_mask] == 0).all()
    # the composite of configs[4] holds both of its parts: the knot's ring low in the grid and the plume rising above it
    dense, vmin, vdim = grid_py.dense_raw(grid_py.read_vrsg(a))                     # [z][y][x]
    occupied_rows = np.nonzero((dense > 0).any(axis=(0, 2)))[0] + vmin[1]
    assert occupied_rows.min() < 16 and occupied_rows.max() > 44
    with pytest.raises(V.VrsError):
        V.write_procedural_vrsg(0, 30, a)
    with pytest.raises(V.VrsError):
        V.write_procedural_vrsg(5, 64, a)


def test_emissive_voxel_lights_from_temperature_grid(V, tmp_path):
    """density + temperature in one file; lights come from temp = log(T) + 273.15 > 275 (vdb.cpp:811,
    Renderer.cpp:1615-1637), in tree order, at most 1001, emission (0.6, 0.2, 0.1)."""
    import vdb_write as W
    rng = np.random.default_rng(2)
    dens = W.Grid("density", background=0.0, voxel_size=0.5, translation=(1.0, 2.0, 3.0))
    temp = W.Grid("temperature", background=0.0, voxel_size=0.5, translation=(1.0, 2.0, 3.0))
    T = {}
    for o in [(0, 0, 0), (8, 0, 0), (-8, 16, 24)]:
        m = rng.uniform(size=512) < 0.5
        dens.set_leaf(o, np.where(m, 1.0, 0.0), m)
        t = np.where(m, rng.uniform(0.5, 20.0, 512), 0.0).astype(np.float32)
        temp.set_leaf(o, t, m)
        T[o] = (t, m)
    path = str(tmp_path / "fire.vdb")
    W.write_vdb(path, [dens, temp])
    lights = V.vdb_emissive_lights(path, "temperature", max_lights=1001)
    # expected set: every active voxel with log(T) + 273.15 > 275
    want = set()
    for o, (t, m) in T.items():
        for off in np.nonzero(m)[0]:
            if np.float32(np.float64(np.log(np.float32(t[off]))) + 273.15) > np.float32(275.0):
                ijk = (o[0] + (off >> 6), o[1] + ((off >> 3) & 7), o[2] + (off & 7))
                want.add(tuple(np.float32(np.float32(np.float64(np.float32(0.05)) * 0.5) * np.float32(c) + np.float32(np.float64(np.float32(0.05)) * tr + np.float64(np.float32(wt))))
                               for c, tr, wt in zip(ijk, (1.0, 2.0, 3.0), (-2.5, 0.5, 0.0))))
    got = set(tuple(np.float32(v) for v in l[:3]) for l in lights)
    assert len(lights) == min(len(want), 1001) and (got <= want)
    assert np.allclose(lights[:, 4:7], [0.6, 0.2, 0.1]) and np.allclose(lights[:, 3], 1.0)
    assert np.allclose(lights[:, 7], 0.2126 * 0.6 + 0.7152 * 0.2 + 0.0722 * 0.1)
    assert len(V.vdb_emissive_lights(path, "temperature", max_lights=5)) == 5
    assert len(V.vdb_emissive_lights(path, "density")) == 0               # log(1) + 273.15 < 275


def test_malformed_vrsg_snapshots_are_rejected(V, tmp_path):
    """the .vrsg reader must not trust indices or sizes from the file (they index host tables in finalize() and device
    tables in the kernels).  vrs_convert_vdb validates a snapshot host-side: corrupt child indices, absurd sizes and
    truncated payloads are VRS_ERR_FORMAT, never a crash."""
    import struct
    import zlib
    src = open(common.asset("cube"), "rb").read()
    assert src[:8] == b"VRSG0001"
    raw_size, zipped = struct.unpack_from("<QQ", src, 8)
    payload = bytearray(zlib.decompress(src[24:24 + zipped]))
    assert len(payload) == raw_size
    out = str(tmp_path / "out.vrsg")

    def write(name, body, claim=None):
        z = zlib.compress(bytes(body), 1)
        p = str(tmp_path / name)
        open(p, "wb").write(b"VRSG0001" + struct.pack("<QQ", len(body) if claim is None else claim, len(z)) + z)
        return p

    V.convert_vdb(write("good.vrsg", payload), out)                       # the untouched payload round-trips
    nroot, n5, n4, nleaf, ntile = struct.unpack_from("<5I", payload, 40)
    assert nroot >= 1 and n5 >= 1 and nleaf > 100
    bad = bytearray(payload)
    struct.pack_into("<i", bad, 60 + 12, n5 + 7)                           # root child -> a node that does not exist
    cases = {"root_child.vrsg": (bad, None)}
    bad = bytearray(payload)
    off = 60 + 16 * nroot                                                  # first sparse i5 table: count, then (slot, value) records
    (count,) = struct.unpack_from("<I", bad, off)
    assert count >= 1
    struct.pack_into("<i", bad, off + 4 + 4, n4 + 1000)                    # i5 child -> a node that does not exist
    cases["i5_child.vrsg"] = (bad, None)
    bad = bytearray(payload)
    struct.pack_into("<I", bad, 40 + 16, 0x7FFFFFFF)                       # ntile absurd
    cases["ntile.vrsg"] = (bad, None)
    cases["truncated.vrsg"] = (payload[:len(payload) // 2], None)
    cases["raw_size.vrsg"] = (payload, 1 << 40)                            # header claims a terabyte
    for name, (body, claim) in cases.items():
        with pytest.raises(V.VrsError) as e:
            V.convert_vdb(write(name, body, claim), out)
        assert e.value.status == 4, (name, e.value)                       # VRS_ERR_FORMAT
