"""TEST INFRASTRUCTURE ONLY — minimal OpenVDB file WRITER used to generate fixtures that exercise the reader paths the two
shipped assets do not (ZIP-compressed buffers, fp32 values, node-mask metadata codes 2..6, root / internal tiles, negative
coordinates, several grids per file, file versions 222-224).  Follows the published layout restated in oracle/vdb_py.py;
the stream it produces is defined by the READ semantics of io::readCompressedValues.
"""
import struct
import zlib

import numpy as np

COMPRESS_ZIP, COMPRESS_ACTIVE_MASK, COMPRESS_BLOSC = 1, 2, 4


def _s(x):
    b = x.encode()
    return struct.pack("<I", len(b)) + b


def _meta(d):
    out = struct.pack("<I", len(d))
    for k, (typ, val) in d.items():
        if typ == "string":
            raw = val.encode()
        elif typ == "int64":
            raw = struct.pack("<q", val)
        elif typ == "vec3i":
            raw = struct.pack("<3i", *val)
        elif typ == "bool":
            raw = b"\x01" if val else b"\x00"
        elif typ == "float":
            raw = struct.pack("<f", val)
        else:
            raise ValueError(typ)
        out += _s(k) + _s(typ) + struct.pack("<I", len(raw)) + raw
    return out


def _mask_bytes(bits):
    return np.packbits(np.asarray(bits, bool), bitorder="little").tobytes()


BLOSC_MODE = {"mode": "openvdb"}     # how _blosc_chunk lays the next chunks out (tests switch it)


def _blosc_chunk(raw, mode):
    """A Blosc 1.x chunk of `raw`.  Modes: "openvdb" = what io::bloscToStream produces (LZ4, byte shuffle, typesize 4,
    blocksize = the buffer rounded down to the typesize, split streams); "nosplit" (flag 0x10); "blocks" (several
    2 KB blocks + a leftover block); "zlib" (codec 3); "memcpy" (stored chunk).  LZ4 streams come from pyarrow's codec."""
    import pyarrow as pa
    typesize, n = 4, len(raw)
    if mode == "memcpy":
        return struct.pack("<BBBBIII", 2, 1, 0x1 | 0x2, typesize, n, n, n + 16) + raw
    codec = 3 if mode == "zlib" else 1
    flags = 0x1 | (codec << 5) | (0x10 if mode == "nosplit" else 0)
    blocksize = 2048 if mode == "blocks" else max(n // typesize * typesize, typesize)
    nblocks = (n + blocksize - 1) // blocksize
    body, starts = b"", []
    for j in range(nblocks):
        blk = raw[j * blocksize:(j + 1) * blocksize]
        bsize, leftover = len(blk), len(blk) != blocksize
        nelem = bsize // typesize
        shuf = np.frombuffer(blk[:nelem * typesize], np.uint8).reshape(nelem, typesize).T.tobytes() + blk[nelem * typesize:]
        nsplits = typesize if (mode != "nosplit" and bsize // typesize >= 128 and not leftover) else 1
        neblock = bsize // nsplits
        starts.append(16 + 4 * nblocks + len(body))
        for k in range(nsplits):
            part = shuf[k * neblock:(k + 1) * neblock]
            z = zlib.compress(part, 6) if codec == 3 else pa.compress(part, codec="lz4_raw", asbytes=True)
            if len(z) >= len(part):
                z = part                                   # stored split stream: cbytes == its raw size
            body += struct.pack("<I", len(z)) + z
    head = struct.pack("<BBBBIII", 2, 1, flags, typesize, n, blocksize, 16 + 4 * nblocks + len(body))
    return head + b"".join(struct.pack("<I", x) for x in starts) + body


def _block(raw, compression):
    if compression & COMPRESS_BLOSC:                       # io::bloscToStream
        if len(raw) <= 48 or BLOSC_MODE["mode"] == "raw":  # BLOSC_MINIMUM_BYTES: stored with a negative size
            return struct.pack("<q", -len(raw)) + raw
        c = _blosc_chunk(raw, BLOSC_MODE["mode"])
        return struct.pack("<q", len(c)) + c
    if compression & COMPRESS_ZIP:
        z = zlib.compress(raw, 6)
        if len(z) < len(raw):
            return struct.pack("<q", len(z)) + z
        return struct.pack("<q", -len(raw)) + raw
    return raw


def _compressed_values(values, value_mask, background, half, compression, skip=None):
    """io::writeCompressedValues: metadata byte, optional inactive values / selection mask, then the (active) values.
    `skip` marks child slots of internal nodes (their stored value is irrelevant)."""
    values = np.asarray(values, np.float32)
    value_mask = np.asarray(value_mask, bool)
    n = len(values)
    consider = ~value_mask if skip is None else (~value_mask & ~np.asarray(skip, bool))
    inact = np.unique(values[consider])
    bg, nbg = np.float32(background), np.float32(-background)
    mask_compressed = bool(compression & COMPRESS_ACTIVE_MASK)
    out = b""
    metadata, v0, v1, sel = 6, None, None, None
    if mask_compressed:
        if len(inact) == 0 or (len(inact) == 1 and inact[0] == bg):
            metadata = 0
        elif len(inact) == 1 and inact[0] == nbg and bg != nbg:
            metadata = 1
        elif len(inact) == 1:
            metadata, v0 = 2, inact[0]
        elif len(inact) == 2 and set(inact.tolist()) == {float(bg), float(nbg)}:
            metadata, sel = 3, (values == bg)
        elif len(inact) == 2 and bg in inact:
            other = inact[0] if inact[1] == bg else inact[1]
            metadata, v0, sel = 4, other, (values == bg)
        elif len(inact) == 2:
            metadata, v0, v1, sel = 5, inact[0], inact[1], (values == inact[1])
        else:
            metadata = 6
    out += struct.pack("<b", metadata)
    if metadata in (2, 4, 5):
        out += struct.pack("<f", float(v0))
        if metadata == 5:
            out += struct.pack("<f", float(v1))
    if metadata in (3, 4, 5):
        out += _mask_bytes(sel & ~value_mask)
    stored = values if (not mask_compressed or metadata == 6) else values[value_mask]
    raw = stored.astype("<f2" if half else "<f4").tobytes()
    return out + _block(raw, compression)


class Grid:
    """Sparse description: leaves {origin(3 ints, multiples of 8): (values[512] offset (x<<6)|(y<<3)|z, mask[512])},
    tiles [(origin, log2dim in {3,7,12}, value, active)]."""

    def __init__(self, name, background=0.0, half=False, compression=COMPRESS_ACTIVE_MASK, voxel_size=0.5, translation=None,
                 grid_class="fog volume", map_type=None, affine=None):
        """`map_type` overrides the transform written ("ScaleMap", "ScaleTranslateMap", "AffineMap", or any name for error
        tests); `affine` is the 4x4 row-major matrix of an AffineMap (row-vector convention: translation in the last row),
        built from voxel_size / translation when omitted; `voxel_size` may be a 3-tuple for the non-uniform Scale maps."""
        self.name, self.background, self.half, self.compression = name, np.float32(background), half, compression
        self.voxel_size, self.translation, self.grid_class = voxel_size, translation, grid_class
        self.map_type, self.affine = map_type, affine
        self.leaves, self.tiles = {}, []

    def set_leaf(self, origin, values, mask):
        assert all(o % 8 == 0 for o in origin)
        v = np.asarray(values, np.float32).reshape(512)
        if self.half:
            v = v.astype(np.float16).astype(np.float32)
        self.leaves[tuple(origin)] = (v, np.asarray(mask, bool).reshape(512))

    def add_tile(self, origin, log2dim, value, active):
        v = np.float32(value)
        if self.half:
            v = np.float32(np.float16(v))
        self.tiles.append((tuple(origin), log2dim, v, bool(active)))


def _grid_bytes(g):
    """-> (topology+header bytes builder) ; returns (pre, topo, buffers)"""
    pre = struct.pack("<I", g.compression)
    pre += _meta({"class": ("string", g.grid_class), "name": ("string", g.name)})
    vs3 = tuple(g.voxel_size) if isinstance(g.voxel_size, (tuple, list)) else (g.voxel_size,) * 3
    scale_vecs = [vs3, vs3, tuple(1 / v for v in vs3), tuple(1 / v ** 2 for v in vs3), tuple(0.5 / v for v in vs3)]
    if g.map_type == "AffineMap":                           # AffineMap::write: Mat4d, 16 doubles, row-major
        m = g.affine
        if m is None:
            t = g.translation or (0.0, 0.0, 0.0)
            m = [[vs3[0], 0, 0, 0], [0, vs3[1], 0, 0], [0, 0, vs3[2], 0], [t[0], t[1], t[2], 1.0]]
        pre += _s("AffineMap") + struct.pack("<16d", *[float(x) for row in m for x in row])
    else:
        if g.translation is None:
            pre += _s(g.map_type or "UniformScaleMap")
            vecs = scale_vecs
        else:
            pre += _s(g.map_type or "UniformScaleTranslateMap")
            vecs = [tuple(g.translation)] + scale_vecs
        for v in vecs:
            pre += struct.pack("<3d", *v)
    bg = g.background
    # ---- organise the tree
    roots = {}
    def root_key(o): return tuple(c & ~4095 for c in o)
    def n5_slot(o): return (((o[0] & 4095) >> 7) << 10) | (((o[1] & 4095) >> 7) << 5) | ((o[2] & 4095) >> 7)
    def n4_slot(o): return (((o[0] & 127) >> 3) << 8) | (((o[1] & 127) >> 3) << 4) | ((o[2] & 127) >> 3)
    root_tiles = []
    for (o, lg, v, a) in g.tiles:
        if lg == 12:
            root_tiles.append((o, v, a))
            continue
        r = roots.setdefault(root_key(o), {"tiles5": {}, "n4": {}})
        if lg == 7:
            r["tiles5"][n5_slot(o)] = (v, a)
        else:
            n4 = r["n4"].setdefault(n5_slot(o), {"tiles": {}, "leaves": {}})
            n4["tiles"][n4_slot(o)] = (v, a)
    for o in g.leaves:
        r = roots.setdefault(root_key(o), {"tiles5": {}, "n4": {}})
        n4 = r["n4"].setdefault(n5_slot(o), {"tiles": {}, "leaves": {}})
        n4["leaves"][n4_slot(o)] = o
    topo = struct.pack("<i", 1) + struct.pack("<f", float(bg)) + struct.pack("<II", len(root_tiles), len(roots))
    for (o, v, a) in sorted(root_tiles):
        topo += struct.pack("<3i", *o) + struct.pack("<f", float(v)) + (b"\x01" if a else b"\x00")
    leaf_order = []
    for rk in sorted(roots):
        r = roots[rk]
        topo += struct.pack("<3i", *rk)
        child = np.zeros(32768, bool); vmask = np.zeros(32768, bool); vals = np.full(32768, bg, np.float32)
        for s in r["n4"]:
            child[s] = True
        for s, (v, a) in r["tiles5"].items():
            assert not child[s]
            vals[s] = v; vmask[s] = a
        topo += _mask_bytes(child) + _mask_bytes(vmask) + _compressed_values(vals, vmask, bg, g.half, g.compression, skip=child)
        for s in sorted(r["n4"]):
            n4 = r["n4"][s]
            child4 = np.zeros(4096, bool); vmask4 = np.zeros(4096, bool); vals4 = np.full(4096, bg, np.float32)
            for t in n4["leaves"]:
                child4[t] = True
            for t, (v, a) in n4["tiles"].items():
                assert not child4[t]
                vals4[t] = v; vmask4[t] = a
            topo += _mask_bytes(child4) + _mask_bytes(vmask4) + _compressed_values(vals4, vmask4, bg, g.half, g.compression, skip=child4)
            for t in sorted(n4["leaves"]):
                o = n4["leaves"][t]
                topo += _mask_bytes(g.leaves[o][1])
                leaf_order.append(o)
    buffers = b""
    for o in leaf_order:
        vals, mask = g.leaves[o]
        buffers += _mask_bytes(mask) + _compressed_values(vals, mask, bg, g.half, g.compression)
    return pre, topo, buffers


def write_vdb(path, grids, version=224):
    head = struct.pack("<q", 0x56444220) + struct.pack("<I", version) + struct.pack("<II", 8, 1) + b"\x01"
    head += b"00000000-0000-0000-0000-000000000000"
    head += _meta({"creator": ("string", "oracle/vdb_write.py")})
    head += struct.pack("<I", len(grids))
    parts = [_grid_bytes(g) for g in grids]
    pos = len(head)
    out = head
    for g, (pre, topo, buf) in zip(grids, parts):
        desc = _s(g.name) + _s("Tree_float_5_4_3" + ("_HalfFloat" if g.half else "")) + _s("")
        grid_pos = pos + len(desc) + 24
        block_pos = grid_pos + len(pre) + len(topo)
        end_pos = block_pos + len(buf)
        out += desc + struct.pack("<qqq", grid_pos, block_pos, end_pos) + pre + topo + buf
        pos = end_pos
    with open(path, "wb") as f:
        f.write(out)
