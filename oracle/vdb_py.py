"""TEST INFRASTRUCTURE ONLY — independent numpy reader for OpenVDB `.vdb` files.

The reference loads grids through OpenVDB >= 8 (README.md:36; CMakeLists.txt:61
`find_package(OpenVDB REQUIRED)`, not vendored, no pinned commit) at the call sites
src/vdb/vdb.cpp:103-208 (io::File::open / getGrids), :747-786 (cbeginValueOn,
ValueAccessor::getValue, indexToWorld) and :1179-1200 (evalActiveVoxelBoundingBox).
OpenVDB is absent from this image, so this module restates the published on-disk
format (file versions 222-224, Tree_float_5_4_3[_HalfFloat], node-mask compression,
optional ZIP blocks; Blosc is not decodable here) and is pinned by the facts measured
from the shipped assets (active-voxel counts == the files' own
`file_voxel_count`, bbox == `file_bbox_min/max`, getValue known answers).

It is deliberately a different implementation from the product's C++ reader
(volume-restir-vulkan_b200/csrc/vdb_reader.cpp) so that one checks the other.
"""
import struct
import zlib

import numpy as np

LOG2 = (5, 4, 3)


class _Stream:
    def __init__(self, data):
        self.d = data
        self.p = 0

    def read(self, n):
        b = self.d[self.p:self.p + n]
        if len(b) != n:
            raise EOFError("vdb: truncated file")
        self.p += n
        return b

    def u32(self):
        return struct.unpack("<I", self.read(4))[0]

    def i32(self):
        return struct.unpack("<i", self.read(4))[0]

    def i64(self):
        return struct.unpack("<q", self.read(8))[0]

    def f32(self):
        return struct.unpack("<f", self.read(4))[0]

    def string(self):
        return self.read(self.u32()).decode("latin-1")

    def vec3d(self):
        return struct.unpack("<3d", self.read(24))


def _read_metamap(s):
    out = {}
    for _ in range(s.u32()):
        name = s.string()
        typ = s.string()
        raw = s.read(s.u32())
        if typ == "string":
            val = raw.decode("latin-1")
        elif typ == "int32":
            val = struct.unpack("<i", raw)[0]
        elif typ == "int64":
            val = struct.unpack("<q", raw)[0]
        elif typ == "float":
            val = struct.unpack("<f", raw)[0]
        elif typ == "double":
            val = struct.unpack("<d", raw)[0]
        elif typ == "bool":
            val = raw != b"\x00"
        elif typ == "vec3i":
            val = struct.unpack("<3i", raw)
        elif typ == "vec3s":
            val = struct.unpack("<3f", raw)
        elif typ == "vec3d":
            val = struct.unpack("<3d", raw)
        else:
            val = raw
        out[name] = val
    return out


COMPRESS_ZIP, COMPRESS_ACTIVE_MASK, COMPRESS_BLOSC = 1, 2, 4


def _blosc_decode(chunk, nbytes_expected):
    """Blosc 1.x chunk (io::bloscFromStream -> blosc_decompress_ctx), restated: 16-byte header, block starts, per block
    `typesize` split streams (or one), byte unshuffle.  LZ4 streams go through pyarrow's lz4_raw codec, so this reader shares
    no decompression code with the product's C++ one."""
    import pyarrow as pa
    flags, typesize = chunk[2], chunk[3] or 1
    nbytes, blocksize, cbytes = struct.unpack_from("<III", chunk, 4)
    if nbytes != nbytes_expected or cbytes > len(chunk):
        raise ValueError("vdb: Blosc chunk header inconsistent with the buffer")
    if flags & 0x2:
        return bytes(chunk[16:16 + nbytes])
    codec = flags >> 5
    if flags & 0x4 or codec not in (1, 3):
        raise NotImplementedError("vdb: Blosc chunk with bit shuffle or a codec other than LZ4 / zlib")
    shuffle, dont_split = bool(flags & 1) and typesize > 1, bool(flags & 0x10)
    nblocks = (nbytes + blocksize - 1) // blocksize if nbytes else 0
    out = bytearray()
    for j in range(nblocks):
        bsize = min(blocksize, nbytes - j * blocksize)
        leftover = bsize != blocksize
        nsplits = typesize if (not dont_split and typesize <= 16 and bsize // typesize >= 128 and not leftover) else 1
        neblock = bsize // nsplits
        p = struct.unpack_from("<I", chunk, 16 + 4 * j)[0]
        block = bytearray()
        for _ in range(nsplits):
            cb = struct.unpack_from("<I", chunk, p)[0]
            p += 4
            part = bytes(chunk[p:p + cb])
            p += cb
            if cb == neblock:
                block += part
            elif codec == 1:
                block += pa.decompress(part, decompressed_size=neblock, codec="lz4_raw", asbytes=True)
            else:
                block += zlib.decompress(part)
        if shuffle:
            nelem = bsize // typesize
            planes = np.frombuffer(bytes(block[:nelem * typesize]), np.uint8).reshape(typesize, nelem)
            block = bytearray(planes.T.tobytes()) + block[nelem * typesize:]
        out += block
    return bytes(out)


def _read_raw_block(s, nbytes, compression):
    if compression & COMPRESS_BLOSC:
        zipped = s.i64()
        if zipped <= 0:
            return s.read(-zipped)
        return _blosc_decode(s.read(zipped), nbytes)
    if compression & COMPRESS_ZIP:
        zipped = s.i64()
        if zipped <= 0:
            return s.read(-zipped)
        return zlib.decompress(s.read(zipped))
    return s.read(nbytes)


def _mask_bits(raw):
    return np.unpackbits(np.frombuffer(raw, dtype=np.uint8), bitorder="little").astype(bool)


def _read_compressed_values(s, count, value_mask, half, compression, background, version):
    """io::readCompressedValues (openvdb/io/Compression.h), restated."""
    metadata = 6
    if version >= 222:
        metadata = struct.unpack("<b", s.read(1))[0]
    inactive1 = np.float32(background)
    inactive0 = np.float32(background) if metadata == 0 else np.float32(-background)
    if metadata in (2, 4, 5):
        inactive0 = np.float32(s.f32())
        if metadata == 5:
            inactive1 = np.float32(s.f32())
    selection = None
    if metadata in (3, 4, 5):
        selection = _mask_bits(s.read(count // 8))
    mask_compressed = bool(compression & COMPRESS_ACTIVE_MASK)
    n_stored = count
    if mask_compressed and metadata != 6 and version >= 222:
        n_stored = int(value_mask.sum())
    item = 2 if half else 4
    raw = _read_raw_block(s, n_stored * item, compression)
    vals = np.frombuffer(raw, dtype="<f2" if half else "<f4", count=n_stored).astype(np.float32)
    if n_stored == count:
        return vals.copy(), metadata
    out = np.empty(count, dtype=np.float32)
    if selection is None:
        out[:] = inactive0
    else:
        out[:] = np.where(selection, inactive1, inactive0)
    out[value_mask] = vals
    return out, metadata


class VdbGrid:
    """One float grid: leaves / tiles exactly as stored, plus derived dense views."""

    def __init__(self):
        self.name = ""
        self.grid_type = ""
        self.meta = {}
        self.compression = 0
        self.voxel_size = 1.0
        self.translation = (0.0, 0.0, 0.0)
        self.background = 0.0
        self.half = False
        self.leaf_origins = []      # (x, y, z)
        self.leaf_masks = []        # 512 bool each, offset (x<<6)|(y<<3)|z
        self.leaf_values = []       # 512 float32 each (inactive voxels filled per metadata rule)
        self.leaf_metadata = []
        self.tiles = []             # (origin xyz, log2dim of extent, value, active)
        self.node_counts = [0, 0, 0, 0]   # root children, internal5, internal4, leaves

    # ---- derived -----------------------------------------------------
    def active_voxel_count(self):
        n = sum(int(m.sum()) for m in self.leaf_masks)
        for (_o, lg, _v, active) in self.tiles:
            if active:
                n += (1 << lg) ** 3
        return n

    def active_bbox(self):
        lo = np.array([2 ** 31 - 1] * 3, dtype=np.int64)
        hi = -lo
        for o, m in zip(self.leaf_origins, self.leaf_masks):
            if not m.any():
                continue
            idx = np.nonzero(m)[0]
            x, y, z = idx >> 6, (idx >> 3) & 7, idx & 7
            lo = np.minimum(lo, [o[0] + x.min(), o[1] + y.min(), o[2] + z.min()])
            hi = np.maximum(hi, [o[0] + x.max(), o[1] + y.max(), o[2] + z.max()])
        for (o, lg, _v, active) in self.tiles:
            if active:
                lo = np.minimum(lo, o)
                hi = np.maximum(hi, np.array(o) + (1 << lg) - 1)
        return tuple(int(v) for v in lo), tuple(int(v) for v in hi)

    def window(self):
        """Leaf-aligned voxel window covering the active bbox: (vmin[3], vdim[3])."""
        lo, hi = self.active_bbox()
        vmin = [(v >> 3) << 3 for v in lo]
        vmax = [((v >> 3) + 1) << 3 for v in hi]
        return vmin, [b - a for a, b in zip(vmin, vmax)]

    def dense_raw(self):
        """Raw values over window(): array [z][y][x] float32 (+ vmin, vdim)."""
        vmin, vdim = self.window()
        cd = [d // 8 for d in vdim]
        cell = np.full((cd[2], cd[1], cd[0]), np.float32(self.background), dtype=np.float32)
        for (o, lg, v, _a) in self.tiles:
            size = 1 << lg
            c0 = [(o[a] - vmin[a]) // 8 for a in range(3)]
            c1 = [c0[a] + max(size // 8, 1) for a in range(3)]
            c0 = [max(c, 0) for c in c0]
            c1 = [min(c1[a], cd[a]) for a in range(3)]
            if all(c1[a] > c0[a] for a in range(3)):
                cell[c0[2]:c1[2], c0[1]:c1[1], c0[0]:c1[0]] = np.float32(v)
        dense = np.repeat(np.repeat(np.repeat(cell, 8, axis=0), 8, axis=1), 8, axis=2)
        for o, vals in zip(self.leaf_origins, self.leaf_values):
            x0, y0, z0 = (o[a] - vmin[a] for a in range(3))
            if x0 < 0 or y0 < 0 or z0 < 0 or x0 >= vdim[0] or y0 >= vdim[1] or z0 >= vdim[2]:
                continue
            blk = vals.reshape(8, 8, 8)            # [x][y][z]
            dense[z0:z0 + 8, y0:y0 + 8, x0:x0 + 8] = blk.transpose(2, 1, 0)
        return np.ascontiguousarray(dense), vmin, vdim

    def get_value(self, i, j, k):
        """ValueAccessor::getValue semantics -> (value, active, level) with level in
        {'leaf','internal4','internal5','background'}."""
        for o, m, v in zip(self.leaf_origins, self.leaf_masks, self.leaf_values):
            if o[0] <= i < o[0] + 8 and o[1] <= j < o[1] + 8 and o[2] <= k < o[2] + 8:
                off = ((i & 7) << 6) | ((j & 7) << 3) | (k & 7)
                return float(v[off]), bool(m[off]), "leaf"
        for (o, lg, v, a) in self.tiles:
            s = 1 << lg
            if o[0] <= i < o[0] + s and o[1] <= j < o[1] + s and o[2] <= k < o[2] + s:
                return float(v), bool(a), {3: "internal4", 7: "internal5", 12: "root"}[lg]
        return float(self.background), False, "background"


def _read_internal(s, g, origin, level, version):
    """InternalNode<...>::readTopology; level 0 = 32^3 children of 128^3, level 1 = 16^3 children of 8^3."""
    log2dim = LOG2[level]
    n = 1 << (3 * log2dim)
    child_mask = _mask_bits(s.read(n // 8))
    value_mask = _mask_bits(s.read(n // 8))
    vals, _md = _read_compressed_values(s, n, value_mask, g.half, g.compression, g.background, version)
    child_log2 = sum(LOG2[level + 1:])           # voxels per child = 2^child_log2
    g.node_counts[1 + level] += 1
    pos = np.arange(n)
    xs = (pos >> (2 * log2dim)) << child_log2
    ys = ((pos >> log2dim) & ((1 << log2dim) - 1)) << child_log2
    zs = (pos & ((1 << log2dim) - 1)) << child_log2
    tile_idx = np.nonzero(~child_mask & ((vals != np.float32(g.background)) | value_mask))[0]
    for t in tile_idx:
        g.tiles.append(((origin[0] + int(xs[t]), origin[1] + int(ys[t]), origin[2] + int(zs[t])), child_log2,
                        float(vals[t]), bool(value_mask[t])))
    for c in np.nonzero(child_mask)[0]:
        co = (origin[0] + int(xs[c]), origin[1] + int(ys[c]), origin[2] + int(zs[c]))
        if level == 0:
            _read_internal(s, g, co, 1, version)
        else:
            g.leaf_origins.append(co)
            g.leaf_masks.append(_mask_bits(s.read(64)))
            g.node_counts[3] += 1


def read_vdb(path, grid_name=None):
    """Parse `path`; return the requested (or first) float grid as a VdbGrid."""
    with open(path, "rb") as f:
        s = _Stream(f.read())
    magic = s.i64()
    if magic != 0x56444220:
        raise ValueError("vdb: bad magic")
    version = s.u32()
    if version < 222:
        raise NotImplementedError("vdb: file version %d < 222" % version)
    s.u32(); s.u32()                     # library major / minor
    has_offsets = s.read(1) != b"\x00"
    s.read(36)                           # uuid
    file_meta = _read_metamap(s)
    ngrids = s.u32()
    grids = []
    for _ in range(ngrids):
        name = s.string()
        gtype = s.string()
        _parent = s.string()
        gpos, bpos, epos = s.i64(), s.i64(), s.i64()
        grids.append((name, gtype, gpos, bpos, epos))
        if not has_offsets:
            raise NotImplementedError("vdb: files without grid offsets")
        s.p = epos
    for (name, gtype, gpos, bpos, epos) in grids:
        base = name.split("\x1e")[0]
        if grid_name is not None and base != grid_name:
            continue
        if not gtype.startswith("Tree_float_5_4_3"):
            if grid_name is None:
                continue
            raise NotImplementedError("vdb: grid type %s" % gtype)
        g = VdbGrid()
        g.name, g.grid_type = base, gtype
        g.half = gtype.endswith("_HalfFloat")
        s.p = gpos
        g.compression = s.u32()
        g.meta = _read_metamap(s)
        g.meta["__file__"] = file_meta
        g.meta["__version__"] = version
        map_type = s.string()
        if map_type == "UniformScaleMap" or map_type == "ScaleMap":
            v = [s.vec3d() for _ in range(5)]
            scale = v[0]
            g.translation = (0.0, 0.0, 0.0)
        elif map_type == "UniformScaleTranslateMap" or map_type == "ScaleTranslateMap":
            v = [s.vec3d() for _ in range(6)]
            g.translation = v[0]
            scale = v[1]
        elif map_type == "AffineMap":                       # Mat4d, row-major, row-vector convention (translation = last row)
            m = np.array(struct.unpack("<16d", s.read(128))).reshape(4, 4)
            scale = (m[0, 0], m[1, 1], m[2, 2])
            off = m[:3, :3] - np.diag(scale)
            if np.abs(off).max() > 1e-9 * abs(scale[0]) or np.abs(m[:3, 3]).max() > 0 or abs(m[3, 3] - 1.0) > 1e-12:
                raise NotImplementedError("vdb: AffineMap with rotation / shear / projection")
            g.translation = tuple(m[3, :3])
        else:
            raise NotImplementedError("vdb: map type %s" % map_type)
        if not (scale[0] > 0 and abs(scale[1] - scale[0]) <= 1e-9 * scale[0] and abs(scale[2] - scale[0]) <= 1e-9 * scale[0]):
            raise NotImplementedError("vdb: non-uniform voxel size %r" % (scale,))
        g.voxel_size = scale[0]
        # ---- topology (Tree::readTopology / RootNode::readTopology)
        _buffer_count = s.i32()
        g.background = struct.unpack("<f", s.read(4))[0]
        ntiles, nchildren = s.u32(), s.u32()
        for _ in range(ntiles):
            o = struct.unpack("<3i", s.read(12))
            val = s.f32()
            active = s.read(1) != b"\x00"
            g.tiles.append((o, 12, val, active))
        for _ in range(nchildren):
            o = struct.unpack("<3i", s.read(12))
            g.node_counts[0] += 1
            _read_internal(s, g, o, 0, version)
        g.topology_end = s.p
        g.block_pos = bpos
        # ---- buffers (Tree::readBuffers -> LeafNode::readBuffers), same traversal order
        s.p = bpos
        for li in range(len(g.leaf_origins)):
            mask = _mask_bits(s.read(64))
            vals, md = _read_compressed_values(s, 512, mask, g.half, g.compression, g.background, version)
            g.leaf_masks[li] = mask
            g.leaf_values.append(vals)
            g.leaf_metadata.append(md)
        g.buffers_end = s.p
        g.end_pos = epos
        return g
    raise KeyError("vdb: no float grid %r in %s" % (grid_name, path))


def density_from_raw(raw, grid_class_level_set, background):
    """fog -> max(v, 0); level set -> clamp(-v / background, 0, 1). fp32 ops."""
    raw = raw.astype(np.float32)
    if grid_class_level_set:
        bg = np.float32(background)
        return np.minimum(np.maximum(-raw / bg, np.float32(0)), np.float32(1)).astype(np.float32)
    return np.maximum(raw, np.float32(0)).astype(np.float32)
