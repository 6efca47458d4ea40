"""TEST INFRASTRUCTURE ONLY — independent numpy reader of the product's `.vrsg` snapshot format
(layout documented in volume-restir-vulkan_b200/csrc/vrs_grid.cpp) producing the oracle's dense window."""
import struct
import zlib

import numpy as np


class VrsgGrid:
    pass


def read_vrsg(path):
    with open(path, "rb") as f:
        data = f.read()
    if data[:8] != b"VRSG0001":
        raise ValueError("not a .vrsg file")
    raw_size, zipped = struct.unpack_from("<QQ", data, 8)
    b = zlib.decompress(data[24:24 + zipped])
    assert len(b) == raw_size
    p = 0

    def take(fmt):
        nonlocal p
        v = struct.unpack_from(fmt, b, p)
        p += struct.calcsize(fmt)
        return v

    g = VrsgGrid()
    (flags,) = take("<I")
    g.level_set, g.half = bool(flags & 1), bool(flags & 2)
    (g.background,) = take("<f")
    (g.voxel_size,) = take("<d")
    g.translation = take("<3d")
    nroot, n5, n4, nleaf, ntile = take("<5I")
    g.root = np.frombuffer(b, "<i4", 4 * nroot, p).reshape(-1, 4).copy(); p += 16 * nroot

    def sparse(n):
        nonlocal p
        (count,) = take("<I")
        arr = np.full(n, -1, np.int32)
        rec = np.frombuffer(b, [("slot", "<u4"), ("val", "<i4")], count, p); p += 8 * count
        arr[rec["slot"]] = rec["val"]
        return arr

    g.i5 = [sparse(32768) for _ in range(n5)]
    g.i4 = [sparse(4096) for _ in range(n4)]
    g.tile_value = np.frombuffer(b, "<f4", ntile, p).copy(); p += 4 * ntile
    g.tile_active = np.frombuffer(b, "u1", ntile, p).copy(); p += ntile
    g.leaf_origin = np.frombuffer(b, "<i4", 3 * nleaf, p).reshape(-1, 3).copy(); p += 12 * nleaf
    g.leaf_mask = np.unpackbits(np.frombuffer(b, "u1", 64 * nleaf, p), bitorder="little").reshape(nleaf, 512).astype(bool); p += 64 * nleaf
    if g.half:
        g.leaf_value = np.frombuffer(b, "<f2", 512 * nleaf, p).astype(np.float32).reshape(nleaf, 512); p += 1024 * nleaf
    else:
        g.leaf_value = np.frombuffer(b, "<f4", 512 * nleaf, p).reshape(nleaf, 512).copy(); p += 2048 * nleaf
    assert p == len(b)
    return g


def active_bbox(g):
    lo = np.array([2 ** 31 - 1] * 3, np.int64); hi = -lo
    idx = np.arange(512)
    off = np.stack([idx >> 6, (idx >> 3) & 7, idx & 7], 1)
    for o, m in zip(g.leaf_origin, g.leaf_mask):
        if m.any():
            q = o[None, :] + off[m]
            lo = np.minimum(lo, q.min(0)); hi = np.maximum(hi, q.max(0))
    # active tiles
    for r in g.root:
        n5 = r[3]
        if n5 < 0:
            if g.tile_active[~n5]:
                lo = np.minimum(lo, r[:3]); hi = np.maximum(hi, r[:3] + 4095)
            continue
        s5 = np.nonzero(g.i5[n5] < 0)[0]
        act = s5[g.tile_active[~g.i5[n5][s5]] != 0]
        for s in act:
            o = r[:3] + np.array([(s >> 10) << 7, ((s >> 5) & 31) << 7, (s & 31) << 7])
            lo = np.minimum(lo, o); hi = np.maximum(hi, o + 127)
        for s in np.nonzero(g.i5[n5] >= 0)[0]:
            o5 = r[:3] + np.array([(s >> 10) << 7, ((s >> 5) & 31) << 7, (s & 31) << 7])
            n4 = g.i5[n5][s]
            s4 = np.nonzero(g.i4[n4] < 0)[0]
            act4 = s4[g.tile_active[~g.i4[n4][s4]] != 0]
            for t in act4:
                o = o5 + np.array([(t >> 8) << 3, ((t >> 4) & 15) << 3, (t & 15) << 3])
                lo = np.minimum(lo, o); hi = np.maximum(hi, o + 7)
    return lo, hi


def dense_raw(g):
    """Raw values over the leaf-aligned active window -> (array [z][y][x], vmin, vdim)."""
    lo, hi = active_bbox(g)
    vmin = [(int(v) >> 3) << 3 for v in lo]
    vmax = [((int(v) >> 3) + 1) << 3 for v in hi]
    vdim = [b - a for a, b in zip(vmin, vmax)]
    cd = [d // 8 for d in vdim]
    cell = np.full((cd[2], cd[1], cd[0]), np.float32(g.background), np.float32)

    def fill(o, size, val):
        c0 = [(int(o[a]) - vmin[a]) // 8 for a in range(3)]
        c1 = [c0[a] + size // 8 for a in range(3)]
        c0 = [max(c, 0) for c in c0]; c1 = [min(c1[a], cd[a]) for a in range(3)]
        if all(c1[a] > c0[a] for a in range(3)):
            cell[c0[2]:c1[2], c0[1]:c1[1], c0[0]:c1[0]] = val

    for r in g.root:
        n5 = r[3]
        if n5 < 0:
            fill(r[:3], 4096, g.tile_value[~n5]); continue
        for s in range(32768):
            c = g.i5[n5][s]
            if c == -1:
                continue
            o5 = r[:3] + np.array([(s >> 10) << 7, ((s >> 5) & 31) << 7, (s & 31) << 7])
            if c < 0:
                fill(o5, 128, g.tile_value[~c]); continue
            slots = g.i4[c]
            for t in np.nonzero((slots < -1))[0]:
                o = o5 + np.array([(t >> 8) << 3, ((t >> 4) & 15) << 3, (t & 15) << 3])
                fill(o, 8, g.tile_value[~slots[t]])
    dense = np.repeat(np.repeat(np.repeat(cell, 8, 0), 8, 1), 8, 2)
    for o, vals in zip(g.leaf_origin, g.leaf_value):
        x0, y0, z0 = (int(o[a]) - vmin[a] for a in range(3))
        dense[z0:z0 + 8, y0:y0 + 8, x0:x0 + 8] = vals.reshape(8, 8, 8).transpose(2, 1, 0)
    return np.ascontiguousarray(dense), vmin, vdim
