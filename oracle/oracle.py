"""TEST INFRASTRUCTURE ONLY — ctypes/numpy front-end of oracle/liboracle.so (the scalar CPU
oracle) and, when present, oracle/_ref/libvrs_ref.so (the reference's own sources compiled
by oracle/ref/build_ref.py).  May be imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libvrs_ref.so")

FLAG_VISIBILITY, FLAG_TEMPORAL, FLAG_SPATIAL, FLAG_ENVIRONMENT = 1, 2, 4, 8
FLAG_FINAL_VISIBILITY, FLAG_FINALIZE_W = 16, 32


class AliasCell(C.Structure):
    _fields_ = [("alias", C.c_int32), ("prob", C.c_float), ("pdf", C.c_float), ("aliasPdf", C.c_float)]


class PointLight(C.Structure):
    _fields_ = [("pos", C.c_float * 4), ("emission_luminance", C.c_float * 4)]


class RestirUniforms(C.Structure):
    _fields_ = [
        ("pointLightCount", C.c_int32), ("triangleLightCount", C.c_int32), ("aliasTableCount", C.c_int32),
        ("environmentalPower", C.c_float), ("fireflyClampThreshold", C.c_float),
        ("spatialNeighbors", C.c_uint32), ("spatialRadius", C.c_float),
        ("initialLightSampleCount", C.c_uint32), ("temporalSampleCountMultiplier", C.c_int32),
        ("_pad0", C.c_uint32), ("screenSize", C.c_uint32 * 2), ("currCamPos", C.c_float * 4),
        ("currFrameProjectionViewMatrix", C.c_float * 16), ("prevCamPos", C.c_float * 4), ("_pad1", C.c_uint32 * 12),
        ("prevFrameProjectionViewMatrix", C.c_float * 16), ("flags", C.c_int32), ("debugMode", C.c_int32),
        ("gamma", C.c_float), ("_pad2", C.c_uint32 * 13),
    ]


class GlobalUniforms(C.Structure):
    _fields_ = [("viewProj", C.c_float * 16), ("viewInverse", C.c_float * 16), ("projInverse", C.c_float * 16)]


class PushConstant(C.Structure):
    _fields_ = [("clearColorRed", C.c_float), ("clearColorGreen", C.c_float), ("clearColorBlue", C.c_float),
                ("frame", C.c_int32), ("initialize", C.c_int32)]


class Scene(C.Structure):
    _fields_ = [
        ("dens", C.c_void_p), ("vmin", C.c_int32 * 3), ("vdim", C.c_int32 * 3), ("bg_density", C.c_float),
        ("A", C.c_float), ("invA", C.c_float), ("B", C.c_float * 3), ("density_scale", C.c_float),
        ("roughness", C.c_float), ("metallic", C.c_float),
        ("lights", C.c_void_p), ("nlights", C.c_int32), ("table", C.c_void_p), ("ntable", C.c_int32),
        ("cellmax", C.c_void_p),
    ]


class GBuf(C.Structure):
    _fields_ = [("worldPos", C.c_void_p), ("albedo", C.c_void_p), ("normal", C.c_void_p), ("matProps", C.c_void_p)]


class ResBuf(C.Structure):
    _fields_ = [("info", C.c_void_p), ("weight", C.c_void_p)]


def build(force=False):
    """Compile liboracle.so (and the reference checker when /root/reference exists)."""
    src = os.path.join(HERE, "vrs_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(HERE, "vrs_oracle.h"))):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/src/shaders") and (force or not os.path.exists(REF_PATH)):
        subprocess.check_call(["python3", os.path.join(HERE, "ref", "build_ref.py")], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        f = C.c_float
        L.orc_rnd.restype = f
        L.orc_luminance_common.restype = f
        L.orc_luminance_common.argtypes = [f, f, f]
        L.orc_luminance_utils.restype = f
        L.orc_luminance_utils.argtypes = [f, f, f]
        L.orc_disney_brdf_luminance.restype = f
        L.orc_disney_brdf_luminance.argtypes = [f] * 7
        L.orc_disney_brdf_color.argtypes = [f, f, f, f, C.c_void_p, f, f, C.c_void_p]
        L.orc_evaluate_phat.restype = f
        L.orc_evaluate_phat.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_evaluate_phat_full.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_alias_table_sample.argtypes = [C.c_void_p, C.c_int, f, f, C.c_void_p, C.c_void_p]
        L.orc_combine_plain.argtypes = [C.c_void_p, C.c_void_p, f, C.c_void_p]
        L.orc_post_shade.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, f, C.c_void_p]
        L.orc_perspectiveVK.argtypes = [f, f, f, f, C.c_void_p]
        L.orc_voxel_albedo.argtypes = [f, C.c_void_p]
        L.orc_neglog1m.restype = f
        L.orc_neglog1m.argtypes = [f]
        L.orc_density_at.restype = f
        L.orc_delta_track.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, f, f, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_ratio_track.restype = f
        L.orc_pixel_seed.restype = C.c_uint32
        _lib = L
    return _lib


_ref = None


def ref():
    """The reference's own sources as a library, or None when it was never built."""
    global _ref
    if _ref is None and os.path.exists(REF_PATH):
        L = C.CDLL(REF_PATH)
        f = C.c_float
        L.ref_rnd.restype = f
        L.ref_luminance_common.restype = f
        L.ref_luminance_common.argtypes = [f, f, f]
        L.ref_luminance_utils.restype = f
        L.ref_luminance_utils.argtypes = [f, f, f]
        L.ref_disneyBrdfLuminance.restype = f
        L.ref_disneyBrdfLuminance.argtypes = [f] * 7
        L.ref_disneyBrdfColor.argtypes = [f, f, f, f, C.c_void_p, f, f, C.c_void_p]
        L.ref_evaluatePHat.restype = f
        L.ref_evaluatePHat.argtypes = [C.c_uint32, C.c_int, C.c_void_p]
        L.ref_evaluatePHatFull.argtypes = [C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_aliasTableSample.argtypes = [f, f, C.c_void_p, C.c_void_p]
        L.ref_combineReservoirs_plain.argtypes = [C.c_void_p, C.c_void_p, f, C.c_void_p]
        L.ref_post_shade.argtypes = [C.c_void_p, C.c_void_p, f, C.c_void_p]
        L.ref_perspectiveVK.argtypes = [f, f, f, f, C.c_void_p]
        L.ref_voxel_albedo.argtypes = [f, C.c_void_p]
        _ref = L
    return _ref


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------ host helpers (oracle side)
def perspectiveVK(fovy, aspect, near, far):
    m = np.zeros(16, np.float32)
    lib().orc_perspectiveVK(fovy, aspect, near, far, _p(m))
    return m


def look_at(eye, center, up):
    m = np.zeros(16, np.float32)
    e, c, u = (np.asarray(v, np.float32) for v in (eye, center, up))
    lib().orc_look_at(_p(e), _p(c), _p(u), _p(m))
    return m


def invert(a):
    a = np.ascontiguousarray(a, np.float32)
    m = np.zeros(16, np.float32)
    lib().orc_invert(_p(a), _p(m))
    return m


def matmul(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    m = np.zeros(16, np.float32)
    lib().orc_matmul(_p(a), _p(b), _p(m))
    return m


def create_alias_table(pdf):
    pdf = np.ascontiguousarray(pdf, np.float32)
    out = np.zeros(len(pdf), dtype=[("alias", "<i4"), ("prob", "<f4"), ("pdf", "<f4"), ("aliasPdf", "<f4")])
    lib().orc_create_alias_table(_p(pdf), len(pdf), _p(out))
    return out


def generate_point_lights(mn, mx, white, n):
    mn, mx = np.asarray(mn, np.float32), np.asarray(mx, np.float32)
    out = np.zeros((n, 8), np.float32)
    lib().orc_generate_point_lights(_p(mn), _p(mx), int(white), n, _p(out))
    return out


class Camera:
    """Uniform producers of Renderer::updateUniformBuffer / updateRestirUniformBuffer
    (src/Renderer.cpp:116-161, 2376-2425) on the oracle side."""

    def __init__(self, eye, center, up=(0, 1, 0), fov=60.0, near=0.1, far=1000.0):
        self.eye, self.center, self.up, self.fov, self.near, self.far = eye, center, up, fov, near, far

    def matrices(self, width, height):
        aspect = np.float32(width) / np.float32(height)
        view = look_at(self.eye, self.center, self.up)
        proj = perspectiveVK(self.fov, float(aspect), self.near, self.far)
        return view, proj, matmul(proj, view), invert(view), invert(proj)


def global_uniforms(cam, W, H):
    view, proj, vp, vinv, pinv = cam.matrices(W, H)
    g = GlobalUniforms()
    g.viewProj[:] = vp.tolist()
    g.viewInverse[:] = vinv.tolist()
    g.projInverse[:] = pinv.tolist()
    return g


def restir_uniforms(cam, prev_cam, W, H, n_lights, M=32, flags=FLAG_VISIBILITY, k=5, radius=30.0,
                    firefly=2.0, temporal_mult=20):
    """Defaults of Renderer::createRestirUniformBuffer (src/Renderer.cpp:2341-2358) unless overridden."""
    u = RestirUniforms()
    u.pointLightCount, u.triangleLightCount, u.aliasTableCount = n_lights, 0, n_lights
    u.environmentalPower, u.fireflyClampThreshold = 1.0, firefly
    u.spatialNeighbors, u.spatialRadius = k, radius
    u.initialLightSampleCount, u.temporalSampleCountMultiplier = M, temporal_mult
    u.screenSize[0], u.screenSize[1] = W, H
    u.currCamPos[:] = [cam.eye[0], cam.eye[1], cam.eye[2], 0.0]
    u.currFrameProjectionViewMatrix[:] = cam.matrices(W, H)[2].tolist()
    pc = prev_cam if prev_cam is not None else cam
    u.prevCamPos[:] = [pc.eye[0], pc.eye[1], pc.eye[2], 0.0]
    u.prevFrameProjectionViewMatrix[:] = pc.matrices(W, H)[2].tolist()
    u.flags, u.debugMode, u.gamma = flags, 0, 4.0
    return u


class OracleScene:
    """Dense-window scene for the oracle. `dens` is [z][y][x] float32 densities."""

    def __init__(self, dens, vmin, voxel_size, translation, lights, bg_density=0.0, world_scale=0.05,
                 world_translate=(-2.5, 0.5, 0.0), density_scale=10.0, roughness=0.9, metallic=0.0001):
        self.dens = np.ascontiguousarray(dens, np.float32)
        vd = self.dens.shape
        assert all(d % 8 == 0 for d in vd) and all(v % 8 == 0 for v in vmin)
        self.lights = np.ascontiguousarray(lights, np.float32).reshape(-1, 8)
        self.table = create_alias_table(self.lights[:, 7])
        self.cellmax = np.zeros((vd[0] // 8, vd[1] // 8, vd[2] // 8), np.float32)
        s = Scene()
        s.dens = self.dens.ctypes.data
        s.vmin[:] = list(vmin)
        s.vdim[:] = [vd[2], vd[1], vd[0]]
        s.bg_density = bg_density
        # world = world_scale * (voxel_size * ijk + translation) + world_translate   (Renderer.cpp:1420-1435)
        f32 = lambda v: np.float64(np.float32(v))      # the product keeps these three as fp32 in vrs_config
        A = np.float32(f32(world_scale) * np.float64(voxel_size))
        s.A = A
        s.invA = np.float32(1.0) / A
        for a in range(3):
            s.B[a] = np.float32(f32(world_scale) * np.float64(translation[a]) + f32(world_translate[a]))
        s.density_scale, s.roughness, s.metallic = density_scale, roughness, metallic
        s.lights, s.nlights = self.lights.ctypes.data, len(self.lights)
        s.table, s.ntable = self.table.ctypes.data, len(self.table)
        s.cellmax = self.cellmax.ctypes.data
        self.c = s
        lib().orc_scene_prepare(C.byref(s))

    def world_bbox(self):
        """fp32 like vrs_get_grid_info (so that lights generated inside the box are the same bits on both sides)."""
        s, f = self.c, np.float32
        lo = [float(f(s.A) * (f(s.vmin[a]) - f(0.5)) + f(s.B[a])) for a in range(3)]
        hi = [float(f(s.A) * (f(s.vmin[a] + s.vdim[a]) - f(0.5)) + f(s.B[a])) for a in range(3)]
        return lo, hi


class Frame:
    """Per-pixel buffers in the reference layouts (planar RGBA32F)."""

    def __init__(self, W, H):
        self.W, self.H = W, H
        z = lambda: np.zeros((H, W, 4), np.float32)
        self.g = [dict(worldPos=z(), albedo=z(), normal=z(), matProps=z()) for _ in range(2)]
        self.res = [dict(info=z(), weight=z()) for _ in range(3)]
        self.accum = z()
        self.trace = np.zeros((H, W, 4), np.uint32)

    @staticmethod
    def gbuf(d):
        return GBuf(d["worldPos"].ctypes.data, d["albedo"].ctypes.data, d["normal"].ctypes.data, d["matProps"].ctypes.data)

    @staticmethod
    def rbuf(d):
        return ResBuf(d["info"].ctypes.data, d["weight"].ctypes.data)


class OracleRenderer:
    """Frame loop of src/main.cpp:301-449 on the oracle: initial(+visibility+temporal) -> spatial x iters -> shade,
    with the reference's ping-pong (Renderer.cpp:108-111, 1977-2040)."""

    def __init__(self, scene, W, H, spatial_iterations=2):
        self.scene, self.W, self.H = scene, W, H
        self.f = Frame(W, H)
        self.cur = 0
        self.iters = spatial_iterations
        self.final_res = 0   # index into f.res of the last frame's final reservoirs

    def render(self, gu, ru, pc, clock, y0=0, y1=None, exchange=None):
        """`exchange(planes, kind)` (multi-rank band mode) is called at the points where libvrs exchanges halo rows:
        kind "spatial" after the initial pass (G-buffer + reservoirs) and between spatial iterations (reservoirs) — only the
        ceil(spatialRadius) rows spatial reuse can reach travel; kind "temporal" after the frame (libvrs: at the start of the
        next one) with the previous G-buffer + final reservoirs over ALL halo rows, which the temporal reprojection may reach
        — the schedule of vrs_render_frame."""
        L, f, s = lib(), self.f, self.scene.c
        y1 = self.H if y1 is None else y1
        cur_g, prev_g = f.g[self.cur], f.g[1 - self.cur]
        prev_r = f.res[self.final_res]
        src = (self.final_res + 1) % 3
        L.orc_pass_initial(C.byref(s), C.byref(gu), C.byref(ru), C.c_uint32(clock), y0, y1, Frame.gbuf(cur_g),
                           Frame.gbuf(prev_g), Frame.rbuf(prev_r), Frame.rbuf(f.res[src]), _p(f.trace))
        spatial = bool(ru.flags & FLAG_SPATIAL) and self.iters > 0
        g_planes = [cur_g[k] for k in ("worldPos", "albedo", "normal", "matProps")]
        r_planes = lambda i: [f.res[i]["info"], f.res[i]["weight"]]
        if exchange and spatial:
            exchange(g_planes + r_planes(src), "spatial")
        if spatial:
            for it in range(self.iters):
                dst = (src + 1) % 3
                L.orc_pass_spatial(C.byref(s), C.byref(ru), C.c_uint32(clock), C.c_uint32(it), y0, y1, Frame.gbuf(cur_g),
                                   Frame.rbuf(f.res[src]), Frame.rbuf(f.res[dst]))
                src = dst
                if exchange and it + 1 < self.iters:
                    exchange(r_planes(src), "spatial")
        L.orc_pass_shade(C.byref(s), C.byref(ru), C.byref(pc), C.c_uint32(clock), y0, y1, Frame.gbuf(cur_g),
                         Frame.rbuf(f.res[src]), _p(f.accum))
        if exchange and (ru.flags & FLAG_TEMPORAL):
            exchange(g_planes + r_planes(src), "temporal")
        self.final_res = src
        self.last_g = self.cur
        self.cur = 1 - self.cur
        return f.accum

    def gbuffer(self):
        return self.f.g[self.last_g]

    def reservoirs(self):
        return self.f.res[self.final_res]


def path_trace(scene, gu, ru, spp, seed_base=1):
    W, H = ru.screenSize[0], ru.screenSize[1]
    out = np.zeros((H, W, 3), np.float32)
    lib().orc_path_trace(C.byref(scene.c), C.byref(gu), C.byref(ru), C.c_uint32(spp), C.c_uint32(seed_base), _p(out))
    return out


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))
