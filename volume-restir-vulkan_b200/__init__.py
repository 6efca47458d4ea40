"""volume-restir-vulkan_b200 — Python host mirror of the C ABI in include/vrs.h (libvrs.so).

The product is the shared library (CUDA kernels for sm_100a + C++ host runtime); this module is
ctypes plumbing plus a `Renderer` whose method names follow the reference's `Renderer`
(src/Renderer.h: createRestirUniformBuffer, updateUniformBuffer, updateRestirUniformBuffer,
updateFrame, resetFrame, updateGBufferFrameIdx) so the frame loop of src/main.cpp:301-449 reads
the same.  There is no CPU fallback: importing works anywhere (so the symbol table can be
checked), but every compute call needs a CUDA device and raises VrsError otherwise.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# VRS_LIB picks another build of the same sources (libvrs_relaxed.so: measurements only, not bit-exact)
LIB_PATH = os.path.join(HERE, os.environ.get("VRS_LIB", "libvrs.so"))

VISIBILITY_REUSE_FLAG, TEMPORAL_REUSE_FLAG, SPATIAL_REUSE_FLAG, USE_ENVIRONMENT_FLAG = 1, 2, 4, 8
FINAL_VISIBILITY_FLAG, FINALIZE_W_FLAG = 16, 32

PEER_BLOB_BYTES = 3200

STATUS = {0: "VRS_OK", 1: "VRS_ERR_INVALID", 2: "VRS_ERR_CUDA", 3: "VRS_ERR_IO", 4: "VRS_ERR_FORMAT",
          5: "VRS_ERR_UNSUPPORTED", 6: "VRS_ERR_COMM", 7: "VRS_ERR_NO_DEVICE"}

# every symbol include/vrs.h declares (checked by tests/test_abi.py against the header text)
EXPORTS = [
    "vrs_default_config", "vrs_default_restir_uniforms", "vrs_create", "vrs_destroy", "vrs_last_error", "vrs_abi_version",
    "vrs_load_vdb", "vrs_load_vrsg", "vrs_convert_vdb", "vrs_make_procedural_grid", "vrs_write_procedural_vrsg", "vrs_get_grid_info", "vrs_grid_get_value",
    "vrs_grid_sample_device", "vrs_set_lights", "vrs_collect_emissive_lights", "vrs_vdb_emissive_lights", "vrs_set_triangle_lights", "vrs_get_alias_table", "vrs_create_alias_table",
    "vrs_generate_point_lights", "vrs_perspectiveVK", "vrs_look_at", "vrs_invert", "vrs_mat4_mul", "vrs_pass_initial",
    "vrs_pass_spatial", "vrs_pass_shade", "vrs_render_frame", "vrs_synchronize", "vrs_read_frame", "vrs_read_gbuffer",
    "vrs_read_reservoirs", "vrs_read_trace", "vrs_read_display", "vrs_present_async", "vrs_present_wait", "vrs_write_image", "vrs_get_timings", "vrs_set_pass_timing", "vrs_stream", "vrs_comm_unique_id",
    "vrs_comm_init", "vrs_peer_export", "vrs_peer_connect", "vrs_peer_connect_local", "vrs_band_for_rank",
    "vrs_resize", "vrs_get_counters", "vrs_set_kernel_timing", "vrs_get_kernel_times", "vrs_render_frame_group",
]


class VrsError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("%s: %s" % (STATUS.get(status, status), message))
        self.status = status


class Config(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("band_y0", C.c_uint32), ("band_y1", C.c_uint32),
                ("halo_rows", C.c_uint32), ("device", C.c_int32), ("spatial_iterations", C.c_uint32),
                ("world_scale", C.c_float), ("world_translate", C.c_float * 3), ("density_scale", C.c_float),
                ("roughness", C.c_float), ("metallic", C.c_float), ("enable_trace", C.c_int32)]


class PointLight(C.Structure):
    _fields_ = [("pos", C.c_float * 4), ("emission_luminance", C.c_float * 4)]


class AliasTableCell(C.Structure):
    _fields_ = [("alias", C.c_int32), ("prob", C.c_float), ("pdf", C.c_float), ("aliasPdf", C.c_float)]


class GlobalUniforms(C.Structure):
    _fields_ = [("viewProj", C.c_float * 16), ("viewInverse", C.c_float * 16), ("projInverse", C.c_float * 16)]


class PushConstantRestir(C.Structure):
    _fields_ = [("clearColorRed", C.c_float), ("clearColorGreen", C.c_float), ("clearColorBlue", C.c_float),
                ("frame", C.c_int32), ("initialize", C.c_int32)]


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


class GridInfo(C.Structure):
    _fields_ = [("bbox_min", C.c_int32 * 3), ("bbox_max", C.c_int32 * 3), ("active_voxels", C.c_uint64),
                ("root_children", C.c_uint32), ("internal5", C.c_uint32), ("internal4", C.c_uint32), ("leaves", C.c_uint32),
                ("tiles", C.c_uint32), ("voxel_size", C.c_double), ("translation", C.c_double * 3),
                ("background", C.c_float), ("is_level_set", C.c_int32), ("max_density", C.c_float),
                ("world_bbox_min", C.c_float * 3), ("world_bbox_max", C.c_float * 3), ("device_bytes", C.c_uint64)]


class Timings(C.Structure):
    _fields_ = [("initial_ms", C.c_float), ("spatial_ms", C.c_float), ("shade_ms", C.c_float), ("exchange_ms", C.c_float),
                ("frame_ms", C.c_float), ("launches", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [("candidates", C.c_uint32), ("hits", C.c_uint32), ("shadow_rays", C.c_uint32),
                ("temporal_out_of_halo", C.c_uint32), ("comm_timeouts", C.c_uint32), ("temporal_reach_rows", C.c_uint32)]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 40), ("ms", C.c_float)]


def build(verbose=False):
    """Compile libvrs.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", HERE, "libvrs.so"], stdout=out)


_lib = None


def lib():
    """Load libvrs.so. Fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VrsError(3, "libvrs.so is missing: run `make -C %s` (there is no CPU fallback)" % HERE)
        L = C.CDLL(LIB_PATH)
        L.vrs_last_error.restype = C.c_char_p
        L.vrs_last_error.argtypes = [C.c_void_p]
        L.vrs_stream.restype = C.c_void_p
        f = C.c_float
        L.vrs_perspectiveVK.argtypes = [f, f, f, f, C.c_void_p]
        L.vrs_create.argtypes = [C.c_void_p, C.c_void_p]
        L.vrs_destroy.argtypes = [C.c_void_p]
        L.vrs_destroy.restype = None
        for name in ["vrs_load_vdb", "vrs_load_vrsg", "vrs_make_procedural_grid", "vrs_write_procedural_vrsg", "vrs_get_grid_info", "vrs_grid_get_value",
                     "vrs_grid_sample_device", "vrs_set_lights", "vrs_collect_emissive_lights", "vrs_vdb_emissive_lights", "vrs_set_triangle_lights", "vrs_get_alias_table",
                     "vrs_pass_initial", "vrs_pass_spatial", "vrs_pass_shade", "vrs_render_frame", "vrs_synchronize",
                     "vrs_read_frame", "vrs_read_gbuffer", "vrs_read_reservoirs", "vrs_read_trace", "vrs_write_image",
                     "vrs_get_timings", "vrs_set_pass_timing", "vrs_comm_init", "vrs_read_display", "vrs_present_async", "vrs_present_wait",
                     "vrs_resize", "vrs_get_counters", "vrs_set_kernel_timing", "vrs_get_kernel_times", "vrs_peer_connect_local"]:
            getattr(L, name).restype = C.c_int
        L.vrs_load_vdb.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.vrs_load_vrsg.argtypes = [C.c_void_p, C.c_char_p]
        L.vrs_convert_vdb.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
        L.vrs_make_procedural_grid.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
        L.vrs_write_procedural_vrsg.argtypes = [C.c_int, C.c_uint32, C.c_char_p]
        L.vrs_get_grid_info.argtypes = [C.c_void_p, C.c_void_p]
        L.vrs_grid_get_value.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.vrs_grid_sample_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        L.vrs_set_lights.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.vrs_set_triangle_lights.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.vrs_collect_emissive_lights.argtypes = [C.c_void_p, C.c_float, C.c_uint32, C.c_void_p, C.c_void_p]
        L.vrs_get_alias_table.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.vrs_pass_initial.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.vrs_pass_spatial.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.vrs_pass_shade.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.vrs_render_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.vrs_synchronize.argtypes = [C.c_void_p]
        L.vrs_read_frame.argtypes = [C.c_void_p, C.c_void_p]
        L.vrs_read_gbuffer.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.vrs_read_reservoirs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.vrs_read_trace.argtypes = [C.c_void_p, C.c_void_p]
        L.vrs_write_image.argtypes = [C.c_void_p, C.c_char_p]
        L.vrs_read_display.argtypes = [C.c_void_p, C.c_void_p]
        L.vrs_present_async.argtypes = [C.c_void_p, C.c_void_p]
        L.vrs_present_wait.argtypes = [C.c_void_p]
        L.vrs_get_timings.argtypes = [C.c_void_p, C.c_void_p]
        L.vrs_set_pass_timing.argtypes = [C.c_void_p, C.c_int]
        L.vrs_stream.argtypes = [C.c_void_p]
        L.vrs_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.vrs_peer_export.argtypes = [C.c_void_p, C.c_void_p]
        L.vrs_peer_export.restype = C.c_int
        L.vrs_peer_connect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.vrs_peer_connect.restype = C.c_int
        L.vrs_peer_connect_local.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.vrs_render_frame_group.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.vrs_render_frame_group.restype = C.c_int
        L.vrs_resize.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.vrs_get_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.vrs_set_kernel_timing.argtypes = [C.c_void_p, C.c_int]
        L.vrs_get_kernel_times.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        L.vrs_create_alias_table.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.vrs_generate_point_lights.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_void_p]
        L.vrs_band_for_rank.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------ host helpers (no device)
def perspectiveVK(fovy, aspect, near, far):
    m = np.zeros(16, np.float32)
    lib().vrs_perspectiveVK(fovy, aspect, near, far, _p(m))
    return m


def look_at(eye, center, up=(0.0, 1.0, 0.0)):
    m = np.zeros(16, np.float32)
    e, c, u = (np.ascontiguousarray(v, np.float32) for v in (eye, center, up))
    lib().vrs_look_at(_p(e), _p(c), _p(u), _p(m))
    return m


def invert(a):
    a = np.ascontiguousarray(a, np.float32)
    m = np.zeros(16, np.float32)
    lib().vrs_invert(_p(a), _p(m))
    return m


def mat4_mul(a, b):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    m = np.zeros(16, np.float32)
    lib().vrs_mat4_mul(_p(a), _p(b), _p(m))
    return m


def create_alias_table(pdf):
    pdf = np.ascontiguousarray(pdf, np.float32)
    out = np.zeros(len(pdf), dtype=[("alias", "<i4"), ("prob", "<f4"), ("pdf", "<f4"), ("aliasPdf", "<f4")])
    lib().vrs_create_alias_table(_p(pdf), len(pdf), _p(out))
    return out


def generate_point_lights(mn, mx, white=True, n=100):
    mn, mx = np.ascontiguousarray(mn, np.float32), np.ascontiguousarray(mx, np.float32)
    out = np.zeros((n, 8), np.float32)
    lib().vrs_generate_point_lights(_p(mn), _p(mx), int(white), n, _p(out))
    return out


def convert_vdb(vdb_path, vrsg_path, grid_name=None):
    s = lib().vrs_convert_vdb(vdb_path.encode(), grid_name.encode() if grid_name else None, vrsg_path.encode())
    if s:
        raise VrsError(s, lib().vrs_last_error(None).decode())


PROCEDURAL_KINDS = {"bunny_cloud": 0, "explosion": 1, "fire": 2, "torus_knot_helix": 3, "fire_torus": 4}


def write_procedural_vrsg(kind, resolution, path):
    """Deterministic stand-in for an asset missing from the reference checkout (.MISSING_LARGE_BLOBS:1-4)."""
    k = PROCEDURAL_KINDS[kind] if isinstance(kind, str) else kind
    s = lib().vrs_write_procedural_vrsg(k, resolution, path.encode())
    if s:
        raise VrsError(s, lib().vrs_last_error(None).decode())


def vdb_emissive_lights(vdb_path, grid_name="temperature", max_lights=1001, world_scale=0.05, world_translate=(-2.5, 0.5, 0.0)):
    """Emissive-voxel lights of Renderer::createRestirLights from a temperature grid of a .vdb file (host-only)."""
    L = lib()
    cfg = Config()
    L.vrs_default_config(C.byref(cfg), 1, 1)
    cfg.world_scale = world_scale
    cfg.world_translate[:] = list(world_translate)
    out = np.zeros((max_lights, 8), np.float32)
    n = C.c_uint32()
    L.vrs_vdb_emissive_lights.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    s = L.vrs_vdb_emissive_lights(vdb_path.encode(), grid_name.encode() if grid_name else None, C.byref(cfg), max_lights, _p(out), C.byref(n))
    if s:
        raise VrsError(s, L.vrs_last_error(None).decode())
    return out[:n.value].copy()


def band_for_rank(height, rank, nranks):
    a, b = C.c_uint32(), C.c_uint32()
    lib().vrs_band_for_rank(height, rank, nranks, C.byref(a), C.byref(b))
    return a.value, b.value


def comm_unique_id():
    buf = (C.c_uint8 * 128)()
    s = lib().vrs_comm_unique_id(buf)
    if s:
        raise VrsError(s, lib().vrs_last_error(None).decode())
    return bytes(buf)


def render_frame_group(renderers, clock):
    """renderFrame() of several band Renderers of one image from one host thread (vrs_render_frame_group): every band
    gets the same camera; phases are interleaved across the bands so that no band waits for rows not yet enqueued."""
    for r in renderers:
        r.updateUniformBuffer(); r.updateRestirUniformBuffer(); r.updateFrame()
        r._last_initialize = r.m_pcRestirPost.initialize
    r0 = renderers[0]
    arr = (C.c_void_p * len(renderers))(*[r._ctx for r in renderers])
    s = lib().vrs_render_frame_group(arr, len(renderers), C.byref(r0.m_globalUniforms), C.byref(r0.m_restirUniforms), C.byref(r0.m_pcRestirPost), clock)
    if s:
        raise VrsError(s, "; ".join(lib().vrs_last_error(r._ctx).decode() for r in renderers))
    for r in renderers:
        r.clock = clock + 1
        if r.m_pcRestirPost.frame > 10:
            r.m_pcRestirPost.initialize = 0


class CameraManip:
    """The slice of nvh::CameraManipulator the hot path consumes: eye / center / up / fov (main.cpp:90-101)."""

    def __init__(self, eye=(1.0, 1.0, 1.0), center=(0.0, 1.0, 0.0), up=(0.0, 1.0, 0.0), fov=60.0):
        self.eye, self.center, self.up, self.fov = tuple(eye), tuple(center), tuple(up), fov

    def setLookat(self, eye, center, up=(0.0, 1.0, 0.0)):
        self.eye, self.center, self.up = tuple(eye), tuple(center), tuple(up)

    def getMatrix(self):
        return look_at(self.eye, self.center, self.up)

    def getFov(self):
        return self.fov


class Renderer:
    """ReSTIR slice of the reference's Renderer (src/Renderer.h:45-333) over libvrs."""

    def __init__(self, width, height, spatial_iterations=2, density_scale=10.0, enable_trace=False, band=None, halo_rows=32,
                 device=-1, world_scale=0.05, world_translate=(-2.5, 0.5, 0.0)):
        L = lib()
        cfg = Config()
        L.vrs_default_config(C.byref(cfg), width, height)
        cfg.spatial_iterations, cfg.density_scale, cfg.enable_trace = spatial_iterations, density_scale, int(enable_trace)
        cfg.halo_rows, cfg.device, cfg.world_scale = halo_rows, device, world_scale
        cfg.world_translate[:] = list(world_translate)
        if band is not None:
            cfg.band_y0, cfg.band_y1 = band
        self.cfg = cfg
        self.width, self.height = width, height
        self.band = band if band is not None else (0, height)
        self.rows = self.band[1] - self.band[0]
        self._ctx = C.c_void_p()
        s = L.vrs_create(C.byref(cfg), C.byref(self._ctx))
        if s:
            raise VrsError(s, L.vrs_last_error(None).decode())
        self.CameraManip = CameraManip()
        self.m_restirUniforms = RestirUniforms()
        L.vrs_default_restir_uniforms(C.byref(self.m_restirUniforms), width, height)
        self.m_globalUniforms = GlobalUniforms()
        self.m_pcRestirPost = PushConstantRestir(0.0, 0.0, 0.0, 0, 1)      # restirPass.h:52
        self._ref_cam = None
        self.clock = 0
        self.n_lights = 0

    # ---- lifetime
    def destroy(self):
        if self._ctx:
            lib().vrs_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def _ck(self, s):
        if s:
            raise VrsError(s, lib().vrs_last_error(self._ctx).decode())

    # ---- scene (main.cpp:221-224, 252-258)
    def loadVDB(self, path, grid_name=None):
        self._ck(lib().vrs_load_vdb(self._ctx, path.encode(), grid_name.encode() if grid_name else None))

    def makeProceduralGrid(self, kind, resolution):
        self._ck(lib().vrs_make_procedural_grid(self._ctx, kind, resolution))

    def gridInfo(self):
        gi = GridInfo()
        self._ck(lib().vrs_get_grid_info(self._ctx, C.byref(gi)))
        return gi

    def gridGetValue(self, i, j, k):
        v, a = C.c_float(), C.c_int32()
        self._ck(lib().vrs_grid_get_value(self._ctx, i, j, k, C.byref(v), C.byref(a)))
        return v.value, bool(a.value)

    def gridSampleDevice(self, ijk):
        ijk = np.ascontiguousarray(ijk, np.int32).reshape(-1, 3)
        out = np.zeros(len(ijk), np.float32)
        self._ck(lib().vrs_grid_sample_device(self._ctx, _p(ijk), len(ijk), _p(out)))
        return out

    def createRestirLights(self, lights):
        """lights: (n, 8) float32 rows {pos.xyzw, emission.rgb, luminance} (PointLight, host_device.h:184-187)."""
        lights = np.ascontiguousarray(lights, np.float32).reshape(-1, 8)
        self._ck(lib().vrs_set_lights(self._ctx, _p(lights), len(lights)))
        self.n_lights = len(lights)
        self.m_restirUniforms.pointLightCount = len(lights)
        self.m_restirUniforms.triangleLightCount = 0
        self.m_restirUniforms.aliasTableCount = len(lights)

    def collectEmissiveLights(self, threshold, max_lights=1001):
        """Voxel lights of Renderer::createRestirLights (Renderer.cpp:1615-1637) from the loaded grid."""
        out = np.zeros((max_lights, 8), np.float32)
        n = C.c_uint32()
        self._ck(lib().vrs_collect_emissive_lights(self._ctx, threshold, max_lights, _p(out), C.byref(n)))
        return out[:n.value].copy()

    def aliasTable(self):
        out = np.zeros(self.n_lights, dtype=[("alias", "<i4"), ("prob", "<f4"), ("pdf", "<f4"), ("aliasPdf", "<f4")])
        self._ck(lib().vrs_get_alias_table(self._ctx, _p(out), self.n_lights))
        return out

    # ---- uniforms (Renderer.cpp:116-161, 2339-2425, 2456-2472)
    def _proj_view(self):
        aspect = float(np.float32(self.width) / np.float32(self.height))
        view = self.CameraManip.getMatrix()
        proj = perspectiveVK(self.CameraManip.getFov(), aspect, 0.1, 1000.0)
        return view, proj

    def createRestirUniformBuffer(self):
        view, proj = self._proj_view()
        u = self.m_restirUniforms
        pv = mat4_mul(proj, view)
        e = self.CameraManip.eye
        u.currCamPos[:] = [e[0], e[1], e[2], 0.0]
        u.prevCamPos[:] = [e[0], e[1], e[2], 0.0]
        u.currFrameProjectionViewMatrix[:] = pv.tolist()
        u.prevFrameProjectionViewMatrix[:] = pv.tolist()

    def updateUniformBuffer(self):
        view, proj = self._proj_view()
        g = self.m_globalUniforms
        g.viewProj[:] = mat4_mul(proj, view).tolist()
        g.viewInverse[:] = invert(view).tolist()
        g.projInverse[:] = invert(proj).tolist()

    def updateRestirUniformBuffer(self):
        u = self.m_restirUniforms
        u.prevCamPos[:] = list(u.currCamPos)
        u.prevFrameProjectionViewMatrix[:] = list(u.currFrameProjectionViewMatrix)
        u.screenSize[0], u.screenSize[1] = self.width, self.height
        view, proj = self._proj_view()
        e = self.CameraManip.eye
        u.currCamPos[:] = [e[0], e[1], e[2], 0.0]
        u.currFrameProjectionViewMatrix[:] = mat4_mul(proj, view).tolist()

    def resetFrame(self):
        self.m_pcRestirPost.frame = -1

    def updateFrame(self):
        cam = (self.CameraManip.getMatrix().tobytes(), self.CameraManip.getFov())
        if cam != self._ref_cam:
            self.resetFrame()
            self._ref_cam = cam
        self.m_pcRestirPost.frame += 1

    # ---- per frame (main.cpp:339-448)
    def renderFrame(self, clock=None):
        """updateUniformBuffer -> updateRestirUniformBuffer -> updateFrame -> RestirPass::run -> SpatialReusePass::run
        -> restirDrawPost -> initialize=0 after frame 10 -> updateGBufferFrameIdx."""
        self.updateUniformBuffer()
        self.updateRestirUniformBuffer()
        self.updateFrame()
        self._last_initialize = self.m_pcRestirPost.initialize
        self.submit(clock)
        if self.m_pcRestirPost.frame > 10:
            self.m_pcRestirPost.initialize = 0

    def submit(self, clock=None):
        if clock is None:
            clock = self.clock
        self._ck(lib().vrs_render_frame(self._ctx, C.byref(self.m_globalUniforms), C.byref(self.m_restirUniforms),
                                        C.byref(self.m_pcRestirPost), clock))
        self.clock = clock + 1

    def passInitial(self, clock):
        self._ck(lib().vrs_pass_initial(self._ctx, C.byref(self.m_globalUniforms), C.byref(self.m_restirUniforms), clock))

    def passSpatial(self, clock, iteration):
        self._ck(lib().vrs_pass_spatial(self._ctx, C.byref(self.m_restirUniforms), clock, iteration))

    def passShade(self, clock):
        self._ck(lib().vrs_pass_shade(self._ctx, C.byref(self.m_restirUniforms), C.byref(self.m_pcRestirPost), clock))

    def synchronize(self):
        self._ck(lib().vrs_synchronize(self._ctx))

    # ---- readback
    def _img(self, dtype=np.float32):
        return np.zeros((self.rows, self.width, 4), dtype)

    def readFrame(self, out=None):
        out = self._img() if out is None else out
        self._ck(lib().vrs_read_frame(self._ctx, _p(out)))
        return out

    def readGBuffer(self):
        a, b, c, d = self._img(), self._img(), self._img(), self._img()
        self._ck(lib().vrs_read_gbuffer(self._ctx, _p(a), _p(b), _p(c), _p(d)))
        return dict(worldPos=a, albedo=b, normal=c, matProps=d)

    def readReservoirs(self):
        a, b = self._img(), self._img()
        self._ck(lib().vrs_read_reservoirs(self._ctx, _p(a), _p(b)))
        return dict(info=a, weight=b)

    def readTrace(self):
        t = self._img(np.uint32)
        self._ck(lib().vrs_read_trace(self._ctx, _p(t)))
        return t

    def readDisplay(self, out=None):
        out = np.zeros((self.rows, self.width, 4), np.uint8) if out is None else out
        self._ck(lib().vrs_read_display(self._ctx, _p(out)))
        return out

    def presentAsync(self, out):
        """Headless swapchain present: tonemap + async device->host copy of the 8-bit frame into `out` (pinned)."""
        self._ck(lib().vrs_present_async(self._ctx, _p(out)))

    def presentWait(self):
        self._ck(lib().vrs_present_wait(self._ctx))

    def writeImage(self, path):
        self._ck(lib().vrs_write_image(self._ctx, path.encode()))

    def setPassTiming(self, enabled):
        self._ck(lib().vrs_set_pass_timing(self._ctx, int(bool(enabled))))

    def timings(self):
        t = Timings()
        self._ck(lib().vrs_get_timings(self._ctx, C.byref(t)))
        return t

    def stream(self):
        return lib().vrs_stream(self._ctx)

    def resize(self, width, height):
        """RestirPass / SpatialReusePass::createRenderPass(VkExtent2D) + the buffer re-creation a window resize implies."""
        self._ck(lib().vrs_resize(self._ctx, width, height))
        self.width, self.height, self.band, self.rows = width, height, (0, height), height
        self.m_restirUniforms.screenSize[0], self.m_restirUniforms.screenSize[1] = width, height
        self._ref_cam = None

    def counters(self):
        c = Counters()
        self._ck(lib().vrs_get_counters(self._ctx, C.byref(c)))
        return c

    def setKernelTiming(self, enabled):
        self._ck(lib().vrs_set_kernel_timing(self._ctx, int(bool(enabled))))

    def kernelTimes(self):
        """[(kernel name, ms)] of the last frame, in launch order (setKernelTiming(True) frames only)."""
        buf = (KernelTime * 64)()
        n = C.c_uint32()
        self._ck(lib().vrs_get_kernel_times(self._ctx, buf, 64, C.byref(n)))
        return [(buf[i].name.decode(), float(buf[i].ms)) for i in range(n.value)]

    def peerConnectLocal(self, up=None, down=None):
        """Wire this band to the contexts rendering the bands above / below it in the same process."""
        self._ck(lib().vrs_peer_connect_local(self._ctx, up._ctx if up is not None else None, down._ctx if down is not None else None))

    def peerExport(self):
        """CUDA-IPC handles of this context's planes (bytes) for the peer-memory halo exchange."""
        buf = (C.c_uint8 * PEER_BLOB_BYTES)()
        self._ck(lib().vrs_peer_export(self._ctx, buf))
        return bytes(buf)

    def peerConnect(self, rank, nranks, blobs):
        """blobs: the peerExport() bytes of every rank, in rank order (all-gathered by the launcher)."""
        joined = b"".join(blobs)
        assert len(joined) == nranks * PEER_BLOB_BYTES
        buf = (C.c_uint8 * len(joined)).from_buffer_copy(joined)
        self._ck(lib().vrs_peer_connect(self._ctx, rank, nranks, buf))

    def commInit(self, unique_id, rank, nranks):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._ck(lib().vrs_comm_init(self._ctx, buf, rank, nranks))
