/* libvrs — C ABI of the B200-native volumetric ReSTIR hot path.
 *
 * Drop-in boundary for the per-frame pass sequence of TheSmokeyGuys/Volume-ReSTIR-Vulkan
 * (initial RIS candidates -> visibility -> temporal reuse -> spatial reuse -> final shading)
 * and for the scene-upload calls that feed it.  The reference has no FFI; its seam is the pair
 * of pass classes plus the Renderer resource methods.  Every entry point below
 * cites the reference interface it replaces (paths relative to the reference root).
 *
 * Conventions: plain C, no exceptions, every call returns vrs_status; matrices are column-major
 * float[16] exactly like nvmath::mat4f (nvmath_types.h:866); host arrays passed in are copied
 * before the call returns; a context is bound to one CUDA device and one stream and is not
 * thread-safe.  There is no CPU fallback: without a CUDA device vrs_create fails.
 */
#ifndef VRS_H
#define VRS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define VRS_ABI_VERSION 2

typedef enum {
  VRS_OK = 0,
  VRS_ERR_INVALID = 1,      /* bad argument / call order */
  VRS_ERR_CUDA = 2,         /* CUDA runtime error (see vrs_last_error) */
  VRS_ERR_IO = 3,           /* file could not be read / written */
  VRS_ERR_FORMAT = 4,       /* unsupported or corrupt .vdb / .vrsg content */
  VRS_ERR_UNSUPPORTED = 5,  /* feature outside the hot path (e.g. triangle lights) */
  VRS_ERR_COMM = 6,         /* NCCL error, or a halo exchange that timed out / left the stored rows (see vrs_get_counters) */
  VRS_ERR_NO_DEVICE = 7
} vrs_status;

/* ---- structs shared with the reference's shaders: src/shaders/host_device.h -------------- */
typedef struct { float pos[4]; float emission_luminance[4]; } vrs_point_light;            /* PointLight :184-187, w = luminance */
typedef struct { float p1[4], p2[4], p3[4], emission_luminance[4], normalArea[4]; } vrs_triangle_light; /* :189-195 */
typedef struct { int32_t alias; float prob, pdf, aliasPdf; } vrs_alias_table_cell;        /* AliasTableCell :197-202 */
typedef struct { float viewProj[16], viewInverse[16], projInverse[16]; } vrs_global_uniforms; /* GlobalUniforms :116-120 */
typedef struct { float clearColorRed, clearColorGreen, clearColorBlue; int32_t frame, initialize; } vrs_push_constant_restir; /* :139-145 */

/* RestirUniforms :204-227 — same field names, order and C++ offsets (sizeof 320). */
typedef struct {
  int32_t  pointLightCount;                /*   0 */
  int32_t  triangleLightCount;             /*   4 */
  int32_t  aliasTableCount;                /*   8 */
  float    environmentalPower;             /*  12 */
  float    fireflyClampThreshold;          /*  16 */
  uint32_t spatialNeighbors;               /*  20 */
  float    spatialRadius;                  /*  24 */
  uint32_t initialLightSampleCount;        /*  28 */
  int32_t  temporalSampleCountMultiplier;  /*  32 */
  uint32_t _pad0;
  uint32_t screenSize[2];                  /*  40 */
  float    currCamPos[4];                  /*  48 */
  float    currFrameProjectionViewMatrix[16]; /* 64 */
  float    prevCamPos[4];                  /* 128 */
  uint32_t _pad1[12];
  float    prevFrameProjectionViewMatrix[16]; /* 192 */
  int32_t  flags;                          /* 256 */
  int32_t  debugMode;                      /* 260 */
  float    gamma;                          /* 264 */
  uint32_t _pad2[13];
} vrs_restir_uniforms;

/* flags: host_device.h:427-430, plus two bits that did not exist in the reference */
#define VRS_RESTIR_VISIBILITY_REUSE_FLAG (1 << 0)
#define VRS_RESTIR_TEMPORAL_REUSE_FLAG   (1 << 1)
#define VRS_RESTIR_SPATIAL_REUSE_FLAG    (1 << 2)
#define VRS_USE_ENVIRONMENT_FLAG         (1 << 3)   /* accepted, ignored (dead in the reference too) */
#define VRS_FINAL_VISIBILITY_FLAG        (1 << 4)   /* NEW: fresh shadow transmittance in the shade pass */
#define VRS_FINALIZE_W_FLAG              (1 << 5)   /* NEW: w = sumW / (M * pHat) after the initial RIS loop */

#define VRS_LIGHT_KIND_POINT 0                       /* structs/light.glsl */
#define VRS_MAX_SPATIAL_NEIGHBORS 16
#define VRS_MAX_SPATIAL_ITERATIONS 4

/* ---- context ----------------------------------------------------------------------------- */
typedef struct vrs_ctx vrs_ctx;

/* Replaces RestirPass::setup/createRenderPass/createPipeline (src/passes/restirPass.h:15-25),
 * SpatialReusePass::setup/... (src/passes/spatialReusePass.h:14-25) and the buffer creation in
 * Renderer::createGBuffers / createRestirBuffer (src/Renderer.cpp:100-106, 673-761). */
typedef struct {
  uint32_t width, height;       /* full image (RestirUniforms::screenSize) */
  uint32_t band_y0, band_y1;    /* rows this context renders; 0,0 = whole image */
  uint32_t halo_rows;           /* rows kept above/below the band for spatial/temporal reads (multi-GPU) */
  int32_t  device;              /* CUDA ordinal, -1 = current */
  uint32_t spatial_iterations;  /* NEW knob (the reference's spatial pass is a single pass-through) */
  /* volume material + placement: Renderer::createVDBBuffer (src/Renderer.cpp:1420-1435, 1494-1509) */
  float    world_scale;         /* 0.05 */
  float    world_translate[3];  /* (-2.5, 0.5, 0) */
  float    density_scale;       /* sigma_t = density * density_scale [1/world unit] (new) */
  float    roughness;           /* 0.9 */
  float    metallic;            /* 0.0001 */
  int32_t  enable_trace;        /* keep a per-pixel u32x4 trace buffer for parity tests */
} vrs_config;

void       vrs_default_config(vrs_config* cfg, uint32_t width, uint32_t height);
/* Defaults of Renderer::createRestirUniformBuffer (src/Renderer.cpp:2341-2358). */
void       vrs_default_restir_uniforms(vrs_restir_uniforms* u, uint32_t width, uint32_t height);
vrs_status vrs_create(const vrs_config* cfg, vrs_ctx** out);
void       vrs_destroy(vrs_ctx* ctx);                                  /* RestirPass::destroy restirPass.h:35 */
/* RestirPass::createRenderPass(VkExtent2D) / SpatialReusePass::createRenderPass (src/passes/restirPass.cpp:79-81,
 * spatialReusePass.cpp:41-43), Renderer::onResize (src/Renderer.cpp:1022-1026) and the createGBuffers / createRestirBuffer it implies: new per-pixel buffers
 * for width x height (whole image, no band), grid and lights stay staged, the temporal history is dropped.
 * Not available on a context that is part of a multi-GPU group (VRS_ERR_INVALID): re-create the group instead. */
vrs_status vrs_resize(vrs_ctx* ctx, uint32_t width, uint32_t height);
const char* vrs_last_error(const vrs_ctx* ctx);                        /* ctx may be NULL: last create error */
int        vrs_abi_version(void);

/* ---- scene upload ------------------------------------------------------------------------ */
/* VDBLoader::Load (src/loaders/VDBLoader.cpp:5-70) + Renderer::createVDBBuffer (src/Renderer.cpp:1408):
 * parse an OpenVDB file, flatten the first (or named) float grid into root/internal/leaf tables plus a
 * dense brick atlas, stage it into HBM.  `.vrsg` files (this library's own flattened snapshot) load the same way. */
vrs_status vrs_load_vdb(vrs_ctx* ctx, const char* path, const char* grid_name);
vrs_status vrs_load_vrsg(vrs_ctx* ctx, const char* path);
/* Host-only conversion `.vdb` -> `.vrsg` (no device needed); a `.vrsg` input is validated and re-written. */
vrs_status vrs_convert_vdb(const char* vdb_path, const char* grid_name, const char* vrsg_path);
/* Deterministic procedural stand-ins for the assets missing from the reference checkout
 * (.MISSING_LARGE_BLOBS:1-4): kind 0 = "bunny_cloud", 1 = "explosion", 2 = "fire", 3 = "torus_knot_helix",
 * 4 = "fire_torus" (the fire + torus_knot_helix composite of BASELINE.json configs[4], merged into one grid). */
vrs_status vrs_make_procedural_grid(vrs_ctx* ctx, int kind, uint32_t resolution);
/* Host-only: generate the same stand-in and write it as a `.vrsg` snapshot (no device needed). */
vrs_status vrs_write_procedural_vrsg(int kind, uint32_t resolution, const char* vrsg_path);

typedef struct {
  int32_t  bbox_min[3], bbox_max[3];    /* active-voxel bounding box (index space) */
  uint64_t active_voxels;
  uint32_t root_children, internal5, internal4, leaves, tiles;
  double   voxel_size; double translation[3];
  float    background; int32_t is_level_set; float max_density;
  float    world_bbox_min[3], world_bbox_max[3];
  uint64_t device_bytes;
} vrs_grid_info;
vrs_status vrs_get_grid_info(const vrs_ctx* ctx, vrs_grid_info* out);
/* ValueAccessor::getValue (src/vdb/vdb.cpp:777-786) on the flattened host tables: raw value + active flag. */
vrs_status vrs_grid_get_value(const vrs_ctx* ctx, int32_t i, int32_t j, int32_t k, float* value, int32_t* active);
/* Same lookup executed ON THE DEVICE through the staged tables (densities, n points). */
vrs_status vrs_grid_sample_device(vrs_ctx* ctx, const int32_t* ijk, uint32_t n, float* density_out);

/* Renderer::createRestirLights (src/Renderer.cpp:1587-1691): upload point lights and build the alias
 * table with createAliasTable semantics (src/utils/restir_utils.cpp:90-155), pdf = emission_luminance.w. */
vrs_status vrs_set_lights(vrs_ctx* ctx, const vrs_point_light* lights, uint32_t n);
/* Emissive-voxel lights of Renderer::createRestirLights (src/Renderer.cpp:1615-1637): walk the active voxels in tree
 * order, turn every voxel whose raw value exceeds `threshold` (the reference tests `temp > 275` on the temperature
 * grid) into a point light at the voxel's world position with emission (0.6, 0.2, 0.1), stop after `max_lights`
 * (the reference stops after 1001).  Host-side; fills `out` and returns the count. */
vrs_status vrs_collect_emissive_lights(const vrs_ctx* ctx, float threshold, uint32_t max_lights, vrs_point_light* out, uint32_t* count);
/* The same rule on a temperature grid read straight from a `.vdb` (multi-grid files: density is rendered, temperature only
 * feeds the lights): vdb/vdb.cpp:809-812 computes temp = log(T) + 273.15 per active voxel and Renderer.cpp:1623 keeps
 * temp > 275.  Host-only; positions use `cfg`'s world placement (only world_scale / world_translate are read). */
vrs_status vrs_vdb_emissive_lights(const char* vdb_path, const char* grid_name, const vrs_config* cfg, uint32_t max_lights,
                                   vrs_point_light* out, uint32_t* count);
vrs_status vrs_set_triangle_lights(vrs_ctx* ctx, const vrs_triangle_light* lights, uint32_t n); /* VRS_ERR_UNSUPPORTED */
vrs_status vrs_get_alias_table(const vrs_ctx* ctx, vrs_alias_table_cell* out, uint32_t n);

/* Host helpers with the reference's semantics (no ctx, no device). */
void vrs_create_alias_table(const float* pdf, uint32_t n, vrs_alias_table_cell* out);             /* restir_utils.cpp:90-155 */
void vrs_generate_point_lights(const float min3[3], const float max3[3], int white, uint32_t n,
                               vrs_point_light* out);                                             /* restir_utils.cpp:22-51 */
void vrs_perspectiveVK(float fovy_deg, float aspect, float n, float f, float out16[16]);          /* nvmath.inl:1149-1183 */
void vrs_look_at(const float eye[3], const float center[3], const float up[3], float out16[16]);  /* nvmath.inl:979-1025 */
void vrs_invert(const float a16[16], float out16[16]);                                            /* nvmath.inl:797-850 */
void vrs_mat4_mul(const float a16[16], const float b16[16], float out16[16]);                     /* nvmath.inl:663-684 */

/* ---- per frame ---------------------------------------------------------------------------- */
/* `clock` stands in for int(clockARB()) (restir.rgen:139, spatialReuse.comp:58): the per-pixel seed of pass p
 * is pcg2d(uvec2(x,y) * (clock*8 + p + 1)).x + .y with p = 0 initial, 1+i spatial iteration i, 5 shade. */

/* RestirPass::run (src/passes/restirPass.cpp:10-59) -> restir.rgen: primary volume event, G-buffer,
 * initial RIS, visibility (transmittance), temporal reuse.  Writes the tmp reservoir. */
vrs_status vrs_pass_initial(vrs_ctx* ctx, const vrs_global_uniforms* gu, const vrs_restir_uniforms* ru, uint32_t clock);
/* SpatialReusePass::run (src/passes/spatialReusePass.cpp:10-29) -> spatialReuse.comp, one iteration. */
vrs_status vrs_pass_spatial(vrs_ctx* ctx, const vrs_restir_uniforms* ru, uint32_t clock, uint32_t iteration);
/* Renderer::restirDrawPost (src/Renderer.cpp:1247-1264) -> restir_post.frag; then flips the ping-pong
 * like Renderer::updateGBufferFrameIdx (src/Renderer.cpp:108-111). */
vrs_status vrs_pass_shade(vrs_ctx* ctx, const vrs_restir_uniforms* ru, const vrs_push_constant_restir* pc, uint32_t clock);
/* The three calls above in main.cpp:405-433 order (spatial x spatial_iterations when the flag is set);
 * asynchronous, replayed from CUDA graphs; consecutive frames overlap (frames in flight, nvvk/appbase_vk.cpp:412-418):
 * vrs_synchronize / any vrs_read_* returns the state after the last frame enqueued. */
vrs_status vrs_render_frame(vrs_ctx* ctx, const vrs_global_uniforms* gu, const vrs_restir_uniforms* ru,
                            const vrs_push_constant_restir* pc, uint32_t clock);
/* The same frame on several contexts driven by ONE host thread — bands of one image wired with vrs_peer_connect_local
 * (several GPUs of one process, or several bands on one GPU).  The frame is enqueued phase by phase across the contexts
 * (initial pass of every band, then the temporal merge of every band, then each spatial iteration, then the shade), so
 * that a band's wait for its neighbours' halo rows is always enqueued after the stores it waits for. */
vrs_status vrs_render_frame_group(vrs_ctx** ctxs, uint32_t n, const vrs_global_uniforms* gu, const vrs_restir_uniforms* ru,
                                  const vrs_push_constant_restir* pc, uint32_t clock);
vrs_status vrs_synchronize(vrs_ctx* ctx);

/* ---- readback (none in the reference: it presents to a swapchain) -------------------------- */
/* All images are band rows [band_y0, band_y1) x width, RGBA32F, reference layouts. */
vrs_status vrs_read_frame(vrs_ctx* ctx, float* rgba);                                   /* accumulation image m_storageImage */
vrs_status vrs_read_gbuffer(vrs_ctx* ctx, float* worldPos, float* albedo, float* normal, float* matProps); /* last rendered frame */
vrs_status vrs_read_reservoirs(vrs_ctx* ctx, float* info, float* weight);               /* last written reservoir buffer */
vrs_status vrs_read_trace(vrs_ctx* ctx, uint32_t* trace4);                              /* enable_trace only */
/* Headless replacement of the swapchain present (nvvk::AppBaseVk::submitFrame, appbase_vk.cpp:437-475): the display
 * image restir_post.frag:104 outputs — pow(accum, 1/0.8), clamped, 8 bits per channel RGBA — for band rows.
 * vrs_present_async tonemaps into one of two device staging buffers on the context stream and copies it to `rgba8`
 * (pinned host memory for real overlap) on a second stream, so the copy of frame i overlaps the render of frame i+1;
 * vrs_present_wait blocks until every outstanding copy has landed. */
vrs_status vrs_read_display(vrs_ctx* ctx, uint8_t* rgba8);
vrs_status vrs_present_async(vrs_ctx* ctx, uint8_t* rgba8);
vrs_status vrs_present_wait(vrs_ctx* ctx);
/* Headless replacement of the swapchain present: write band rows as PFM (linear) or PPM (pow(c, 1/0.8), restir_post.frag:104). */
vrs_status vrs_write_image(vrs_ctx* ctx, const char* path);

/* ---- introspection for benches / tests ------------------------------------------------------ */
typedef struct { float initial_ms, spatial_ms, shade_ms, exchange_ms, frame_ms; uint32_t launches; } vrs_timings;
/* CUDA-event times of the last vrs_render_frame on the context stream (valid after vrs_synchronize). */
vrs_status vrs_get_timings(vrs_ctx* ctx, vrs_timings* out);
/* Per-pass CUDA events inside every frame (off by default: vrs_get_timings fails with VRS_ERR_INVALID until they are
   enabled and a frame ran).  While enabled, frames do not overlap: the front half of frame n+1 (everything of the
   initial pass that does not read frame n) otherwise runs on a second stream while the back half of frame n executes. */
vrs_status vrs_set_pass_timing(vrs_ctx* ctx, int enabled);
void*      vrs_stream(vrs_ctx* ctx);                                                    /* cudaStream_t */
/* Work counters of the last frame (valid after vrs_synchronize) and the cumulative health counters of the halo exchange:
 * temporal_out_of_halo = hit pixels whose temporal reprojection fell on a row of the image this context does not store
 * (halo_rows too small for the camera motion: those merges were skipped, the frame differs from a single-GPU frame);
 * comm_timeouts = halo waits that gave up (a neighbour never published its rows); temporal_reach_rows = the largest vertical
 * distance (rows) any temporal reprojection has moved so far — what a launcher sizes band heights / halos from. */
typedef struct { uint32_t candidates, hits, shadow_rays, temporal_out_of_halo, comm_timeouts, temporal_reach_rows; } vrs_counters;
vrs_status vrs_get_counters(vrs_ctx* ctx, vrs_counters* out);
/* Per-kernel CUDA-event times of the last frame.  While enabled, frames are launched kernel by kernel (no CUDA graph)
 * with an event after every kernel on the context stream; meant for a short probe next to the timed loops. */
typedef struct { char name[40]; float ms; } vrs_kernel_time;
vrs_status vrs_set_kernel_timing(vrs_ctx* ctx, int enabled);
vrs_status vrs_get_kernel_times(vrs_ctx* ctx, vrs_kernel_time* out, uint32_t capacity, uint32_t* count);

/* ---- multi-GPU: screen-space bands, grid replicated, halo rows exchanged over NCCL ---------- */
/* 128-byte ncclUniqueId produced on rank 0 and broadcast by the launcher (torch.distributed / MPI / file). */
vrs_status vrs_comm_unique_id(uint8_t id128[128]);
vrs_status vrs_comm_init(vrs_ctx* ctx, const uint8_t id128[128], int rank, int nranks);
/* Peer-memory halo exchange (preferred on NVLink / NVSwitch boxes): no NCCL.  Every rank exports the CUDA IPC handles of
 * its per-pixel planes, the launcher all-gathers the blobs (any transport), every rank opens its two neighbours' planes
 * and from then on ONE kernel per exchange stores the boundary rows straight into the neighbours' halo rows over NVLink
 * and releases a system-scope flag; consumers acquire the flag in a one-thread wait kernel.  Being plain kernels, the
 * exchange is part of the captured CUDA graph of the frame.  A context takes ONE transport, once: vrs_peer_connect or
 * vrs_comm_init on a context that is already connected returns VRS_ERR_INVALID. */
#define VRS_PEER_BLOB_BYTES 3200
vrs_status vrs_peer_export(vrs_ctx* ctx, uint8_t blob[VRS_PEER_BLOB_BYTES]);
vrs_status vrs_peer_connect(vrs_ctx* ctx, int rank, int nranks, const uint8_t* all_blobs /* nranks x VRS_PEER_BLOB_BYTES */);
/* The same wiring between contexts of ONE process (any mix of devices with peer access, or several bands on one device):
 * `up` / `down` are the contexts that render the bands above / below this one (NULL at the image edges).  Frames of the
 * connected contexts must all be enqueued before any of them is synchronized (each waits for its neighbours' rows). */
vrs_status vrs_peer_connect_local(vrs_ctx* ctx, vrs_ctx* up, vrs_ctx* down);
/* Even band split of `height` rows over nranks (helper for launchers). */
void       vrs_band_for_rank(uint32_t height, int rank, int nranks, uint32_t* y0, uint32_t* y1);

#ifdef __cplusplus
}
#endif
#endif /* VRS_H */
