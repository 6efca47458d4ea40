/* TEST INFRASTRUCTURE ONLY — scalar CPU oracle for the volumetric ReSTIR hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker or the reported CPU baseline.
 * The product (libvrs.so) never links, loads or calls anything under oracle/.
 *
 * Parity pinning: the reference-derived math (RNG, alias table, Disney BRDF, p-hat,
 * reservoir update/combine, final shade, nvmath camera) is pinned against the
 * reference's own sources executed through oracle/_ref/libvrs_ref.so and against
 * the committed vectors in tests/golden/ref_vectors.json.  The volumetric front-end
 * (sparse-grid DDA, delta / ratio tracking, gradient normal) has NO reference
 * implementation: for those functions parity is UNPINNED by
 * the reference; the spec is DESIGN.md §3 and this file is its scalar statement.
 *
 * Build: g++ -O2 -fopenmp -ffp-contract=off (no fast-math).  All arithmetic fp32,
 * evaluated in the order written; the CUDA kernels are compiled with -fmad=false.
 */
#ifndef VRS_ORACLE_H
#define VRS_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int32_t alias; float prob, pdf, aliasPdf; } orc_alias_cell;      /* host_device.h:197-202 */
typedef struct { float pos[4]; float emission_luminance[4]; } orc_point_light;   /* host_device.h:184-187 */

/* RestirUniforms, host_device.h:204-227 (C++ side layout, sizeof 320). */
typedef struct {
  int32_t pointLightCount, triangleLightCount, aliasTableCount;
  float environmentalPower, fireflyClampThreshold;
  uint32_t spatialNeighbors; float spatialRadius;
  uint32_t initialLightSampleCount; int32_t temporalSampleCountMultiplier;
  uint32_t _pad0;
  uint32_t screenSize[2];
  float currCamPos[4];
  float currFrameProjectionViewMatrix[16];
  float prevCamPos[4];
  uint32_t _pad1[12];
  float prevFrameProjectionViewMatrix[16];
  int32_t flags, debugMode; float gamma;
  uint32_t _pad2[13];
} orc_restir_uniforms;

typedef struct { float viewProj[16], viewInverse[16], projInverse[16]; } orc_global_uniforms; /* host_device.h:116-120 */
typedef struct { float clearColorRed, clearColorGreen, clearColorBlue; int32_t frame, initialize; } orc_push_constant; /* :139-145 */

/* Dense density window (the oracle's own grid representation: one float per voxel
 * over the leaf-aligned active bounding box; everything outside = bg_density). */
typedef struct {
  const float* dens;       /* [vdim[2]][vdim[1]][vdim[0]], x fastest */
  int32_t vmin[3];         /* voxel coordinate of dens[0]; multiple of 8 */
  int32_t vdim[3];         /* multiples of 8 */
  float bg_density;
  float A, invA, B[3];     /* world = A*ijk + B */
  float density_scale;     /* sigma_t = density * density_scale */
  float roughness, metallic;
  const orc_point_light* lights; int32_t nlights;
  const orc_alias_cell* table; int32_t ntable;
  float* cellmax;          /* scratch [vdim/8 ...] filled by orc_scene_prepare */
} orc_scene;

typedef struct {            /* planar RGBA32F images, reference layouts */
  float* worldPos; float* albedo; float* normal; float* matProps;
} orc_gbuffer;
typedef struct { float* info; float* weight; } orc_reservoirs;

/* ---- reference-derived math ---- */
void     orc_pcg2d(uint32_t x, uint32_t y, uint32_t* out2);
uint32_t orc_lcg(uint32_t* state);
float    orc_rnd(uint32_t* state);
uint32_t orc_pixel_seed(uint32_t x, uint32_t y, uint32_t clock, uint32_t pass);
float    orc_luminance_common(float r, float g, float b);
float    orc_luminance_utils(float r, float g, float b);
float    orc_disney_brdf_luminance(float cosIn, float cosOut, float cosHalf, float cosInHalf, float lum, float rough, float metal);
void     orc_disney_brdf_color(float cosIn, float cosOut, float cosHalf, float cosInHalf, const float* albedo3, float rough, float metal, float* out3);
/* ginfo16: camPos3 worldPos3 normal3 albedo4 albedoLum roughness metallic */
float    orc_evaluate_phat(const orc_point_light* lights, uint32_t idx, const float* ginfo16);
void     orc_evaluate_phat_full(const orc_point_light* lights, uint32_t idx, const float* ginfo16, float* out3);
void     orc_create_alias_table(const float* pdf, int n, orc_alias_cell* out);
void     orc_alias_table_sample(const orc_alias_cell* t, int n, float r1, float r2, uint32_t* index, float* prob);
void     orc_generate_point_lights(const float* mn3, const float* mx3, int white, uint32_t n, orc_point_light* out);
/* res8: M lightIndex lightKind sampleSeed (u32 bits) pHat sumWeights w pad */
void     orc_initial_ris(const orc_scene* s, const float* ginfo16, int count, uint32_t* seed, uint32_t* res8);
void     orc_combine_geom(const orc_scene* s, uint32_t* self8, const uint32_t* other8, const float* g16, const float* og16, uint32_t* seed);
void     orc_combine_plain(uint32_t* self8, const uint32_t* other8, float pHat, uint32_t* seed);
void     orc_post_shade(const orc_scene* s, const uint32_t* res8, const float* ginfo16, float thr, float* out3);
void     orc_perspectiveVK(float fovy, float aspect, float n, float f, float* out16);
void     orc_look_at(const float* eye, const float* center, const float* up, float* out16);
void     orc_invert(const float* a16, float* out16);
void     orc_matmul(const float* a16, const float* b16, float* out16);
void     orc_voxel_albedo(float density, float* out4);

/* ---- volumetric front-end (new design) ---- */
float    orc_neglog1m(float u);
void     orc_scene_prepare(orc_scene* s);
float    orc_density_at(const orc_scene* s, int i, int j, int k);
/* returns 1 on real collision; fills t, voxel ijk, counters[2] = {tentative collisions, cells entered} */
int      orc_delta_track(const orc_scene* s, const float* org3, const float* dir3, float tmin, float tmax,
                         uint32_t* seed, float* t_hit, int32_t* voxel3, uint32_t* counters2);
float    orc_ratio_track(const orc_scene* s, const float* p3, const float* l3, uint32_t* seed, uint32_t* counters2);

/* ---- passes over whole images (OpenMP over rows) ---- */
/* y0..y1: row range to process (multi-GPU band tests); images are full-size. */
void orc_pass_initial(const orc_scene* s, const orc_global_uniforms* gu, const orc_restir_uniforms* ru,
                      uint32_t clock, int y0, int y1,
                      orc_gbuffer cur, orc_gbuffer prev, orc_reservoirs prevRes, orc_reservoirs outRes,
                      uint32_t* trace4);
void orc_pass_spatial(const orc_scene* s, const orc_restir_uniforms* ru, uint32_t clock, uint32_t iteration,
                      int y0, int y1, orc_gbuffer cur, orc_reservoirs inRes, orc_reservoirs outRes);
void orc_pass_shade(const orc_scene* s, const orc_restir_uniforms* ru, const orc_push_constant* pc, uint32_t clock,
                    int y0, int y1, orc_gbuffer cur, orc_reservoirs res, float* accum);
/* brute-force estimator of the same integrand: spp independent (primary event, light) samples per pixel */
void orc_path_trace(const orc_scene* s, const orc_global_uniforms* gu, const orc_restir_uniforms* ru,
                    uint32_t spp, uint32_t seed_base, float* out_rgb /* W*H*3 */);
void orc_set_num_threads(int n);
int  orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
