// Device-side math of the volumetric ReSTIR hot path (sm_100a).
//
// Everything here is fp32 evaluated in the order written; the translation unit is compiled with
// -fmad=false (no FMA contraction), IEEE division and square root, no fast-math, so that integer state
// (RNG streams, voxel / leaf indices, chosen light indices) is bit-exact against the CPU oracle.
// Citations are relative to the reference root (TheSmokeyGuys/Volume-ReSTIR-Vulkan).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vrs {

// ------------------------------------------------------------------ scene tables staged in HBM
struct GridDev {
  int vmin[3], vdim[3], cdim[3];  // leaf-aligned active window (voxels / 8^3 cells)
  float bg_density;
  float A, invA, B[3];            // world = A * ijk + B
  float density_scale;
  float roughness, metallic;
  int nroot;
  const int4* root;               // {origin xyz, child}: child >= 0 internal5 node, < 0 ~tile
  const int* i5;                  // [n5][32768]  >= 0 internal4 node, < 0 ~tile
  const int* i4;                  // [n4][4096]   >= 0 leaf, < 0 ~tile
  const float* tile_density;      // [ntile], tile 0 = background
  const float* leaf_max;          // [nleaf] majorant density of the brick
  const float* atlas;             // [nleaf][512] densities, offset (x<<6)|(y<<3)|z
  // Dense directory over the window's 8^3 cells (the top tree levels flattened once more): the raymarch reads one
  // L1/L2-resident entry per cell instead of walking root -> internal5 -> internal4.
  const float2* dir;              // [cdim z][cdim y][cdim x] {majorant density of the cell (0 = empty), leaf index (>= 0) or ~tile (< 0) as bits}:
                                  // one 8-byte load per visited cell serves both the majorant test and the later brick lookup
};

struct LightsDev {
  const float4* lights;           // 2 x float4 per PointLight (host_device.h:184-187)
  const float4* alias;            // AliasTableCell as 16 bytes {alias, prob, pdf, aliasPdf} (:197-202)
  int nlights, ntable;
};

struct FrameParams {
  float viewInverse[16], projInverse[16];
  float prevVP[16];
  float camPos[3];
  uint32_t W, H;
  uint32_t M;                     // initialLightSampleCount
  int temporalMult;
  uint32_t spatialNeighbors;
  float spatialRadius;
  float fireflyClamp;
  int flags;
  uint32_t clock;
  float clear[3];
  int frame, initialize;
  float cullVP[16];               // world -> clip of the primary rays (inverse of viewInverse / projInverse), for k_cover
  int cull;                       // 1 = cullVP is valid, screen-space coverage culling may be used
  float roughness, metallic;      // the volume's material constants (what restir.rgen:197 stores in the matProps image of every hit)
};

struct Planes {                   // band-local RGBA32F planes (reference layouts)
  float4* worldPos; float4* albedo; float4* normal; float4* mat;
};
struct ResPlanes { float4* info; float4* weight; };

// Where the temporal pass finds the previous frame's pixel of image row fy: rows [own_y0, own_y1) in this context's own
// planes; with peer memory, rows [up_y0, own_y0) in the up neighbour's planes (its first stored row is up_row0) and rows
// [own_y1, down_y1) in the down neighbour's.  Null plane pointers = no neighbour on that side.
struct PrevAccess {
  int own_y0, own_y1;
  Planes up, down;
  ResPlanes upR, downR;
  int up_y0, up_row0, down_y1, down_row0;
};

// Per-frame work lists (all indices are band-local pixel indices or hit slots).  Only pixels whose primary ray has a
// real collision ("hits", listed in pixel order) carry G-buffer / reservoir data; for every other pixel the only word
// anybody reads is worldPos.w = 0, which is what keeps the frame's HBM traffic proportional to the
// part of the screen the volume covers.
struct Queues {
  uint32_t* counters;     // [0] candidates [1] hits [2] shadow rays [3] primary queue head [4] shadow queue head [5] coverage mask unusable
  uint8_t* cover;         // per 8x8-pixel screen tile: 1 = some non-empty cell of the grid may project into it (k_cover)
  uint32_t* cand;         // pixels whose primary ray enters the grid window
  float4* cand_ray;       // 2 x float4 per candidate: clipped primary ray {o.xyz, t0}, {d.xyz, t1}
  uint8_t* flag;          // per pixel: 1 = real collision found by k_primary
  uint32_t* block_count;  // hit compaction: per 2048-pixel block count, then exclusive offset
  uint32_t* hit_pix;      // pixel of hit slot s, ascending
  uint32_t* hit_seed;     // RNG state carried between the kernels of the pass
  float* hit_T;           // shadow transmittance estimate (1 when no shadow ray was needed)
  uint32_t* shadow;       // hit slots that need a shadow ray
  float4* shadow_ray;     // 2 x float4 per shadow-queue entry
};

// One peer-memory halo exchange: boundary rows -> the up / down neighbour's halo rows (peer pointers, null = no neighbour).
// The worldPos plane travels whole (its w is the hit marker every consumer tests first); the other planes only where the
// pixel is a hit — a miss pixel's remaining planes are never read, and NVLink stores are what this costs.
struct HaloPush {
  const float4* wp_src; float4* wp_up; float4* wp_down;   // worldPos plane of the frame; copied when copy_wp, always the hit test
  int copy_wp;
  const float4* src[4];
  float4* up_dst[4];
  float4* down_dst[4];
  int nplanes;
  size_t up_src_off, up_dst_off, up_count;         // in float4 elements (= pixels)
  size_t down_src_off, down_dst_off, down_count;
  unsigned* up_flag;                               // "from below" flag in the up neighbour's memory
  unsigned* down_flag;                             // "from above" flag in the down neighbour's memory
  unsigned* serial;                                // this rank's exchange counter (device)
  unsigned* block_counter;
};

// ------------------------------------------------------------------ small vector helpers
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 mul(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 muls(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 divs(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 normalize(V3 a) { return divs(a, sqrtf(dot(a, a))); }
__device__ __forceinline__ float gmax(float a, float b) { return a < b ? b : a; }   // GLSL max
__device__ __forceinline__ float gmin(float a, float b) { return b < a ? b : a; }   // GLSL min
__device__ __forceinline__ float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
__device__ __forceinline__ float gmix(float x, float y, float a) { return x * (1.0f - a) + y * a; }

#define VRS_PI 3.1415926535897932384626433832795f   /* headers/math.glsl:1 */

// ------------------------------------------------------------------ RNG: headers/random.glsl
__device__ __forceinline__ uint32_t lcg(uint32_t& prev) {          // :58-63
  prev = 1664525u * prev + 1013904223u;
  return prev & 0x00FFFFFFu;
}
__device__ __forceinline__ float rnd(uint32_t& seed) {             // :90-99, RAND_LCG
  return float(lcg(seed)) / 16777216.0f;
}
__device__ __forceinline__ uint32_t pixel_seed(uint32_t px, uint32_t py, uint32_t clock, uint32_t pass) {
  // pcg2d (:73-87) of pixel * K, K = clock*8 + pass + 1 replacing int(clockARB()) (restir.rgen:139)
  uint32_t K = clock * 8u + pass + 1u;
  uint32_t x = px * K, y = py * K;
  x = x * 1664525u + 1013904223u;
  y = y * 1664525u + 1013904223u;
  x += y * 1664525u;
  y += x * 1664525u;
  x = x ^ (x >> 16u);
  y = y ^ (y >> 16u);
  x += y * 1664525u;
  y += x * 1664525u;
  x = x ^ (x >> 16u);
  y = y ^ (y >> 16u);
  return x + y;
}

__device__ __forceinline__ float luminance_common(float r, float g, float b) {   // headers/common.glsl:5-7
  return 0.2126f * r + 0.7152f * g + 0.0722f * b;
}
__device__ __forceinline__ float luminance_utils(V3 v) {                          // headers/restirUtils.glsl:6-8
  return dot(v, v3(0.212671f, 0.715160f, 0.072169f));
}

// ------------------------------------------------------------------ Disney BRDF: headers/disneyBRDF.glsl
__device__ __forceinline__ float schlickFresnel(float c) {                       // :6-10
  float m = gclamp(1.0f - c, 0.0f, 1.0f);
  float sm = m * m;
  return sm * sm * m;
}
__device__ __forceinline__ float GTR2(float NdotH, float a) {                    // :13-17
  float a2 = a * a;
  float t = 1.0f + (a2 - 1.0f) * NdotH * NdotH;
  return a2 / (VRS_PI * t * t);
}
__device__ __forceinline__ float smithG_GGX(float NdotV, float alphaG) {         // :19-23
  float a = alphaG * alphaG;
  float b = NdotV * NdotV;
  return 1.0f / (fabsf(NdotV) + gmax(sqrtf(a + b - a * b), 0.0001f));
}
__device__ __forceinline__ float diffuseFactor(float cosIn, float cosOut, float cosInHalf, float roughness, float metallic) {  // :25-33
  float fresnelIn = schlickFresnel(cosIn);
  float fresnelOut = schlickFresnel(cosOut);
  float fd90 = 0.5f + 2.0f * cosInHalf * cosInHalf * roughness;
  float fd = gmix(1.0f, fd90, fresnelIn) * gmix(1.0f, fd90, fresnelOut);
  return fd * (1.0f - metallic) / VRS_PI;
}
__device__ __forceinline__ void specularFactors(float cosIn, float cosOut, float cosHalf, float cosInHalf, float roughness,
                                                float& fresnelInHalf, float& GsDs) {                                          // :47-62
  fresnelInHalf = schlickFresnel(cosInHalf);
  float a = gmax(0.001f, roughness * roughness);
  float Ds = GTR2(cosHalf, a);
  float Gs = smithG_GGX(cosIn, a);
  Gs *= smithG_GGX(cosOut, a);
  GsDs = Gs * Ds;
}
__device__ __forceinline__ float disneyBrdfLuminance(float cosIn, float cosOut, float cosHalf, float cosInHalf, float lum,
                                                     float roughness, float metallic) {                                       // :98-110
  if (cosIn < 0.0f) return 0.0f;
  float diffuse = lum * diffuseFactor(cosIn, cosOut, cosInHalf, roughness, metallic);
  float fih, gsds;
  specularFactors(cosIn, cosOut, cosHalf, cosInHalf, roughness, fih, gsds);
  float specLum = gmix(0.04f, lum, metallic);
  float Fs = gmix(specLum, 1.0f, fih);
  float specular = Fs * gsds;
  return diffuse + specular;
}
__device__ __forceinline__ V3 disneyBrdfColor(float cosIn, float cosOut, float cosHalf, float cosInHalf, V3 albedo,
                                              float roughness, float metallic) {                                              // :86-97
  if (cosIn < 0.0f) return v3(0.0f, 0.0f, 0.0f);
  V3 diffuse = muls(albedo, diffuseFactor(cosIn, cosOut, cosInHalf, roughness, metallic));
  float fih, gsds;
  specularFactors(cosIn, cosOut, cosHalf, cosInHalf, roughness, fih, gsds);
  V3 specColor = v3(gmix(0.04f, albedo.x, metallic), gmix(0.04f, albedo.y, metallic), gmix(0.04f, albedo.z, metallic));
  V3 Fs = v3(gmix(specColor.x, 1.0f, fih), gmix(specColor.y, 1.0f, fih), gmix(specColor.z, 1.0f, fih));
  V3 specular = muls(Fs, gsds);
  return add(diffuse, specular);
}

// ------------------------------------------------------------------ GeometryInfo / Reservoir: structs/restirStructs.glsl
struct GInfo {
  V3 camPos, worldPos, normal;
  float albedo[4];
  float albedoLum, roughness, metallic;
  uint32_t sampleSeed;
};
struct Res {
  uint32_t M, lightIndex; int32_t lightKind; uint32_t sampleSeed;
  float pHat, sumWeights, w;
};
__device__ __forceinline__ Res newReservoir() {                                  // reservoir.glsl:92-100
  Res r; r.M = 0; r.lightIndex = 0; r.lightKind = 0; r.sampleSeed = 0; r.pHat = 0.0f; r.sumWeights = 0.0f; r.w = 0.0f;
  return r;
}
__device__ __forceinline__ Res unpackReservoir(float4 info, float4 weight) {     // reservoir.glsl:4-16
  Res r;
  r.M = __float_as_uint(info.x); r.lightIndex = __float_as_uint(info.y);
  r.lightKind = __float_as_int(info.z); r.sampleSeed = __float_as_uint(info.w);
  r.pHat = weight.x; r.sumWeights = weight.y; r.w = weight.z;
  return r;
}
__device__ __forceinline__ void packReservoir(const Res& r, float4& info, float4& weight) {   // reservoir.glsl:18-28
  info = make_float4(__uint_as_float(r.M), __uint_as_float(r.lightIndex), __int_as_float(r.lightKind), __uint_as_float(r.sampleSeed));
  weight = make_float4(r.pHat, r.sumWeights, r.w, 0.0f);
}

// ------------------------------------------------------------------ p-hat: headers/restirUtils.glsl (point lights)
// The terms of evaluatePHat / disneyBrdf* that depend only on the shading point (wo, cosOut and everything derived
// from them) are computed once per pixel in ShadePre; the per-light remainder performs the same fp32 operations on
// the same values in the same order, so the result is bit-identical to evaluating the reference expression per light.
struct ShadePre {
  V3 wo;
  float cosOut, fresnelOut, smithOut, a;
};
__device__ __forceinline__ ShadePre shade_pre(const GInfo& g) {
  ShadePre p;
  p.wo = normalize(sub(g.camPos, g.worldPos));                        // restirUtils.glsl:62
  p.cosOut = dot(g.normal, p.wo);                                     // :65
  p.fresnelOut = schlickFresnel(p.cosOut);                            // disneyBRDF.glsl:28
  p.a = gmax(0.001f, g.roughness * g.roughness);                      // disneyBRDF.glsl:53
  p.smithOut = smithG_GGX(p.cosOut, p.a);                             // disneyBRDF.glsl:59
  return p;
}
struct PHatGeom { float cosIn, cosHalf, cosInHalf, geometry; bool back; };
__device__ __forceinline__ PHatGeom phat_geometry(V3 lightPos, const GInfo& g, const ShadePre& pre) {  // :39-71
  PHatGeom o;
  V3 wi = sub(lightPos, g.worldPos);
  o.back = dot(wi, g.normal) < 0.0f;
  float sqrDist = dot(wi, wi);
  wi = divs(wi, sqrtf(sqrDist));
  o.cosIn = dot(g.normal, wi);
  V3 halfVec = normalize(add(wi, pre.wo));
  o.cosHalf = dot(g.normal, halfVec);
  o.cosInHalf = dot(wi, halfVec);
  o.geometry = 1.0f * o.cosIn / sqrDist;
  return o;
}
// disneyBrdfDiffuseFactor (:25-33) and disneyBrdfSpecularFactors (:47-62) with the per-pixel terms taken from ShadePre
__device__ __forceinline__ float diffuseFactorPre(const PHatGeom& q, const ShadePre& pre, float roughness, float metallic) {
  float fresnelIn = schlickFresnel(q.cosIn);
  float fd90 = 0.5f + 2.0f * q.cosInHalf * q.cosInHalf * roughness;
  float fd = gmix(1.0f, fd90, fresnelIn) * gmix(1.0f, fd90, pre.fresnelOut);
  return fd * (1.0f - metallic) / VRS_PI;
}
__device__ __forceinline__ void specularFactorsPre(const PHatGeom& q, const ShadePre& pre, float& fresnelInHalf, float& GsDs) {
  fresnelInHalf = schlickFresnel(q.cosInHalf);
  float Ds = GTR2(q.cosHalf, pre.a);
  float Gs = smithG_GGX(q.cosIn, pre.a);
  Gs *= pre.smithOut;
  GsDs = Gs * Ds;
}
// evaluatePHat (:36-78) for a point light already fetched: position and luminance (PointLight.emission_luminance.w)
__device__ __forceinline__ float evaluatePHatLight(V3 lightPos, float lum, const GInfo& g, const ShadePre& pre) {
  PHatGeom q = phat_geometry(lightPos, g, pre);
  if (q.back) return 0.0f;
  float brdf = 0.0f;                                                                  // disneyBrdfLuminance :98-110
  if (!(q.cosIn < 0.0f)) {
    float diffuse = g.albedoLum * diffuseFactorPre(q, pre, g.roughness, g.metallic);
    float fih, gsds;
    specularFactorsPre(q, pre, fih, gsds);
    float specLum = gmix(0.04f, g.albedoLum, g.metallic);
    float Fs = gmix(specLum, 1.0f, fih);
    brdf = diffuse + Fs * gsds;
  }
  return lum * brdf * q.geometry;
}
__device__ __forceinline__ float evaluatePHat(const LightsDev& L, uint32_t lightIdx, const GInfo& g, const ShadePre& pre) {   // :36-78
  float4 lp = __ldg(&L.lights[2 * lightIdx]);
  float4 le = __ldg(&L.lights[2 * lightIdx + 1]);
  return evaluatePHatLight(v3(lp.x, lp.y, lp.z), le.w, g, pre);
}
__device__ __forceinline__ float evaluatePHat(const LightsDev& L, uint32_t lightIdx, const GInfo& g) {
  return evaluatePHat(L, lightIdx, g, shade_pre(g));
}
__device__ __forceinline__ V3 evaluatePHatFull(const LightsDev& L, uint32_t lightIdx, const GInfo& g) {  // :80-122
  float4 lp = __ldg(&L.lights[2 * lightIdx]);
  float4 le = __ldg(&L.lights[2 * lightIdx + 1]);
  ShadePre pre = shade_pre(g);
  PHatGeom q = phat_geometry(v3(lp.x, lp.y, lp.z), g, pre);
  if (q.back) return v3(0.0f, 0.0f, 0.0f);
  V3 brdf = v3(0.0f, 0.0f, 0.0f);                                                     // disneyBrdfColor :86-97
  if (!(q.cosIn < 0.0f)) {
    V3 albedo = v3(g.albedo[0], g.albedo[1], g.albedo[2]);
    V3 diffuse = muls(albedo, diffuseFactorPre(q, pre, g.roughness, g.metallic));
    float fih, gsds;
    specularFactorsPre(q, pre, fih, gsds);
    V3 specColor = v3(gmix(0.04f, albedo.x, g.metallic), gmix(0.04f, albedo.y, g.metallic), gmix(0.04f, albedo.z, g.metallic));
    V3 Fs = v3(gmix(specColor.x, 1.0f, fih), gmix(specColor.y, 1.0f, fih), gmix(specColor.z, 1.0f, fih));
    brdf = add(diffuse, muls(Fs, gsds));
  }
  return muls(mul(v3(le.x, le.y, le.z), brdf), q.geometry);
}

// ------------------------------------------------------------------ reservoir ops: headers/reservoir.glsl
__device__ __forceinline__ void updateReservoir(Res& res, uint32_t lightIdx, int32_t lightKind, float weight, float pHat, float w,
                                                uint32_t& seed, uint32_t sampleSeed) {            // :30-43
  res.sumWeights += weight;
  float replacePossibility = weight / res.sumWeights;
  if (rnd(seed) < replacePossibility) {
    res.lightIndex = lightIdx; res.lightKind = lightKind; res.pHat = pHat; res.w = w; res.sampleSeed = sampleSeed;
  }
}
__device__ __forceinline__ void addSampleToReservoir(const LightsDev& L, Res& res, uint32_t lightIdx, int32_t lightKind,
                                                     float lightPdf, const GInfo& g, const ShadePre& pre, uint32_t& seed) {   // :45-54
  float pHat = evaluatePHat(L, lightIdx, g, pre);
  float weight = pHat / lightPdf;
  res.M += 1;
  float w = (res.sumWeights + weight) / (float(res.M) * pHat);
  updateReservoir(res, lightIdx, lightKind, weight, pHat, w, seed, g.sampleSeed);
}
__device__ __forceinline__ void combineReservoirsGeom(const LightsDev& L, Res& self, const Res& other, const GInfo& g,
                                                      const GInfo& og, uint32_t& seed) {          // :56-76
  uint32_t Z = self.M;
  self.M += other.M;
  float pHat = evaluatePHat(L, other.lightIndex, g);
  float weight = pHat * other.w * float(other.M);
  if (weight > 0.0f) updateReservoir(self, other.lightIndex, other.lightKind, weight, pHat, other.w, seed, other.sampleSeed);
  pHat = evaluatePHat(L, self.lightIndex, og);
  if (pHat > 0.0f) Z += other.M;
  if (self.w > 0.0f) self.w = self.sumWeights / (float(Z) * self.pHat);
}

// aliasTableSample (restir.rgen:97-110) in two halves so that callers can overlap the cell fetch with other work
__device__ __forceinline__ uint32_t aliasColumn(const LightsDev& L, float r1) {
  uint32_t col = uint32_t(float(L.ntable) * r1);
  uint32_t last = uint32_t(L.ntable - 1);
  if (last < col) col = last;
  return col;
}
__device__ __forceinline__ void aliasPick(float4 c, uint32_t col, float r2, uint32_t& index, float& prob) {
  if (c.y > r2) { index = col; prob = c.z; } else { index = (uint32_t)__float_as_int(c.x); prob = c.w; }
}
__device__ __forceinline__ void aliasTableSample(const LightsDev& L, float r1, float r2, uint32_t& index, float& prob) {
  const uint32_t col = aliasColumn(L, r1);
  aliasPick(__ldg(&L.alias[col]), col, r2, index, prob);
}

// ------------------------------------------------------------------ sparse grid lookups (flattened Tree_float_5_4_3)
// Returns leaf index (>= 0) or ~tile (< 0) for the 8^3 cell containing absolute voxel (x,y,z).
__device__ __forceinline__ int cell_lookup(const GridDev& G, int x, int y, int z) {
  int kx = x & ~4095, ky = y & ~4095, kz = z & ~4095;
  int n5 = -1;
  for (int r = 0; r < G.nroot; ++r) {
    int4 e = __ldg(&G.root[r]);
    if (e.x == kx && e.y == ky && e.z == kz) { n5 = e.w; break; }
  }
  if (n5 < 0) return n5;                                      // ~tile (missing root key = ~0 = background)
  int s5 = (((x & 4095) >> 7) << 10) | (((y & 4095) >> 7) << 5) | ((z & 4095) >> 7);
  int n4 = __ldg(&G.i5[(size_t)n5 * 32768 + s5]);
  if (n4 < 0) return n4;
  int s4 = (((x & 127) >> 3) << 8) | (((y & 127) >> 3) << 4) | ((z & 127) >> 3);
  return __ldg(&G.i4[(size_t)n4 * 4096 + s4]);
}
__device__ __forceinline__ float density_at(const GridDev& G, int i, int j, int k) {
  int x = i - G.vmin[0], y = j - G.vmin[1], z = k - G.vmin[2];
  if (x < 0 || y < 0 || z < 0 || x >= G.vdim[0] || y >= G.vdim[1] || z >= G.vdim[2]) return G.bg_density;
  int c = cell_lookup(G, i, j, k);
  if (c < 0) return __ldg(&G.tile_density[~c]);
  return __ldg(&G.atlas[(size_t)c * 512 + (((i & 7) << 6) | ((j & 7) << 3) | (k & 7))]);
}

__device__ __forceinline__ float neglog1m(float u) {   // -ln(1-u) with + - * / only
  float x = 1.0f - u;
  uint32_t bits = __float_as_uint(x);
  int e = int(bits >> 23) - 127;
  float m = __uint_as_float((bits & 0x007FFFFFu) | 0x3F800000u);
  if (m > 1.41421356f) { m = m * 0.5f; e = e + 1; }
  float f = m - 1.0f;
  float s = f / (2.0f + f);
  float z = s * s;
  float p = 0.18181818f;
  p = p * z + 0.22222222f;
  p = p * z + 0.28571429f;
  p = p * z + 0.4f;
  p = p * z + 0.66666667f;
  float lnm = 2.0f * s + s * z * p;
  return -(float(e) * 0.69314718f + lnm);
}

struct TrackResult { bool hit; float t; int vox[3]; float T; uint32_t ntent, ncells; };

// A ray clipped to the grid window, in voxel space (o + d t), ready to march: what k_classify / k_ris hand to the
// persistent raymarch kernels through the work queues (2 x float4 per ray).
struct RaySeg { float o[3], d[3], t0, t1; };

// Voxel-space ray + slab test against the window.  false = nothing to march.
__device__ __forceinline__ bool clip_ray(const GridDev& G, V3 org, V3 dir, float tmin, float tmax, RaySeg& r) {
  r.o[0] = (org.x - G.B[0]) * G.invA + 0.5f; r.o[1] = (org.y - G.B[1]) * G.invA + 0.5f; r.o[2] = (org.z - G.B[2]) * G.invA + 0.5f;
  r.d[0] = dir.x * G.invA; r.d[1] = dir.y * G.invA; r.d[2] = dir.z * G.invA;
  float t0 = tmin, t1 = tmax;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float lo = float(G.vmin[a]), hi = float(G.vmin[a] + G.vdim[a]);
    if (r.d[a] == 0.0f) {
      if (r.o[a] < lo || !(r.o[a] < hi)) return false;
    } else {
      float inv = 1.0f / r.d[a];
      float ta = (lo - r.o[a]) * inv, tb = (hi - r.o[a]) * inv;
      float tnr = gmin(ta, tb), tf = gmax(ta, tb);
      t0 = gmax(t0, tnr); t1 = gmin(t1, tf);
    }
  }
  r.t0 = t0; r.t1 = t1;
  return t0 < t1;
}

enum { RAY_DONE = 0, RAY_SKIP = 1, RAY_COLLIDE = 2 };

// Residual-optical-depth tracking over the dense cell directory, as two kinds of unit step so that a warp can batch
// them (march_loop below) — the same arithmetic and RNG draws, in the same order, as the oracle's nested loops:
//   cell_step()    visit one 8^3 cell: one directory load; the carried sample tau either crosses the cell
//                  (tau -= (tcell - t) * mu) or runs out inside it                       (cheap, no RNG)
//   collide_step() the tentative collision where tau ran out: brick load, accept test / T update, fresh tau
template <int MODE>
struct Ray {
  float o[3], d[3], tn[3], dt[3];
  float t, t1, tcell, mu_d, mu, tau, T;
  int c[3], sgn[3];
  int cell, axis, leaf;
  uint32_t ntent, ncells;
  bool last, hit;
  int vox[3];

  __device__ __forceinline__ void start(const GridDev& G, const RaySeg& r, uint32_t& seed) {
    hit = false; T = 1.0f; ntent = 0; ncells = 0; vox[0] = vox[1] = vox[2] = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      o[a] = r.o[a]; d[a] = r.d[a];
      float q = o[a] + d[a] * r.t0;
      int v = int(floorf(q)) - G.vmin[a];
      int ci = v >> 3;
      if (ci < 0) ci = 0;
      if (ci > G.cdim[a] - 1) ci = G.cdim[a] - 1;
      c[a] = ci;
      sgn[a] = d[a] > 0.0f ? 1 : -1;
      if (d[a] == 0.0f) { tn[a] = __int_as_float(0x7f800000); dt[a] = 0.0f; }
      else {
        float inv = 1.0f / d[a];
        float bound = float(G.vmin[a] + (c[a] + (d[a] > 0.0f ? 1 : 0)) * 8);
        tn[a] = (bound - o[a]) * inv;
        dt[a] = fabsf(inv) * 8.0f;
      }
    }
    cell = (c[2] * G.cdim[1] + c[1]) * G.cdim[0] + c[0];
    t = r.t0; t1 = r.t1;
    tau = neglog1m(rnd(seed));
  }

  // The unit steps are split so that the scheduler (march_loop) runs ONE copy of the common tail for whichever kind
  // of step the warp executes, and so that empty and non-empty cells take the same path (no divergence between them):
  //   cell_head()    enter the current cell: exit time / axis, one directory load (majorant, 0 = empty)
  //   collide()      the tentative collision where tau ran out: brick load, accept test / T update, fresh tau
  //   tail()         does the carried sample run out before the cell's exit?  else consume it and step to the next cell
  __device__ __forceinline__ void cell_head(const GridDev& G) {
    ncells += 1;
    axis = 0; tcell = tn[0];
    if (tn[1] < tcell) { tcell = tn[1]; axis = 1; }
    if (tn[2] < tcell) { tcell = tn[2]; axis = 2; }
    last = false;
    if (!(tcell < t1)) { tcell = t1; last = true; }
    const float2 e = __ldg(&G.dir[cell]);                    // `cell` = (c[2] * cdim[1] + c[1]) * cdim[0] + c[0], kept incrementally
    mu_d = e.x; leaf = __float_as_int(e.y);
    mu = mu_d > 0.0f ? mu_d * G.density_scale : 0.0f;        // empty cell: seg = 0 below, tau is untouched (x - 0 = x)
  }

  __device__ __forceinline__ int tail(const GridDev& G) {
    const float seg = (tcell - t) * mu;
    if (tau < seg) return RAY_COLLIDE;
    tau = tau - seg;
    t = tcell;
    if (last) return RAY_DONE;
    // leave the cell along `axis`, branch-free: lanes of a warp leave along different axes, and a three-way branch would run
    // its arms one after the other with a third of the lanes each (ncu: 1.7 - 5.7 lanes per instruction in those arms)
    const int m0 = -(int)(axis == 0), m1 = -(int)(axis == 1), m2 = -(int)(axis == 2);
    c[0] += sgn[0] & m0; c[1] += sgn[1] & m1; c[2] += sgn[2] & m2;
    cell += (sgn[0] & m0) + ((sgn[1] * G.cdim[0]) & m1) + ((sgn[2] * (G.cdim[0] * G.cdim[1])) & m2);
    const float n0 = tn[0] + dt[0], n1 = tn[1] + dt[1], n2 = tn[2] + dt[2];
    tn[0] = axis == 0 ? n0 : tn[0]; tn[1] = axis == 1 ? n1 : tn[1]; tn[2] = axis == 2 ? n2 : tn[2];
    const bool out = (unsigned)c[0] >= (unsigned)G.cdim[0] || (unsigned)c[1] >= (unsigned)G.cdim[1] || (unsigned)c[2] >= (unsigned)G.cdim[2];   // only the coordinate that moved can have left the window
    return out ? RAY_DONE : RAY_SKIP;
  }

  // false = the ray ended at this collision (primary: real collision found; shadow: opaque)
  __device__ __forceinline__ bool collide(const GridDev& G, uint32_t& seed) {
    t = t + tau / mu;
    ntent += 1;
    int vlo0 = G.vmin[0] + c[0] * 8, vlo1 = G.vmin[1] + c[1] * 8, vlo2 = G.vmin[2] + c[2] * 8;
    int vx = int(floorf(o[0] + d[0] * t)), vy = int(floorf(o[1] + d[1] * t)), vz = int(floorf(o[2] + d[2] * t));
    vx = vx < vlo0 ? vlo0 : (vx > vlo0 + 7 ? vlo0 + 7 : vx);
    vy = vy < vlo1 ? vlo1 : (vy > vlo1 + 7 ? vlo1 + 7 : vy);
    vz = vz < vlo2 ? vlo2 : (vz > vlo2 + 7 ? vlo2 + 7 : vz);
    float dens = leaf < 0 ? mu_d : __ldg(&G.atlas[(size_t)leaf * 512 + (((vx & 7) << 6) | ((vy & 7) << 3) | (vz & 7))]);
    if (MODE == 0) {
      float u2 = rnd(seed);
      if (u2 * mu_d < dens) { hit = true; vox[0] = vx; vox[1] = vy; vox[2] = vz; return false; }
    } else {
      // (dens = 0: 0 / mu_d = 0, T * 1 = T, nothing changes — and a zero numerator would send the IEEE division to its slow path)
      if (dens != 0.0f) {
        T = T * (1.0f - dens / mu_d);
        if (!(T > 1e-5f)) { T = 0.0f; return false; }        // opaque for every practical purpose: stop marching
      }
    }
    tau = neglog1m(rnd(seed));
    return true;
  }

  __device__ __forceinline__ int cell_step(const GridDev& G) { cell_head(G); return tail(G); }
  __device__ __forceinline__ int collide_step(const GridDev& G, uint32_t& seed) { return collide(G, seed) ? tail(G) : RAY_DONE; }
};

// Nested-loop form for one-ray-per-thread callers (k_shade's optional final visibility): runs a Ray to completion.
template <int MODE>
__device__ __forceinline__ TrackResult track(const GridDev& G, V3 org, V3 dir, float tmin, float tmax, uint32_t& seed) {
  TrackResult R; R.hit = false; R.t = 0.0f; R.vox[0] = R.vox[1] = R.vox[2] = 0; R.T = 1.0f; R.ntent = 0; R.ncells = 0;
  RaySeg seg;
  if (!clip_ray(G, org, dir, tmin, tmax, seg)) return R;
  Ray<MODE> ray;
  ray.start(G, seg, seed);
  int st = RAY_SKIP;
  while (st != RAY_DONE) st = st == RAY_SKIP ? ray.cell_step(G) : ray.collide_step(G, seed);
  R.hit = ray.hit; R.t = ray.t; R.vox[0] = ray.vox[0]; R.vox[1] = ray.vox[1]; R.vox[2] = ray.vox[2]; R.T = ray.T;
  R.ntent = ray.ntent; R.ncells = ray.ncells;
  return R;
}

// warp-cooperative work fetch: lanes with `want` take consecutive indices from *head (one atomic per warp)
__device__ __forceinline__ uint32_t warp_fetch(uint32_t* head, bool want) {
  const unsigned full = 0xffffffffu;
  unsigned b = __ballot_sync(full, want);
  if (b == 0) return 0xFFFFFFFFu;
  int lane = threadIdx.x & 31;
  int leader = __ffs(b) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(head, (uint32_t)__popc(b));
  base = __shfl_sync(full, base, leader);
  return want ? base + (uint32_t)__popc(b & ((1u << lane) - 1u)) : 0xFFFFFFFFu;
}
// warp-aggregated append from divergent code: returns this lane's slot
__device__ __forceinline__ uint32_t warp_append(uint32_t* counter) {
  unsigned m = __activemask();
  int lane = threadIdx.x & 31;
  int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(m, base, leader);
  return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}

// Warp-level scheduler of the persistent raymarch kernels.  Each lane owns one ray (Job supplies fetch / retire).
// Per iteration the warp executes ONE kind of unit step, the kind more lanes are waiting for, so divergent lanes are
// batched instead of serialised; lanes whose ray retired are refilled from the job queue once enough are idle.
// `lane_limit` (<= 32) caps how many lanes of a warp hold rays: when a launch has few rays (a narrow band of a multi-GPU
// frame) they are spread over more warps, so that the per-ray serial chain is hidden by other warps instead of
// setting the kernel's duration.  Which ray runs in which lane never changes a result.
template <int MODE, class Job>
__device__ __forceinline__ void march_loop(const GridDev& G, Job& job, uint32_t* head, uint32_t njobs, int refill_min_idle, int cells_per_decision,
                                           int lane_limit) {
  const unsigned full = 0xffffffffu;
  if (lane_limit > 32) lane_limit = 32;
  const bool eligible = (int)(threadIdx.x & 31) < lane_limit;
  if (refill_min_idle > (lane_limit + 1) / 2) refill_min_idle = (lane_limit + 1) / 2;
  Ray<MODE> ray;
  int st = RAY_DONE;
  uint32_t seed = 0;
  bool queue_empty = false;
  for (;;) {
    const unsigned bs = __ballot_sync(full, st == RAY_SKIP), bc = __ballot_sync(full, st == RAY_COLLIDE);
    const int nS = __popc(bs), nC = __popc(bc);
    if (!queue_empty && (lane_limit - nS - nC >= refill_min_idle || (bs | bc) == 0)) {   // only lanes below the limit ever hold rays
      const bool want = st == RAY_DONE && eligible;
      const uint32_t j = warp_fetch(head, want);
      if (want && j < njobs) st = job.fetch(G, j, ray, seed) ? RAY_SKIP : RAY_DONE;
      queue_empty = __any_sync(full, want && j >= njobs);
      continue;
    }
    if ((bs | bc) == 0) break;
    const bool cells = nS >= nC;
#pragma unroll 1
    for (int r = 0; r < cells_per_decision; ++r) {         // a few cell visits per scheduling decision (they are cheap)
      bool run = false;
      if (cells) {
        if (st == RAY_SKIP) { ray.cell_head(G); run = true; }
      } else if (st == RAY_COLLIDE) {
        run = ray.collide(G, seed);
        if (!run) { st = RAY_DONE; job.retire(G, ray, seed); }
      }
      if (run) { st = ray.tail(G); if (st == RAY_DONE) job.retire(G, ray, seed); }
      if (!cells) break;
    }
  }
}
// lanes per warp for `njobs` rays on a grid of `nwarps` persistent warps: fill `target_warps` warps before widening
__device__ __forceinline__ int lanes_for(uint32_t njobs, uint32_t nwarps, uint32_t target_warps) {
  if (target_warps > nwarps) target_warps = nwarps;
  if (target_warps == 0) return 32;
  uint32_t l = (njobs + target_warps - 1) / target_warps;
  return l < 4u ? 4 : (l > 32u ? 32 : (int)l);
}
// Pixels per group for the cooperative kernels, whose time per group is proportional to its size: small launches are
// spread over `target_warps` warps (as lanes_for); large ones get equal groups, so that every warp of the resident grid
// runs the same number of equally sized groups (no partial last round).
__device__ __forceinline__ uint32_t group_size_for(uint32_t nitems, uint32_t nwarps, uint32_t target_warps, uint32_t max_group) {
  if (target_warps > nwarps) target_warps = nwarps;
  if (target_warps == 0 || nwarps == 0) return max_group;
  uint32_t g = (nitems + target_warps - 1) / target_warps;
  if (g > max_group) {
    const uint32_t per_warp = (nitems + nwarps - 1) / nwarps;
    const uint32_t rounds = (per_warp + max_group - 1) / max_group;
    g = rounds ? (per_warp + rounds - 1) / rounds : max_group;
  }
  return g < 4u ? 4u : (g > max_group ? max_group : g);
}

__device__ __forceinline__ float ratio_track(const GridDev& G, V3 P, V3 L, uint32_t& seed) {
  V3 dir = sub(L, P);
  float dist = sqrtf(dot(dir, dir));
  if (!(dist > 0.0f)) return 1.0f;
  dir = divs(dir, dist);
  TrackResult r = track<1>(G, P, dir, 0.0f, dist, seed);
  return r.T;
}

// voxel material: vdb/vdb.cpp:814-821 + Renderer.cpp:1494-1500 (nvmath::normalize = multiply by reciprocal norm)
__device__ __forceinline__ float4 voxel_albedo(float v) {
  float n3 = sqrtf(100.0f * 100.0f + 100.0f * 100.0f + 100.0f * 100.0f);
  float s = 100.0f * (1.0f / n3);
  float c = s * v * 1000.0f;
  float n4 = sqrtf(c * c + c * c + c * c + 1.0f * 1.0f);
  float inv = n4 > 10e-6f ? 1.0f / n4 : 0.0f;
  return make_float4(c * inv, c * inv, c * inv, 1.0f * inv);
}

__device__ __forceinline__ void mat_vec(const float* m, float x, float y, float z, float w, float* o) {   // nvmath.inl:483-492
#pragma unroll
  for (int r = 0; r < 4; ++r) o[r] = m[0 + r] * x + m[4 + r] * y + m[8 + r] * z + m[12 + r] * w;
}

// GeometryInfo of a hit pixel from the G-buffer planes (spatialReuse.comp:65-74 / restir_post.frag:60-67).  The matProps plane of
// a volume hit always holds {roughness, metallic, 1, 1} of the grid, so the two constants come from the frame parameters
// instead of a fourth 16-byte gather per pixel (the plane itself is still written for readback).
__device__ __forceinline__ GInfo ginfo_from_planes(const Planes& gb, size_t idx, const FrameParams& F) {
  const float* camPos = F.camPos;
  float4 p = gb.worldPos[idx], a = gb.albedo[idx], n = gb.normal[idx];
  const float2 m = make_float2(F.roughness, F.metallic);
  GInfo g;
  g.albedo[0] = a.x; g.albedo[1] = a.y; g.albedo[2] = a.z; g.albedo[3] = a.w;
  g.normal = v3(n.x, n.y, n.z);
  g.worldPos = v3(p.x, p.y, p.z);
  g.roughness = m.x; g.metallic = m.y;
  g.albedoLum = luminance_common(a.x, a.y, a.z);
  g.camPos = v3(camPos[0], camPos[1], camPos[2]);
  g.sampleSeed = 0;
  return g;
}

}  // namespace vrs
