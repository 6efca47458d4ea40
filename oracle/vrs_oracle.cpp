// TEST INFRASTRUCTURE ONLY — see vrs_oracle.h for the usage rules and the parity-pinning statement.
// Scalar restatement of the Volume-ReSTIR hot path.  Citations are relative to /root/reference/.
#include "vrs_oracle.h"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <queue>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct V3 { float x, y, z; };
inline V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
inline V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 mul(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 muls(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 divs(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 normalize(V3 a) { return divs(a, sqrtf(dot(a, a))); }       // GLSL normalize fixed as v / sqrt(dot)
inline float gmax(float a, float b) { return a < b ? b : a; }          // GLSL max: y if x < y
inline float gmin(float a, float b) { return b < a ? b : a; }          // GLSL min: y if y < x
inline float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
inline float gmix(float x, float y, float a) { return x * (1.0f - a) + y * a; }   // GLSL mix
inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

const float kPi = 3.1415926535897932384626433832795f;   // headers/math.glsl:1 (fp32 literal in GLSL)

// ---------------------------------------------------------------- RNG: headers/random.glsl
inline uint32_t lcg(uint32_t& prev) {            // random.glsl:58-63
  prev = 1664525u * prev + 1013904223u;
  return prev & 0x00FFFFFFu;
}
inline float rnd(uint32_t& seed) {               // random.glsl:90-99 (RAND_METHOD == RAND_LCG, :37)
  return float(lcg(seed)) / float(0x01000000);
}
inline void pcg2d(uint32_t& x, uint32_t& y) {    // random.glsl:73-87
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
}
// restir.rgen:139-140 / spatialReuse.comp:58-59 with int(clockARB()) replaced by
// K = clock * 8 + pass + 1.
inline uint32_t pixel_seed(uint32_t px, uint32_t py, uint32_t clock, uint32_t pass) {
  uint32_t K = clock * 8u + pass + 1u;
  uint32_t x = px * K, y = py * K;
  pcg2d(x, y);
  return x + y;
}

// ---------------------------------------------------------------- luminance (two variants)
inline float luminance_common(float r, float g, float b) {     // headers/common.glsl:5-7
  return 0.2126f * r + 0.7152f * g + 0.0722f * b;
}
inline float luminance_utils(V3 v) {                             // headers/restirUtils.glsl:6-8
  return dot(v, v3(0.212671f, 0.715160f, 0.072169f));
}

// ---------------------------------------------------------------- Disney BRDF: headers/disneyBRDF.glsl
inline float schlickFresnel(float c) {                           // :6-10
  float m = gclamp(1.0f - c, 0.0f, 1.0f);
  float sm = m * m;
  return sm * sm * m;
}
inline float GTR2(float NdotH, float a) {                        // :13-17
  float a2 = a * a;
  float t = 1.0f + (a2 - 1.0f) * NdotH * NdotH;
  return a2 / (kPi * t * t);
}
inline float smithG_GGX(float NdotV, float alphaG) {             // :19-23
  float a = alphaG * alphaG;
  float b = NdotV * NdotV;
  return 1.0f / (fabsf(NdotV) + gmax(sqrtf(a + b - a * b), 0.0001f));
}
inline float diffuseFactor(float cosIn, float cosOut, float cosInHalf, float roughness, float metallic) {  // :25-33
  float fresnelIn = schlickFresnel(cosIn);
  float fresnelOut = schlickFresnel(cosOut);
  float fd90 = 0.5f + 2.0f * cosInHalf * cosInHalf * roughness;
  float fd = gmix(1.0f, fd90, fresnelIn) * gmix(1.0f, fd90, fresnelOut);
  return fd * (1.0f - metallic) / kPi;
}
inline void specularFactors(float cosIn, float cosOut, float cosHalf, float cosInHalf, float roughness,
                            float& fresnelInHalf, float& GsDs) {                                          // :47-62
  fresnelInHalf = schlickFresnel(cosInHalf);
  float a = gmax(0.001f, roughness * roughness);   // pow(roughness, 2.0), :53
  float Ds = GTR2(cosHalf, a);
  float Gs = smithG_GGX(cosIn, a);
  Gs *= smithG_GGX(cosOut, a);
  GsDs = Gs * Ds;
}
inline float disneyBrdfLuminance(float cosIn, float cosOut, float cosHalf, float cosInHalf, float lum,
                                 float roughness, float metallic) {                                       // :98-110
  if (cosIn < 0.0f) return 0.0f;
  float diffuse = lum * diffuseFactor(cosIn, cosOut, cosInHalf, roughness, metallic);                     // :40-45
  float fih, gsds;
  specularFactors(cosIn, cosOut, cosHalf, cosInHalf, roughness, fih, gsds);
  float specLum = gmix(0.04f, lum, metallic);                                                             // :74-85
  float Fs = gmix(specLum, 1.0f, fih);
  float specular = Fs * gsds;
  return diffuse + specular;
}
inline V3 disneyBrdfColor(float cosIn, float cosOut, float cosHalf, float cosInHalf, V3 albedo, float roughness,
                          float metallic) {                                                               // :86-97
  if (cosIn < 0.0f) return v3(0.0f, 0.0f, 0.0f);
  V3 diffuse = muls(albedo, diffuseFactor(cosIn, cosOut, cosInHalf, roughness, metallic));                // :34-38
  float fih, gsds;
  specularFactors(cosIn, cosOut, cosHalf, cosInHalf, roughness, fih, gsds);
  V3 specColor = v3(gmix(0.04f, albedo.x, metallic), gmix(0.04f, albedo.y, metallic), gmix(0.04f, albedo.z, metallic));
  V3 Fs = v3(gmix(specColor.x, 1.0f, fih), gmix(specColor.y, 1.0f, fih), gmix(specColor.z, 1.0f, fih));   // :63-73
  V3 specular = muls(Fs, gsds);
  return add(diffuse, specular);
}

// ---------------------------------------------------------------- GeometryInfo / Reservoir: structs/restirStructs.glsl
struct GInfo {       // :1-11 (emissive is never read on this path)
  V3 camPos, worldPos, normal;
  float albedo[4];
  float albedoLum, roughness, metallic;
  uint32_t sampleSeed;
};
struct Res {         // :13-23 ; lightPos/info are not persisted (reservoir.glsl:18-28) — lightPos is re-derived from lightIndex
  uint32_t M, lightIndex; int32_t lightKind; uint32_t sampleSeed;
  float pHat, sumWeights, w;
};
inline Res newReservoir() {   // reservoir.glsl:92-100 (+ defined values for the fields GLSL leaves undefined)
  Res r; r.M = 0; r.lightIndex = 0; r.lightKind = 0; r.sampleSeed = 0; r.pHat = 0.0f; r.sumWeights = 0.0f; r.w = 0.0f;
  return r;
}
inline GInfo ginfo_from16(const float* g) {
  GInfo o; o.camPos = v3(g[0], g[1], g[2]); o.worldPos = v3(g[3], g[4], g[5]); o.normal = v3(g[6], g[7], g[8]);
  o.albedo[0] = g[9]; o.albedo[1] = g[10]; o.albedo[2] = g[11]; o.albedo[3] = g[12];
  o.albedoLum = g[13]; o.roughness = g[14]; o.metallic = g[15]; o.sampleSeed = 0;
  return o;
}
inline Res res_from8(const uint32_t* r) {
  Res o; o.M = r[0]; o.lightIndex = r[1]; o.lightKind = (int32_t)r[2]; o.sampleSeed = r[3];
  o.pHat = u2f(r[4]); o.sumWeights = u2f(r[5]); o.w = u2f(r[6]);
  return o;
}
inline void res_to8(const Res& o, uint32_t* r) {     // packReservoirStruct, reservoir.glsl:18-28
  r[0] = o.M; r[1] = o.lightIndex; r[2] = (uint32_t)o.lightKind; r[3] = o.sampleSeed;
  r[4] = f2u(o.pHat); r[5] = f2u(o.sumWeights); r[6] = f2u(o.w); r[7] = 0;
}

// ---------------------------------------------------------------- p-hat: headers/restirUtils.glsl (point lights)
inline float evaluatePHat(const orc_point_light* lights, uint32_t lightIdx, const GInfo& g) {   // :36-78
  const orc_point_light& L = lights[lightIdx];
  V3 wi = sub(v3(L.pos[0], L.pos[1], L.pos[2]), g.worldPos);
  float emissionLum = L.emission_luminance[3];
  float LdotN = 1.0f;
  if (dot(wi, g.normal) < 0.0f) return 0.0f;
  float sqrDist = dot(wi, wi);
  wi = divs(wi, sqrtf(sqrDist));
  V3 wo = normalize(sub(g.camPos, g.worldPos));
  float cosIn = dot(g.normal, wi);
  float cosOut = dot(g.normal, wo);
  V3 halfVec = normalize(add(wi, wo));
  float cosHalf = dot(g.normal, halfVec);
  float cosInHalf = dot(wi, halfVec);
  float geometry = LdotN * cosIn / sqrDist;
  return emissionLum * disneyBrdfLuminance(cosIn, cosOut, cosHalf, cosInHalf, g.albedoLum, g.roughness, g.metallic) * geometry;
}
inline V3 evaluatePHatFull(const orc_point_light* lights, uint32_t lightIdx, const GInfo& g) {   // :80-122
  const orc_point_light& L = lights[lightIdx];
  V3 wi = sub(v3(L.pos[0], L.pos[1], L.pos[2]), g.worldPos);
  V3 emission = v3(L.emission_luminance[0], L.emission_luminance[1], L.emission_luminance[2]);
  float LdotN = 1.0f;
  if (dot(wi, g.normal) < 0.0f) return v3(0.0f, 0.0f, 0.0f);
  float sqrDist = dot(wi, wi);
  wi = divs(wi, sqrtf(sqrDist));
  V3 wo = normalize(sub(g.camPos, g.worldPos));
  float cosIn = dot(g.normal, wi);
  float cosOut = dot(g.normal, wo);
  V3 halfVec = normalize(add(wi, wo));
  float cosHalf = dot(g.normal, halfVec);
  float cosInHalf = dot(wi, halfVec);
  float geometry = LdotN * cosIn / sqrDist;
  V3 brdf = disneyBrdfColor(cosIn, cosOut, cosHalf, cosInHalf, v3(g.albedo[0], g.albedo[1], g.albedo[2]), g.roughness, g.metallic);
  return muls(mul(emission, brdf), geometry);
}

// ---------------------------------------------------------------- reservoir ops: headers/reservoir.glsl
inline void updateReservoir(Res& res, uint32_t lightIdx, int32_t lightKind, float weight, float pHat, float w,
                            uint32_t& seed, uint32_t sampleSeed) {                                // :30-43
  res.sumWeights += weight;
  float replacePossibility = weight / res.sumWeights;
  if (rnd(seed) < replacePossibility) {
    res.lightIndex = lightIdx; res.lightKind = lightKind; res.pHat = pHat; res.w = w; res.sampleSeed = sampleSeed;
  }
}
inline void addSampleToReservoir(const orc_point_light* lights, Res& res, uint32_t lightIdx, int32_t lightKind,
                                 float lightPdf, const GInfo& g, uint32_t& seed) {                 // :45-54
  float pHat = evaluatePHat(lights, lightIdx, g);
  float weight = pHat / lightPdf;
  res.M += 1;
  float w = (res.sumWeights + weight) / (float(res.M) * pHat);
  updateReservoir(res, lightIdx, lightKind, weight, pHat, w, seed, g.sampleSeed);
}
inline void combineReservoirsGeom(const orc_point_light* lights, Res& self, const Res& other, const GInfo& g,
                                  const GInfo& og, uint32_t& seed) {                               // :56-76
  uint32_t Z = self.M;
  self.M += other.M;
  float pHat = evaluatePHat(lights, other.lightIndex, g);
  float weight = pHat * other.w * float(other.M);
  if (weight > 0.0f) updateReservoir(self, other.lightIndex, other.lightKind, weight, pHat, other.w, seed, other.sampleSeed);
  pHat = evaluatePHat(lights, self.lightIndex, og);
  if (pHat > 0.0f) Z += other.M;
  if (self.w > 0.0f) self.w = self.sumWeights / (float(Z) * self.pHat);
}
inline void combineReservoirsPlain(Res& self, const Res& other, float pHat, uint32_t& seed) {      // :78-90
  self.M += other.M;
  float weight = pHat * other.w * float(other.M);
  if (weight > 0.0f) updateReservoir(self, other.lightIndex, other.lightKind, weight, pHat, other.w, seed, other.sampleSeed);
  if (self.w > 0.0f) self.w = self.sumWeights / (float(self.M) * self.pHat);
}

// ---------------------------------------------------------------- alias table
inline void aliasTableSample(const orc_alias_cell* t, int n, float r1, float r2, uint32_t& index, float& prob) {   // restir.rgen:97-110
  uint32_t col = uint32_t(float(n) * r1);
  uint32_t last = uint32_t(n - 1);
  if (last < col) col = last;
  const orc_alias_cell& c = t[col];
  if (c.prob > r2) { index = col; prob = c.pdf; } else { index = (uint32_t)c.alias; prob = c.aliasPdf; }
}

// ---------------------------------------------------------------- volumetric front-end (no reference counterpart)
inline float neglog1m(float u) {
  // -ln(1-u) for u = k/2^24 using only + - * / (identical on CPU and GPU without FMA contraction)
  float x = 1.0f - u;
  uint32_t bits = f2u(x);
  int e = int(bits >> 23) - 127;
  float m = u2f((bits & 0x007FFFFFu) | 0x3F800000u);
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

inline int floordiv8(int v) { return v >> 3; }   // arithmetic shift = floor for negatives

struct Grid {
  const orc_scene* s;
  int cdim[3];
  inline float density(int i, int j, int k) const {
    int x = i - s->vmin[0], y = j - s->vmin[1], z = k - s->vmin[2];
    if (x < 0 || y < 0 || z < 0 || x >= s->vdim[0] || y >= s->vdim[1] || z >= s->vdim[2]) return s->bg_density;
    return s->dens[(size_t(z) * s->vdim[1] + y) * s->vdim[0] + x];
  }
  inline float cellmax(int cx, int cy, int cz) const {   // cell coords relative to the window
    return s->cellmax[(size_t(cz) * cdim[1] + cy) * cdim[0] + cx];
  }
};
inline Grid make_grid(const orc_scene* s) {
  Grid g; g.s = s; g.cdim[0] = s->vdim[0] / 8; g.cdim[1] = s->vdim[1] / 8; g.cdim[2] = s->vdim[2] / 8; return g;
}

// Shared DDA over 8^3-voxel cells with per-cell majorants.
// MODE 0: delta tracking (first real collision), MODE 1: ratio tracking (transmittance estimate).
// One exponential optical-depth sample tau is carried across cells ("residual" tracking): a cell with majorant mu
// consumes (tcell - t) * mu of it; only when tau runs out inside a cell does a tentative collision happen there,
// after which a fresh tau is drawn.  Exit times advance incrementally (tn[axis] += dt[axis]).
struct TrackResult { bool hit; float t; int vox[3]; float T; uint32_t ntent, ncells; };

template <int MODE>
inline TrackResult track(const Grid& g, V3 org, V3 dir, float tmin, float tmax, uint32_t& seed) {
  const orc_scene* s = g.s;
  TrackResult R; R.hit = false; R.t = 0.0f; R.vox[0] = R.vox[1] = R.vox[2] = 0; R.T = 1.0f; R.ntent = 0; R.ncells = 0;
  // voxel-space ray: q(t) = o + d t, voxel = floor(q); voxel ijk is centred on its index position (+0.5 shift)
  float o[3] = { (org.x - s->B[0]) * s->invA + 0.5f, (org.y - s->B[1]) * s->invA + 0.5f, (org.z - s->B[2]) * s->invA + 0.5f };
  float d[3] = { dir.x * s->invA, dir.y * s->invA, dir.z * s->invA };
  float t0 = tmin, t1 = tmax;
  float inv[3];
  for (int a = 0; a < 3; ++a) {
    float lo = float(s->vmin[a]), hi = float(s->vmin[a] + s->vdim[a]);
    if (d[a] == 0.0f) {
      inv[a] = 0.0f;
      if (o[a] < lo || !(o[a] < hi)) return R;
    } else {
      inv[a] = 1.0f / d[a];
      float ta = (lo - o[a]) * inv[a], tb = (hi - o[a]) * inv[a];
      float tn = gmin(ta, tb), tf = gmax(ta, tb);
      t0 = gmax(t0, tn); t1 = gmin(t1, tf);
    }
  }
  if (!(t0 < t1)) return R;
  int c[3], step[3];
  float tn[3], dt[3];
  for (int a = 0; a < 3; ++a) {
    float q = o[a] + d[a] * t0;
    int v = int(floorf(q)) - s->vmin[a];
    int ci = floordiv8(v);
    if (ci < 0) ci = 0;
    if (ci > g.cdim[a] - 1) ci = g.cdim[a] - 1;
    c[a] = ci;
    step[a] = d[a] > 0.0f ? 1 : -1;
    if (d[a] == 0.0f) { tn[a] = INFINITY; dt[a] = 0.0f; }
    else {
      float bound = float(s->vmin[a] + (c[a] + (step[a] > 0 ? 1 : 0)) * 8);
      tn[a] = (bound - o[a]) * inv[a];
      dt[a] = fabsf(inv[a]) * 8.0f;
    }
  }
  float t = t0;
  float tau = neglog1m(rnd(seed));
  for (;;) {
    R.ncells += 1;
    int axis = 0; float tcell = tn[0];
    if (tn[1] < tcell) { tcell = tn[1]; axis = 1; }
    if (tn[2] < tcell) { tcell = tn[2]; axis = 2; }
    bool last = false;
    if (!(tcell < t1)) { tcell = t1; last = true; }
    float mu_d = g.cellmax(c[0], c[1], c[2]);
    if (mu_d > 0.0f) {
      float mu = mu_d * s->density_scale;
      int vlo[3] = { s->vmin[0] + c[0] * 8, s->vmin[1] + c[1] * 8, s->vmin[2] + c[2] * 8 };
      for (;;) {
        float seg = (tcell - t) * mu;
        if (!(tau < seg)) { tau = tau - seg; break; }
        t = t + tau / mu;
        R.ntent += 1;
        int vx[3];
        for (int a = 0; a < 3; ++a) {
          int v = int(floorf(o[a] + d[a] * t));
          if (v < vlo[a]) v = vlo[a];
          if (v > vlo[a] + 7) v = vlo[a] + 7;
          vx[a] = v;
        }
        float dens = g.density(vx[0], vx[1], vx[2]);
        if (MODE == 0) {
          float u2 = rnd(seed);
          if (u2 * mu_d < dens) { R.hit = true; R.t = t; R.vox[0] = vx[0]; R.vox[1] = vx[1]; R.vox[2] = vx[2]; return R; }
        } else {
          R.T = R.T * (1.0f - dens / mu_d);
          if (!(R.T > 1e-5f)) { R.T = 0.0f; return R; }     // opaque for every practical purpose: stop marching
        }
        tau = neglog1m(rnd(seed));
      }
    }
    t = tcell;
    if (last) return R;
    c[axis] += step[axis];
    if (c[axis] < 0 || c[axis] >= g.cdim[axis]) return R;
    tn[axis] = tn[axis] + dt[axis];
  }
}

inline float ratio_track(const Grid& g, V3 P, V3 L, uint32_t& seed, uint32_t* counters) {
  V3 dir = sub(L, P);
  float dist = sqrtf(dot(dir, dir));
  if (!(dist > 0.0f)) return 1.0f;
  dir = divs(dir, dist);
  TrackResult r = track<1>(g, P, dir, 0.0f, dist, seed);
  if (counters) { counters[0] = r.ntent; counters[1] = r.ncells; }
  return r.T;
}

// voxel material: vdb/vdb.cpp:814-821 (smoke colour) + Renderer.cpp:1494-1500 (baseColor, metallic, roughness)
inline void voxel_albedo(float v, float* out4) {
  float n3 = sqrtf(100.0f * 100.0f + 100.0f * 100.0f + 100.0f * 100.0f);   // nvmath::normalize(vec3(100,100,100)), nvmath.inl:1222-1231
  float s = 100.0f * (1.0f / n3);
  float c = s * v * 1000.0f;
  float n4 = sqrtf(c * c + c * c + c * c + 1.0f * 1.0f);                   // nvmath::normalize(vec4(c,c,c,1)), nvmath.inl:1234-1243
  float inv = n4 > 10e-6f ? 1.0f / n4 : 0.0f;
  out4[0] = c * inv; out4[1] = c * inv; out4[2] = c * inv; out4[3] = 1.0f * inv;
}

inline void mat_vec(const float* m, const float* v, float* o) {   // nvmath.inl:483-492, column-major aRC = m[C*4+R]
  for (int r = 0; r < 4; ++r) o[r] = m[0 + r] * v[0] + m[4 + r] * v[1] + m[8 + r] * v[2] + m[12 + r] * v[3];
}

struct PixelG { float worldPos[4], albedo[4], normal[4], mat[4]; };

inline void store4(float* img, size_t idx, const float* v) { memcpy(img + idx * 4, v, 16); }
inline void load4(const float* img, size_t idx, float* v) { memcpy(v, img + idx * 4, 16); }

inline GInfo ginfo_from_images(const orc_gbuffer& gb, size_t idx, const float* camPos) {   // spatialReuse.comp:65-74 / restir_post.frag:60-67
  float p[4], a[4], n[4], m[4];
  load4(gb.worldPos, idx, p); load4(gb.albedo, idx, a); load4(gb.normal, idx, n); load4(gb.matProps, idx, m);
  GInfo g;
  g.albedo[0] = a[0]; g.albedo[1] = a[1]; g.albedo[2] = a[2]; g.albedo[3] = a[3];
  g.normal = v3(n[0], n[1], n[2]);
  g.worldPos = v3(p[0], p[1], p[2]);
  g.roughness = m[0]; g.metallic = m[1];
  g.albedoLum = luminance_common(a[0], a[1], a[2]);
  g.camPos = v3(camPos[0], camPos[1], camPos[2]);
  g.sampleSeed = 0;
  return g;
}
inline Res res_from_images(const orc_reservoirs& r, size_t idx) {   // unpackReservoirStruct, reservoir.glsl:4-16
  uint32_t a[4]; float b[4];
  memcpy(a, r.info + idx * 4, 16); memcpy(b, r.weight + idx * 4, 16);
  Res o; o.M = a[0]; o.lightIndex = a[1]; o.lightKind = (int32_t)a[2]; o.sampleSeed = a[3];
  o.pHat = b[0]; o.sumWeights = b[1]; o.w = b[2];
  return o;
}
inline void res_to_images(const Res& o, const orc_reservoirs& r, size_t idx) {   // packReservoirStruct, reservoir.glsl:18-28
  uint32_t a[4] = { o.M, o.lightIndex, (uint32_t)o.lightKind, o.sampleSeed };
  float b[4] = { o.pHat, o.sumWeights, o.w, 0.0f };
  memcpy(r.info + idx * 4, a, 16); memcpy(r.weight + idx * 4, b, 16);
}

const int FLAG_VISIBILITY = 1 << 0, FLAG_TEMPORAL = 1 << 1, FLAG_SPATIAL = 1 << 2;     // host_device.h:427-430
const int FLAG_FINAL_VISIBILITY = 1 << 4, FLAG_FINALIZE_W = 1 << 5;                    // new (include/vrs.h)
const uint32_t PASS_INITIAL = 0, PASS_SPATIAL0 = 1, PASS_SHADE = 5;
const int MAX_NEIGHBORS = 16;

inline void primary_ray(const orc_global_uniforms* gu, uint32_t x, uint32_t y, uint32_t W, uint32_t H, V3& org, V3& dir) {
  // restir.rgen:142-148
  float ux = float(x) / float(W), uy = float(y) / float(H);
  float dx = ux * 2.0f - 1.0f, dy = uy * 2.0f - 1.0f;
  float o4[4], t4[4], d4[4];
  float v0[4] = { 0.0f, 0.0f, 0.0f, 1.0f };
  mat_vec(gu->viewInverse, v0, o4);
  float v1[4] = { dx, dy, 1.0f, 1.0f };
  mat_vec(gu->projInverse, v1, t4);
  V3 tn = normalize(v3(t4[0], t4[1], t4[2]));
  float v2[4] = { tn.x, tn.y, tn.z, 0.0f };
  mat_vec(gu->viewInverse, v2, d4);
  org = v3(o4[0], o4[1], o4[2]); dir = v3(d4[0], d4[1], d4[2]);
}

// Primary volume event -> Payload (raycommon.glsl:24-32 contract)
inline bool primary_event(const Grid& g, V3 org, V3 dir, uint32_t& seed, PixelG& px, uint32_t* trace) {
  const orc_scene* s = g.s;
  memset(&px, 0, sizeof(px));
  TrackResult r = track<0>(g, org, dir, 0.0001f, 100000.0f, seed);    // restir.rgen:164-166 ray range
  if (trace) { trace[1] = r.ntent; trace[2] = r.ncells; trace[0] = 0xFFFFFFFFu; }
  px.normal[3] = 1.0f; px.mat[2] = 1.0f; px.mat[3] = 1.0f;            // restir.rgen:195-197 constants
  if (!r.hit) return false;
  V3 P = add(org, muls(dir, r.t));
  int i = r.vox[0], j = r.vox[1], k = r.vox[2];
  if (trace) trace[0] = uint32_t(i - s->vmin[0]) + uint32_t(s->vdim[0]) * (uint32_t(j - s->vmin[1]) + uint32_t(s->vdim[1]) * uint32_t(k - s->vmin[2]));
  float dens = g.density(i, j, k);
  V3 grad = v3(g.density(i + 1, j, k) - g.density(i - 1, j, k), g.density(i, j + 1, k) - g.density(i, j - 1, k),
               g.density(i, j, k + 1) - g.density(i, j, k - 1));
  float gg = dot(grad, grad);
  V3 n;
  if (gg > 0.0f) { float l = sqrtf(gg); n = v3(-grad.x / l, -grad.y / l, -grad.z / l); }
  else n = v3(-dir.x, -dir.y, -dir.z);
  px.worldPos[0] = P.x; px.worldPos[1] = P.y; px.worldPos[2] = P.z; px.worldPos[3] = 1.0f;
  voxel_albedo(dens, px.albedo);
  px.normal[0] = n.x; px.normal[1] = n.y; px.normal[2] = n.z;
  px.mat[0] = s->roughness; px.mat[1] = s->metallic;
  return true;
}

}  // namespace

// =================================================================== C exports
extern "C" {

void orc_pcg2d(uint32_t x, uint32_t y, uint32_t* o) { pcg2d(x, y); o[0] = x; o[1] = y; }
uint32_t orc_lcg(uint32_t* s) { return lcg(*s); }
float orc_rnd(uint32_t* s) { return rnd(*s); }
uint32_t orc_pixel_seed(uint32_t x, uint32_t y, uint32_t clock, uint32_t pass) { return pixel_seed(x, y, clock, pass); }
float orc_luminance_common(float r, float g, float b) { return luminance_common(r, g, b); }
float orc_luminance_utils(float r, float g, float b) { return luminance_utils(v3(r, g, b)); }
float orc_disney_brdf_luminance(float a, float b, float c, float d, float lum, float rough, float metal) {
  return disneyBrdfLuminance(a, b, c, d, lum, rough, metal);
}
void orc_disney_brdf_color(float a, float b, float c, float d, const float* alb, float rough, float metal, float* out) {
  V3 r = disneyBrdfColor(a, b, c, d, v3(alb[0], alb[1], alb[2]), rough, metal); out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
float orc_evaluate_phat(const orc_point_light* l, uint32_t idx, const float* g16) { return evaluatePHat(l, idx, ginfo_from16(g16)); }
void orc_evaluate_phat_full(const orc_point_light* l, uint32_t idx, const float* g16, float* out) {
  V3 r = evaluatePHatFull(l, idx, ginfo_from16(g16)); out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

void orc_create_alias_table(const float* pdf, int n, orc_alias_cell* out) {   // utils/restir_utils.cpp:90-155
  std::queue<int> bigger, smaller;
  std::vector<float> lp(pdf, pdf + n);
  float powerSum = 0.0f;
  for (int i = 0; i < n; ++i) powerSum += pdf[i];
  for (int i = 0; i < n; ++i) { out[i].alias = -1; out[i].prob = 0.0f; out[i].pdf = 0.0f; out[i].aliasPdf = 0.0f; }
  for (int i = 0; i < n; ++i) {
    out[i].pdf = lp[i] / powerSum;
    lp[i] = float(n) * lp[i] / powerSum;
    if (lp[i] >= 1.0f) bigger.push(i); else smaller.push(i);
  }
  while (!bigger.empty() && !smaller.empty()) {
    int gq = bigger.front(); bigger.pop();
    int l = smaller.front(); smaller.pop();
    out[l].prob = lp[l];
    out[l].alias = gq;
    lp[gq] = (lp[gq] + lp[l]) - 1.0f;
    if (lp[gq] < 1.0f) smaller.push(gq); else bigger.push(gq);
  }
  while (!bigger.empty()) { int gq = bigger.front(); bigger.pop(); out[gq].prob = 1.0f; out[gq].alias = gq; }
  while (!smaller.empty()) { int l = smaller.front(); smaller.pop(); out[l].prob = 1.0f; out[l].alias = l; }
  for (int i = 0; i < n; ++i) out[i].aliasPdf = out[out[i].alias].pdf;
}
void orc_alias_table_sample(const orc_alias_cell* t, int n, float r1, float r2, uint32_t* index, float* prob) {
  aliasTableSample(t, n, r1, r2, *index, *prob);
}

// utils/restir_utils.cpp:22-51 with libstdc++ semantics spelled out:
//   std::default_random_engine = minstd_rand0 (x <- 16807 x mod 2^31-1, seed 1);
//   uniform_real_distribution<float>(a,b)(g) = (b-a) * generate_canonical<float,24>(g) + a,
//   generate_canonical = float(g() - 1) / 2147483646.0f, clamped below 1;
//   constructor arguments are evaluated right-to-left by g++ (Z, Y, X then B, G, R).
void orc_generate_point_lights(const float* mn, const float* mx, int white, uint32_t n, orc_point_light* out) {
  uint64_t st = 1;
  auto canon = [&st]() {
    st = (st * 16807ull) % 2147483647ull;
    float r = float(uint32_t(st) - 1u) / 2147483646.0f;
    if (r >= 1.0f) r = nextafterf(1.0f, 0.0f);
    return r;
  };
  for (uint32_t i = 0; i < n; ++i) {
    float z = (mx[2] - mn[2]) * canon() + mn[2];
    float y = (mx[1] - mn[1]) * canon() + mn[1];
    float x = (mx[0] - mn[0]) * canon() + mn[0];
    out[i].pos[0] = x; out[i].pos[1] = y; out[i].pos[2] = z; out[i].pos[3] = 1.0f;
    float r = 1.0f, g = 1.0f, b = 1.0f;
    if (!white) { b = (1.0f - 0.0f) * canon() + 0.0f; g = (1.0f - 0.0f) * canon() + 0.0f; r = (1.0f - 0.0f) * canon() + 0.0f; }
    out[i].emission_luminance[0] = r; out[i].emission_luminance[1] = g; out[i].emission_luminance[2] = b;
    out[i].emission_luminance[3] = luminance_common(r, g, b);
  }
}

void orc_initial_ris(const orc_scene* s, const float* g16, int count, uint32_t* seed, uint32_t* res8) {   // restir.rgen:203-227
  GInfo g = ginfo_from16(g16);
  Res res = newReservoir();
  if (dot(g.normal, g.normal) != 0.0f) {
    for (int i = 0; i < count; ++i) {
      g.sampleSeed = *seed;
      float r1 = rnd(*seed), r2 = rnd(*seed);
      uint32_t idx; float pdf;
      aliasTableSample(s->table, s->ntable, r1, r2, idx, pdf);
      addSampleToReservoir(s->lights, res, idx, 0, pdf, g, *seed);
    }
  }
  res_to8(res, res8);
}
void orc_combine_geom(const orc_scene* s, uint32_t* self8, const uint32_t* other8, const float* g16, const float* og16, uint32_t* seed) {
  Res a = res_from8(self8), b = res_from8(other8);
  combineReservoirsGeom(s->lights, a, b, ginfo_from16(g16), ginfo_from16(og16), *seed);
  res_to8(a, self8);
}
void orc_combine_plain(uint32_t* self8, const uint32_t* other8, float pHat, uint32_t* seed) {
  Res a = res_from8(self8), b = res_from8(other8);
  combineReservoirsPlain(a, b, pHat, *seed);
  res_to8(a, self8);
}
void orc_post_shade(const orc_scene* s, const uint32_t* res8, const float* g16, float thr, float* out) {   // restir_post.frag:78-92
  Res res = res_from8(res8); GInfo g = ginfo_from16(g16); g.sampleSeed = res.sampleSeed;
  V3 pHat = evaluatePHatFull(s->lights, res.lightIndex, g);
  V3 c = add(v3(0.0f, 0.0f, 0.0f), muls(pHat, res.w));
  if (g.albedo[3] > 0.5f) c = v3(g.albedo[0], g.albedo[1], g.albedo[2]);
  float lum = luminance_utils(c);
  if (lum > thr) c = muls(c, thr / lum);
  c = v3(0.0f < c.x ? c.x : 0.0f, 0.0f < c.y ? c.y : 0.0f, 0.0f < c.z ? c.z : 0.0f);
  out[0] = c.x; out[1] = c.y; out[2] = c.z;
}

// nvmath.inl:1149-1183
void orc_perspectiveVK(float fovy, float aspect, float nearPlane, float farPlane, float* M) {
  float f = farPlane, n = nearPlane;
  float nv_pi = float(3.14159265358979323846264338327950288419716939937510582);
  float t = n * tanf(fovy * (nv_pi / float(180)) * float(0.5));
  float b = -t;
  float l = b * aspect;
  float r = t * aspect;
  for (int i = 0; i < 16; ++i) M[i] = 0.0f;
  M[0] = (2 * n) / (r - l);
  M[5] = -(2 * n) / (t - b);
  M[8] = (r + l) / (r - l);
  M[9] = (t + b) / (t - b);
  M[10] = -(f) / (f - n);
  M[11] = -1;
  M[14] = (f * n) / (n - f);
}
// nvmath.inl:979-1025 (vector3::normalize = multiply by reciprocal norm, :377-388)
void orc_look_at(const float* eye, const float* center, const float* up, float* M) {
  auto nrm = [](float* v) {
    float norm = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    norm = norm > 10e-6f ? 1.0f / norm : 0.0f;
    v[0] *= norm; v[1] *= norm; v[2] *= norm;
  };
  float z[3] = { eye[0] - center[0], eye[1] - center[1], eye[2] - center[2] };
  nrm(z);
  float y[3] = { up[0], up[1], up[2] };
  float x[3] = { y[1] * z[2] - y[2] * z[1], y[2] * z[0] - y[0] * z[2], y[0] * z[1] - y[1] * z[0] };
  float y2[3] = { z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0] };
  nrm(x); nrm(y2);
  M[0] = x[0]; M[4] = x[1]; M[8] = x[2];  M[12] = -x[0] * eye[0] - x[1] * eye[1] - x[2] * eye[2];
  M[1] = y2[0]; M[5] = y2[1]; M[9] = y2[2]; M[13] = -y2[0] * eye[0] - y2[1] * eye[1] - y2[2] * eye[2];
  M[2] = z[0]; M[6] = z[1]; M[10] = z[2]; M[14] = -z[0] * eye[0] - z[1] * eye[1] - z[2] * eye[2];
  M[3] = 0.0f; M[7] = 0.0f; M[11] = 0.0f; M[15] = 1.0f;
}
// nvmath.inl:777-850 (cofactor expansion)
void orc_invert(const float* A, float* B) {
#define E(r, c) A[(c) * 4 + (r)]
  auto det2 = [](float a1, float a2, float b1, float b2) { return a1 * b2 - b1 * a2; };
  auto det3 = [&det2](float a1, float a2, float a3, float b1, float b2, float b3, float c1, float c2, float c3) {
    return a1 * det2(b2, b3, c2, c3) - b1 * det2(a2, a3, c2, c3) + c1 * det2(a2, a3, b2, b3);
  };
  float b00 = det3(E(1,1), E(2,1), E(3,1), E(1,2), E(2,2), E(3,2), E(1,3), E(2,3), E(3,3));
  float b10 = -det3(E(1,0), E(2,0), E(3,0), E(1,2), E(2,2), E(3,2), E(1,3), E(2,3), E(3,3));
  float b20 = det3(E(1,0), E(2,0), E(3,0), E(1,1), E(2,1), E(3,1), E(1,3), E(2,3), E(3,3));
  float b30 = -det3(E(1,0), E(2,0), E(3,0), E(1,1), E(2,1), E(3,1), E(1,2), E(2,2), E(3,2));
  float b01 = -det3(E(0,1), E(2,1), E(3,1), E(0,2), E(2,2), E(3,2), E(0,3), E(2,3), E(3,3));
  float b11 = det3(E(0,0), E(2,0), E(3,0), E(0,2), E(2,2), E(3,2), E(0,3), E(2,3), E(3,3));
  float b21 = -det3(E(0,0), E(2,0), E(3,0), E(0,1), E(2,1), E(3,1), E(0,3), E(2,3), E(3,3));
  float b31 = det3(E(0,0), E(2,0), E(3,0), E(0,1), E(2,1), E(3,1), E(0,2), E(2,2), E(3,2));
  float b02 = det3(E(0,1), E(1,1), E(3,1), E(0,2), E(1,2), E(3,2), E(0,3), E(1,3), E(3,3));
  float b12 = -det3(E(0,0), E(1,0), E(3,0), E(0,2), E(1,2), E(3,2), E(0,3), E(1,3), E(3,3));
  float b22 = det3(E(0,0), E(1,0), E(3,0), E(0,1), E(1,1), E(3,1), E(0,3), E(1,3), E(3,3));
  float b32 = -det3(E(0,0), E(1,0), E(3,0), E(0,1), E(1,1), E(3,1), E(0,2), E(1,2), E(3,2));
  float b03 = -det3(E(0,1), E(1,1), E(2,1), E(0,2), E(1,2), E(2,2), E(0,3), E(1,3), E(2,3));
  float b13 = det3(E(0,0), E(1,0), E(2,0), E(0,2), E(1,2), E(2,2), E(0,3), E(1,3), E(2,3));
  float b23 = -det3(E(0,0), E(1,0), E(2,0), E(0,1), E(1,1), E(2,1), E(0,3), E(1,3), E(2,3));
  float b33 = det3(E(0,0), E(1,0), E(2,0), E(0,1), E(1,1), E(2,1), E(0,2), E(1,2), E(2,2));
  float det = (E(0,0) * b00) + (E(0,1) * b10) + (E(0,2) * b20) + (E(0,3) * b30);
  float oodet = 1.0f / det;
#undef E
  float b[16] = { b00, b10, b20, b30, b01, b11, b21, b31, b02, b12, b22, b32, b03, b13, b23, b33 };
  for (int i = 0; i < 16; ++i) B[i] = b[i] * oodet;
}
void orc_matmul(const float* A, const float* Bm, float* C) {   // nvmath.inl:663-684
  float out[16];
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r)
      out[c * 4 + r] = A[0 + r] * Bm[c * 4 + 0] + A[4 + r] * Bm[c * 4 + 1] + A[8 + r] * Bm[c * 4 + 2] + A[12 + r] * Bm[c * 4 + 3];
  memcpy(C, out, sizeof(out));
}
void orc_voxel_albedo(float density, float* out4) { voxel_albedo(density, out4); }

float orc_neglog1m(float u) { return neglog1m(u); }

void orc_scene_prepare(orc_scene* s) {
  int cd[3] = { s->vdim[0] / 8, s->vdim[1] / 8, s->vdim[2] / 8 };
#pragma omp parallel for
  for (int cz = 0; cz < cd[2]; ++cz)
    for (int cy = 0; cy < cd[1]; ++cy)
      for (int cx = 0; cx < cd[0]; ++cx) {
        float m = 0.0f;
        for (int z = 0; z < 8; ++z)
          for (int y = 0; y < 8; ++y)
            for (int x = 0; x < 8; ++x) {
              float v = s->dens[(size_t(cz * 8 + z) * s->vdim[1] + (cy * 8 + y)) * s->vdim[0] + (cx * 8 + x)];
              if (v > m) m = v;
            }
        s->cellmax[(size_t(cz) * cd[1] + cy) * cd[0] + cx] = m;
      }
}
float orc_density_at(const orc_scene* s, int i, int j, int k) { return make_grid(s).density(i, j, k); }

int orc_delta_track(const orc_scene* s, const float* o, const float* d, float tmin, float tmax, uint32_t* seed,
                    float* t_hit, int32_t* voxel3, uint32_t* counters2) {
  Grid g = make_grid(s);
  TrackResult r = track<0>(g, v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2]), tmin, tmax, *seed);
  *t_hit = r.t; voxel3[0] = r.vox[0]; voxel3[1] = r.vox[1]; voxel3[2] = r.vox[2];
  counters2[0] = r.ntent; counters2[1] = r.ncells;
  return r.hit ? 1 : 0;
}
float orc_ratio_track(const orc_scene* s, const float* p, const float* l, uint32_t* seed, uint32_t* counters2) {
  Grid g = make_grid(s);
  return ratio_track(g, v3(p[0], p[1], p[2]), v3(l[0], l[1], l[2]), *seed, counters2);
}

// ------------------------------------------------------------------ pass 0: restir.rgen main (:136-290)
void orc_pass_initial(const orc_scene* s, const orc_global_uniforms* gu, const orc_restir_uniforms* ru, uint32_t clock,
                      int y0, int y1, orc_gbuffer cur, orc_gbuffer prev, orc_reservoirs prevRes, orc_reservoirs outRes,
                      uint32_t* trace4) {
  const uint32_t W = ru->screenSize[0], H = ru->screenSize[1];
  Grid g = make_grid(s);
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = y0; y < y1; ++y) {
    for (uint32_t x = 0; x < W; ++x) {
      size_t idx = size_t(y) * W + x;
      uint32_t seed = pixel_seed(x, uint32_t(y), clock, PASS_INITIAL);     // :139-140
      V3 org, dir;
      primary_ray(gu, x, uint32_t(y), W, H, org, dir);
      PixelG px; uint32_t tr[4] = { 0, 0, 0, 0 };
      bool exist = primary_event(g, org, dir, seed, px, tr);
      store4(cur.worldPos, idx, px.worldPos); store4(cur.albedo, idx, px.albedo);   // :193-197
      store4(cur.normal, idx, px.normal); store4(cur.matProps, idx, px.mat);
      Res res = newReservoir();
      if (exist) {
        GInfo gi;
        gi.albedo[0] = px.albedo[0]; gi.albedo[1] = px.albedo[1]; gi.albedo[2] = px.albedo[2]; gi.albedo[3] = px.albedo[3];
        gi.normal = v3(px.normal[0], px.normal[1], px.normal[2]);
        gi.worldPos = v3(px.worldPos[0], px.worldPos[1], px.worldPos[2]);
        gi.metallic = px.mat[1]; gi.roughness = px.mat[0];
        gi.albedoLum = luminance_common(gi.albedo[0], gi.albedo[1], gi.albedo[2]);   // :182
        gi.camPos = v3(ru->currCamPos[0], ru->currCamPos[1], ru->currCamPos[2]);      // :183
        gi.sampleSeed = 0;
        if (dot(gi.normal, gi.normal) != 0.0f) {                                       // :205
          for (uint32_t i = 0; i < ru->initialLightSampleCount; ++i) {                 // :206-226
            gi.sampleSeed = seed;                                                      // :213
            float r1 = rnd(seed), r2 = rnd(seed);                                      // :116 (left-to-right)
            uint32_t sel; float pdf;
            aliasTableSample(s->table, s->ntable, r1, r2, sel, pdf);
            addSampleToReservoir(s->lights, res, sel, 0, pdf, gi, seed);               // :224-225
          }
        }
        if ((ru->flags & FLAG_FINALIZE_W) != 0 && res.w > 0.0f) {                      // new: standard RIS weight
          res.w = res.sumWeights / (float(res.M) * res.pHat);
        }
        if ((ru->flags & FLAG_VISIBILITY) != 0 && res.w > 0.0f) {                      // :229-235, binary test -> transmittance
          const orc_point_light& L = s->lights[res.lightIndex];
          float T = ratio_track(g, gi.worldPos, v3(L.pos[0], L.pos[1], L.pos[2]), seed, nullptr);
          res.w = res.w * T;
          res.sumWeights = res.sumWeights * T;
        }
        if ((ru->flags & FLAG_TEMPORAL) != 0) {                                        // :237-284 (commented block, intent)
          float P4[4] = { gi.worldPos.x, gi.worldPos.y, gi.worldPos.z, 1.0f }, q[4];
          mat_vec(ru->prevFrameProjectionViewMatrix, P4, q);
          q[0] = q[0] / q[3]; q[1] = q[1] / q[3]; q[2] = q[2] / q[3];
          q[0] = (q[0] + 1.0f) * 0.5f * float(W);
          q[1] = (q[1] + 1.0f) * 0.5f * float(H);
          if (q[0] > 0.0f && q[1] > 0.0f && q[0] < float(W) && q[1] < float(H)) {
            int fx = int(q[0]), fy = int(q[1]);
            size_t pidx = size_t(fy) * W + size_t(fx);
            GInfo pg = ginfo_from_images(prev, pidx, ru->currCamPos);                  // prevGInfo.camPos = gInfo.camPos (:259)
            V3 pd = sub(gi.worldPos, pg.worldPos);
            // a previous miss holds no data; the reference would reject it anyway (cleared normal fails :274)
            if (!(prev.worldPos[pidx * 4 + 3] < 0.5f) && dot(pd, pd) < 0.01f) {
              V3 ad = v3(gi.albedo[0] - pg.albedo[0], gi.albedo[1] - pg.albedo[1], gi.albedo[2] - pg.albedo[2]);
              if (dot(ad, ad) < 0.01f) {
                if (dot(gi.normal, pg.normal) > 0.5f) {
                  Res pr = res_from_images(prevRes, pidx);                             // at prevFrag
                  uint32_t cap = uint32_t(ru->temporalSampleCountMultiplier) * res.M;
                  if (cap < pr.M) pr.M = cap;
                  combineReservoirsGeom(s->lights, res, pr, gi, pg, seed);
                }
              }
            }
          }
        }
      }
      res_to_images(res, outRes, idx);                                                 // :286-289 (also on miss: SURVEY App. C-5)
      if (trace4) { tr[3] = seed; memcpy(trace4 + idx * 4, tr, 16); }
    }
  }
}

// ------------------------------------------------------------------ pass 1+i: spatial reuse (spatialReuse.comp + reservoir.glsl:56-76)
void orc_pass_spatial(const orc_scene* s, const orc_restir_uniforms* ru, uint32_t clock, uint32_t iteration, int y0, int y1,
                      orc_gbuffer cur, orc_reservoirs inRes, orc_reservoirs outRes) {
  const uint32_t W = ru->screenSize[0], H = ru->screenSize[1];
  const float radius = ru->spatialRadius;
  uint32_t k = ru->spatialNeighbors; if (k > (uint32_t)MAX_NEIGHBORS) k = MAX_NEIGHBORS;
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = y0; y < y1; ++y) {
    for (uint32_t x = 0; x < W; ++x) {
      size_t idx = size_t(y) * W + x;
      uint32_t seed = pixel_seed(x, uint32_t(y), clock, PASS_SPATIAL0 + iteration);    // spatialReuse.comp:58-59
      Res res = res_from_images(inRes, idx);
      float exist = cur.worldPos[idx * 4 + 3];
      if (!(exist < 0.5f)) {                                                            // :76-79
        GInfo gi = ginfo_from_images(cur, idx, ru->currCamPos);
        uint32_t Z = res.M;
        size_t nb_idx[MAX_NEIGHBORS]; uint32_t nb_M[MAX_NEIGHBORS]; int nacc = 0;
        // the 2k offset draws come first, the selection draws of updateReservoir after them: a neighbour's position in
        // the RNG stream then does not depend on the neighbours before it
        float off[MAX_NEIGHBORS][2];
        for (uint32_t i = 0; i < k; ++i) { off[i][0] = rnd(seed); off[i][1] = rnd(seed); }
        for (uint32_t i = 0; i < k; ++i) {
          float r1 = off[i][0], r2 = off[i][1];
          float dx = (r1 * 2.0f - 1.0f) * radius, dy = (r2 * 2.0f - 1.0f) * radius;
          if (dx * dx + dy * dy > radius * radius) continue;
          int ox = int(dx), oy = int(dy);
          if (ox == 0 && oy == 0) continue;
          int nx = int(x) + ox, ny = y + oy;
          if (nx < 0 || ny < 0 || nx >= int(W) || ny >= int(H)) continue;
          size_t nidx = size_t(ny) * W + size_t(nx);
          if (cur.worldPos[nidx * 4 + 3] < 0.5f) continue;
          GInfo ng = ginfo_from_images(cur, nidx, ru->currCamPos);
          V3 pd = sub(gi.worldPos, ng.worldPos);
          if (!(dot(pd, pd) < 0.01f)) continue;
          V3 ad = v3(gi.albedo[0] - ng.albedo[0], gi.albedo[1] - ng.albedo[1], gi.albedo[2] - ng.albedo[2]);
          if (!(dot(ad, ad) < 0.01f)) continue;
          if (!(dot(gi.normal, ng.normal) > 0.5f)) continue;
          Res nr = res_from_images(inRes, nidx);
          // reservoir.glsl:61-68
          res.M += nr.M;
          float pHat = evaluatePHat(s->lights, nr.lightIndex, gi);
          float weight = pHat * nr.w * float(nr.M);
          if (weight > 0.0f) updateReservoir(res, nr.lightIndex, nr.lightKind, weight, pHat, nr.w, seed, nr.sampleSeed);
          nb_idx[nacc] = nidx; nb_M[nacc] = nr.M; ++nacc;
        }
        if (nacc > 0) {
          for (int j = 0; j < nacc; ++j) {                                              // reservoir.glsl:70-73, deferred to the final sample
            GInfo ng = ginfo_from_images(cur, nb_idx[j], ru->currCamPos);
            float pHat = evaluatePHat(s->lights, res.lightIndex, ng);
            if (pHat > 0.0f) Z += nb_M[j];
          }
          if (res.w > 0.0f) res.w = res.sumWeights / (float(Z) * res.pHat);             // :74-75
        }
      }
      res_to_images(res, outRes, idx);
    }
  }
}

// ------------------------------------------------------------------ pass 5: restir_post.frag main (:57-105)
void orc_pass_shade(const orc_scene* s, const orc_restir_uniforms* ru, const orc_push_constant* pc, uint32_t clock, int y0, int y1,
                    orc_gbuffer cur, orc_reservoirs rs, float* accum) {
  const uint32_t W = ru->screenSize[0];
  Grid g = make_grid(s);
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = y0; y < y1; ++y) {
    for (uint32_t x = 0; x < W; ++x) {
      size_t idx = size_t(y) * W + x;
      GInfo gi = ginfo_from_images(cur, idx, ru->currCamPos);
      Res res = res_from_images(rs, idx);
      gi.sampleSeed = res.sampleSeed;
      float exist = cur.worldPos[idx * 4 + 3];
      V3 c;
      if (exist < 0.5f) {
        c = v3(pc->clearColorRed, pc->clearColorGreen, pc->clearColorBlue);            // miss: defined value
      } else {
        V3 pHat = evaluatePHatFull(s->lights, res.lightIndex, gi);
        c = add(v3(0.0f, 0.0f, 0.0f), muls(pHat, res.w));                              // :80-81
        if ((ru->flags & FLAG_FINAL_VISIBILITY) != 0 && res.w > 0.0f) {                // new: unbiased final transmittance
          uint32_t seed = pixel_seed(x, uint32_t(y), clock, PASS_SHADE);
          const orc_point_light& L = s->lights[res.lightIndex];
          float T = ratio_track(g, gi.worldPos, v3(L.pos[0], L.pos[1], L.pos[2]), seed, nullptr);
          c = muls(c, T);
        }
        if (gi.albedo[3] > 0.5f) c = v3(gi.albedo[0], gi.albedo[1], gi.albedo[2]);     // :82-84
        float lum = luminance_utils(c);                                                // :86-90
        if (lum > ru->fireflyClampThreshold) c = muls(c, ru->fireflyClampThreshold / lum);
        c = v3(0.0f < c.x ? c.x : 0.0f, 0.0f < c.y ? c.y : 0.0f, 0.0f < c.z ? c.z : 0.0f);   // :92
      }
      float* a = accum + idx * 4;
      if (pc->frame < 1 || pc->initialize == 1) {                                      // :94-102
        a[0] = c.x; a[1] = c.y; a[2] = c.z; a[3] = 1.0f;
      } else {
        float w = 1.0f / float(pc->frame);
        a[0] = gmix(a[0], c.x, w); a[1] = gmix(a[1], c.y, w); a[2] = gmix(a[2], c.z, w); a[3] = 1.0f;
      }
    }
  }
}

// ------------------------------------------------------------------ brute-force reference estimator
// E[ f(P, y) * Le(y) * G / pdf(y) * T(P, y) ] over (primary event P by delta tracking, light y by alias table,
// T by ratio tracking); the same integrand the ReSTIR passes estimate (including the reference's emissive
// override), without reuse and without the firefly clamp.
void orc_path_trace(const orc_scene* s, const orc_global_uniforms* gu, const orc_restir_uniforms* ru, uint32_t spp,
                    uint32_t seed_base, float* out) {
  const uint32_t W = ru->screenSize[0], H = ru->screenSize[1];
  Grid g = make_grid(s);
#pragma omp parallel for schedule(dynamic, 2)
  for (int y = 0; y < int(H); ++y) {
    for (uint32_t x = 0; x < W; ++x) {
      double acc[3] = { 0, 0, 0 };
      V3 org, dir;
      primary_ray(gu, x, uint32_t(y), W, H, org, dir);
      for (uint32_t sp = 0; sp < spp; ++sp) {
        uint32_t a = x + 7919u * (sp + 1u) + seed_base, b = uint32_t(y) + 104729u * (sp + 1u) + seed_base * 31u;
        pcg2d(a, b);
        uint32_t seed = a + b;
        PixelG px;
        if (!primary_event(g, org, dir, seed, px, nullptr)) continue;
        GInfo gi;
        gi.albedo[0] = px.albedo[0]; gi.albedo[1] = px.albedo[1]; gi.albedo[2] = px.albedo[2]; gi.albedo[3] = px.albedo[3];
        gi.normal = v3(px.normal[0], px.normal[1], px.normal[2]);
        gi.worldPos = v3(px.worldPos[0], px.worldPos[1], px.worldPos[2]);
        gi.metallic = px.mat[1]; gi.roughness = px.mat[0];
        gi.albedoLum = luminance_common(gi.albedo[0], gi.albedo[1], gi.albedo[2]);
        gi.camPos = v3(ru->currCamPos[0], ru->currCamPos[1], ru->currCamPos[2]);
        gi.sampleSeed = 0;
        if (gi.albedo[3] > 0.5f) {                  // restir_post.frag:82-84 emissive override is part of the shaded quantity
          acc[0] += gi.albedo[0]; acc[1] += gi.albedo[1]; acc[2] += gi.albedo[2];
          continue;
        }
        float r1 = rnd(seed), r2 = rnd(seed);
        uint32_t sel; float pdf;
        aliasTableSample(s->table, s->ntable, r1, r2, sel, pdf);
        V3 f = evaluatePHatFull(s->lights, sel, gi);
        if (f.x == 0.0f && f.y == 0.0f && f.z == 0.0f) continue;
        float T = 1.0f;
        if ((ru->flags & (FLAG_VISIBILITY | FLAG_FINAL_VISIBILITY)) != 0) {
          const orc_point_light& L = s->lights[sel];
          T = ratio_track(g, gi.worldPos, v3(L.pos[0], L.pos[1], L.pos[2]), seed, nullptr);
        }
        acc[0] += double(f.x) * T / pdf; acc[1] += double(f.y) * T / pdf; acc[2] += double(f.z) * T / pdf;
      }
      size_t idx = size_t(y) * W + x;
      out[idx * 3 + 0] = float(acc[0] / spp); out[idx * 3 + 1] = float(acc[1] / spp); out[idx * 3 + 2] = float(acc[2] / spp);
    }
  }
}

// OMP_NUM_THREADS may have been pinned to 1 by a launcher (torch.distributed.run does that): the caller states the count.
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
