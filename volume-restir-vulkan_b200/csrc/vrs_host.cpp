// Host-side helpers with the reference's semantics: alias table and light generation
// (src/utils/restir_utils.cpp), camera matrices (nvpro_core nvmath), uniform defaults (src/Renderer.cpp:2341-2358).
#include <cmath>
#include <cstring>
#include <deque>
#include <random>
#include <vector>

#include "../../include/vrs.h"

extern "C" {

int vrs_abi_version(void) { return VRS_ABI_VERSION; }

// createAliasTable, src/utils/restir_utils.cpp:90-155: Vose construction with FIFO work lists, fp32 throughout.
void vrs_create_alias_table(const float* pdf, uint32_t n, vrs_alias_table_cell* out) {
  std::deque<int> large, small;
  std::vector<float> scaled(n);
  float power_sum = 0.f;
  for (uint32_t i = 0; i < n; ++i) power_sum += pdf[i];
  for (uint32_t i = 0; i < n; ++i) {
    out[i].alias = -1; out[i].prob = 0.f; out[i].aliasPdf = 0.f;
    out[i].pdf = pdf[i] / power_sum;
    scaled[i] = float(n) * pdf[i] / power_sum;
    (scaled[i] >= 1.f ? large : small).push_back((int)i);
  }
  while (!large.empty() && !small.empty()) {
    int g = large.front(); large.pop_front();
    int l = small.front(); small.pop_front();
    out[l].prob = scaled[l];
    out[l].alias = g;
    scaled[g] = (scaled[g] + scaled[l]) - 1.f;
    (scaled[g] < 1.f ? small : large).push_back(g);
  }
  for (int g : large) { out[g].prob = 1.f; out[g].alias = g; }
  for (int l : small) { out[l].prob = 1.f; out[l].alias = l; }
  for (uint32_t i = 0; i < n; ++i) out[i].aliasPdf = out[out[i].alias].pdf;
}

// generatePointLights, src/utils/restir_utils.cpp:22-51.  The reference draws inside a constructor argument list
// (`vec4(distX(rand), distY(rand), distZ(rand), 1)`), whose evaluation order C++ leaves open; g++ — the compiler
// the reference builds with here — goes right to left, so the draw order is Z, Y, X, then B, G, R.
void vrs_generate_point_lights(const float mn[3], const float mx[3], int white, uint32_t n, vrs_point_light* out) {
  std::uniform_real_distribution<float> distR(0.0f, 1.0f), distG(0.0f, 1.0f), distB(0.0f, 1.0f);
  std::uniform_real_distribution<float> distX(mn[0], mx[0]), distY(mn[1], mx[1]), distZ(mn[2], mx[2]);
  std::default_random_engine rand;
  for (uint32_t i = 0; i < n; ++i) {
    float z = distZ(rand), y = distY(rand), x = distX(rand);
    out[i].pos[0] = x; out[i].pos[1] = y; out[i].pos[2] = z; out[i].pos[3] = 1.0f;
    float r = 1.0f, g = 1.0f, b = 1.0f;
    if (!white) { b = distB(rand); g = distG(rand); r = distR(rand); }
    out[i].emission_luminance[0] = r; out[i].emission_luminance[1] = g; out[i].emission_luminance[2] = b;
    out[i].emission_luminance[3] = 0.2126f * r + 0.7152f * g + 0.0722f * b;   // shader::luminance, headers/common.glsl:5-7
  }
}

// nvmath::perspectiveVK, nvmath.inl:1149-1183 (column-major: element (r,c) at [c*4+r])
void vrs_perspectiveVK(float fovy, float aspect, float n, float f, float M[16]) {
  const float to_rad = float(3.14159265358979323846264338327950288419716939937510582) / float(180);
  float t = n * tanf(fovy * to_rad * float(0.5));
  float b = -t, l = b * aspect, r = t * aspect;
  memset(M, 0, 64);
  M[0] = (2 * n) / (r - l);
  M[5] = -(2 * n) / (t - b);
  M[8] = (r + l) / (r - l);
  M[9] = (t + b) / (t - b);
  M[10] = -(f) / (f - n);
  M[11] = -1;
  M[14] = (f * n) / (n - f);
}

static void unit3(float* v) {   // vector3::normalize, nvmath.inl:377-388
  float norm = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  norm = norm > 10e-6f ? 1.0f / norm : 0.0f;
  v[0] *= norm; v[1] *= norm; v[2] *= norm;
}
static void cross3(const float* v, const float* w, float* u) {
  u[0] = v[1] * w[2] - v[2] * w[1]; u[1] = v[2] * w[0] - v[0] * w[2]; u[2] = v[0] * w[1] - v[1] * w[0];
}
// nvmath::look_at, nvmath.inl:979-1025
void vrs_look_at(const float eye[3], const float center[3], const float up[3], float M[16]) {
  float z[3] = {eye[0] - center[0], eye[1] - center[1], eye[2] - center[2]};
  unit3(z);
  float x[3], y[3];
  cross3(up, z, x);
  cross3(z, x, y);
  unit3(x); unit3(y);
  const float* rows[3] = {x, y, z};
  for (int r = 0; r < 3; ++r) {
    M[0 + r] = rows[r][0]; M[4 + r] = rows[r][1]; M[8 + r] = rows[r][2];
    M[12 + r] = -rows[r][0] * eye[0] - rows[r][1] * eye[1] - rows[r][2] * eye[2];
  }
  M[3] = 0.f; M[7] = 0.f; M[11] = 0.f; M[15] = 1.f;
}

// nvmath::invert(matrix4), nvmath.inl:797-850: adjugate by 3x3 cofactors, then scale by 1/det
void vrs_invert(const float A[16], float B[16]) {
  auto a = [A](int r, int c) { return A[c * 4 + r]; };
  auto det2 = [](float a1, float a2, float b1, float b2) { return a1 * b2 - b1 * a2; };
  auto det3 = [&](float a1, float a2, float a3, float b1, float b2, float b3, float c1, float c2, float c3) {
    return a1 * det2(b2, b3, c2, c3) - b1 * det2(a2, a3, c2, c3) + c1 * det2(a2, a3, b2, b3);
  };
  float out[16];
  // B(r,c) = (-1)^(r+c) * minor of A with row c and column r removed, rows/cols listed in nvmath's argument order
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      int rr[3], cc[3], k = 0;
      for (int i = 0; i < 4; ++i) if (i != c) rr[k++] = i;   // rows of A kept (row index != c)
      k = 0;
      for (int i = 0; i < 4; ++i) if (i != r) cc[k++] = i;   // columns of A kept (col index != r)
      float m = det3(a(rr[0], cc[0]), a(rr[1], cc[0]), a(rr[2], cc[0]), a(rr[0], cc[1]), a(rr[1], cc[1]), a(rr[2], cc[1]),
                     a(rr[0], cc[2]), a(rr[1], cc[2]), a(rr[2], cc[2]));
      out[c * 4 + r] = ((r + c) & 1) ? -m : m;
    }
  float det = (a(0, 0) * out[0]) + (a(0, 1) * out[1]) + (a(0, 2) * out[2]) + (a(0, 3) * out[3]);
  float oodet = 1.0f / det;
  for (int i = 0; i < 16; ++i) B[i] = out[i] * oodet;
}

// matrix4::operator*, nvmath.inl:663-684
void vrs_mat4_mul(const float A[16], const float Bm[16], float C[16]) {
  float out[16];
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r)
      out[c * 4 + r] = A[0 + r] * Bm[c * 4 + 0] + A[4 + r] * Bm[c * 4 + 1] + A[8 + r] * Bm[c * 4 + 2] + A[12 + r] * Bm[c * 4 + 3];
  memcpy(C, out, 64);
}

void vrs_default_config(vrs_config* c, uint32_t width, uint32_t height) {
  memset(c, 0, sizeof(*c));
  c->width = width; c->height = height; c->band_y0 = 0; c->band_y1 = 0; c->halo_rows = 32; c->device = -1;
  c->spatial_iterations = 2;
  c->world_scale = 0.05f;                                              // Renderer.cpp:1420
  c->world_translate[0] = -2.5f; c->world_translate[1] = 0.5f; c->world_translate[2] = 0.0f;   // :1423
  c->density_scale = 10.0f;
  c->roughness = 0.9f; c->metallic = 0.0001f;                          // :1498-1500
  c->enable_trace = 0;
}

void vrs_default_restir_uniforms(vrs_restir_uniforms* u, uint32_t width, uint32_t height) {   // Renderer.cpp:2341-2358
  memset(u, 0, sizeof(*u));
  u->debugMode = 0; u->gamma = 4.0f;
  u->screenSize[0] = width; u->screenSize[1] = height;
  u->flags = VRS_RESTIR_VISIBILITY_REUSE_FLAG | VRS_RESTIR_TEMPORAL_REUSE_FLAG | VRS_RESTIR_SPATIAL_REUSE_FLAG;
  u->spatialNeighbors = 4; u->spatialRadius = 30.0f; u->initialLightSampleCount = 1u << 6;
  u->environmentalPower = 1.0f; u->fireflyClampThreshold = 2.0f; u->temporalSampleCountMultiplier = 20;
}

void vrs_band_for_rank(uint32_t height, int rank, int nranks, uint32_t* y0, uint32_t* y1) {
  uint64_t h = height;
  *y0 = (uint32_t)(h * (uint64_t)rank / (uint64_t)nranks);
  *y1 = (uint32_t)(h * (uint64_t)(rank + 1) / (uint64_t)nranks);
}

}  // extern "C"
