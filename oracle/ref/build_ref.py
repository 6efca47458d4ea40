#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY. Builds oracle/_ref/libvrs_ref.so from the REFERENCE'S OWN SOURCES.

Nothing is copied into the repository: the reference files are read where they
lie under /root/reference, rewritten lexically into a temporary directory, and
compiled by g++; only the shared library lands in oracle/_ref/ (git-ignored).

Two translation units:

  ref_glsl.cpp  the reference GLSL headers
                  src/shaders/headers/{common,math,random,disneyBRDF,restirUtils,reservoir}.glsl
                  src/shaders/structs/{light,restirStructs}.glsl
                plus three text ranges of entry shaders
                  src/shaders/restir.rgen:97-134   (aliasTableSample, SceneSample)
                  src/shaders/restir.rgen:205-227  (initial RIS loop body)
                  src/shaders/restir_post.frag:78-92 (shade, emissive override, firefly clamp)
                compiled as C++ against oracle/ref/glsl_shim.h.
                Lexical rewrite (semantics-preserving for GLSL 4.60):
                  * `inout T x`/`out T x` -> `T& x`; `in T x` -> `const T& x`
                  * un-suffixed floating literals get an `f` (GLSL literals are fp32)
                  * `.xyz`/`.xy` rvalue swizzles -> `.xyz()`/`.xy()`
                  * `F(rnd(seed), rnd(seed), ...` -> order-independent helpers that
                    realise GLSL's left-to-right argument evaluation (4.60 §6.1.1)
  ref_host.cpp  src/utils/restir_utils.cpp:22-51 (generatePointLights) and :90-155
                (createAliasTable) verbatim, against the reference's own
                src/shaders/host_device.h + src/utils/shader_functions.hpp and
                nvpro_core's nvmath (perspectiveVK / look_at / invert).

Flags: -O2 -ffp-contract=off (no FMA contraction, no fast-math).
"""
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VRS_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(os.path.dirname(HERE), "_ref")
OUT = os.path.join(OUT_DIR, "libvrs_ref.so")


def read_lines(rel, first=None, last=None, expect_first=None, expect_last=None):
    with open(os.path.join(REF, rel)) as f:
        lines = f.read().split("\n")
    if first is None:
        return "\n".join(lines)
    sel = lines[first - 1:last]
    if expect_first is not None:
        assert expect_first in sel[0], (rel, first, sel[0])
    if expect_last is not None:
        assert expect_last in sel[-1], (rel, last, sel[-1])
    return "\n".join(sel)


def strip_comments(s):
    s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
    s = re.sub(r"//[^\n]*", "", s)
    return s


_FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?)(?![\w.])")


def glsl_to_cpp(s):
    s = strip_comments(s)
    s = re.sub(r"^\s*#\s*include[^\n]*$", "", s, flags=re.M)
    s = re.sub(r"^\s*#\s*(version|extension)[^\n]*$", "", s, flags=re.M)
    s = re.sub(r"\binout[ \t]+(\w+)[ \t]+(\w+)", r"\1& \2", s)
    s = re.sub(r"\bout[ \t]+(\w+)[ \t]+(\w+)", r"\1& \2", s)
    s = re.sub(r"\bin[ \t]+(\w+)[ \t]+(\w+)", r"const \1& \2", s)
    s = _FLOAT_LIT.sub(r"\1f", s)
    s = re.sub(r"\.xyz\b(?!\s*\()", ".xyz()", s)
    s = re.sub(r"\.xy\b(?!\s*\()", ".xy()", s)
    # GLSL evaluates call arguments left to right; C++ does not promise it.
    s = re.sub(r"\(\s*rnd\(seed\),\s*rnd\(seed\),", "(GLSL_LR_FIRST(seed), GLSL_LR_SECOND(seed),", s)
    return s


LR_HELPERS = r"""
// Left-to-right evaluation of `f(rnd(seed), rnd(seed), ...)` whatever order the
// C++ compiler picks for the two argument expressions.
static thread_local float lr_a, lr_b; static thread_local int lr_have = 0;
float rnd(uint& seed);
static inline void lr_fill(uint& seed) { lr_a = rnd(seed); lr_b = rnd(seed); }
static inline float GLSL_LR_FIRST(uint& seed)  { if (lr_have) { lr_have = 0; return lr_a; } lr_fill(seed); lr_have = 1; return lr_a; }
static inline float GLSL_LR_SECOND(uint& seed) { if (lr_have) { lr_have = 0; return lr_b; } lr_fill(seed); lr_have = 1; return lr_b; }
"""

GLSL_EXPORTS = r"""
PointLightsSSBO pointLights; TriangleLightsSSBO triangleLights; AliasTableSSBO aliasTable;
RestirUniformSubset restirUniform;

struct RefGInfo { float camPos[3], worldPos[3], normal[3], albedo[4], emissive[3];
                  float albedoLum, roughness, metallic; uint sampleSeed; };
struct RefRes { float lightPos[3]; uint numStreamSamples, lightIndex; int lightKind; uint sampleSeed;
                float pHat, sumWeights, w; };

static GeometryInfo toG(const RefGInfo* g) {
  GeometryInfo o;
  o.camPos = vec3(g->camPos[0], g->camPos[1], g->camPos[2]);
  o.worldPos = vec3(g->worldPos[0], g->worldPos[1], g->worldPos[2]);
  o.normal = vec3(g->normal[0], g->normal[1], g->normal[2]);
  o.albedo = vec4(g->albedo[0], g->albedo[1], g->albedo[2], g->albedo[3]);
  o.emissive = vec3(g->emissive[0], g->emissive[1], g->emissive[2]);
  o.albedoLum = g->albedoLum; o.roughness = g->roughness; o.metallic = g->metallic;
  o.sampleSeed = g->sampleSeed;
  return o;
}
static Reservoir toR(const RefRes* r) {
  Reservoir o;
  o.lightPos = vec3(r->lightPos[0], r->lightPos[1], r->lightPos[2]);
  o.numStreamSamples = r->numStreamSamples; o.lightIndex = r->lightIndex; o.lightKind = r->lightKind;
  o.sampleSeed = r->sampleSeed; o.pHat = r->pHat; o.sumWeights = r->sumWeights; o.w = r->w;
  return o;
}
static void fromR(const Reservoir& o, RefRes* r) {
  r->lightPos[0] = o.lightPos.x; r->lightPos[1] = o.lightPos.y; r->lightPos[2] = o.lightPos.z;
  r->numStreamSamples = o.numStreamSamples; r->lightIndex = o.lightIndex; r->lightKind = o.lightKind;
  r->sampleSeed = o.sampleSeed; r->pHat = o.pHat; r->sumWeights = o.sumWeights; r->w = o.w;
}

// restir.rgen:205-227 wrapped as a function (text inserted by build_ref.py)
struct RisUniforms { int initialLightSampleCount; };
static void ref_ris_body(GeometryInfo& gInfo, Reservoir& res, uint& seed, int count) {
  RisUniformsShadow
@RIS_LOOP@
}
// restir_post.frag:78-92 wrapped as a function
static void ref_post_body(const Reservoir& res, GeometryInfo& gInfo, float fireflyClampThreshold, vec3& outColor) {
  struct { float fireflyClampThreshold; } uniforms = { fireflyClampThreshold };
@POST_BODY@
}

extern "C" {
void ref_set_scene(const PointLight* pl, int npl, const TriangleLight* tl, const AliasTableCell* at, int nat) {
  pointLights.lights = pl; triangleLights.lights = tl; aliasTable.aliasCol = at;
  restirUniform.aliasTableCount = nat; restirUniform.pointLightCount = npl;
}
void ref_pcg2d(uint x, uint y, uint* o) { uvec2 v = pcg2d(uvec2(x, y)); o[0] = v.x; o[1] = v.y; }
uint ref_lcg(uint* s) { return lcg(*s); }
uint ref_pcg(uint* s) { return pcg(*s); }
uint ref_tea(uint a, uint b) { return tea(a, b); }
float ref_rnd(uint* s) { return rnd(*s); }
float ref_luminance_common(float r, float g, float b) { return luminance(r, g, b); }
float ref_luminance_utils(float r, float g, float b) { return luminance(vec3(r, g, b)); }
float ref_disneyBrdfLuminance(float cosIn, float cosOut, float cosHalf, float cosInHalf, float lum, float rough, float metal) {
  return disneyBrdfLuminance(cosIn, cosOut, cosHalf, cosInHalf, lum, rough, metal);
}
void ref_disneyBrdfColor(float cosIn, float cosOut, float cosHalf, float cosInHalf, const float* albedo, float rough, float metal, float* out) {
  vec3 c = disneyBrdfColor(cosIn, cosOut, cosHalf, cosInHalf, vec3(albedo[0], albedo[1], albedo[2]), rough, metal);
  out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
float ref_evaluatePHat(uint idx, int kind, const RefGInfo* g) { return evaluatePHat(idx, kind, toG(g)); }
void ref_evaluatePHatFull(uint idx, int kind, const RefGInfo* g, float* out) {
  vec3 c = evaluatePHatFull(idx, kind, toG(g)); out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
void ref_offsetRay(const float* p, const float* n, float* out) {
  vec3 o = OffsetRay(vec3(p[0], p[1], p[2]), vec3(n[0], n[1], n[2])); out[0] = o.x; out[1] = o.y; out[2] = o.z;
}
void ref_newReservoir(RefRes* r) { Reservoir o = newReservoir(); o.lightIndex = 0; o.lightKind = 0; o.sampleSeed = 0; o.lightPos = vec3(0.0f); fromR(o, r); }
void ref_updateReservoir(RefRes* r, uint idx, int kind, float weight, float pHat, float w, const float* lp, uint* seed, uint sampleSeed) {
  Reservoir o = toR(r); updateReservoir(o, idx, kind, weight, pHat, w, vec3(lp[0], lp[1], lp[2]), *seed, sampleSeed); fromR(o, r);
}
void ref_addSampleToReservoir(RefRes* r, uint idx, int kind, float pdf, const float* lp, const RefGInfo* g, uint* seed) {
  Reservoir o = toR(r); addSampleToReservoir(o, idx, kind, pdf, vec3(lp[0], lp[1], lp[2]), toG(g), *seed); fromR(o, r);
}
void ref_combineReservoirs_geom(RefRes* self, const RefRes* other, const RefGInfo* g, const RefGInfo* og, uint* seed) {
  Reservoir o = toR(self); combineReservoirs(o, toR(other), toG(g), toG(og), *seed); fromR(o, self);
}
void ref_combineReservoirs_plain(RefRes* self, const RefRes* other, float pHat, uint* seed) {
  Reservoir o = toR(self); combineReservoirs(o, toR(other), pHat, *seed); fromR(o, self);
}
void ref_packReservoir(const RefRes* r, float* info, float* weight) {
  vec4 a, b; packReservoirStruct(toR(r), a, b);
  info[0] = a.x; info[1] = a.y; info[2] = a.z; info[3] = a.w; weight[0] = b.x; weight[1] = b.y; weight[2] = b.z; weight[3] = b.w;
}
void ref_unpackReservoir(const float* info, const float* weight, RefRes* r) {
  Reservoir o = unpackReservoirStruct(vec4(info[0], info[1], info[2], info[3]), vec4(weight[0], weight[1], weight[2], weight[3]));
  o.lightPos = vec3(0.0f); fromR(o, r);
}
void ref_aliasTableSample(float r1, float r2, uint* index, float* prob) { aliasTableSample(r1, r2, *index, *prob); }
// restir.rgen:203-227: newReservoir + the RIS loop, then pack (restir.rgen:286-287).
void ref_initial_ris(const RefGInfo* g, int count, uint* seed, RefRes* out) {
  GeometryInfo gi = toG(g); Reservoir res = newReservoir();
  res.lightIndex = 0; res.lightKind = 0; res.sampleSeed = 0; res.lightPos = vec3(0.0f);
  ref_ris_body(gi, res, *seed, count); fromR(res, out);
}
void ref_post_shade(const RefRes* r, const RefGInfo* g, float thr, float* out) {
  GeometryInfo gi = toG(g); gi.sampleSeed = r->sampleSeed; vec3 c(0.0f);
  ref_post_body(toR(r), gi, thr, c); out[0] = c.x; out[1] = c.y; out[2] = c.z;
}
void ref_post_accumulate(const float* oldc, const float* newc, int frame, int initialize, float* out) {
  // restir_post.frag:94-102
  vec3 o(oldc[0], oldc[1], oldc[2]), n(newc[0], newc[1], newc[2]);
  vec3 r = n;
  if (!(frame < 1 || initialize == 1)) { float a = 1.0f / float(frame); r = mix(o, n, a); }
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
}
"""

HOST_TU = r"""
#include <cstdint>
#include <cstddef>
#include <queue>
#include <random>
#include <vector>
#include "shaders/host_device.h"
#include "utils/shader_functions.hpp"
@HOST_FUNCS@
extern "C" {
void ref_createAliasTable(const float* pdf, int n, AliasTableCell* out) {
  std::vector<float> v(pdf, pdf + n);
  std::vector<AliasTableCell> t = createAliasTable(v);
  for (int i = 0; i < n; ++i) out[i] = t[i];
}
void ref_generatePointLights(const float* mn, const float* mx, int white, uint32_t n, float* out) {
  std::vector<PointLight> l = generatePointLights(nvmath::vec3(mn[0], mn[1], mn[2]), nvmath::vec3(mx[0], mx[1], mx[2]), white != 0, n);
  for (uint32_t i = 0; i < n; ++i) {
    out[8 * i + 0] = l[i].pos.x; out[8 * i + 1] = l[i].pos.y; out[8 * i + 2] = l[i].pos.z; out[8 * i + 3] = l[i].pos.w;
    out[8 * i + 4] = l[i].emission_luminance.x; out[8 * i + 5] = l[i].emission_luminance.y;
    out[8 * i + 6] = l[i].emission_luminance.z; out[8 * i + 7] = l[i].emission_luminance.w;
  }
}
void ref_perspectiveVK(float fovy, float aspect, float n, float f, float* out) {
  nvmath::mat4f m = nvmath::perspectiveVK(fovy, aspect, n, f); for (int i = 0; i < 16; ++i) out[i] = m.mat_array[i];
}
void ref_look_at(const float* e, const float* c, const float* u, float* out) {
  nvmath::mat4f m = nvmath::look_at(nvmath::vec3f(e[0], e[1], e[2]), nvmath::vec3f(c[0], c[1], c[2]), nvmath::vec3f(u[0], u[1], u[2]));
  for (int i = 0; i < 16; ++i) out[i] = m.mat_array[i];
}
void ref_invert(const float* a, float* out) {
  nvmath::mat4f m; for (int i = 0; i < 16; ++i) m.mat_array[i] = a[i];
  nvmath::mat4f r = nvmath::invert(m); for (int i = 0; i < 16; ++i) out[i] = r.mat_array[i];
}
void ref_matmul(const float* a, const float* b, float* out) {
  nvmath::mat4f x, y; for (int i = 0; i < 16; ++i) { x.mat_array[i] = a[i]; y.mat_array[i] = b[i]; }
  nvmath::mat4f r = x * y; for (int i = 0; i < 16; ++i) out[i] = r.mat_array[i];
}
void ref_mat_vec(const float* a, const float* v, float* out) {
  nvmath::mat4f x; for (int i = 0; i < 16; ++i) x.mat_array[i] = a[i];
  nvmath::vec4f r = x * nvmath::vec4f(v[0], v[1], v[2], v[3]); out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
// voxel material: the reference's expressions at vdb/vdb.cpp:816-817 (smoke colour) and
// Renderer.cpp:1494-1496 (pbrBaseColorFactor), evaluated with the reference's nvmath.
void ref_voxel_albedo(float value, float* out4) {
  nvmath::vec3f smokeColor = nvmath::normalize(nvmath::vec3f(100, 100, 100)) * (float)value * 1000.f;
  nvmath::vec4f c = nvmath::normalize(nvmath::vec4(smokeColor[0], smokeColor[1], smokeColor[2], 1));
  out4[0] = c.x; out4[1] = c.y; out4[2] = c.z; out4[3] = c.w;
}
// layout pins for include/vrs.h (host_device.h:116-120,139-145,184-227)
void ref_struct_layout(int* o) {
  o[0] = (int)sizeof(RestirUniforms); o[1] = (int)offsetof(RestirUniforms, spatialNeighbors);
  o[2] = (int)offsetof(RestirUniforms, screenSize); o[3] = (int)offsetof(RestirUniforms, currCamPos);
  o[4] = (int)offsetof(RestirUniforms, currFrameProjectionViewMatrix); o[5] = (int)offsetof(RestirUniforms, prevCamPos);
  o[6] = (int)offsetof(RestirUniforms, prevFrameProjectionViewMatrix); o[7] = (int)offsetof(RestirUniforms, flags);
  o[8] = (int)offsetof(RestirUniforms, gamma); o[9] = (int)sizeof(GlobalUniforms); o[10] = (int)sizeof(PointLight);
  o[11] = (int)sizeof(TriangleLight); o[12] = (int)sizeof(AliasTableCell); o[13] = (int)sizeof(PushConstantRestir);
  o[14] = (int)offsetof(RestirUniforms, initialLightSampleCount); o[15] = (int)offsetof(RestirUniforms, temporalSampleCountMultiplier);
}
}
"""


def build():
    if not os.path.isdir(os.path.join(REF, "src", "shaders")):
        print("build_ref: %s not present; skipping (prebuilt oracle/_ref is used if it exists)" % REF)
        return False
    sh = "src/shaders/"
    parts = []
    for rel in ["structs/light.glsl", "headers/common.glsl", "headers/math.glsl", "headers/random.glsl",
                "headers/disneyBRDF.glsl", "structs/restirStructs.glsl", "headers/restirUtils.glsl",
                "headers/reservoir.glsl"]:
        parts.append("// ---- %s%s ----\n" % (sh, rel) + glsl_to_cpp(read_lines(sh + rel)))
    parts.append(glsl_to_cpp(read_lines(sh + "restir.rgen", 97, 134, "void aliasTableSample", "}")))
    ris = glsl_to_cpp(read_lines(sh + "restir.rgen", 205, 227, "if (dot(gInfo.normal, gInfo.normal)", "}"))
    post = glsl_to_cpp(read_lines(sh + "restir_post.frag", 78, 92, "uint lightIndex", "outColor = max"))
    exports = GLSL_EXPORTS.replace("RisUniformsShadow", "RisUniforms restirUniform = { count };")
    exports = exports.replace("@RIS_LOOP@", ris).replace("@POST_BODY@", post)
    glsl_tu = ('#include "glsl_shim.h"\n#define CPP_FUNCTION inline\n#define COMMON_HOST_DEVICE 1\n' + LR_HELPERS +
               "\n".join(parts) + exports)

    host_funcs = (read_lines("src/utils/restir_utils.cpp", 22, 51, "generatePointLights", "}") + "\n" +
                  read_lines("src/utils/restir_utils.cpp", 90, 155, "createAliasTable", "}"))
    host_tu = HOST_TU.replace("@HOST_FUNCS@", host_funcs)

    os.makedirs(OUT_DIR, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="vrs_ref_") as tmp:
        with open(os.path.join(tmp, "ref_glsl.cpp"), "w") as f:
            f.write(glsl_tu)
        with open(os.path.join(tmp, "ref_host.cpp"), "w") as f:
            f.write(host_tu)
        common = ["g++", "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-w"]
        subprocess.check_call(common + ["-I", HERE, "-c", os.path.join(tmp, "ref_glsl.cpp"), "-o", os.path.join(tmp, "a.o")])
        subprocess.check_call(common + ["-I", os.path.join(REF, "src"), "-I", os.path.join(REF, "external", "nvpro_core"),
                                        "-I", os.path.join(REF, "external", "nvpro_core", "nvp"),
                                        "-c", os.path.join(tmp, "ref_host.cpp"), "-o", os.path.join(tmp, "b.o")])
        subprocess.check_call(["g++", "-shared", "-o", OUT, os.path.join(tmp, "a.o"), os.path.join(tmp, "b.o")])
        if os.environ.get("VRS_REF_KEEP"):
            import shutil
            shutil.copy(os.path.join(tmp, "ref_glsl.cpp"), "/tmp/ref_glsl_debug.cpp")
    print("build_ref: wrote", OUT)
    return True


if __name__ == "__main__":
    sys.exit(0 if build() or os.path.exists(OUT) else 1)
