// TEST INFRASTRUCTURE ONLY — never linked into the product library.
//
// Minimal GLSL-on-C++ vocabulary so that the reference's own shader headers
// (/root/reference/src/shaders/headers/*.glsl, structs/*.glsl) can be compiled
// by g++ after the purely lexical rewrite done in build_ref.py (parameter
// qualifiers -> C++ references, fp literals -> float, swizzles -> methods).
// Semantics follow the GLSL 4.60 spec where it fixes them:
//   mix(x,y,a)   = x*(1-a) + y*a                      (spec 8.3)
//   clamp(x,a,b) = min(max(x,a),b)                    (spec 8.3)
//   normalize(v) = v / sqrt(dot(v,v))   (spec leaves the form open; we fix it)
//   dot          = left-to-right sum of products
// All arithmetic is fp32; build with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#undef M_PI
#undef M_1_PI

typedef unsigned int uint;

struct vec2 {
  float x, y;
  vec2() : x(0), y(0) {}
  vec2(float a, float b) : x(a), y(b) {}
};
struct vec3 {
  float x, y, z;
  vec3() : x(0), y(0), z(0) {}
  explicit vec3(float a) : x(a), y(a), z(a) {}
  vec3(float a, float b, float c) : x(a), y(b), z(c) {}
  vec3& operator/=(float s) { x = x / s; y = y / s; z = z / s; return *this; }
  vec3& operator*=(float s) { x = x * s; y = y * s; z = z * s; return *this; }
  vec3& operator+=(const vec3& o) { x = x + o.x; y = y + o.y; z = z + o.z; return *this; }
};
struct vec4 {
  float x, y, z, w;
  vec4() : x(0), y(0), z(0), w(0) {}
  explicit vec4(float a) : x(a), y(a), z(a), w(a) {}
  vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
  vec4(const vec3& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
  vec3 xyz() const { return vec3(x, y, z); }
  vec2 xy() const { return vec2(x, y); }
};
struct ivec3 {
  int x, y, z;
  ivec3(float a, float b, float c) : x(int(a)), y(int(b)), z(int(c)) {}
};
struct uvec2 {
  uint x, y;
  uvec2() : x(0), y(0) {}
  uvec2(uint a, uint b) : x(a), y(b) {}
};

inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }

inline uvec2 operator*(const uvec2& a, uint s) { return uvec2(a.x * s, a.y * s); }
inline uvec2 operator+(const uvec2& a, uint s) { return uvec2(a.x + s, a.y + s); }
inline uvec2 operator^(const uvec2& a, const uvec2& b) { return uvec2(a.x ^ b.x, a.y ^ b.y); }
inline uvec2 operator>>(const uvec2& a, uint s) { return uvec2(a.x >> s, a.y >> s); }

inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float sqrt(float x) { return ::sqrtf(x); }
inline float pow(float x, float y) { return ::powf(x, y); }
inline float abs(float x) { return ::fabsf(x); }
inline float max(float a, float b) { return a < b ? b : a; }  // GLSL: y if x < y
inline float min(float a, float b) { return b < a ? b : a; }  // GLSL: y if y < x
inline uint min(uint a, uint b) { return b < a ? b : a; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline vec3 mix(const vec3& x, const vec3& y, float a) {
  return vec3(mix(x.x, y.x, a), mix(x.y, y.y, a), mix(x.z, y.z, a));
}
inline vec3 max(const vec3& a, const vec3& b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline vec3 normalize(const vec3& v) { return v / ::sqrtf(dot(v, v)); }
inline float length(const vec3& v) { return ::sqrtf(dot(v, v)); }

inline uint floatBitsToUint(float f) { uint u; std::memcpy(&u, &f, 4); return u; }
inline int floatBitsToInt(float f) { int u; std::memcpy(&u, &f, 4); return u; }
inline float uintBitsToFloat(uint u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float intBitsToFloat(int u) { float f; std::memcpy(&f, &u, 4); return f; }

// Plain-C mirrors of the reference's host/device structs
// (src/shaders/host_device.h:184-202), scalar block layout.
struct PointLight { vec4 pos; vec4 emission_luminance; };
struct TriangleLight { vec4 p1, p2, p3, emission_luminance, normalArea; };
struct AliasTableCell { int alias; float prob; float pdf; float aliasPdf; };

// Storage-buffer globals the shaders index (restir.rgen:21-32).
struct PointLightsSSBO { const PointLight* lights; };
struct TriangleLightsSSBO { const TriangleLight* lights; };
struct AliasTableSSBO { const AliasTableCell* aliasCol; };
struct RestirUniformSubset { int aliasTableCount; int pointLightCount; };
extern PointLightsSSBO pointLights;
extern TriangleLightsSSBO triangleLights;
extern AliasTableSSBO aliasTable;
extern RestirUniformSubset restirUniform;
