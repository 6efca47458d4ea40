// C++ host façade over the C ABI (include/vrs.h) with the reference's class and method names, so that the call
// order of the reference's main loop (src/main.cpp:209-284 init, :301-449 per frame) carries over unchanged.
// Only the ReSTIR slice of the reference's Renderer exists here (src/Renderer.h:64-165); Vulkan plumbing has no
// equivalent on a B200 and is gone.  Header-only; everything forwards to libvrs.so.
#pragma once
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "vrs.h"

namespace vrs_host {

inline void check(vrs_status s, vrs_ctx* ctx, const char* what) {
  if (s != VRS_OK) throw std::runtime_error(std::string(what) + ": " + vrs_last_error(ctx));
}

// nvh::CameraManipulator subset (main.cpp:90-101; default fov 60, nvh/cameramanipulator.hpp:93)
struct CameraManipulator {
  float eye[3] = {1, 1, 1}, center[3] = {0, 1, 0}, up[3] = {0, 1, 0}, fov = 60.0f;
  void setLookat(const float e[3], const float c[3], const float u[3]) { memcpy(eye, e, 12); memcpy(center, c, 12); memcpy(up, u, 12); }
  void getMatrix(float out16[16]) const { vrs_look_at(eye, center, up, out16); }
  float getFov() const { return fov; }
};

// src/loaders/VDBLoader.hpp:8-25
class VDBLoader {
 public:
  void Load(const std::string& filename) { file_ = filename; is_vdb_loaded_ = true; }
  bool IsVDBLoaded() const { return is_vdb_loaded_; }
  const std::string& file() const { return file_; }
 private:
  std::string file_; bool is_vdb_loaded_ = false;
};

class Renderer;

// src/passes/restirPass.h:13-63 — the Vulkan pipeline objects are gone; run() enqueues kernel A on the context stream.
class RestirPass {
 public:
  void setup(Renderer* r) { r_ = r; }
  void run();
  void destroy() {}
 private:
  Renderer* r_ = nullptr;
};
// src/passes/spatialReusePass.h:12-47
class SpatialReusePass {
 public:
  void setup(Renderer* r) { r_ = r; }
  void run();
  void destroy() {}
 private:
  Renderer* r_ = nullptr;
};

class Renderer {
 public:
  CameraManipulator CameraManip;

  void setup(uint32_t width, uint32_t height, uint32_t spatial_iterations = 2, float density_scale = 10.0f) {
    vrs_config cfg; vrs_default_config(&cfg, width, height);
    cfg.spatial_iterations = spatial_iterations; cfg.density_scale = density_scale;
    check(vrs_create(&cfg, &ctx_), nullptr, "vrs_create");
    width_ = width; height_ = height; spatial_iterations_ = spatial_iterations;
    vrs_default_restir_uniforms(&m_restirUniforms, width, height);
    m_restirPass.setup(this); m_spatialReusePass.setup(this);
  }
  void destroy() { if (ctx_) vrs_destroy(ctx_); ctx_ = nullptr; }
  // Renderer::onResize (Renderer.cpp:1022-1026) + the passes' createRenderPass(VkExtent2D) (restirPass.cpp:79-81, spatialReusePass.cpp:41-43)
  void onResize(int w, int h) {
    check(vrs_resize(ctx_, (uint32_t)w, (uint32_t)h), ctx_, "vrs_resize");
    width_ = (uint32_t)w; height_ = (uint32_t)h;
    m_restirUniforms.screenSize[0] = width_; m_restirUniforms.screenSize[1] = height_;
    have_ref_ = false;
  }
  ~Renderer() { destroy(); }

  // Renderer::createVDBBuffer (Renderer.cpp:1408-1582): flatten + stage the grid instead of building spheres
  void createVDBBuffer(const VDBLoader& loader, const char* grid = nullptr) { check(vrs_load_vdb(ctx_, loader.file().c_str(), grid), ctx_, "vrs_load_vdb"); }
  // Renderer::createRestirLights (Renderer.cpp:1587-1691)
  void createRestirLights(const std::vector<vrs_point_light>& lights) {
    check(vrs_set_lights(ctx_, lights.data(), (uint32_t)lights.size()), ctx_, "vrs_set_lights");
    m_restirUniforms.pointLightCount = (int32_t)lights.size(); m_restirUniforms.triangleLightCount = 0;
    m_restirUniforms.aliasTableCount = (int32_t)lights.size();
  }
  // Renderer.cpp:2339-2374
  void createRestirUniformBuffer() {
    float pv[16]; projView(pv);
    for (int a = 0; a < 3; ++a) m_restirUniforms.currCamPos[a] = m_restirUniforms.prevCamPos[a] = CameraManip.eye[a];
    memcpy(m_restirUniforms.currFrameProjectionViewMatrix, pv, 64); memcpy(m_restirUniforms.prevFrameProjectionViewMatrix, pv, 64);
  }
  // Renderer.cpp:116-161
  void updateUniformBuffer() {
    float view[16], proj[16]; CameraManip.getMatrix(view);
    vrs_perspectiveVK(CameraManip.getFov(), (float)width_ / (float)height_, 0.1f, 1000.0f, proj);
    vrs_mat4_mul(proj, view, m_globalUniforms.viewProj); vrs_invert(view, m_globalUniforms.viewInverse); vrs_invert(proj, m_globalUniforms.projInverse);
  }
  // Renderer.cpp:2376-2425
  void updateRestirUniformBuffer() {
    memcpy(m_restirUniforms.prevCamPos, m_restirUniforms.currCamPos, 16);
    memcpy(m_restirUniforms.prevFrameProjectionViewMatrix, m_restirUniforms.currFrameProjectionViewMatrix, 64);
    m_restirUniforms.screenSize[0] = width_; m_restirUniforms.screenSize[1] = height_;
    for (int a = 0; a < 3; ++a) m_restirUniforms.currCamPos[a] = CameraManip.eye[a];
    projView(m_restirUniforms.currFrameProjectionViewMatrix);
  }
  // Renderer.cpp:2456-2472
  void updateFrame() {
    float m[16]; CameraManip.getMatrix(m);
    if (!have_ref_ || memcmp(refCam_, m, 64) != 0 || refFov_ != CameraManip.getFov()) { resetFrame(); memcpy(refCam_, m, 64); refFov_ = CameraManip.getFov(); have_ref_ = true; }
    m_pcRestirPost.frame++;
  }
  void resetFrame() { m_pcRestirPost.frame = -1; }
  RestirPass& getRestirPass() { return m_restirPass; }
  SpatialReusePass& getSpatialReusePass() { return m_spatialReusePass; }
  vrs_push_constant_restir* getRestirPostPipelinePC() { return &m_pcRestirPost; }
  // Renderer::restirDrawPost (Renderer.cpp:1247-1264) + updateGBufferFrameIdx (:108-111, done inside vrs_pass_shade)
  void restirDrawPost() { check(vrs_pass_shade(ctx_, &m_restirUniforms, &m_pcRestirPost, clock_), ctx_, "vrs_pass_shade"); }
  void updateGBufferFrameIdx() { ++clock_; }
  void submitFrame() { check(vrs_synchronize(ctx_), ctx_, "vrs_synchronize"); }
  void writeImage(const std::string& path) { check(vrs_write_image(ctx_, path.c_str()), ctx_, "vrs_write_image"); }

  vrs_ctx* ctx() { return ctx_; }
  uint32_t clock() const { return clock_; }
  uint32_t spatialIterations() const { return spatial_iterations_; }
  vrs_restir_uniforms m_restirUniforms{};
  vrs_global_uniforms m_globalUniforms{};
  vrs_push_constant_restir m_pcRestirPost{0.f, 0.f, 0.f, 0, 1};
  uint32_t spatial_iterations_ = 2;

 private:
  void projView(float out[16]) {
    float view[16], proj[16]; CameraManip.getMatrix(view);
    vrs_perspectiveVK(CameraManip.getFov(), (float)width_ / (float)height_, 0.1f, 1000.0f, proj);
    vrs_mat4_mul(proj, view, out);
  }
  vrs_ctx* ctx_ = nullptr;
  uint32_t width_ = 0, height_ = 0, clock_ = 0;
  float refCam_[16] = {0}, refFov_ = 0.f; bool have_ref_ = false;
  RestirPass m_restirPass; SpatialReusePass m_spatialReusePass;
};

inline void RestirPass::run() {
  check(vrs_pass_initial(r_->ctx(), &r_->m_globalUniforms, &r_->m_restirUniforms, r_->clock()), r_->ctx(), "vrs_pass_initial");
}
inline void SpatialReusePass::run() {
  if (!(r_->m_restirUniforms.flags & VRS_RESTIR_SPATIAL_REUSE_FLAG)) return;
  for (uint32_t it = 0; it < r_->spatialIterations(); ++it)
    check(vrs_pass_spatial(r_->ctx(), &r_->m_restirUniforms, r_->clock(), it), r_->ctx(), "vrs_pass_spatial");
}

}  // namespace vrs_host
