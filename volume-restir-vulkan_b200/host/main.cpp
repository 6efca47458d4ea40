// Headless driver with the reference's frame loop (src/main.cpp:209-284, 301-449): load a VDB, create lights,
// orbit the camera, run RestirPass -> SpatialReusePass -> restirDrawPost per frame, write the frame buffer.
//   vrs_render <file.vdb|file.vrsg> [width height frames lights out.ppm]
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "Renderer.h"

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s <file.vdb|file.vrsg> [width height frames lights out.ppm]\n", argv[0]); return 2; }
  uint32_t W = argc > 2 ? atoi(argv[2]) : 1280, H = argc > 3 ? atoi(argv[3]) : 720;     // main.cpp:80-81
  int frames = argc > 4 ? atoi(argv[4]) : 60;
  uint32_t nlights = argc > 5 ? atoi(argv[5]) : 64;
  const char* out = argc > 6 ? argv[6] : "frame.ppm";
  try {
    vrs_host::VDBLoader loader; loader.Load(argv[1]);                                   // main.cpp:221-223
    vrs_host::Renderer renderer; renderer.setup(W, H);
    renderer.createVDBBuffer(loader);                                                   // main.cpp:224
    vrs_grid_info gi; vrs_get_grid_info(renderer.ctx(), &gi);
    float ctr[3], ext = 0.f;
    for (int a = 0; a < 3; ++a) { ctr[a] = 0.5f * (gi.world_bbox_min[a] + gi.world_bbox_max[a]); float e = 0.5f * (gi.world_bbox_max[a] - gi.world_bbox_min[a]); ext += e * e; }
    ext = sqrtf(ext);
    std::vector<vrs_point_light> lights(nlights);
    vrs_generate_point_lights(gi.world_bbox_min, gi.world_bbox_max, 0, nlights, lights.data());
    renderer.createRestirLights(lights);                                                // main.cpp:253
    renderer.m_restirUniforms.initialLightSampleCount = 32; renderer.m_restirUniforms.spatialNeighbors = 5;
    float up[3] = {0, 1, 0}, eye[3] = {ctr[0] + 1.25f * ext, ctr[1], ctr[2]};
    renderer.CameraManip.setLookat(eye, ctr, up);
    renderer.createRestirUniformBuffer();                                               // main.cpp:256
    for (int f = 0; f < frames; ++f) {
      float a = 6.0f * f * 3.14159265f / 180.0f;
      eye[0] = ctr[0] + 1.25f * ext * cosf(a); eye[2] = ctr[2] + 1.25f * ext * sinf(a);
      renderer.CameraManip.setLookat(eye, ctr, up);
      renderer.updateUniformBuffer();                                                   // main.cpp:343
      renderer.updateRestirUniformBuffer();                                             // main.cpp:344
      renderer.updateFrame();                                                           // main.cpp:345
      renderer.getRestirPass().run();                                                   // main.cpp:405-409
      renderer.getSpatialReusePass().run();                                             // main.cpp:410-413
      renderer.restirDrawPost();                                                        // main.cpp:431
      if (renderer.getRestirPostPipelinePC()->frame > 10) renderer.getRestirPostPipelinePC()->initialize = 0;   // main.cpp:441-443
      renderer.submitFrame();                                                           // main.cpp:447
      renderer.updateGBufferFrameIdx();                                                 // main.cpp:448
    }
    renderer.writeImage(out);
    printf("wrote %s (%ux%u, %d frames, %u leaves, %llu active voxels)\n", out, W, H, frames, gi.leaves, (unsigned long long)gi.active_voxels);
  } catch (const std::exception& e) { fprintf(stderr, "error: %s\n", e.what()); return 1; }
  return 0;
}
