// Launch interface between the host runtime (vrs_api.cu) and the kernels (vrs_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vrs {
struct GridDev;
struct LightsDev;
struct FrameParams;
struct Planes;
struct ResPlanes;
struct Queues;
struct HaloPush;
struct PrevAccess;

// Optional per-kernel timing of an eagerly launched frame: one event after every kernel on the launching stream.
struct KTimer {
  static constexpr int MAX = 47;
  cudaEvent_t ev[MAX + 1];
  const char* name[MAX];
  int n;
  bool on;
};
inline void ktick(KTimer* kt, cudaStream_t s, const char* name) {
  if (!kt || !kt->on || kt->n >= KTimer::MAX) return;
  kt->name[kt->n] = name;
  cudaEventRecord(kt->ev[++kt->n], s);
}

void launch_front_trace(cudaStream_t st, const GridDev& G, const FrameParams& F, const FrameParams* dF, Planes cur, const Queues& Q, uint32_t* trace,
                        int y0, int y1, int store_y0, int store_y1, int persistent_blocks, KTimer* kt);
void launch_front_ris(cudaStream_t st, const GridDev& G, const LightsDev& L, const FrameParams& F, const FrameParams* dF, Planes cur,
                      ResPlanes outR, const Queues& Q, uint32_t* trace, int store_y0, int persistent_blocks, KTimer* kt);
void launch_initial_finish(cudaStream_t st, const LightsDev& L, const FrameParams& F, const FrameParams* dF, Planes cur, Planes prev, ResPlanes prevR,
                           ResPlanes outR, const Queues& Q, uint32_t* trace, int store_y0, const PrevAccess& PA, unsigned* out_of_halo, KTimer* kt);
int front_trace_launches(bool culling);
int front_ris_launches(int flags, const LightsDev& L);
int front_shadow_launches(int flags);
void launch_front_shadow(cudaStream_t st, const GridDev& G, const FrameParams& F, const Queues& Q, int persistent_blocks, KTimer* kt);
void launch_spatial(cudaStream_t s, const LightsDev& L, const FrameParams* dF, Planes cur, ResPlanes inR, ResPlanes outR, const Queues& Q,
                    uint32_t iteration, int store_y0, int store_y1, KTimer* kt);
void launch_shade(cudaStream_t s, const GridDev& G, const LightsDev& L, const FrameParams& F, const FrameParams* dF, Planes cur, ResPlanes rs,
                  float4* accum, int y0, int y1, int store_y0, KTimer* kt);
void launch_halo_push(cudaStream_t s, const HaloPush& H, int blocks, KTimer* kt);
void launch_halo_wait(cudaStream_t s, const unsigned* serial, const unsigned* from_up, const unsigned* from_down, unsigned* error, KTimer* kt);
void launch_export(cudaStream_t s, Planes cur, ResPlanes rs, float4* out6, size_t first_pix, size_t n);
void launch_display(cudaStream_t s, const float4* accum, uchar4* out, size_t n);
void launch_sample_density(cudaStream_t s, const GridDev& G, const int* ijk, uint32_t n, float* out);
}  // namespace vrs
