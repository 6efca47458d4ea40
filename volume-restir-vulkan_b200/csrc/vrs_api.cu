// C ABI implementation: context, HBM-resident buffers, pass sequencing on one CUDA stream, readbacks.
// Mirrors the resource ownership of the reference's Renderer (src/Renderer.cpp:100-111 G-buffers x2 ping-pong,
// :673-761 reservoirs x2 + tmp + storage image, :1587-1691 lights + alias table, :1977-2040 ping-pong wiring)
// and the call order of src/main.cpp:405-448.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "../../include/vrs.h"
#include "vrs_comm.h"
#include "vrs_device.cuh"
#include "vrs_grid.h"
#include "vrs_kernels.h"

using namespace vrs;

#define VRS_PARAM_SLOTS 8
// Frames in flight (the reference keeps 2-3 swapchain images in flight, nvvk/appbase_vk.cpp:412-418).  A frame is four
// stages on four streams: A (coverage, classification, primary event, hit list), B (RIS candidates), C (shadow rays) — nothing
// in A, B or C reads the previous frame — and the back half (temporal merge, spatial reuse, shade, halo exchanges).  A(n + 3),
// B(n + 2), C(n + 1) and back(n) may execute at the same time; each kernel's ramp and tail is filled by the other stages' kernels.
// Buffers are sized for that: 6 G-buffers (frame n writes n % 6 and reads (n - 1) % 6 as "previous"), 6 pairs of reservoir
// buffers (frame n ping-pongs inside pair n % 6 and reads the final one of frame n - 1), 4 sets of work queues and 4 device
// parameter blocks (n % 4).  A(n) waits for back(n - 4), B(n) for A(n), C(n) for B(n), back(n) for C(n) and, by stream order,
// back(n - 1).  (Five G-buffers / pairs would do for one GPU.  The sixth covers several GPUs: a neighbour reads this context's
// frame n - 1 planes in place during ITS temporal pass of frame n, and this context's A(n + 5) — the first writer of those
// planes — cannot start before its back(n + 1), whose first phase waits for the flag that neighbour publishes after that pass.)
#define VRS_NG 6
#define VRS_NR 12
#define VRS_NQ 4
#define VRS_NFRONT 3
struct GraphEntry { cudaGraphExec_t exec = nullptr; uint32_t launches = 0; };
struct FrameIdx { int g = 0, gprev = 0, ra = 0, rb = 0, q = 0; };

struct vrs_ctx {
  vrs_config cfg;
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  uint32_t W = 0, H = 0;
  int band_y0 = 0, band_y1 = 0, store_y0 = 0, store_y1 = 0;
  size_t npix = 0;                       // stored pixels = (store_y1 - store_y0) * W

  float4* g_planes[VRS_NG][4] = {{nullptr}};  // worldPos, albedo, normal, matProps
  float4* r_planes[VRS_NR][2] = {{nullptr}};  // info, weight
  float4* accum = nullptr;
  uint32_t* trace = nullptr;
  uint64_t frame_no = 0;                 // frames begun so far: the frame being built is frame_no (Renderer::m_currentGBufferFrameIdx generalised)
  FrameIdx cur{};                        // buffer indices of the frame being built (per-pass API) / last built
  int last_g = 0;                        // G-buffer of the last completed frame
  int final_r = 0;                       // final reservoirs of the last completed frame (temporal input of the next)
  int src_r = 0;                         // most recently written reservoir buffer inside the frame
  int last_q = 0;                        // queue set of the last frame (vrs_get_counters)
  cudaStream_t fstream[VRS_NFRONT] = {nullptr};          // streams of the front stages A, B, C (frames in flight)
  cudaEvent_t ev_stage_done[VRS_NFRONT][VRS_NQ] = {{nullptr}}, ev_back_done[VRS_NQ] = {nullptr};
  bool back_recorded[VRS_NQ] = {false, false, false, false};
  int depth = 4;                         // stages that may overlap: 4 = A | B | C | back, 3 = A | B+C | back, 2 = A+B+C | back (VRS_PIPELINE)
  bool halo_pending = false;             // a halo push has been enqueued whose consumer-side wait has not been yet
  bool pipeline = true;                  // VRS_PIPELINE=0 switches the overlap of frames off
  bool replaying = false;                // a captured stage is being replayed: bodies only advance host-side state

  HostGrid host_grid;
  bool has_grid = false;
  GridDev grid{};
  std::vector<void*> grid_allocs;
  uint64_t grid_bytes = 0;
  float max_density = 0.f;

  LightsDev lights{};
  std::vector<vrs_alias_table_cell> alias_host;
  void* d_lights = nullptr; void* d_alias = nullptr;

  uchar4* display[2] = {nullptr, nullptr};   // device staging of the 8-bit display image (double-buffered)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t display_ready[2] = {nullptr, nullptr}, copy_done[2] = {nullptr, nullptr};
  uint32_t present_count = 0;
  FrameParams* d_params = nullptr;           // VRS_NQ device-resident copies (frame n reads copy n % VRS_NQ)
  FrameParams* h_params = nullptr;           // pinned ring feeding it
  cudaEvent_t param_ev[VRS_PARAM_SLOTS] = {nullptr};
  uint64_t param_serial = 0;
  std::map<uint64_t, GraphEntry> graphs;     // captured stages, keyed by buffer-rotation phase + structural flags (graph_key)
  bool capturing = false;
  Queues queues[VRS_NQ]{};
  int persistent_blocks = 148 * 12;

  cudaEvent_t ev[8] = {nullptr};
  vrs_timings timings{};
  bool timings_valid = false;
  bool pass_timing = false;                  // record the per-pass events inside each frame (vrs_set_pass_timing); off: frames overlap
  Comm* comm = nullptr;
  // peer-memory exchange (vrs_peer_connect): neighbours' planes opened through CUDA IPC
  struct Peer { bool present = false; float4* g[VRS_NG][4] = {{nullptr}}; float4* r[VRS_NR][2] = {{nullptr}}; unsigned* flags = nullptr; int store_y0 = 0, store_y1 = 0, band_y0 = 0, band_y1 = 0; std::vector<void*> opened; };
  Peer peer_up, peer_down;
  bool peer_mode = false;
  unsigned* xflags = nullptr;                // [0] flag written by the up neighbour, [1] by the down neighbour, [2] serial, [3] block counter, [4] wait timed out,
                                             // [5] temporal reprojections nobody could supply, [6] largest vertical reprojection distance (rows)
  cudaStream_t comm_stream = nullptr;        // halo exchanges run here so that they can overlap the next kernels
  cudaEvent_t ev_halo_src = nullptr, ev_halo_done = nullptr;
  bool history_valid = false;                // false: the previous frame's buffers do not belong to this scene / size (first frame, new lights, new grid, resize)
  KTimer kt{};                               // optional per-kernel event timing (vrs_set_kernel_timing): frames then launch eagerly
  bool kt_events = false;
  uint32_t comm_timeouts = 0;
};

static std::string g_create_error;

#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                 \
      return VRS_ERR_CUDA;                                                                           \
    }                                                                                                \
  } while (0)

static vrs_status fail(vrs_ctx* ctx, vrs_status s, const std::string& msg) {
  if (ctx) ctx->err = msg; else g_create_error = msg;
  return s;
}

extern "C" {

const char* vrs_last_error(const vrs_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

// Everything whose size depends on the image: per-pixel planes, work queues, display staging (vrs_create / vrs_resize).
static void free_frame_buffers(vrs_ctx* ctx) {
  for (int i = 0; i < VRS_NG; ++i) for (int p = 0; p < 4; ++p) { cudaFree(ctx->g_planes[i][p]); ctx->g_planes[i][p] = nullptr; }
  for (int i = 0; i < VRS_NR; ++i) for (int p = 0; p < 2; ++p) { cudaFree(ctx->r_planes[i][p]); ctx->r_planes[i][p] = nullptr; }
  cudaFree(ctx->accum); ctx->accum = nullptr; cudaFree(ctx->trace); ctx->trace = nullptr;
  for (int k = 0; k < VRS_NQ; ++k) {
    Queues& Q = ctx->queues[k];
    cudaFree(Q.counters); cudaFree(Q.cand); cudaFree(Q.cand_ray); cudaFree(Q.flag); cudaFree(Q.block_count); cudaFree(Q.hit_pix);
    cudaFree(Q.hit_seed); cudaFree(Q.hit_T); cudaFree(Q.shadow); cudaFree(Q.shadow_ray); cudaFree(Q.cover);
    memset(&Q, 0, sizeof(Q));
  }
  for (int i = 0; i < 2; ++i) { cudaFree(ctx->display[i]); ctx->display[i] = nullptr; }
}
static vrs_status alloc_frame_buffers(vrs_ctx* ctx) {
  const vrs_config* cfg = &ctx->cfg;
  ctx->W = cfg->width; ctx->H = cfg->height;
  ctx->band_y0 = (int)cfg->band_y0; ctx->band_y1 = cfg->band_y1 ? (int)cfg->band_y1 : (int)cfg->height;
  if (ctx->band_y1 > (int)ctx->H || ctx->band_y0 >= ctx->band_y1) return fail(ctx, VRS_ERR_INVALID, "bad band");
  ctx->store_y0 = ctx->band_y0 - (int)cfg->halo_rows; if (ctx->store_y0 < 0) ctx->store_y0 = 0;
  ctx->store_y1 = ctx->band_y1 + (int)cfg->halo_rows; if (ctx->store_y1 > (int)ctx->H) ctx->store_y1 = (int)ctx->H;
  ctx->npix = (size_t)(ctx->store_y1 - ctx->store_y0) * ctx->W;
  if (ctx->npix >= ((size_t)1 << 32)) return fail(ctx, VRS_ERR_UNSUPPORTED, "more than 2^32 stored pixels per context");
  auto alloc = [&](void** p, size_t bytes) {
    if (cudaMalloc(p, bytes) != cudaSuccess) { ctx->err = "cudaMalloc failed (" + std::to_string(bytes) + " bytes)"; return false; }
    return cudaMemset(*p, 0, bytes) == cudaSuccess;
  };
  // planes are exported through CUDA IPC (vrs_peer_export): whole multiples of 2 MB, so that each is an allocation block of its own
  const size_t plane_bytes = ((ctx->npix * 16 + ((size_t)2 << 20) - 1) >> 21) << 21;
  for (int i = 0; i < VRS_NG; ++i) for (int p = 0; p < 4; ++p) if (!alloc((void**)&ctx->g_planes[i][p], plane_bytes)) return VRS_ERR_CUDA;
  for (int i = 0; i < VRS_NR; ++i) for (int p = 0; p < 2; ++p) if (!alloc((void**)&ctx->r_planes[i][p], plane_bytes)) return VRS_ERR_CUDA;
  if (!alloc((void**)&ctx->accum, ctx->npix * 16)) return VRS_ERR_CUDA;
  if (cfg->enable_trace && !alloc((void**)&ctx->trace, ctx->npix * 16)) return VRS_ERR_CUDA;
  const size_t ncompact = (ctx->npix + 2047) / 2048 + 1;
  for (int k = 0; k < VRS_NQ; ++k) {
    Queues& Q = ctx->queues[k];
    if (!alloc((void**)&Q.counters, 64) || !alloc((void**)&Q.cand, ctx->npix * 4) || !alloc((void**)&Q.cand_ray, ctx->npix * 32) ||
        !alloc((void**)&Q.flag, ctx->npix + 16) || !alloc((void**)&Q.block_count, ncompact * 4) || !alloc((void**)&Q.hit_pix, ctx->npix * 4) ||
        !alloc((void**)&Q.hit_seed, ctx->npix * 4) || !alloc((void**)&Q.hit_T, ctx->npix * 4) || !alloc((void**)&Q.shadow, ctx->npix * 4) ||
        !alloc((void**)&Q.shadow_ray, ctx->npix * 32) || !alloc((void**)&Q.cover, ((size_t)(ctx->W + 7) / 8) * ((size_t)(ctx->H + 7) / 8) + 16))
      return VRS_ERR_CUDA;
  }
  for (int i = 0; i < 2; ++i) if (!alloc((void**)&ctx->display[i], ctx->npix * 4)) return VRS_ERR_CUDA;
  ctx->frame_no = 0; ctx->cur = FrameIdx(); ctx->last_g = ctx->final_r = ctx->src_r = ctx->last_q = 0;
  for (int i = 0; i < VRS_NQ; ++i) ctx->back_recorded[i] = false;
  ctx->halo_pending = false;
  ctx->present_count = 0;
  ctx->history_valid = false;
  return VRS_OK;
}

vrs_status vrs_create(const vrs_config* cfg, vrs_ctx** out) {
  if (!cfg || !out || cfg->width == 0 || cfg->height == 0) return fail(nullptr, VRS_ERR_INVALID, "vrs_create: bad config");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, VRS_ERR_NO_DEVICE, "vrs_create: no CUDA device (libvrs has no CPU fallback)");
  vrs_ctx* ctx = new vrs_ctx();
  ctx->cfg = *cfg;
  auto bail = [&](vrs_status s) { g_create_error = ctx->err; vrs_destroy(ctx); return s; };
  if (cfg->device >= 0) ctx->device = cfg->device; else if (cudaGetDevice(&ctx->device) != cudaSuccess) ctx->device = 0;
  if (cudaSetDevice(ctx->device) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return bail(VRS_ERR_CUDA); }
  if (cfg->spatial_iterations > VRS_MAX_SPATIAL_ITERATIONS) { ctx->err = "vrs_create: spatial_iterations > 4"; return bail(VRS_ERR_INVALID); }
  auto alloc = [&](void** p, size_t bytes) {
    if (cudaMalloc(p, bytes) != cudaSuccess) { ctx->err = "cudaMalloc failed (" + std::to_string(bytes) + " bytes)"; return false; }
    return cudaMemset(*p, 0, bytes) == cudaSuccess;
  };
  // The back half is the frame's critical chain (frame n + 1's temporal merge needs frame n's last spatial iteration, and on
  // several GPUs the neighbours wait for its pushes): its stream outranks the front stages', so that its blocks are placed first
  // whenever SM resources free up (VRS_NO_PRIORITY=1: all streams equal, for measurements).
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (getenv("VRS_NO_PRIORITY")) prio_hi = prio_lo;
  if (cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
      cudaStreamCreateWithPriority(&ctx->fstream[0], cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
      cudaStreamCreateWithPriority(&ctx->fstream[1], cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
      cudaStreamCreateWithPriority(&ctx->fstream[2], cudaStreamNonBlocking, prio_lo) != cudaSuccess) { ctx->err = "stream create failed"; return bail(VRS_ERR_CUDA); }
  for (int i = 0; i < VRS_NQ; ++i)
    if (cudaEventCreateWithFlags(&ctx->ev_stage_done[0][i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_stage_done[1][i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_stage_done[2][i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_back_done[i], cudaEventDisableTiming) != cudaSuccess) { ctx->err = "event create failed"; return bail(VRS_ERR_CUDA); }
  ctx->pipeline = !(getenv("VRS_PIPELINE") && getenv("VRS_PIPELINE")[0] == '0');
  ctx->depth = getenv("VRS_PIPELINE") && getenv("VRS_PIPELINE")[0] >= '2' && getenv("VRS_PIPELINE")[0] <= '4' ? getenv("VRS_PIPELINE")[0] - '0' : 4;
  if (vrs_status s = alloc_frame_buffers(ctx)) return bail(s);
  {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess) ctx->persistent_blocks = prop.multiProcessorCount * (getenv("VRS_BLOCKS_PER_SM") ? atoi(getenv("VRS_BLOCKS_PER_SM")) : 12);
  }
  for (int i = 0; i < 8; ++i) if (cudaEventCreate(&ctx->ev[i]) != cudaSuccess) { ctx->err = "event create failed"; return bail(VRS_ERR_CUDA); }
  // The exchange flags are exported through CUDA IPC: a handle names a whole allocation block, and small cudaMalloc requests are
  // carved out of shared 2 MB blocks — an own 2 MB allocation makes the opened pointer land on the flags, not on a block base.
  if (!alloc((void**)&ctx->xflags, (size_t)2 << 20)) return bail(VRS_ERR_CUDA);
  if (!alloc((void**)&ctx->d_params, sizeof(FrameParams) * VRS_NQ) ||
      cudaHostAlloc((void**)&ctx->h_params, sizeof(FrameParams) * VRS_PARAM_SLOTS, cudaHostAllocDefault) != cudaSuccess) { ctx->err = "param alloc failed"; return bail(VRS_ERR_CUDA); }
  for (int i = 0; i < VRS_PARAM_SLOTS; ++i)
    if (cudaEventCreateWithFlags(&ctx->param_ev[i], cudaEventDisableTiming) != cudaSuccess) { ctx->err = "event create failed"; return bail(VRS_ERR_CUDA); }
  // (copy_stream and comm_stream are created on first use: a context that never presents / never uses NCCL holds one stream,
  // which keeps several band contexts of one process on distinct hardware queues)
  if (cudaEventCreateWithFlags(&ctx->ev_halo_src, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_halo_done, cudaEventDisableTiming) != cudaSuccess) { ctx->err = "stream create failed"; return bail(VRS_ERR_CUDA); }
  for (int i = 0; i < 2; ++i)
    if (cudaEventCreateWithFlags(&ctx->display_ready[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->copy_done[i], cudaEventDisableTiming) != cudaSuccess) { ctx->err = "event create failed"; return bail(VRS_ERR_CUDA); }
  cudaDeviceSynchronize();
  *out = ctx;
  return VRS_OK;
}

static void sync_front_streams(vrs_ctx* ctx) { for (int i = 0; i < VRS_NFRONT; ++i) if (ctx->fstream[i]) cudaStreamSynchronize(ctx->fstream[i]); }
static void invalidate_graphs(vrs_ctx* ctx) {
  sync_front_streams(ctx);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  ctx->graphs.clear();
}

static void free_grid(vrs_ctx* ctx) {
  for (void* p : ctx->grid_allocs) cudaFree(p);
  ctx->grid_allocs.clear(); ctx->has_grid = false; ctx->grid_bytes = 0;
}

void vrs_destroy(vrs_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  sync_front_streams(ctx);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->comm) comm_destroy(ctx->comm);
  for (int i = 0; i < VRS_NQ; ++i) {
    for (int k = 0; k < VRS_NFRONT; ++k) if (ctx->ev_stage_done[k][i]) cudaEventDestroy(ctx->ev_stage_done[k][i]);
    if (ctx->ev_back_done[i]) cudaEventDestroy(ctx->ev_back_done[i]);
  }
  for (vrs_ctx::Peer* p : {&ctx->peer_up, &ctx->peer_down}) for (void* q : p->opened) cudaIpcCloseMemHandle(q);
  cudaFree(ctx->xflags);
  free_grid(ctx);
  free_frame_buffers(ctx);
  cudaFree(ctx->d_lights); cudaFree(ctx->d_alias);
  if (ctx->kt_events) for (int i = 0; i <= KTimer::MAX; ++i) cudaEventDestroy(ctx->kt.ev[i]);
  for (int i = 0; i < 8; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (auto& kv : ctx->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  for (int i = 0; i < VRS_PARAM_SLOTS; ++i) if (ctx->param_ev[i]) cudaEventDestroy(ctx->param_ev[i]);
  cudaFree(ctx->d_params); if (ctx->h_params) cudaFreeHost(ctx->h_params);
  for (int i = 0; i < 2; ++i) {
    if (ctx->display_ready[i]) cudaEventDestroy(ctx->display_ready[i]);
    if (ctx->copy_done[i]) cudaEventDestroy(ctx->copy_done[i]);
  }
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  if (ctx->comm_stream) { cudaStreamSynchronize(ctx->comm_stream); cudaStreamDestroy(ctx->comm_stream); }
  if (ctx->ev_halo_src) cudaEventDestroy(ctx->ev_halo_src);
  if (ctx->ev_halo_done) cudaEventDestroy(ctx->ev_halo_done);
  for (int i = 0; i < VRS_NFRONT; ++i) if (ctx->fstream[i]) cudaStreamDestroy(ctx->fstream[i]);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

// ------------------------------------------------------------------------------------------ grid staging
static vrs_status upload_grid(vrs_ctx* ctx) {
  cudaSetDevice(ctx->device);
  invalidate_graphs(ctx);   // frames in flight read the old tables, and captured stages point at them: drain, then drop
  free_grid(ctx);
  const HostGrid& h = ctx->host_grid;
  if (h.nleaf() == 0 && h.active_voxels == 0) return fail(ctx, VRS_ERR_FORMAT, "grid has no active voxels");
  GridDev& G = ctx->grid;
  memset(&G, 0, sizeof(G));
  for (int a = 0; a < 3; ++a) {
    G.vmin[a] = (h.bbox_min[a] >> 3) << 3;
    int vmax = ((h.bbox_max[a] >> 3) + 1) << 3;
    G.vdim[a] = vmax - G.vmin[a];
    G.cdim[a] = G.vdim[a] / 8;
  }
  G.bg_density = h.density_from_raw(h.background);
  // world = world_scale * (voxel_size * ijk + translation) + world_translate   (Renderer.cpp:1420-1435)
  G.A = (float)((double)ctx->cfg.world_scale * h.voxel_size);
  G.invA = 1.0f / G.A;
  for (int a = 0; a < 3; ++a) G.B[a] = (float)((double)ctx->cfg.world_scale * h.translation[a] + (double)ctx->cfg.world_translate[a]);
  G.density_scale = ctx->cfg.density_scale;
  G.roughness = ctx->cfg.roughness; G.metallic = ctx->cfg.metallic;
  // densities: atlas + tile table + per-brick majorants
  std::vector<float> atlas(h.leaf_value.size()), leaf_max(h.nleaf()), tile_density(h.tile_value.size());
  float gmaxd = 0.f;
  for (size_t l = 0; l < h.nleaf(); ++l) {
    float m = 0.f;
    for (int o = 0; o < 512; ++o) { float d = h.density_from_raw(h.leaf_value[l * 512 + o]); atlas[l * 512 + o] = d; if (d > m) m = d; }
    leaf_max[l] = m; if (m > gmaxd) gmaxd = m;
  }
  for (size_t t = 0; t < h.tile_value.size(); ++t) { tile_density[t] = h.density_from_raw(h.tile_value[t]); if (tile_density[t] > gmaxd) gmaxd = tile_density[t]; }
  ctx->max_density = gmaxd;
  // dense directory over the window's cells (the top tree levels flattened once more for the raymarch)
  const size_t ncell = (size_t)G.cdim[0] * G.cdim[1] * G.cdim[2];
  if (ncell > ((size_t)1 << 30)) { free_grid(ctx); return fail(ctx, VRS_ERR_UNSUPPORTED, "grid window exceeds 2^30 cells"); }
  std::vector<float> dir(2 * ncell);      // {majorant, leaf bits} per cell (GridDev::dir)
  for (int cz = 0; cz < G.cdim[2]; ++cz)
    for (int cy = 0; cy < G.cdim[1]; ++cy)
      for (int cx = 0; cx < G.cdim[0]; ++cx) {
        const int32_t x = G.vmin[0] + cx * 8, y = G.vmin[1] + cy * 8, z = G.vmin[2] + cz * 8;
        int32_t c = ~0;
        const int32_t kx = x & ~4095, ky = y & ~4095, kz = z & ~4095;
        for (size_t r = 0; r < h.root.size() / 4; ++r)
          if (h.root[4 * r] == kx && h.root[4 * r + 1] == ky && h.root[4 * r + 2] == kz) { c = h.root[4 * r + 3]; break; }
        if (c >= 0) {
          c = h.i5[(size_t)c * 32768 + ((((x & 4095) >> 7) << 10) | (((y & 4095) >> 7) << 5) | ((z & 4095) >> 7))];
          if (c >= 0) c = h.i4[(size_t)c * 4096 + ((((x & 127) >> 3) << 8) | (((y & 127) >> 3) << 4) | ((z & 127) >> 3))];
        }
        const size_t ci = ((size_t)cz * G.cdim[1] + cy) * G.cdim[0] + cx;
        dir[2 * ci] = c < 0 ? tile_density[~c] : leaf_max[c];
        memcpy(&dir[2 * ci + 1], &c, 4);
      }
  // One slab for every grid table, so that a single L2 access-policy window can pin the whole grid: the per-pixel
  // buffers stream through L2 every frame (hundreds of MB) and would otherwise evict the few MB every ray keeps re-reading.
  struct Part { const void* src; size_t bytes; const void** dst; };
  G.nroot = (int)(h.root.size() / 4);
  Part parts[] = {{h.root.data(), h.root.size() * 4, (const void**)&G.root}, {h.i5.data(), h.i5.size() * 4, (const void**)&G.i5},
                  {h.i4.data(), h.i4.size() * 4, (const void**)&G.i4}, {tile_density.data(), tile_density.size() * 4, (const void**)&G.tile_density},
                  {leaf_max.data(), leaf_max.size() * 4, (const void**)&G.leaf_max}, {dir.data(), dir.size() * 4, (const void**)&G.dir}, {atlas.data(), atlas.size() * 4, (const void**)&G.atlas}};
  size_t total = 0;
  for (const Part& p : parts) total += (p.bytes + 255) & ~(size_t)255;
  char* slab = nullptr;
  if (cudaMalloc((void**)&slab, total ? total : 256) != cudaSuccess) { free_grid(ctx); return fail(ctx, VRS_ERR_CUDA, "grid staging failed (cudaMalloc)"); }
  ctx->grid_allocs.push_back(slab); ctx->grid_bytes = total;
  size_t off = 0;
  for (const Part& p : parts) {
    if (p.bytes && cudaMemcpy(slab + off, p.src, p.bytes, cudaMemcpyHostToDevice) != cudaSuccess) { free_grid(ctx); return fail(ctx, VRS_ERR_CUDA, "grid staging failed (copy)"); }
    *p.dst = slab + off;
    off += (p.bytes + 255) & ~(size_t)255;
  }
  // persisting-L2 window over the slab (hot part first: directory + atlas sit at the end, tables are tiny except i5)
  {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0 && !getenv("VRS_NO_L2_WINDOW")) {
      size_t want = total < (size_t)prop.persistingL2CacheMaxSize ? total : (size_t)prop.persistingL2CacheMaxSize;
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
      cudaStreamAttrValue attr; memset(&attr, 0, sizeof(attr));
      size_t win = total < (size_t)prop.accessPolicyMaxWindowSize ? total : (size_t)prop.accessPolicyMaxWindowSize;
      attr.accessPolicyWindow.base_ptr = slab;
      attr.accessPolicyWindow.num_bytes = win;
      attr.accessPolicyWindow.hitRatio = win <= want ? 1.0f : (float)want / (float)win;
      attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
      for (int i = 0; i < VRS_NFRONT; ++i) cudaStreamSetAttribute(ctx->fstream[i], cudaStreamAttributeAccessPolicyWindow, &attr);
      cudaGetLastError();       // best effort: the window is an optimisation, never an error
    }
  }
  ctx->has_grid = true;
  ctx->history_valid = false;   // the previous frame's G-buffer / reservoirs describe another volume
  return VRS_OK;
}

// No exception crosses the C ABI: a hostile file can still make a reader ask for an absurd allocation.
#define VRS_GUARD(ctx_, body)                                                                                     \
  try { body }                                                                                                    \
  catch (const std::bad_alloc&) { return fail(ctx_, VRS_ERR_FORMAT, "file asks for more memory than can be allocated (corrupt?)"); } \
  catch (const std::exception& e) { return fail(ctx_, VRS_ERR_FORMAT, std::string("reader failed: ") + e.what()); }

vrs_status vrs_load_vdb(vrs_ctx* ctx, const char* path, const char* grid_name) {
  if (!ctx || !path) return VRS_ERR_INVALID;
  std::string p(path), err;
  if (p.size() > 5 && p.compare(p.size() - 5, 5, ".vrsg") == 0) return vrs_load_vrsg(ctx, path);
  VRS_GUARD(ctx,
    if (!read_vdb(p, grid_name, ctx->host_grid, err)) {
      bool io = err.rfind("cannot open", 0) == 0 || err.rfind("short read", 0) == 0;
      return fail(ctx, io ? VRS_ERR_IO : VRS_ERR_FORMAT, err);
    }
    return upload_grid(ctx);)
}
vrs_status vrs_load_vrsg(vrs_ctx* ctx, const char* path) {
  if (!ctx || !path) return VRS_ERR_INVALID;
  std::string err;
  VRS_GUARD(ctx,
    if (!read_vrsg(path, ctx->host_grid, err)) return fail(ctx, err.rfind("cannot open", 0) == 0 ? VRS_ERR_IO : VRS_ERR_FORMAT, err);
    return upload_grid(ctx);)
}
vrs_status vrs_convert_vdb(const char* vdb_path, const char* grid_name, const char* vrsg_path) {
  if (!vdb_path || !vrsg_path) return VRS_ERR_INVALID;
  HostGrid g; std::string err;
  const std::string in(vdb_path);
  const bool is_vrsg = in.size() > 5 && in.compare(in.size() - 5, 5, ".vrsg") == 0;      // a snapshot is validated and re-written
  VRS_GUARD(nullptr,
    if (!(is_vrsg ? read_vrsg(in, g, err) : read_vdb(in, grid_name, g, err))) return fail(nullptr, err.rfind("cannot open", 0) == 0 ? VRS_ERR_IO : VRS_ERR_FORMAT, err);
    if (!write_vrsg(vrsg_path, g, err)) return fail(nullptr, VRS_ERR_IO, err);
    return VRS_OK;)
}
vrs_status vrs_make_procedural_grid(vrs_ctx* ctx, int kind, uint32_t resolution) {
  if (!ctx) return VRS_ERR_INVALID;
  std::string err;
  if (!make_procedural(kind, resolution, ctx->host_grid, err)) return fail(ctx, VRS_ERR_INVALID, err);
  return upload_grid(ctx);
}

vrs_status vrs_write_procedural_vrsg(int kind, uint32_t resolution, const char* vrsg_path) {
  if (!vrsg_path) return VRS_ERR_INVALID;
  HostGrid g; std::string err;
  if (!make_procedural(kind, resolution, g, err)) return fail(nullptr, VRS_ERR_INVALID, err);
  if (!write_vrsg(vrsg_path, g, err)) return fail(nullptr, VRS_ERR_IO, err);
  return VRS_OK;
}

vrs_status vrs_get_grid_info(const vrs_ctx* ctx, vrs_grid_info* o) {
  if (!ctx || !o || !ctx->has_grid) return VRS_ERR_INVALID;
  const HostGrid& h = ctx->host_grid;
  memset(o, 0, sizeof(*o));
  for (int a = 0; a < 3; ++a) { o->bbox_min[a] = h.bbox_min[a]; o->bbox_max[a] = h.bbox_max[a]; o->translation[a] = h.translation[a]; }
  o->active_voxels = h.active_voxels; o->root_children = h.root_children; o->internal5 = (uint32_t)h.n5(); o->internal4 = (uint32_t)h.n4();
  o->leaves = (uint32_t)h.nleaf(); o->tiles = (uint32_t)h.tile_value.size(); o->voxel_size = h.voxel_size; o->background = h.background;
  o->is_level_set = h.level_set ? 1 : 0; o->max_density = ctx->max_density; o->device_bytes = ctx->grid_bytes;
  const GridDev& G = ctx->grid;
  for (int a = 0; a < 3; ++a) {
    o->world_bbox_min[a] = G.A * ((float)G.vmin[a] - 0.5f) + G.B[a];
    o->world_bbox_max[a] = G.A * ((float)(G.vmin[a] + G.vdim[a]) - 0.5f) + G.B[a];
  }
  return VRS_OK;
}
vrs_status vrs_grid_get_value(const vrs_ctx* ctx, int32_t i, int32_t j, int32_t k, float* value, int32_t* active) {
  if (!ctx || !ctx->has_grid || !value) return VRS_ERR_INVALID;
  bool a = false;
  *value = ctx->host_grid.get_value(i, j, k, &a);
  if (active) *active = a ? 1 : 0;
  return VRS_OK;
}
vrs_status vrs_grid_sample_device(vrs_ctx* ctx, const int32_t* ijk, uint32_t n, float* out) {
  if (!ctx || !ctx->has_grid || !ijk || !out) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (n == 0) return VRS_OK;
  int* d_ijk = nullptr; float* d_out = nullptr;
  cudaError_t e = cudaMalloc(&d_ijk, (size_t)n * 12);
  if (e == cudaSuccess) e = cudaMalloc(&d_out, (size_t)n * 4);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_ijk, ijk, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) { launch_sample_density(ctx->stream, ctx->grid, d_ijk, n, d_out); e = cudaGetLastError(); }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_ijk); cudaFree(d_out);
  if (e != cudaSuccess) { ctx->err = std::string("vrs_grid_sample_device: ") + cudaGetErrorString(e); return VRS_ERR_CUDA; }
  return VRS_OK;
}

// ------------------------------------------------------------------------------------------ lights
vrs_status vrs_set_lights(vrs_ctx* ctx, const vrs_point_light* lights, uint32_t n) {
  if (!ctx || !lights || n == 0) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  std::vector<float> pdf(n);
  for (uint32_t i = 0; i < n; ++i) pdf[i] = lights[i].emission_luminance[3];      // Renderer.cpp:1653-1657
  ctx->alias_host.resize(n);
  vrs_create_alias_table(pdf.data(), n, ctx->alias_host.data());
  invalidate_graphs(ctx);   // captured frames point at the old light buffers
  cudaFree(ctx->d_lights); cudaFree(ctx->d_alias); ctx->d_lights = ctx->d_alias = nullptr;
  CK(cudaMalloc(&ctx->d_lights, (size_t)n * sizeof(vrs_point_light)));
  CK(cudaMalloc(&ctx->d_alias, (size_t)n * sizeof(vrs_alias_table_cell)));
  CK(cudaMemcpy(ctx->d_lights, lights, (size_t)n * sizeof(vrs_point_light), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->d_alias, ctx->alias_host.data(), (size_t)n * sizeof(vrs_alias_table_cell), cudaMemcpyHostToDevice));
  ctx->lights.lights = (const float4*)ctx->d_lights; ctx->lights.alias = (const float4*)ctx->d_alias;
  ctx->lights.nlights = (int)n; ctx->lights.ntable = (int)n;
  ctx->history_valid = false;   // stored reservoirs hold indices into the old light table: no temporal merge on the next frame
  return VRS_OK;
}
vrs_status vrs_collect_emissive_lights(const vrs_ctx* ctx, float threshold, uint32_t max_lights, vrs_point_light* out, uint32_t* count) {
  if (!ctx || !ctx->has_grid || !out || !count) return VRS_ERR_INVALID;
  const HostGrid& h = ctx->host_grid;
  const GridDev& G = ctx->grid;
  uint32_t n = 0;
  for (size_t l = 0; l < h.nleaf() && n < max_lights; ++l)
    for (int o = 0; o < 512 && n < max_lights; ++o) {
      if (!((h.leaf_mask[l * 8 + (o >> 6)] >> (o & 63)) & 1)) continue;
      if (!(h.leaf_value[l * 512 + o] > threshold)) continue;
      const int ijk[3] = {h.leaf_origin[3 * l] + (o >> 6), h.leaf_origin[3 * l + 1] + ((o >> 3) & 7), h.leaf_origin[3 * l + 2] + (o & 7)};
      vrs_point_light& p = out[n++];
      for (int a = 0; a < 3; ++a) p.pos[a] = G.A * (float)ijk[a] + G.B[a];            // sphere centre, Renderer.cpp:1427-1431
      p.pos[3] = 1.0f;
      p.emission_luminance[0] = 0.6f; p.emission_luminance[1] = 0.2f; p.emission_luminance[2] = 0.1f;   // :1627
      p.emission_luminance[3] = 0.2126f * 0.6f + 0.7152f * 0.2f + 0.0722f * 0.1f;                        // shader::luminance
    }
  *count = n;
  return VRS_OK;
}
vrs_status vrs_vdb_emissive_lights(const char* vdb_path, const char* grid_name, const vrs_config* cfg, uint32_t max_lights,
                                   vrs_point_light* out, uint32_t* count) {
  if (!vdb_path || !cfg || !out || !count) return VRS_ERR_INVALID;
  HostGrid h; std::string err;
  VRS_GUARD(nullptr,
    if (!read_vdb(vdb_path, grid_name, h, err)) return fail(nullptr, err.rfind("cannot open", 0) == 0 ? VRS_ERR_IO : VRS_ERR_FORMAT, err);)
  const float A = (float)((double)cfg->world_scale * h.voxel_size);
  float B[3];
  for (int a = 0; a < 3; ++a) B[a] = (float)((double)cfg->world_scale * h.translation[a] + (double)cfg->world_translate[a]);
  uint32_t n = 0;
  for (size_t l = 0; l < h.nleaf() && n < max_lights; ++l)
    for (int o = 0; o < 512 && n < max_lights; ++o) {
      if (!((h.leaf_mask[l * 8 + (o >> 6)] >> (o & 63)) & 1)) continue;
      const float temp = (float)((double)logf(h.leaf_value[l * 512 + o]) + 273.15);            // vdb.cpp:811
      if (!(temp > 275.0f)) continue;                                                           // Renderer.cpp:1623
      const int ijk[3] = {h.leaf_origin[3 * l] + (o >> 6), h.leaf_origin[3 * l + 1] + ((o >> 3) & 7), h.leaf_origin[3 * l + 2] + (o & 7)};
      vrs_point_light& p = out[n++];
      for (int a = 0; a < 3; ++a) p.pos[a] = A * (float)ijk[a] + B[a];
      p.pos[3] = 1.0f;
      p.emission_luminance[0] = 0.6f; p.emission_luminance[1] = 0.2f; p.emission_luminance[2] = 0.1f;   // Renderer.cpp:1627
      p.emission_luminance[3] = 0.2126f * 0.6f + 0.7152f * 0.2f + 0.0722f * 0.1f;
    }
  *count = n;
  return VRS_OK;
}
vrs_status vrs_set_triangle_lights(vrs_ctx* ctx, const vrs_triangle_light*, uint32_t) {
  return fail(ctx, VRS_ERR_UNSUPPORTED, "triangle lights belong to the mesh path (SURVEY.md §8f rank 4), not the volume hot path");
}
vrs_status vrs_get_alias_table(const vrs_ctx* ctx, vrs_alias_table_cell* out, uint32_t n) {
  if (!ctx || !out || n != ctx->alias_host.size()) return VRS_ERR_INVALID;
  memcpy(out, ctx->alias_host.data(), n * sizeof(vrs_alias_table_cell));
  return VRS_OK;
}

// ------------------------------------------------------------------------------------------ per-frame
static Planes planes_of(vrs_ctx* ctx, int i) { Planes p; p.worldPos = ctx->g_planes[i][0]; p.albedo = ctx->g_planes[i][1]; p.normal = ctx->g_planes[i][2]; p.mat = ctx->g_planes[i][3]; return p; }
static ResPlanes res_of(vrs_ctx* ctx, int i) { ResPlanes r; r.info = ctx->r_planes[i][0]; r.weight = ctx->r_planes[i][1]; return r; }

// 4x4 inverse in double precision (Gauss-Jordan with partial pivoting, column-major in and out); false = singular
static bool invert4d(const double* m, double* out) {
  double a[4][8];
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { a[r][c] = m[c * 4 + r]; a[r][4 + c] = r == c ? 1.0 : 0.0; }
  for (int i = 0; i < 4; ++i) {
    int piv = i;
    for (int r = i + 1; r < 4; ++r) if (std::fabs(a[r][i]) > std::fabs(a[piv][i])) piv = r;
    if (!(std::fabs(a[piv][i]) > 1e-300)) return false;
    if (piv != i) for (int c = 0; c < 8; ++c) std::swap(a[i][c], a[piv][c]);
    const double d = 1.0 / a[i][i];
    for (int c = 0; c < 8; ++c) a[i][c] *= d;
    for (int r = 0; r < 4; ++r) if (r != i) { const double f = a[r][i]; if (f != 0.0) for (int c = 0; c < 8; ++c) a[r][c] -= f * a[i][c]; }
  }
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out[c * 4 + r] = a[r][4 + c];
  return true;
}

static vrs_status make_params(vrs_ctx* ctx, const vrs_global_uniforms* gu, const vrs_restir_uniforms* ru, const vrs_push_constant_restir* pc,
                              uint32_t clock, FrameParams& F) {
  if (!ctx->has_grid) return fail(ctx, VRS_ERR_INVALID, "no grid loaded");
  if (!ctx->d_lights) return fail(ctx, VRS_ERR_INVALID, "no lights set");
  if (ru->screenSize[0] != ctx->W || ru->screenSize[1] != ctx->H) return fail(ctx, VRS_ERR_INVALID, "RestirUniforms.screenSize != context size");
  if (ru->aliasTableCount != ctx->lights.ntable || ru->pointLightCount != ctx->lights.nlights)
    return fail(ctx, VRS_ERR_INVALID, "RestirUniforms light counts != uploaded lights");
  if (ru->triangleLightCount > 1) return fail(ctx, VRS_ERR_UNSUPPORTED, "triangle lights are outside the volume hot path");
  memset(&F, 0, sizeof(F));
  if (gu) {
    memcpy(F.viewInverse, gu->viewInverse, 64); memcpy(F.projInverse, gu->projInverse, 64);
    // world -> clip of the primary rays, from the very matrices the rays are built with (coverage culling, k_cover)
    double vi[16], pi[16], v[16], p[16];
    for (int i = 0; i < 16; ++i) { vi[i] = gu->viewInverse[i]; pi[i] = gu->projInverse[i]; }
    if (invert4d(vi, v) && invert4d(pi, p)) {
      for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) {          // column-major: (P * V)[r][c] = sum_k P[r][k] V[k][c]
        double acc = 0.0;
        for (int k = 0; k < 4; ++k) acc += p[k * 4 + r] * v[c * 4 + k];
        F.cullVP[c * 4 + r] = (float)acc;
      }
      F.cull = 1;
      for (int i = 0; i < 16; ++i) if (!std::isfinite(F.cullVP[i])) F.cull = 0;
    }
  }
  memcpy(F.prevVP, ru->prevFrameProjectionViewMatrix, 64);
  for (int a = 0; a < 3; ++a) F.camPos[a] = ru->currCamPos[a];
  F.W = ctx->W; F.H = ctx->H; F.M = ru->initialLightSampleCount; F.temporalMult = ru->temporalSampleCountMultiplier;
  F.spatialNeighbors = ru->spatialNeighbors; F.spatialRadius = ru->spatialRadius; F.fireflyClamp = ru->fireflyClampThreshold;
  F.flags = ru->flags; F.clock = clock;
  F.roughness = ctx->grid.roughness; F.metallic = ctx->grid.metallic;
  if (!ctx->history_valid) F.flags &= ~VRS_RESTIR_TEMPORAL_REUSE_FLAG;   // first frame / new lights / new grid / resize: nothing valid to merge
  if (pc) { F.clear[0] = pc->clearColorRed; F.clear[1] = pc->clearColorGreen; F.clear[2] = pc->clearColorBlue; F.frame = pc->frame; F.initialize = pc->initialize; }
  return VRS_OK;
}

// Per-frame values travel through device-resident FrameParams blocks (fed from a ring of pinned host slots), so the
// kernels' launch arguments never change and the frame's stages can be replayed as CUDA graphs.  Frame n uses block n % VRS_NQ.
static vrs_status upload_params(vrs_ctx* ctx, const FrameParams& F, int q, cudaStream_t st) {
  const int slot = (int)(ctx->param_serial % VRS_PARAM_SLOTS);
  if (ctx->param_serial >= VRS_PARAM_SLOTS) CK(cudaEventSynchronize(ctx->param_ev[slot]));   // slot free again (normally long since)
  ctx->h_params[slot] = F;
  CK(cudaMemcpyAsync(ctx->d_params + q, &ctx->h_params[slot], sizeof(FrameParams), cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(ctx->param_ev[slot], st));
  ctx->param_serial++;
  return VRS_OK;
}

static FrameIdx frame_idx(uint64_t n) {
  FrameIdx f;
  f.g = (int)(n % VRS_NG); f.gprev = (int)((n + VRS_NG - 1) % VRS_NG);
  f.ra = 2 * (int)(n % (VRS_NR / 2)); f.rb = f.ra + 1;
  f.q = (int)(n % VRS_NQ);
  return f;
}

// ---- halo exchange, producer and consumer side.  halo_push sends rows of the given planes to both neighbours and leaves a
// pending mark; halo_wait (placed by the next phase, right before the first kernel that reads halo rows) joins it.
// `max_rows` limits the rows sent per side: spatial reuse needs ceil(spatialRadius) rows, the temporal reprojection every
// halo row the neighbour stores.
static vrs_status halo_push(vrs_ctx* ctx, cudaStream_t st, bool gbuf, int g_index, int r_index, int max_rows) {
  if (ctx->peer_mode) {
    // (g_index always names the frame's G-buffer: its worldPos.w is the hit test.  gbuf = false and r_index < 0: no payload, the kernel only publishes the serial — "my planes of this frame are final")
    // one kernel stores the boundary rows into both neighbours' halo rows (NVLink P2P / same-device stores) and publishes the serial
    HaloPush H; memset(&H, 0, sizeof(H));
    const vrs_ctx::Peer& U = ctx->peer_up; const vrs_ctx::Peer& D = ctx->peer_down;
    auto add = [&](float4* mine, float4* up, float4* down) { H.src[H.nplanes] = mine; H.up_dst[H.nplanes] = up; H.down_dst[H.nplanes] = down; H.nplanes++; };
    // worldPos (hit marker) whole, albedo + normal + the reservoir pair only for hit pixels; matProps is a per-grid constant
    // nobody reads from the plane (ginfo_from_planes)
    H.wp_src = ctx->g_planes[g_index][0]; H.wp_up = U.present ? U.g[g_index][0] : nullptr; H.wp_down = D.present ? D.g[g_index][0] : nullptr;
    H.copy_wp = gbuf ? 1 : 0;
    if (gbuf) for (int p = 1; p < 3; ++p) add(ctx->g_planes[g_index][p], U.present ? U.g[g_index][p] : nullptr, D.present ? D.g[g_index][p] : nullptr);
    if (r_index >= 0) for (int p = 0; p < 2; ++p) add(ctx->r_planes[r_index][p], U.present ? U.r[r_index][p] : nullptr, D.present ? D.r[r_index][p] : nullptr);
    const bool payload = gbuf || r_index >= 0;
    const int band_h = ctx->band_y1 - ctx->band_y0;
    if (ctx->peer_up.present) {          // my first rows -> the rows just below the up neighbour's band
      int rows = ctx->peer_up.store_y1 - ctx->peer_up.band_y1; if (rows > band_h) rows = band_h;
      if (rows > max_rows) rows = max_rows;
      if (!payload) rows = 0;
      H.up_count = (size_t)rows * ctx->W; H.up_src_off = (size_t)(ctx->band_y0 - ctx->store_y0) * ctx->W;
      H.up_dst_off = (size_t)(ctx->band_y0 - ctx->peer_up.store_y0) * ctx->W; H.up_flag = ctx->peer_up.flags + 1;     // "written by the down neighbour"
    }
    if (ctx->peer_down.present) {        // my last rows -> the rows just above the down neighbour's band
      int rows = ctx->peer_down.band_y0 - ctx->peer_down.store_y0; if (rows > band_h) rows = band_h;
      if (rows > max_rows) rows = max_rows;
      if (!payload) rows = 0;
      H.down_count = (size_t)rows * ctx->W; H.down_src_off = (size_t)(ctx->band_y1 - rows - ctx->store_y0) * ctx->W;
      H.down_dst_off = (size_t)(ctx->band_y1 - rows - ctx->peer_down.store_y0) * ctx->W; H.down_flag = ctx->peer_down.flags + 0;   // "written by the up neighbour"
    }
    H.serial = ctx->xflags + 2; H.block_counter = ctx->xflags + 3;
    // one pixel per thread, a few pixels per thread for large pushes; a flag-only push is one block
    size_t blocks = (H.up_count + H.down_count + 1023) / 1024;
    const size_t max_blocks = (size_t)(ctx->persistent_blocks / 3);
    if (blocks < 1) blocks = 1;
    if (blocks > max_blocks) blocks = max_blocks;
    launch_halo_push(st, H, (int)blocks, &ctx->kt);
    CK(cudaGetLastError());
    ctx->timings.launches += 1;
    ctx->halo_pending = true;
    return VRS_OK;
  }
  if (!ctx->comm) return VRS_OK;
  std::vector<float4*> planes;
  if (gbuf) for (int p = 0; p < 4; ++p) planes.push_back(ctx->g_planes[g_index][p]);
  if (r_index >= 0) for (int p = 0; p < 2; ++p) planes.push_back(ctx->r_planes[r_index][p]);
  std::string err;
  CK(cudaEventRecord(ctx->ev_halo_src, st));
  CK(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_halo_src, 0));
  if (!comm_exchange_halo(ctx->comm, ctx->comm_stream, planes, ctx->W, ctx->band_y0, ctx->band_y1, ctx->store_y0, ctx->store_y1, (int)ctx->H, max_rows, err))
    return fail(ctx, VRS_ERR_COMM, err);
  CK(cudaEventRecord(ctx->ev_halo_done, ctx->comm_stream));
  ctx->halo_pending = true;
  return VRS_OK;
}
static vrs_status halo_wait(vrs_ctx* ctx, cudaStream_t st) {
  if (!ctx->halo_pending) return VRS_OK;
  ctx->halo_pending = false;
  if (ctx->peer_mode) {
    launch_halo_wait(st, ctx->xflags + 2, ctx->peer_up.present ? ctx->xflags + 0 : nullptr, ctx->peer_down.present ? ctx->xflags + 1 : nullptr, ctx->xflags + 4, &ctx->kt);
    CK(cudaGetLastError());
    ctx->timings.launches += 1;
  } else if (ctx->comm) {
    CK(cudaStreamWaitEvent(st, ctx->ev_halo_done, 0));
  }
  return VRS_OK;
}

// timing events must become event-record nodes when the frame is being captured into a graph
static cudaError_t mark(vrs_ctx* ctx, int i, cudaStream_t st) {
  if (!ctx->pass_timing) return cudaSuccess;
  return ctx->capturing ? cudaEventRecordWithFlags(ctx->ev[i], st, cudaEventRecordExternal) : cudaEventRecord(ctx->ev[i], st);
}
static bool culling_on(vrs_ctx* ctx, const FrameParams& F) {
  return F.cull && !ctx->trace && !getenv("VRS_NO_CULL") && (long long)ctx->grid.cdim[0] * ctx->grid.cdim[1] * ctx->grid.cdim[2] <= (1ll << 27);
}

// ---- the frame as phases.  Front: everything of the initial pass that is independent of the previous frame.  Back phases:
//   0             [join the pending halo push] temporal merge + visibility (k_finish); with spatial reuse on several GPUs, push
//                 the G-buffer + tmp reservoir rows spatial reuse can reach
//   1 .. iters    [join] spatial iteration i - 1; push the new reservoir rows when another iteration follows
//   iters + 1     shade; on several GPUs push what the next frame's temporal reprojection reads (all halo rows)
// A phase never waits for a push of the same phase, so phases of several contexts may be interleaved by one host thread in
// any order that keeps the phase number non-decreasing (vrs_render_frame_group).
static vrs_status enqueue_front_a(vrs_ctx* ctx, const FrameParams& F, const FrameIdx& fi, cudaStream_t st) {
  CK(mark(ctx, 0, st));
  launch_front_trace(st, ctx->grid, F, ctx->d_params + fi.q, planes_of(ctx, fi.g), ctx->queues[fi.q], ctx->trace,
                     ctx->band_y0, ctx->band_y1, ctx->store_y0, ctx->store_y1, ctx->persistent_blocks, &ctx->kt);
  CK(cudaGetLastError());
  ctx->timings.launches += (uint32_t)front_trace_launches(culling_on(ctx, F));
  return VRS_OK;
}
static vrs_status enqueue_front_b(vrs_ctx* ctx, const FrameParams& F, const FrameIdx& fi, cudaStream_t st) {
  launch_front_ris(st, ctx->grid, ctx->lights, F, ctx->d_params + fi.q, planes_of(ctx, fi.g), res_of(ctx, fi.ra), ctx->queues[fi.q], ctx->trace,
                   ctx->store_y0, ctx->persistent_blocks, &ctx->kt);
  CK(cudaGetLastError());
  ctx->timings.launches += (uint32_t)front_ris_launches(F.flags, ctx->lights);
  return VRS_OK;
}
static vrs_status enqueue_front_c(vrs_ctx* ctx, const FrameParams& F, const FrameIdx& fi, cudaStream_t st) {
  launch_front_shadow(st, ctx->grid, F, ctx->queues[fi.q], ctx->persistent_blocks, &ctx->kt);
  CK(cudaGetLastError());
  ctx->timings.launches += (uint32_t)front_shadow_launches(F.flags);
  ctx->src_r = fi.ra;
  return VRS_OK;
}
static vrs_status enqueue_front_stage(vrs_ctx* ctx, int stage, const FrameParams& F, const FrameIdx& fi, cudaStream_t st) {
  return stage == 0 ? enqueue_front_a(ctx, F, fi, st) : stage == 1 ? enqueue_front_b(ctx, F, fi, st) : enqueue_front_c(ctx, F, fi, st);
}
static vrs_status enqueue_front(vrs_ctx* ctx, const FrameParams& F, const FrameIdx& fi, cudaStream_t st) {
  for (int k = 0; k < VRS_NFRONT; ++k) { vrs_status s = enqueue_front_stage(ctx, k, F, fi, st); if (s) return s; }
  return VRS_OK;
}
static int back_phases(vrs_ctx* ctx, const FrameParams& F) {
  const bool spatial = (F.flags & VRS_RESTIR_SPATIAL_REUSE_FLAG) != 0 && ctx->cfg.spatial_iterations > 0;
  return 2 + (spatial ? (int)ctx->cfg.spatial_iterations : 0);
}
static vrs_status enqueue_spatial(vrs_ctx* ctx, const FrameParams& F, const FrameIdx& fi, uint32_t iteration, cudaStream_t st) {
  const int dst = ctx->src_r == fi.ra ? fi.rb : fi.ra;
  launch_spatial(st, ctx->lights, ctx->d_params + fi.q, planes_of(ctx, fi.g), res_of(ctx, ctx->src_r), res_of(ctx, dst), ctx->queues[fi.q],
                 iteration, ctx->store_y0, ctx->store_y1, &ctx->kt);
  CK(cudaGetLastError());
  ctx->src_r = dst;
  ctx->timings.launches += 1;
  return VRS_OK;
}
static vrs_status enqueue_back_phase(vrs_ctx* ctx, const FrameParams& F, const FrameIdx& fi, int phase, bool want_temporal_push, cudaStream_t st) {
  vrs_status s;
  const bool spatial = (F.flags & VRS_RESTIR_SPATIAL_REUSE_FLAG) != 0 && ctx->cfg.spatial_iterations > 0;
  const int iters = spatial ? (int)ctx->cfg.spatial_iterations : 0;
  const bool multi = ctx->comm != nullptr || ctx->peer_mode;
  int sp_rows = (int)ceilf(F.spatialRadius); if (sp_rows < 1) sp_rows = 1;      // |int(dy)| <= radius: the rows spatial reuse can reach
  if (phase == 0) {
    const bool needs_finish = (F.flags & (VRS_RESTIR_VISIBILITY_REUSE_FLAG | VRS_RESTIR_TEMPORAL_REUSE_FLAG)) != 0;
    // joins the exchange that closed the previous frame (peer memory: the neighbours' "final" flag; NCCL: their halo rows)
    if ((s = halo_wait(ctx, st))) return s;
    PrevAccess PA; memset(&PA, 0, sizeof(PA));
    if (ctx->peer_mode) {
      // peer memory: a reprojection that leaves this band reads the neighbour's planes of the previous frame in place
      PA.own_y0 = ctx->band_y0; PA.own_y1 = ctx->band_y1;
      auto peer_planes = [&](const vrs_ctx::Peer& P, Planes& pl, ResPlanes& rp) {
        pl.worldPos = P.g[fi.gprev][0]; pl.albedo = P.g[fi.gprev][1]; pl.normal = P.g[fi.gprev][2]; pl.mat = P.g[fi.gprev][3];
        rp.info = P.r[ctx->final_r][0]; rp.weight = P.r[ctx->final_r][1];
      };
      if (ctx->peer_up.present) { peer_planes(ctx->peer_up, PA.up, PA.upR); PA.up_y0 = ctx->peer_up.band_y0; PA.up_row0 = ctx->peer_up.store_y0; }
      if (ctx->peer_down.present) { peer_planes(ctx->peer_down, PA.down, PA.downR); PA.down_y1 = ctx->peer_down.band_y1; PA.down_row0 = ctx->peer_down.store_y0; }
    } else {
      // one GPU: every row; NCCL: the halo rows were shipped at the end of the previous frame
      PA.own_y0 = ctx->store_y0; PA.own_y1 = ctx->store_y1;
    }
    launch_initial_finish(st, ctx->lights, F, ctx->d_params + fi.q, planes_of(ctx, fi.g), planes_of(ctx, fi.gprev), res_of(ctx, ctx->final_r), res_of(ctx, fi.ra),
                          ctx->queues[fi.q], ctx->trace, ctx->store_y0, PA, ctx->xflags + 5, &ctx->kt);   // main.cpp:405-409
    CK(cudaGetLastError());
    if (needs_finish) ctx->timings.launches += 1;
    CK(mark(ctx, 1, st));
    if (multi && spatial && (s = halo_push(ctx, st, true, fi.g, ctx->src_r, sp_rows))) return s;
    CK(mark(ctx, 2, st));
  } else if (phase <= iters) {                                                                       // main.cpp:410-413
    if ((s = halo_wait(ctx, st))) return s;
    if ((s = enqueue_spatial(ctx, F, fi, (uint32_t)(phase - 1), st))) return s;
    if (multi && phase < iters && (s = halo_push(ctx, st, false, fi.g, ctx->src_r, sp_rows))) return s;
  } else {
    CK(mark(ctx, 3, st));
    // What the NEXT frame's temporal pass reads of this frame.  Peer memory: neighbours read this band's planes in place, so only
    // a flag travels ("this frame's G-buffer and final reservoirs are complete; my own temporal pass no longer reads yours of
    // the frame before") — it is published before the shade pass, which touches neither.  NCCL: the G-buffer + final reservoirs
    // of every halo row are shipped.  Either way the consumer-side wait is phase 0 of the next frame.
    const bool push_t = multi && want_temporal_push;
    if (push_t && ctx->peer_mode && (s = halo_push(ctx, st, false, fi.g, -1, 0))) return s;
    launch_shade(st, ctx->grid, ctx->lights, F, ctx->d_params + fi.q, planes_of(ctx, fi.g), res_of(ctx, ctx->src_r), ctx->accum, ctx->band_y0,
                 ctx->band_y1, ctx->store_y0, &ctx->kt);                                             // main.cpp:416-433
    CK(cudaGetLastError());
    ctx->timings.launches += 1;
    CK(mark(ctx, 4, st));
    if (push_t && !ctx->peer_mode && (s = halo_push(ctx, st, true, fi.g, ctx->src_r, 1 << 30))) return s;
  }
  return VRS_OK;
}
// host-side state of a completed frame (Renderer::updateGBufferFrameIdx, Renderer.cpp:108-111, generalised)
static void finish_frame_state(vrs_ctx* ctx, const FrameIdx& fi) {
  // a later front half that reuses this frame's queue set / parameter block / G-buffer slot waits for this event
  if (cudaEventRecord(ctx->ev_back_done[fi.q], ctx->stream) == cudaSuccess) ctx->back_recorded[fi.q] = true;
  ctx->final_r = ctx->src_r; ctx->last_g = fi.g; ctx->last_q = fi.q;
  ctx->frame_no++;
  ctx->history_valid = true;
}

// Graph cache: the launch sequence of a stage depends only on the buffer rotation (frame_no % 12 = lcm of the ring lengths), on where the previous
// frame left its final reservoirs, on the structural flags and on whether a halo push is pending.
static uint64_t graph_key(vrs_ctx* ctx, const FrameParams& F, int half, bool want_temporal_push) {
  // half: 1 = back, 2 + k = front stage k
  uint64_t k = (uint64_t)(ctx->frame_no % 12) | ((uint64_t)(F.flags & 0x3f) << 8) | ((uint64_t)(ctx->cfg.spatial_iterations & 7) << 14) |
               ((uint64_t)(F.cull ? 1 : 0) << 17) | ((uint64_t)(ctx->pass_timing ? 1 : 0) << 18) | ((uint64_t)half << 24);
  if (half == 1) k |= ((uint64_t)(ctx->final_r & 15) << 4) | ((uint64_t)(ctx->peer_mode ? 1 : 0) << 19) | ((uint64_t)(ctx->halo_pending ? 1 : 0) << 20) |
                      ((uint64_t)(want_temporal_push ? 1 : 0) << 21);
  return k;
}
static vrs_status run_captured(vrs_ctx* ctx, uint64_t key, cudaStream_t st, const std::function<vrs_status()>& body) {
  auto it = ctx->graphs.find(key);
  if (it == ctx->graphs.end()) {
    cudaGraph_t graph = nullptr;
    const uint32_t before = ctx->timings.launches;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    ctx->capturing = true;
    vrs_status s = body();
    ctx->capturing = false;
    cudaError_t e = cudaStreamEndCapture(st, &graph);
    if (s) { if (graph) cudaGraphDestroy(graph); return s; }
    if (e != cudaSuccess) { ctx->err = std::string("graph capture: ") + cudaGetErrorString(e); return VRS_ERR_CUDA; }
    GraphEntry ge;
    e = cudaGraphInstantiate(&ge.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { ctx->err = std::string("graph instantiate: ") + cudaGetErrorString(e); return VRS_ERR_CUDA; }
    ge.launches = ctx->timings.launches - before;
    it = ctx->graphs.emplace(key, ge).first;
  } else {
    // replay: the body only advances host-side state (buffer rotation, pending marks); it enqueues nothing
    ctx->replaying = true;
    vrs_status s = body();
    ctx->replaying = false;
    if (s) return s;
    ctx->timings.launches += it->second.launches;
  }
  CK(cudaGraphLaunch(it->second.exec, st));
  return VRS_OK;
}

vrs_status vrs_pass_initial(vrs_ctx* ctx, const vrs_global_uniforms* gu, const vrs_restir_uniforms* ru, uint32_t clock) {
  if (!ctx || !gu || !ru) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  FrameParams F; vrs_status s = make_params(ctx, gu, ru, nullptr, clock, F); if (s) return s;
  const FrameIdx fi = frame_idx(ctx->frame_no);
  ctx->cur = fi;
  sync_front_streams(ctx);                               // the per-pass calls run on the main stream, strictly in order
  if ((s = upload_params(ctx, F, fi.q, ctx->stream))) return s;
  if ((s = enqueue_front(ctx, F, fi, ctx->stream))) return s;
  return enqueue_back_phase(ctx, F, fi, 0, false, ctx->stream);
}
vrs_status vrs_pass_spatial(vrs_ctx* ctx, const vrs_restir_uniforms* ru, uint32_t clock, uint32_t iteration) {
  if (!ctx || !ru || iteration >= VRS_MAX_SPATIAL_ITERATIONS) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  FrameParams F; vrs_status s = make_params(ctx, nullptr, ru, nullptr, clock, F); if (s) return s;
  if ((s = upload_params(ctx, F, ctx->cur.q, ctx->stream))) return s;
  return enqueue_spatial(ctx, F, ctx->cur, iteration, ctx->stream);
}
vrs_status vrs_pass_shade(vrs_ctx* ctx, const vrs_restir_uniforms* ru, const vrs_push_constant_restir* pc, uint32_t clock) {
  if (!ctx || !ru || !pc) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  FrameParams F; vrs_status s = make_params(ctx, nullptr, ru, pc, clock, F); if (s) return s;
  if ((s = upload_params(ctx, F, ctx->cur.q, ctx->stream))) return s;
  launch_shade(ctx->stream, ctx->grid, ctx->lights, F, ctx->d_params + ctx->cur.q, planes_of(ctx, ctx->cur.g), res_of(ctx, ctx->src_r), ctx->accum, ctx->band_y0,
               ctx->band_y1, ctx->store_y0, &ctx->kt);
  CK(cudaGetLastError());
  finish_frame_state(ctx, ctx->cur);
  return VRS_OK;
}

// One frame in main.cpp:405-433 order.
vrs_status vrs_render_frame(vrs_ctx* ctx, const vrs_global_uniforms* gu, const vrs_restir_uniforms* ru, const vrs_push_constant_restir* pc, uint32_t clock) {
  if (!ctx || !gu || !ru || !pc) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  FrameParams F; vrs_status s = make_params(ctx, gu, ru, pc, clock, F); if (s) return s;
  ctx->timings.launches = 0;
  static const bool no_graph = getenv("VRS_NO_GRAPH") != nullptr;
  // NCCL send/recv under stream capture hangs here (NCCL 2.28.9): NCCL contexts launch eagerly.  Per-kernel timing and
  // per-pass timing want one stream with nothing overlapping the kernels they bracket.
  const bool eager = no_graph || ctx->comm || ctx->kt.on;
  // (the diagnostic trace buffer is one per context, written by every stage: frames must not overlap while it is on)
  const bool overlap = ctx->pipeline && !ctx->kt.on && !ctx->pass_timing && !ctx->trace;
  const FrameIdx fi = frame_idx(ctx->frame_no);
  ctx->cur = fi;
  const bool want_temporal_push = (ru->flags & VRS_RESTIR_TEMPORAL_REUSE_FLAG) != 0;
  // stream of front stage k: depth 4 = one stream each, 3 = B and C share one, 2 = all three share one
  cudaStream_t sst[VRS_NFRONT];
  for (int k = 0; k < VRS_NFRONT; ++k) sst[k] = !overlap ? ctx->stream : ctx->fstream[ctx->depth >= 4 ? k : (ctx->depth == 3 ? (k ? 1 : 0) : 0)];
  if (ctx->kt.on) { ctx->kt.n = 0; CK(cudaEventRecord(ctx->kt.ev[0], ctx->stream)); }
  // stage A needs the queue set / parameter block that frame n - 4 used, and a G-buffer slot nobody reads any more
  if (overlap && ctx->back_recorded[fi.q]) CK(cudaStreamWaitEvent(sst[0], ctx->ev_back_done[fi.q], 0));
  if ((s = upload_params(ctx, F, fi.q, sst[0]))) return s;
  for (int k = 0; k < VRS_NFRONT; ++k) {
    if (overlap && k > 0 && sst[k] != sst[k - 1]) { CK(cudaEventRecord(ctx->ev_stage_done[k - 1][fi.q], sst[k - 1])); CK(cudaStreamWaitEvent(sst[k], ctx->ev_stage_done[k - 1][fi.q], 0)); }
    if (eager) s = enqueue_front_stage(ctx, k, F, fi, sst[k]);
    else s = run_captured(ctx, graph_key(ctx, F, 2 + k, false), sst[k], [&]() -> vrs_status {
      if (ctx->replaying) { if (k == VRS_NFRONT - 1) ctx->src_r = fi.ra; return VRS_OK; }
      return enqueue_front_stage(ctx, k, F, fi, sst[k]);
    });
    if (s) return s;
  }
  if (overlap) { CK(cudaEventRecord(ctx->ev_stage_done[VRS_NFRONT - 1][fi.q], sst[VRS_NFRONT - 1])); CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_stage_done[VRS_NFRONT - 1][fi.q], 0)); }
  // ---- back half
  const int nph = back_phases(ctx, F);
  auto back = [&]() -> vrs_status {
    if (ctx->replaying) {       // host-side state only: reservoir rotation and the pending mark of the last push
      const bool spatial = (F.flags & VRS_RESTIR_SPATIAL_REUSE_FLAG) != 0 && ctx->cfg.spatial_iterations > 0;
      if (spatial) for (uint32_t i = 0; i < ctx->cfg.spatial_iterations; ++i) ctx->src_r = ctx->src_r == fi.ra ? fi.rb : fi.ra;
      ctx->halo_pending = (ctx->comm != nullptr || ctx->peer_mode) && want_temporal_push;
      return VRS_OK;
    }
    for (int ph = 0; ph < nph; ++ph) { vrs_status s2 = enqueue_back_phase(ctx, F, fi, ph, want_temporal_push, ctx->stream); if (s2) return s2; }
    return VRS_OK;
  };
  s = eager ? back() : run_captured(ctx, graph_key(ctx, F, 1, want_temporal_push), ctx->stream, back);
  if (s) return s;
  finish_frame_state(ctx, fi);
  ctx->timings_valid = ctx->pass_timing;
  return VRS_OK;
}

// The same frame on several contexts driven by ONE host thread (bands of one image: several GPUs of one process, or several
// bands on one GPU): phase k of every context is enqueued before phase k + 1 of any, so a context's wait for its neighbours'
// rows is always enqueued after the pushes it waits for — safe even when the contexts' streams share a hardware queue.
vrs_status vrs_render_frame_group(vrs_ctx** ctxs, uint32_t n, const vrs_global_uniforms* gu, const vrs_restir_uniforms* ru,
                                  const vrs_push_constant_restir* pc, uint32_t clock) {
  if (!ctxs || n == 0 || !gu || !ru || !pc) return VRS_ERR_INVALID;
  std::vector<FrameParams> F(n);
  std::vector<FrameIdx> fi(n);
  const bool want_temporal_push = (ru->flags & VRS_RESTIR_TEMPORAL_REUSE_FLAG) != 0;
  for (uint32_t i = 0; i < n; ++i) {
    vrs_ctx* ctx = ctxs[i];
    if (!ctx) return VRS_ERR_INVALID;
    if (ctx->comm) return fail(ctx, VRS_ERR_INVALID, "vrs_render_frame_group drives peer-memory contexts (vrs_peer_connect_local), not NCCL ones");
    cudaSetDevice(ctx->device);
    vrs_status s = make_params(ctx, gu, ru, pc, clock, F[i]); if (s) return s;
    ctx->timings.launches = 0;
    fi[i] = frame_idx(ctx->frame_no); ctx->cur = fi[i];
    sync_front_streams(ctx);
    if ((s = upload_params(ctx, F[i], fi[i].q, ctx->stream))) return s;
    if ((s = enqueue_front(ctx, F[i], fi[i], ctx->stream))) return s;
  }
  const int nph = back_phases(ctxs[0], F[0]);
  for (int ph = 0; ph < nph; ++ph)
    for (uint32_t i = 0; i < n; ++i) {
      vrs_ctx* ctx = ctxs[i];
      cudaSetDevice(ctx->device);
      if (back_phases(ctx, F[i]) != nph) return fail(ctx, VRS_ERR_INVALID, "contexts of a group must share spatial_iterations");
      vrs_status s = enqueue_back_phase(ctx, F[i], fi[i], ph, want_temporal_push, ctx->stream); if (s) return s;
    }
  for (uint32_t i = 0; i < n; ++i) { finish_frame_state(ctxs[i], fi[i]); ctxs[i]->timings_valid = false; }
  return VRS_OK;
}

vrs_status vrs_synchronize(vrs_ctx* ctx) {
  if (!ctx) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  for (int i = 0; i < VRS_NFRONT; ++i) CK(cudaStreamSynchronize(ctx->fstream[i]));
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->comm_stream) CK(cudaStreamSynchronize(ctx->comm_stream));
  if (ctx->peer_mode) {          // did a halo wait give up?  (k_halo_wait sets xflags[4]; stale halo rows must not pass as VRS_OK)
    unsigned err = 0;
    CK(cudaMemcpy(&err, ctx->xflags + 4, sizeof(err), cudaMemcpyDeviceToHost));
    if (err) {
      ctx->comm_timeouts += err;
      CK(cudaMemset(ctx->xflags + 4, 0, sizeof(unsigned)));
      return fail(ctx, VRS_ERR_COMM, "halo exchange timed out: a neighbouring band never published its rows; the frame was computed on stale halo rows");
    }
  }
  return VRS_OK;
}

vrs_status vrs_set_pass_timing(vrs_ctx* ctx, int enabled) {
  if (!ctx) return VRS_ERR_INVALID;
  ctx->pass_timing = enabled != 0;
  if (!ctx->pass_timing) ctx->timings_valid = false;
  return VRS_OK;
}

vrs_status vrs_get_timings(vrs_ctx* ctx, vrs_timings* out) {
  if (!ctx || !out || !ctx->timings_valid) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  CK(cudaEventSynchronize(ctx->ev[4]));
  float a, b, c, d;
  CK(cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1])); CK(cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]));
  CK(cudaEventElapsedTime(&c, ctx->ev[2], ctx->ev[3])); CK(cudaEventElapsedTime(&d, ctx->ev[3], ctx->ev[4]));
  ctx->timings.initial_ms = a; ctx->timings.spatial_ms = c; ctx->timings.shade_ms = d; ctx->timings.exchange_ms = b;
  ctx->timings.frame_ms = a + b + c + d;
  *out = ctx->timings;
  return VRS_OK;
}
void* vrs_stream(vrs_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// ------------------------------------------------------------------------------------------ readback
static vrs_status read_plane(vrs_ctx* ctx, const void* dev, void* host, size_t elem) {
  if (!host) return VRS_OK;
  size_t off = (size_t)(ctx->band_y0 - ctx->store_y0) * ctx->W * elem;
  size_t bytes = (size_t)(ctx->band_y1 - ctx->band_y0) * ctx->W * elem;
  CK(cudaMemcpyAsync(host, (const char*)dev + off, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return VRS_OK;
}
vrs_status vrs_read_frame(vrs_ctx* ctx, float* rgba) {
  if (!ctx || !rgba) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  vrs_status s = read_plane(ctx, ctx->accum, rgba, 16); if (s) return s;
  CK(cudaStreamSynchronize(ctx->stream));
  return VRS_OK;
}
// G-buffer / reservoir readback goes through k_export: miss pixels carry no data on the device
static vrs_status export_planes(vrs_ctx* ctx, int g_index, int r_index, float* dst[6]) {
  const size_t first = (size_t)(ctx->band_y0 - ctx->store_y0) * ctx->W, n = (size_t)(ctx->band_y1 - ctx->band_y0) * ctx->W;
  float4* tmp = nullptr;
  CK(cudaMalloc(&tmp, n * 16 * 6));
  launch_export(ctx->stream, planes_of(ctx, g_index), res_of(ctx, r_index), tmp, first, n);
  cudaError_t e = cudaGetLastError();
  for (int p = 0; p < 6 && e == cudaSuccess; ++p)
    if (dst[p]) e = cudaMemcpyAsync(dst[p], tmp + (size_t)p * n, n * 16, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(tmp);
  if (e != cudaSuccess) { ctx->err = std::string("readback: ") + cudaGetErrorString(e); return VRS_ERR_CUDA; }
  return VRS_OK;
}
vrs_status vrs_read_gbuffer(vrs_ctx* ctx, float* worldPos, float* albedo, float* normal, float* matProps) {
  if (!ctx) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  float* dst[6] = {worldPos, albedo, normal, matProps, nullptr, nullptr};
  return export_planes(ctx, ctx->last_g, ctx->src_r, dst);
}
vrs_status vrs_read_reservoirs(vrs_ctx* ctx, float* info, float* weight) {
  if (!ctx) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  float* dst[6] = {nullptr, nullptr, nullptr, nullptr, info, weight};
  return export_planes(ctx, ctx->last_g, ctx->src_r, dst);
}
vrs_status vrs_read_trace(vrs_ctx* ctx, uint32_t* trace4) {
  if (!ctx || !trace4 || !ctx->trace) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  vrs_status s = read_plane(ctx, ctx->trace, trace4, 16); if (s) return s;
  CK(cudaStreamSynchronize(ctx->stream));
  return VRS_OK;
}

vrs_status vrs_present_async(vrs_ctx* ctx, uint8_t* rgba8) {
  if (!ctx || !rgba8) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (!ctx->copy_stream) CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  const int b = (int)(ctx->present_count & 1u);
  if (ctx->present_count >= 2) CK(cudaStreamWaitEvent(ctx->stream, ctx->copy_done[b], 0));   // staging buffer is free again
  const size_t off = (size_t)(ctx->band_y0 - ctx->store_y0) * ctx->W, n = (size_t)(ctx->band_y1 - ctx->band_y0) * ctx->W;
  launch_display(ctx->stream, ctx->accum + off, ctx->display[b], n);
  CK(cudaGetLastError());
  CK(cudaEventRecord(ctx->display_ready[b], ctx->stream));
  CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->display_ready[b], 0));
  CK(cudaMemcpyAsync(rgba8, ctx->display[b], n * 4, cudaMemcpyDeviceToHost, ctx->copy_stream));
  CK(cudaEventRecord(ctx->copy_done[b], ctx->copy_stream));
  ctx->present_count++;
  return VRS_OK;
}
vrs_status vrs_present_wait(vrs_ctx* ctx) {
  if (!ctx) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (ctx->copy_stream) CK(cudaStreamSynchronize(ctx->copy_stream));
  return VRS_OK;
}
vrs_status vrs_read_display(vrs_ctx* ctx, uint8_t* rgba8) {
  vrs_status s = vrs_present_async(ctx, rgba8); if (s) return s;
  return vrs_present_wait(ctx);
}

vrs_status vrs_write_image(vrs_ctx* ctx, const char* path) {
  if (!ctx || !path) return VRS_ERR_INVALID;
  size_t rows = (size_t)(ctx->band_y1 - ctx->band_y0), n = rows * ctx->W;
  std::vector<float> img(n * 4);
  vrs_status s = vrs_read_frame(ctx, img.data()); if (s) return s;
  std::string p(path);
  FILE* f = fopen(path, "wb");
  if (!f) return fail(ctx, VRS_ERR_IO, "cannot write " + p);
  if (p.size() > 4 && p.compare(p.size() - 4, 4, ".pfm") == 0) {
    fprintf(f, "PF\n%u %zu\n-1.0\n", ctx->W, rows);
    for (size_t y = rows; y-- > 0;) for (uint32_t x = 0; x < ctx->W; ++x) fwrite(&img[(y * ctx->W + x) * 4], 4, 3, f);
  } else {
    fprintf(f, "P6\n%u %zu\n255\n", ctx->W, rows);
    for (size_t i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) {
      float v = powf(img[i * 4 + c] < 0.f ? 0.f : img[i * 4 + c], 1.0f / 0.8f);      // restir_post.frag:104
      v = v > 1.f ? 1.f : v;
      fputc((int)(v * 255.f + 0.5f), f);
    }
  }
  fclose(f);
  return VRS_OK;
}

// ------------------------------------------------------------------------------------------ multi-GPU
// ---- peer-memory exchange set-up: blob = { band_y0, band_y1, store_y0, store_y1 (int32) | 49 x cudaIpcMemHandle_t }
vrs_status vrs_peer_export(vrs_ctx* ctx, uint8_t blob[VRS_PEER_BLOB_BYTES]) {
  if (!ctx || !blob) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  static_assert(16 + (VRS_NG * 4 + VRS_NR * 2 + 1) * sizeof(cudaIpcMemHandle_t) <= VRS_PEER_BLOB_BYTES, "blob too small");
  memset(blob, 0, VRS_PEER_BLOB_BYTES);
  int32_t hdr[4] = {ctx->band_y0, ctx->band_y1, ctx->store_y0, ctx->store_y1};
  memcpy(blob, hdr, 16);
  cudaIpcMemHandle_t* h = (cudaIpcMemHandle_t*)(blob + 16);
  int k = 0;
  for (int i = 0; i < VRS_NG; ++i) for (int p = 0; p < 4; ++p) CK(cudaIpcGetMemHandle(&h[k++], ctx->g_planes[i][p]));
  for (int i = 0; i < VRS_NR; ++i) for (int p = 0; p < 2; ++p) CK(cudaIpcGetMemHandle(&h[k++], ctx->r_planes[i][p]));
  CK(cudaIpcGetMemHandle(&h[k++], ctx->xflags));
  return VRS_OK;
}
vrs_status vrs_peer_connect(vrs_ctx* ctx, int rank, int nranks, const uint8_t* all_blobs) {
  if (!ctx || !all_blobs || rank < 0 || rank >= nranks) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (ctx->peer_mode || ctx->comm) return fail(ctx, VRS_ERR_INVALID, "context is already connected to its neighbours");
  if (nranks > 1 && ctx->band_y1 - ctx->band_y0 < (int)ctx->cfg.halo_rows)
    return fail(ctx, VRS_ERR_INVALID, "band shorter than halo_rows: halos are exchanged with adjacent ranks only");
  auto open_peer = [&](int r, vrs_ctx::Peer& P) -> vrs_status {
    const uint8_t* blob = all_blobs + (size_t)r * VRS_PEER_BLOB_BYTES;
    int32_t hdr[4]; memcpy(hdr, blob, 16);
    P.band_y0 = hdr[0]; P.band_y1 = hdr[1]; P.store_y0 = hdr[2]; P.store_y1 = hdr[3];
    const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)(blob + 16);
    int k = 0;
    auto open = [&](void** out) -> vrs_status { cudaIpcMemHandle_t hh; memcpy(&hh, &h[k++], sizeof(hh)); CK(cudaIpcOpenMemHandle(out, hh, cudaIpcMemLazyEnablePeerAccess)); P.opened.push_back(*out); return VRS_OK; };
    vrs_status s;
    for (int i = 0; i < VRS_NG; ++i) for (int p = 0; p < 4; ++p) if ((s = open((void**)&P.g[i][p]))) return s;
    for (int i = 0; i < VRS_NR; ++i) for (int p = 0; p < 2; ++p) if ((s = open((void**)&P.r[i][p]))) return s;
    if ((s = open((void**)&P.flags))) return s;
    P.present = true;
    return VRS_OK;
  };
  vrs_status s;
  if (rank > 0 && (s = open_peer(rank - 1, ctx->peer_up))) return s;
  if (rank + 1 < nranks && (s = open_peer(rank + 1, ctx->peer_down))) return s;
  ctx->peer_mode = nranks > 1;
  return VRS_OK;
}

vrs_status vrs_comm_unique_id(uint8_t id128[128]) {
  std::string err;
  if (!comm_unique_id(id128, err)) return fail(nullptr, VRS_ERR_COMM, err);
  return VRS_OK;
}
vrs_status vrs_comm_init(vrs_ctx* ctx, const uint8_t id128[128], int rank, int nranks) {
  if (!ctx || !id128 || rank < 0 || rank >= nranks) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (nranks > 1 && ctx->band_y1 - ctx->band_y0 < (int)ctx->cfg.halo_rows)
    return fail(ctx, VRS_ERR_INVALID, "band shorter than halo_rows: halos are exchanged with adjacent ranks only");
  if (ctx->peer_mode || ctx->comm) return fail(ctx, VRS_ERR_INVALID, "context is already connected to its neighbours");
  std::string err;
  if (!ctx->comm_stream) CK(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
  ctx->comm = comm_create(id128, rank, nranks, err);
  if (!ctx->comm) return fail(ctx, VRS_ERR_COMM, err);
  return VRS_OK;
}

vrs_status vrs_peer_connect_local(vrs_ctx* ctx, vrs_ctx* up, vrs_ctx* down) {
  if (!ctx) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if ((up || down) && ctx->band_y1 - ctx->band_y0 < (int)ctx->cfg.halo_rows)
    return fail(ctx, VRS_ERR_INVALID, "band shorter than halo_rows: halos are exchanged with adjacent bands only");
  auto wire = [&](vrs_ctx* o, vrs_ctx::Peer& P) -> vrs_status {
    if (!o) return VRS_OK;
    if (o->W != ctx->W || o->H != ctx->H) return fail(ctx, VRS_ERR_INVALID, "peer context renders another image size");
    if (o->device != ctx->device) {
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, ctx->device, o->device));
      if (!can) return fail(ctx, VRS_ERR_UNSUPPORTED, "no peer access between the devices of the two contexts");
      cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ctx->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); return VRS_ERR_CUDA; }
      cudaGetLastError();
    }
    P.band_y0 = o->band_y0; P.band_y1 = o->band_y1; P.store_y0 = o->store_y0; P.store_y1 = o->store_y1;
    for (int i = 0; i < VRS_NG; ++i) for (int p = 0; p < 4; ++p) P.g[i][p] = o->g_planes[i][p];
    for (int i = 0; i < VRS_NR; ++i) for (int p = 0; p < 2; ++p) P.r[i][p] = o->r_planes[i][p];
    P.flags = o->xflags;
    P.present = true;
    return VRS_OK;
  };
  vrs_status s;
  if ((s = wire(up, ctx->peer_up))) return s;
  if ((s = wire(down, ctx->peer_down))) return s;
  if (up && up->band_y1 != ctx->band_y0) return fail(ctx, VRS_ERR_INVALID, "the up neighbour's band does not end where this band starts");
  if (down && down->band_y0 != ctx->band_y1) return fail(ctx, VRS_ERR_INVALID, "the down neighbour's band does not start where this band ends");
  invalidate_graphs(ctx);
  ctx->peer_mode = up || down;
  return VRS_OK;
}

vrs_status vrs_resize(vrs_ctx* ctx, uint32_t width, uint32_t height) {
  if (!ctx || width == 0 || height == 0) return VRS_ERR_INVALID;
  if (ctx->comm || ctx->peer_mode) return fail(ctx, VRS_ERR_INVALID, "vrs_resize: context belongs to a multi-GPU group (its neighbours hold pointers into its planes)");
  cudaSetDevice(ctx->device);
  invalidate_graphs(ctx);
  if (ctx->copy_stream) CK(cudaStreamSynchronize(ctx->copy_stream));
  free_frame_buffers(ctx);
  ctx->cfg.width = width; ctx->cfg.height = height; ctx->cfg.band_y0 = 0; ctx->cfg.band_y1 = 0;
  vrs_status s = alloc_frame_buffers(ctx);
  if (s) return s;
  ctx->timings_valid = false;
  CK(cudaDeviceSynchronize());
  return VRS_OK;
}

vrs_status vrs_get_counters(vrs_ctx* ctx, vrs_counters* out) {
  if (!ctx || !out) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  uint32_t q[8] = {0}; unsigned x[8] = {0};
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpy(q, ctx->queues[ctx->last_q].counters, sizeof(q), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(x, ctx->xflags, sizeof(x), cudaMemcpyDeviceToHost));
  out->candidates = q[0]; out->hits = q[1]; out->shadow_rays = q[2];
  out->temporal_out_of_halo = x[5];
  out->comm_timeouts = ctx->comm_timeouts + x[4];
  out->temporal_reach_rows = x[6];
  return VRS_OK;
}

vrs_status vrs_set_kernel_timing(vrs_ctx* ctx, int enabled) {
  if (!ctx) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  if (enabled && !ctx->kt_events) {
    for (int i = 0; i <= KTimer::MAX; ++i) CK(cudaEventCreate(&ctx->kt.ev[i]));
    ctx->kt_events = true;
  }
  ctx->kt.on = enabled != 0; ctx->kt.n = 0;
  return VRS_OK;
}
vrs_status vrs_get_kernel_times(vrs_ctx* ctx, vrs_kernel_time* out, uint32_t capacity, uint32_t* count) {
  if (!ctx || !out || !count || !ctx->kt.on) return VRS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  CK(cudaStreamSynchronize(ctx->stream));
  uint32_t n = 0;
  for (int i = 0; i < ctx->kt.n && n < capacity; ++i, ++n) {
    memset(&out[n], 0, sizeof(out[n]));
    strncpy(out[n].name, ctx->kt.name[i], sizeof(out[n].name) - 1);
    CK(cudaEventElapsedTime(&out[n].ms, ctx->kt.ev[i], ctx->kt.ev[i + 1]));
  }
  *count = n;
  return VRS_OK;
}

}  // extern "C"
