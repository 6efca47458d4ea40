// Hot-path kernels (sm_100a): initial (primary volume event + RIS + visibility + temporal), spatial reuse,
// final shade.  No tensor cores / RT cores: nothing here is a dense contraction (BASELINE.json north_star).
// Compile with -fmad=false (see vrs_device.cuh).
#include "vrs_device.cuh"
#include "vrs_kernels.h"

#include <cstdlib>

namespace vrs {

static constexpr uint32_t PASS_INITIAL = 0, PASS_SPATIAL0 = 1, PASS_SHADE = 5;
static constexpr int FLAG_VISIBILITY = 1 << 0, FLAG_TEMPORAL = 1 << 1;
static constexpr int FLAG_FINAL_VISIBILITY = 1 << 4, FLAG_FINALIZE_W = 1 << 5;
static constexpr int MAX_NEIGHBORS = 16;

// -------------------------------------------------------------------------------------------------
// Initial pass — restir.rgen main (:136-290) on a volume, as a chain of kernels over device-side work lists.
// The pass is instruction-issue bound (ncu, profiles/): what matters is how many of the 32 lanes do useful work and
// how few bytes the empty part of the screen costs, so each stage runs on the smallest, densest set it can.  The host
// runtime runs the chain as three stages on three streams plus the back half of the frame (vrs_api.cu, frames in flight):
//   stage A  k_cover        non-empty cells of the grid -> screen tiles that can contain a collision
//            k_classify     every pixel: writes worldPos = 0 (the miss marker, 16 B/px); in covered tiles the primary ray
//                           (:142-148) vs. the grid window, queues the rays that enter it
//            k_primary      persistent warps: residual delta tracking of queued rays, unit steps batched per warp, lane
//                           refill; a real collision leaves {t, cell, RNG state, voxel in cell} in the pixel's worldPos slot
//            k_hit_compact  stream compaction of the hit flags -> hit list, pixel order inside 2048-pixel blocks
//   stage B  k_ris_*        per hit: G-buffer stores (:193-197), M-candidate RIS (:203-227), shadow-ray set-up
//   stage C  k_shadow       persistent warps: ratio-tracking transmittance toward the selected light (:229-235)
//   back     k_finish       per hit: apply the transmittance, temporal merge (:237-284), pack (:286-289)
// -------------------------------------------------------------------------------------------------
enum { Q_CAND = 0, Q_HIT = 1, Q_SHADOW = 2, Q_PRIMARY_HEAD = 3, Q_SHADOW_HEAD = 4, Q_COVER_ALL = 5 };
static constexpr int COVER_TILE = 8;            // pixels per side of a coverage tile
static constexpr int REFILL_MIN_IDLE = 16;
static constexpr int CELLS_PER_DECISION = 2;
static constexpr uint32_t RIS_SMALL_LAUNCH = 100000u;   // hit pixels below which the cooperative RIS kernel is used even with small light tables
static constexpr int MIN_WARPS_PER_SM = 16;     // small launches are spread over at least this many warps per SM
static constexpr int COMPACT_BLOCK = 2048;      // pixels per compaction block (256 threads x 8 flags)

__device__ __forceinline__ void primary_ray(const FrameParams& F, int x, int y, V3& org, V3& dir) {   // :142-148
  float ux = float(x) / float(F.W), uy = float(y) / float(F.H);
  float dx = ux * 2.0f - 1.0f, dy = uy * 2.0f - 1.0f;
  float o4[4], t4[4], d4[4];
  mat_vec(F.viewInverse, 0.0f, 0.0f, 0.0f, 1.0f, o4);
  mat_vec(F.projInverse, dx, dy, 1.0f, 1.0f, t4);
  V3 tn = normalize(v3(t4[0], t4[1], t4[2]));
  mat_vec(F.viewInverse, tn.x, tn.y, tn.z, 0.0f, d4);
  org = v3(o4[0], o4[1], o4[2]); dir = v3(d4[0], d4[1], d4[2]);
}


// A-1 k_cover: screen-space bound of the occupied part of the grid.  Every non-empty 8^3 cell of the directory is
// projected (8 corners through the inverse of the primary-ray matrices, padded by 2 pixels) and the 8x8-pixel tiles its
// bounding rectangle touches are marked.  A primary ray can only collide inside a non-empty cell, so a pixel in an
// unmarked tile is a miss without marching: k_classify does not queue it.  Cells that reach behind the eye, or that
// would cover a large part of the screen (camera inside the volume), switch the mask off for the frame instead.
__global__ void __launch_bounds__(128) k_cover(const GridDev G, const FrameParams* __restrict__ Fp, Queues Q, int tiles_x, int tiles_y, int band_ty0, int band_ty1) {
  const FrameParams& F = *Fp;
  // 8 lanes per cell: lane k projects corner k, the bounding rectangle is reduced with shuffles, the lanes share the tile loop
  const int ncell = G.cdim[0] * G.cdim[1] * G.cdim[2];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int ci = t >> 3, k = t & 7;
  const bool live = ci < ncell && __ldg(&G.dir[ci < ncell ? ci : 0]).x > 0.0f;
  float minx = 3.0e38f, maxx = -3.0e38f, miny = 3.0e38f, maxy = -3.0e38f;
  bool bad = false;
  if (live) {
    const int c[3] = {ci % G.cdim[0], (ci / G.cdim[0]) % G.cdim[1], ci / (G.cdim[0] * G.cdim[1])};
    float w3[3];
    for (int a = 0; a < 3; ++a) w3[a] = (float(G.vmin[a] + 8 * (c[a] + ((k >> a) & 1))) - 0.5f) * G.A + G.B[a];   // voxel ijk covers [ijk - 1/2, ijk + 1/2)
    float q[4];
    mat_vec(F.cullVP, w3[0], w3[1], w3[2], 1.0f, q);
    if (!(q[3] > 1e-6f)) bad = true;
    else {
      const float px = (q[0] / q[3] + 1.0f) * 0.5f * float(F.W), py = (q[1] / q[3] + 1.0f) * 0.5f * float(F.H);   // pixel x <-> NDC 2x/W - 1 (:142-145)
      if (!(px == px && py == py)) bad = true;
      minx = maxx = px; miny = maxy = py;
    }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {                          // all 8 lanes of a cell share `live`
    minx = fminf(minx, __shfl_xor_sync(0xffffffffu, minx, o)); maxx = fmaxf(maxx, __shfl_xor_sync(0xffffffffu, maxx, o));
    miny = fminf(miny, __shfl_xor_sync(0xffffffffu, miny, o)); maxy = fmaxf(maxy, __shfl_xor_sync(0xffffffffu, maxy, o));
    bad = bad || __shfl_xor_sync(0xffffffffu, (int)bad, o) != 0;
  }
  if (!live) return;
  if (bad) { if (k == 0) Q.counters[Q_COVER_ALL] = 1u; return; }
  minx = floorf(minx) - 2.0f; miny = floorf(miny) - 2.0f; maxx = ceilf(maxx) + 2.0f; maxy = ceilf(maxy) + 2.0f;
  if (maxx < 0.0f || maxy < 0.0f || minx > float(F.W - 1) || miny > float(F.H - 1)) return;                       // off screen
  const int tx0 = (int)fmaxf(minx, 0.0f) / COVER_TILE;
  int ty0 = (int)fmaxf(miny, 0.0f) / COVER_TILE;
  const int tx1 = (int)fminf(maxx, float(F.W - 1)) / COVER_TILE;
  int ty1 = (int)fminf(maxy, float(F.H - 1)) / COVER_TILE;
  if (ty0 < band_ty0) ty0 = band_ty0;                         // only the tile rows of this context's band are ever tested (k_classify)
  if (ty1 > band_ty1) ty1 = band_ty1;
  if (ty1 < ty0) return;
  const int nx = tx1 - tx0 + 1, ny = ty1 - ty0 + 1;
  if ((long long)nx * ny > 4096) { if (k == 0) Q.counters[Q_COVER_ALL] = 1u; return; }
  for (int i = k; i < nx * ny; i += 8) {
    const int ty = ty0 + i / nx, tx = tx0 + i % nx;
    if (ty < tiles_y && tx < tiles_x) Q.cover[(size_t)ty * tiles_x + tx] = 1;
  }
}

__global__ void __launch_bounds__(256) k_classify(const GridDev G, const FrameParams* __restrict__ Fp, Planes cur, Queues Q,
                                                  uint32_t* __restrict__ trace, int y0, int y1, int store_y0, int tiles_x) {
  const FrameParams& F = *Fp;
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = y0 + blockIdx.y * 8 + threadIdx.y;
  bool enters = false;
  size_t idx = 0;
  RaySeg seg;
  if (x < (int)F.W && y < y1) {
    idx = (size_t)(y - store_y0) * F.W + x;
    // a pixel in a tile no occupied cell projects into is a miss whatever its ray does: skip the ray set-up as well
    const bool covered = !(tiles_x > 0 && Q.counters[Q_COVER_ALL] == 0u && Q.cover[(size_t)(y / COVER_TILE) * tiles_x + x / COVER_TILE] == 0);
    if (covered) {
      V3 org, dir; primary_ray(F, x, y, org, dir);
      enters = clip_ray(G, org, dir, 0.0001f, 100000.0f, seg);                     // :164-166 ray range
    }
    cur.worldPos[idx] = make_float4(0.f, 0.f, 0.f, 0.f);                           // miss until k_primary says otherwise
    Q.flag[idx] = 0;
    if (trace) {
      trace[idx * 4 + 0] = 0xFFFFFFFFu; trace[idx * 4 + 1] = 0u; trace[idx * 4 + 2] = 0u;
      trace[idx * 4 + 3] = pixel_seed((uint32_t)x, (uint32_t)y, F.clock, PASS_INITIAL);
    }
  }
  const unsigned b = __ballot_sync(0xffffffffu, enters);
  if (b) {
    const int lane = (threadIdx.y * 32 + threadIdx.x) & 31;
    uint32_t base = 0;
    if (lane == __ffs(b) - 1) base = atomicAdd(&Q.counters[Q_CAND], (uint32_t)__popc(b));
    base = __shfl_sync(0xffffffffu, base, __ffs(b) - 1);
    if (enters) {
      const uint32_t j = base + __popc(b & ((1u << lane) - 1u));
      Q.cand[j] = (uint32_t)idx;
      Q.cand_ray[2 * j] = make_float4(seg.o[0], seg.o[1], seg.o[2], seg.t0);
      Q.cand_ray[2 * j + 1] = make_float4(seg.d[0], seg.d[1], seg.d[2], seg.t1);
    }
  }
}

struct PrimaryJob {
  const FrameParams& F; Planes cur; Queues Q; uint32_t* trace; int store_y0;
  uint32_t idx;
  __device__ __forceinline__ bool fetch(const GridDev& G, uint32_t j, Ray<0>& ray, uint32_t& seed) {
    idx = Q.cand[j];
    const float4 a = Q.cand_ray[2 * j], b = Q.cand_ray[2 * j + 1];
    RaySeg seg; seg.o[0] = a.x; seg.o[1] = a.y; seg.o[2] = a.z; seg.t0 = a.w; seg.d[0] = b.x; seg.d[1] = b.y; seg.d[2] = b.z; seg.t1 = b.w;
    seed = pixel_seed(idx % F.W, idx / F.W + (uint32_t)store_y0, F.clock, PASS_INITIAL);   // :139-140
    ray.start(G, seg, seed);
    return true;
  }
  __device__ __forceinline__ void retire(const GridDev& G, const Ray<0>& ray, uint32_t seed) {
    if (ray.hit) {
      // scratch until k_ris: {t, cell of the directory (< 2^30), RNG state, voxel inside the cell (9 bits)} — a voxel index over the
      // whole window would not fit 32 bits for large sparse grids
      const uint32_t off = (uint32_t)(((ray.vox[0] & 7) << 6) | ((ray.vox[1] & 7) << 3) | (ray.vox[2] & 7));
      cur.worldPos[idx] = make_float4(ray.t, __uint_as_float((uint32_t)ray.cell), __uint_as_float(seed), __uint_as_float(off));
      Q.flag[idx] = 1;
    }
    if (trace) { trace[(size_t)idx * 4 + 1] = ray.ntent; trace[(size_t)idx * 4 + 2] = ray.ncells; trace[(size_t)idx * 4 + 3] = seed; }
  }
};

__global__ void __launch_bounds__(128, 9) k_primary(const GridDev G, const FrameParams* __restrict__ Fp, Planes cur, Queues Q,
                                                 uint32_t* __restrict__ trace, int store_y0, int refill) {
  const FrameParams& F = *Fp;
  PrimaryJob job{F, cur, Q, trace, store_y0, 0u};
  const uint32_t njobs = Q.counters[Q_CAND], nwarps = gridDim.x * (blockDim.x >> 5);
  march_loop<0>(G, job, &Q.counters[Q_PRIMARY_HEAD], njobs, refill & 0xff, (refill >> 8) & 0xff, lanes_for(njobs, nwarps, (uint32_t)(refill >> 16)));
}

// ---- ordered compaction of the hit flags (count per block, scan the block counts, scatter in pixel order)
__device__ __forceinline__ uint32_t flags8(const uint8_t* flag, size_t first, size_t n) {   // 8 consecutive flags as a bit mask
  uint32_t m = 0;
  if (first + 8 <= n) {
    const uint2 v = *reinterpret_cast<const uint2*>(flag + first);
#pragma unroll
    for (int k = 0; k < 4; ++k) { m |= ((v.x >> (8 * k)) & 0xffu) ? (1u << k) : 0u; m |= ((v.y >> (8 * k)) & 0xffu) ? (1u << (4 + k)) : 0u; }
  } else {
    for (int k = 0; k < 8; ++k) if (first + k < n && flag[first + k]) m |= 1u << k;
  }
  return m;
}
// One-kernel compaction: every 2048-pixel block keeps its hits contiguous and in pixel order (that is what the
// consumers' coalescing needs); the blocks' base offsets come from one atomicAdd each, so the order of the blocks in the
// list is arbitrary — results do not depend on it, every hit pixel is processed independently.
__global__ void __launch_bounds__(256) k_hit_compact(const uint8_t* __restrict__ flag, size_t npix, uint32_t* __restrict__ counters,
                                                     uint32_t* __restrict__ hit_pix) {
  __shared__ uint32_t s_warp[8];
  __shared__ uint32_t s_base;
  const size_t base = (size_t)blockIdx.x * COMPACT_BLOCK + (size_t)threadIdx.x * 8;
  const uint32_t m = base < npix ? flags8(flag, base, npix) : 0u;
  const uint32_t c = __popc(m);
  uint32_t incl = c;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t total = 0;
    for (int w = 0; w < 8; ++w) total += s_warp[w];
    s_base = total ? atomicAdd(&counters[Q_HIT], total) : 0u;
  }
  __syncthreads();
  uint32_t off = s_base + incl - c;
  for (int w = 0; w < warp; ++w) off += s_warp[w];
  uint32_t mm = m;
  while (mm) { const int k = __ffs(mm) - 1; mm &= mm - 1; hit_pix[off++] = (uint32_t)(base + k); }
}

// ---- A3, the parts every form of the RIS stage shares: what a hit pixel needs before its candidate loop, and what happens to
// the finished reservoir.
// ris_hit_setup: world position of the collision k_primary left in the pixel's worldPos slot as {t, voxel code, RNG state, 1},
// gradient normal (6 density lookups), voxel material, the G-buffer stores of restir.rgen:193-197, GeometryInfo (:174-186).
__device__ __forceinline__ void ris_hit_setup(const GridDev& G, const FrameParams& F, const Planes& cur, const Queues& Q, uint32_t s, int store_y0,
                                              uint32_t& idx, uint32_t& vcode, uint32_t& seed, GInfo& gi) {
  idx = Q.hit_pix[s];
  const int x = (int)(idx % F.W), y = (int)(idx / F.W) + store_y0;
  const float4 scratch = cur.worldPos[idx];
  seed = __float_as_uint(scratch.z);
  V3 org, dir; primary_ray(F, x, y, org, dir);
  const V3 P = add(org, muls(dir, scratch.x));
  const uint32_t cell = __float_as_uint(scratch.y), off = __float_as_uint(scratch.w);
  const int i = G.vmin[0] + 8 * (int)(cell % (uint32_t)G.cdim[0]) + (int)(off >> 6);
  const int j = G.vmin[1] + 8 * (int)((cell / (uint32_t)G.cdim[0]) % (uint32_t)G.cdim[1]) + (int)((off >> 3) & 7u);
  const int k = G.vmin[2] + 8 * (int)(cell / ((uint32_t)G.cdim[0] * (uint32_t)G.cdim[1])) + (int)(off & 7u);
  // (parity traces only; wraps above 2^32 voxels exactly like the oracle's uint32 expression)
  vcode = uint32_t(i - G.vmin[0]) + uint32_t(G.vdim[0]) * (uint32_t(j - G.vmin[1]) + uint32_t(G.vdim[1]) * uint32_t(k - G.vmin[2]));
  const float dens = density_at(G, i, j, k);
  V3 grad = v3(density_at(G, i + 1, j, k) - density_at(G, i - 1, j, k), density_at(G, i, j + 1, k) - density_at(G, i, j - 1, k),
               density_at(G, i, j, k + 1) - density_at(G, i, j, k - 1));
  const float gg = dot(grad, grad);
  V3 n;
  if (gg > 0.0f) { float l = sqrtf(gg); n = v3(-grad.x / l, -grad.y / l, -grad.z / l); }
  else n = v3(-dir.x, -dir.y, -dir.z);
  const float4 al = voxel_albedo(dens);
  cur.worldPos[idx] = make_float4(P.x, P.y, P.z, 1.0f); cur.albedo[idx] = al;                          // :193-197
  cur.normal[idx] = make_float4(n.x, n.y, n.z, 1.0f); cur.mat[idx] = make_float4(G.roughness, G.metallic, 1.0f, 1.0f);
  gi.albedo[0] = al.x; gi.albedo[1] = al.y; gi.albedo[2] = al.z; gi.albedo[3] = al.w;
  gi.normal = n; gi.worldPos = P; gi.metallic = G.metallic; gi.roughness = G.roughness;
  gi.albedoLum = luminance_common(al.x, al.y, al.z);                                                   // :182
  gi.camPos = v3(F.camPos[0], F.camPos[1], F.camPos[2]);                                               // :183
  gi.sampleSeed = 0;
}
// ris_hit_store: optional FINALIZE_W, pack (:286-289), the state k_shadow / k_finish continue from, and the shadow ray of
// ratio_track() toward the selected light (:229-235), clipped and queued for k_shadow.
__device__ __forceinline__ void ris_hit_store(const GridDev& G, const LightsDev& L, const FrameParams& F, const ResPlanes& outR, const Queues& Q,
                                              uint32_t* __restrict__ trace, int needs_finish, uint32_t s, uint32_t idx, uint32_t vcode, uint32_t seed,
                                              V3 P, Res res) {
  if ((F.flags & FLAG_FINALIZE_W) != 0 && res.w > 0.0f) res.w = res.sumWeights / (float(res.M) * res.pHat);
  float4 a, b; packReservoir(res, a, b);
  outR.info[idx] = a; outR.weight[idx] = b;
  Q.hit_seed[s] = seed; Q.hit_T[s] = 1.0f;
  if (trace) { trace[(size_t)idx * 4 + 0] = vcode; if (!needs_finish) trace[(size_t)idx * 4 + 3] = seed; }
  if ((F.flags & FLAG_VISIBILITY) != 0 && res.w > 0.0f) {
    const float4 lp = __ldg(&L.lights[2 * res.lightIndex]);
    V3 sd = sub(v3(lp.x, lp.y, lp.z), P);
    const float dist = sqrtf(dot(sd, sd));
    RaySeg seg;
    bool march = dist > 0.0f;
    if (march) { sd = divs(sd, dist); march = clip_ray(G, P, sd, 0.0f, dist, seg); }
    if (march) {                          // otherwise T = 1 and the RNG state is untouched
      const uint32_t q = warp_append(&Q.counters[Q_SHADOW]);
      Q.shadow[q] = s;
      Q.shadow_ray[2 * q] = make_float4(seg.o[0], seg.o[1], seg.o[2], seg.t0);
      Q.shadow_ray[2 * q + 1] = make_float4(seg.d[0], seg.d[1], seg.d[2], seg.t1);
    }
  }
}

// Thread form of the RIS stage: one thread per hit pixel runs the M-candidate loop serially.  The launcher uses it for light
// tables that stay L1-resident (<= 64 KB) on large launches; k_ris_coop below takes the others and produces the same bits.
__global__ void __launch_bounds__(128, 8) k_ris_thread(const GridDev G, const LightsDev L, const FrameParams* __restrict__ Fp, Planes cur, ResPlanes outR,
                                             Queues Q, uint32_t* __restrict__ trace, int store_y0, int needs_finish, uint32_t min_hits) {
  const FrameParams& F = *Fp;
  const uint32_t nhit = Q.counters[Q_HIT];
  if (nhit < min_hits) return;                               // small launches are left to k_ris_coop (launch_initial)
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < nhit; s += gridDim.x * blockDim.x) {
    uint32_t idx, vcode, seed; GInfo gi;
    ris_hit_setup(G, F, cur, Q, s, store_y0, idx, vcode, seed, gi);
    const V3 P = gi.worldPos;
    Res res = newReservoir();
    if (dot(gi.normal, gi.normal) != 0.0f) {                                                           // :205
      const ShadePre pre = shade_pre(gi);
      uint32_t selM = 0u; float selSumW = 0.0f;
      for (uint32_t c = 0; c < F.M; ++c) {                                                             // :206-226
        const uint32_t sampleSeed = seed;                                                              // :213
        float r1 = rnd(seed), r2 = rnd(seed);                                                          // :116, GLSL left-to-right
        uint32_t sel; float pdf;
        aliasTableSample(L, r1, r2, sel, pdf);
        // addSampleToReservoir + updateReservoir (reservoir.glsl:45-54, 30-43).  The w the reference forms for every
        // candidate, (sumW + weight) / (M * pHat), only survives for the selected one: remember its sumW and M and do
        // that division once after the loop (same operands, same result).
        const float pHat = evaluatePHat(L, sel, gi, pre);
        res.M += 1;
        const float u = rnd(seed);
        // pHat = 0 (light behind the surface) with a positive pdf: weight = 0 / pdf = 0, the sum does not move and nothing can be
        // selected — same bits as the divisions, which would take div.rn's slow path for the zero numerator
        if (!(pHat == 0.0f && pdf > 0.0f)) {
          const float weight = pHat / pdf;
          res.sumWeights += weight;
          const float replacePossibility = weight / res.sumWeights;
          if (u < replacePossibility) {
            res.lightIndex = sel; res.lightKind = 0; res.pHat = pHat; res.sampleSeed = sampleSeed;
            selM = res.M; selSumW = res.sumWeights;
          }
        }
      }
      if (selM != 0u) res.w = selSumW / (float(selM) * res.pHat);                                      // reservoir.glsl:51
    }
    ris_hit_store(G, L, F, outR, Q, trace, needs_finish, s, idx, vcode, seed, P, res);
  }
}


// ---- A3, cooperative form.  The candidates of restir.rgen:206-226 are independent except for the running sum of the
// streaming reservoir, so a warp takes a group of up to 32 hit pixels and turns the work 90 degrees twice:
//   step 1  lane = pixel      gradient normal, G-buffer stores, per-pixel shading terms -> shared memory
//   phase A lane = candidate  (one pixel at a time) LCG jump to draw 3c, alias sample, light fetch, the
//                             dot(wi, n) < 0 early-out of restirUtils.glsl:42-44; survivors are appended to a work list
//   phase B lane = list entry the expensive remainder of evaluatePHat on 32 surviving (pixel, candidate) pairs of
//                             ANY pixels of the group: no lane idles behind a culled candidate
//   step 3  lane = pixel      the serial part of updateReservoir (sum, weight / sum, compare with the third draw)
// Every candidate's arithmetic is the reference expression on the same values, and the running sum is added in
// candidate order, so the reservoir is bit-identical to the serial loop; w of the selected candidate is
// (sumW at selection) / (float(c + 1) * pHat), exactly what reservoir.glsl:51 left in it.
// The group size shrinks when a launch has few hits (multi-GPU bands), which shortens the serial chain per warp.
struct RisSmem {
  float w[32 * 33];                                  // candidate weights [pixel][candidate of the chunk], padded
  uint32_t l_pc[64]; float l_pdf[64], l_lp[3][64], l_lew[64];   // phase-B work list (ring): (pixel, candidate), pdf, light position, luminance
  float P[3][32], n[3][32], wo[3][32], fresnelOut[32], smithOut[32], albedoLum[32];
  uint32_t seed0[32];                                // RNG state at the first candidate of the current chunk
};

__global__ void __launch_bounds__(128, 6) k_ris_coop(const GridDev G, const LightsDev L, const FrameParams* __restrict__ Fp, Planes cur, ResPlanes outR,
                                                  Queues Q, uint32_t* __restrict__ trace, int store_y0, int needs_finish, int target_warps, uint32_t max_hits) {
  if (Q.counters[Q_HIT] >= max_hits) return;                 // large launches with L1-resident light tables go to k_ris_thread
  __shared__ RisSmem sm_all[4];
  RisSmem& sm = sm_all[threadIdx.x >> 5];
  const FrameParams& F = *Fp;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const uint32_t nhit = Q.counters[Q_HIT];
  const uint32_t nwarps = gridDim.x * 4u, warp = blockIdx.x * 4u + (threadIdx.x >> 5);
  const uint32_t gsz = group_size_for(nhit, nwarps, (uint32_t)target_warps, 32u);
  const uint32_t M = F.M;
  const float pre_a = gmax(0.001f, G.roughness * G.roughness);                     // disneyBRDF.glsl:53 (roughness is per grid)

  for (uint32_t g0 = warp * gsz; g0 < nhit; g0 += nwarps * gsz) {
    const uint32_t npix = nhit - g0 < gsz ? nhit - g0 : gsz;
    const uint32_t s = g0 + (uint32_t)lane;
    const bool pix = (uint32_t)lane < npix;
    // ---------------- step 1: lane = pixel
    uint32_t idx = 0, vcode = 0, seed = 0;
    bool valid = false;
    V3 Ppix = v3(0.f, 0.f, 0.f), Npix = v3(0.f, 0.f, 0.f);
    if (pix) {
      GInfo gi;
      ris_hit_setup(G, F, cur, Q, s, store_y0, idx, vcode, seed, gi);
      const V3 P = gi.worldPos, n = gi.normal;
      Ppix = P; Npix = n;
      valid = dot(gi.normal, gi.normal) != 0.0f;                                                       // :205
      const ShadePre pre = shade_pre(gi);
      sm.P[0][lane] = P.x; sm.P[1][lane] = P.y; sm.P[2][lane] = P.z;
      sm.n[0][lane] = n.x; sm.n[1][lane] = n.y; sm.n[2][lane] = n.z;
      sm.wo[0][lane] = pre.wo.x; sm.wo[1][lane] = pre.wo.y; sm.wo[2][lane] = pre.wo.z;
      sm.fresnelOut[lane] = pre.fresnelOut; sm.smithOut[lane] = pre.smithOut;
      sm.albedoLum[lane] = gi.albedoLum;
      sm.seed0[lane] = seed;
    }
    __syncwarp();

    float sumW = 0.0f, sel_sumW = 0.0f;
    uint32_t selc = 0xFFFFFFFFu, sel_seed = 0u;
    for (uint32_t c0 = 0; c0 < M; c0 += 32u) {
      const uint32_t cn = M - c0 < 32u ? M - c0 : 32u;
      uint32_t head = 0u, tail = 0u;
      uint32_t zero_mask = 0u;                         // candidates of this lane's pixel whose weight is exactly 0 (light behind the surface)
      // Phase A, lane = pixel: every lane walks the candidates of its own pixel (own RNG stream, own P and n in registers) as a
      // three-stage software pipeline, so that the two dependent table fetches of a candidate (alias cell -> light) are in flight
      // while the lane works on the candidates before it:
      //   stage 1 draws r1, r2 of candidate `it` and requests its alias cell       (consumed one iteration later)
      //   stage 2 picks the light of candidate `it - 1` and requests it            (consumed one iteration later)
      //   stage 3 culls candidate `it - 2` when the light is behind the surface, else appends it to the phase-B list
      uint32_t sA = seed;                              // runs ahead of `seed`, which step 3 advances
      uint32_t colA = 0u; float r2A = 0.0f, pdfL = 0.0f, lewL = 0.0f;
      float4 cellA = make_float4(0.f, 0.f, 0.f, 0.f), lpL = make_float4(0.f, 0.f, 0.f, 0.f);
      for (uint32_t it = 0; it < cn + 2u; ++it) {
        uint32_t selN = 0u; float pdfN = 0.0f, lewN = 0.0f; float4 lpN = lpL;
        if (it >= 1u && it <= cn) {                                                                    // stage 2
          aliasPick(cellA, colA, r2A, selN, pdfN);
          lpN = __ldg(&L.lights[2 * selN]);
          lewN = __ldg(reinterpret_cast<const float*>(L.lights) + 8 * (size_t)selN + 7);
        }
        uint32_t colN = 0u; float r2N = 0.0f; float4 cellN = cellA;
        if (it < cn) {                                                                                 // stage 1
          const float r1 = rnd(sA); r2N = rnd(sA);                                                     // :116, GLSL left-to-right
          lcg(sA);                                                                                     // the selection draw of updateReservoir
          colN = aliasColumn(L, r1);
          cellN = __ldg(&L.alias[colN]);
        }
        if (it >= 2u) {                                                                                // stage 3
          const uint32_t cc = it - 2u;
          const V3 wi = sub(v3(lpL.x, lpL.y, lpL.z), Ppix);                                            // restirUtils.glsl:41-44
          const bool back = dot(wi, Npix) < 0.0f;
          if (valid && back) {                                                                         // pHat = 0 -> weight = 0 / pdf
            if (pdfL > 0.0f) zero_mask |= 1u << cc;
            else sm.w[lane * 33 + (int)cc] = 0.0f / pdfL;
          }
          const bool keep = valid && !back;
          const unsigned m = __ballot_sync(full, keep);
          if (keep) {
            const uint32_t e = (tail + (uint32_t)__popc(m & lt_mask)) & 63u;
            sm.l_pc[e] = ((uint32_t)lane << 8) | cc; sm.l_pdf[e] = pdfL;
            sm.l_lp[0][e] = lpL.x; sm.l_lp[1][e] = lpL.y; sm.l_lp[2][e] = lpL.z; sm.l_lew[e] = lewL;
          }
          tail += (uint32_t)__popc(m);
        }
        const bool flush = it == cn + 1u;
        __syncwarp();
        while (tail - head >= 32u || (flush && tail != head)) {
          // ---------------- phase B: lane = work-list entry (no global loads: the light travels in the list)
          const uint32_t cnt = tail - head < 32u ? tail - head : 32u;
          if ((uint32_t)lane < cnt) {
            const uint32_t e = (head + (uint32_t)lane) & 63u;
            const uint32_t pc = sm.l_pc[e];
            const float pdf = sm.l_pdf[e];
            const int p = (int)(pc >> 8), c = (int)(pc & 255u);
            GInfo g;
            g.worldPos = v3(sm.P[0][p], sm.P[1][p], sm.P[2][p]);
            g.normal = v3(sm.n[0][p], sm.n[1][p], sm.n[2][p]);
            g.albedoLum = sm.albedoLum[p]; g.roughness = G.roughness; g.metallic = G.metallic;
            ShadePre pre;
            pre.wo = v3(sm.wo[0][p], sm.wo[1][p], sm.wo[2][p]);
            pre.fresnelOut = sm.fresnelOut[p]; pre.smithOut = sm.smithOut[p]; pre.a = pre_a; pre.cosOut = 0.0f;
            const float pHat = evaluatePHatLight(v3(sm.l_lp[0][e], sm.l_lp[1][e], sm.l_lp[2][e]), sm.l_lew[e], g, pre);
            sm.w[p * 33 + c] = pHat / pdf;                                                             // reservoir.glsl:47
          }
          head += cnt;
          __syncwarp();
        }
        pdfL = pdfN; lpL = lpN; lewL = lewN;
        colA = colN; r2A = r2N; cellA = cellN;
      }
      // ---------------- step 3: lane = pixel, the serial part of updateReservoir (reservoir.glsl:30-43) for this chunk
      if (valid) {
        for (uint32_t cc = 0; cc < cn; ++cc) {
          const uint32_t sb = seed;                                                                    // gi.sampleSeed (:213)
          lcg(seed); lcg(seed);
          const float u = rnd(seed);
          // A zero weight (three quarters of the candidates on the bunny: light behind the surface) changes nothing: sumW + 0 = sumW
          // and rnd < 0 / sumW is false (also for 0 / 0 = NaN); only the draw is consumed.  Skipping the division there matters:
          // a zero numerator takes div.rn's slow path (ncu: 15 % of this kernel's instructions were that subroutine).
          if ((zero_mask >> cc) & 1u) continue;
          const float wt = sm.w[lane * 33 + (int)cc];
          if (wt != 0.0f) {
            sumW += wt;
            const float replacePossibility = wt / sumW;
            if (u < replacePossibility) { selc = c0 + cc; sel_seed = sb; sel_sumW = sumW; }
          }
        }
      }
      __syncwarp();
    }

    // ---------------- lane = pixel: the selected candidate, visibility set-up, stores
    if (pix) {
      const V3 P = v3(sm.P[0][lane], sm.P[1][lane], sm.P[2][lane]);
      Res res = newReservoir();
      if (valid) {
        res.M = M; res.sumWeights = sumW;
        if (selc != 0xFFFFFFFFu) {
          uint32_t st = sel_seed;
          const float r1 = rnd(st), r2 = rnd(st);
          uint32_t sel; float pdf;
          aliasTableSample(L, r1, r2, sel, pdf);
          GInfo g;
          g.worldPos = P; g.normal = v3(sm.n[0][lane], sm.n[1][lane], sm.n[2][lane]);
          g.albedoLum = sm.albedoLum[lane]; g.roughness = G.roughness; g.metallic = G.metallic;
          ShadePre pre;
          pre.wo = v3(sm.wo[0][lane], sm.wo[1][lane], sm.wo[2][lane]);
          pre.fresnelOut = sm.fresnelOut[lane]; pre.smithOut = sm.smithOut[lane]; pre.a = pre_a; pre.cosOut = 0.0f;
          const float pHat = evaluatePHat(L, sel, g, pre);
          res.lightIndex = sel; res.lightKind = 0; res.pHat = pHat; res.sampleSeed = sel_seed;
          res.w = sel_sumW / (float(selc + 1u) * pHat);                                                // reservoir.glsl:51
        }
      }
      ris_hit_store(G, L, F, outR, Q, trace, needs_finish, s, idx, vcode, seed, P, res);
    }
    __syncwarp();
  }
}

struct ShadowJob {
  Queues Q;
  uint32_t s;
  __device__ __forceinline__ bool fetch(const GridDev& G, uint32_t j, Ray<1>& ray, uint32_t& seed) {
    s = Q.shadow[j];
    seed = Q.hit_seed[s];
    const float4 a = Q.shadow_ray[2 * j], b = Q.shadow_ray[2 * j + 1];
    RaySeg seg; seg.o[0] = a.x; seg.o[1] = a.y; seg.o[2] = a.z; seg.t0 = a.w; seg.d[0] = b.x; seg.d[1] = b.y; seg.d[2] = b.z; seg.t1 = b.w;
    ray.start(G, seg, seed);
    return true;
  }
  __device__ __forceinline__ void retire(const GridDev&, const Ray<1>& ray, uint32_t seed) { Q.hit_T[s] = ray.T; Q.hit_seed[s] = seed; }
};

__global__ void __launch_bounds__(128, 9) k_shadow(const GridDev G, Queues Q, int refill) {
  ShadowJob job{Q, 0u};
  const uint32_t njobs = Q.counters[Q_SHADOW], nwarps = gridDim.x * (blockDim.x >> 5);
  march_loop<1>(G, job, &Q.counters[Q_SHADOW_HEAD], njobs, refill & 0xff, (refill >> 8) & 0xff, lanes_for(njobs, nwarps, (uint32_t)(refill >> 16)));
}

// A/B forms (VRS_MARCH=s): one thread per ray, plain nested loops, no warp-level scheduling
__global__ void __launch_bounds__(128, 8) k_shadow_simple(const GridDev G, Queues Q) {
  const uint32_t njobs = Q.counters[Q_SHADOW];
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < njobs; j += gridDim.x * blockDim.x) {
    ShadowJob job{Q, 0u};
    Ray<1> ray; uint32_t seed;
    job.fetch(G, j, ray, seed);
    int st = RAY_SKIP;
    while (st != RAY_DONE) st = st == RAY_SKIP ? ray.cell_step(G) : ray.collide_step(G, seed);
    job.retire(G, ray, seed);
  }
}
__global__ void __launch_bounds__(128, 8) k_primary_simple(const GridDev G, const FrameParams* __restrict__ Fp, Planes cur, Queues Q,
                                                           uint32_t* __restrict__ trace, int store_y0) {
  const FrameParams& F = *Fp;
  const uint32_t njobs = Q.counters[Q_CAND];
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < njobs; j += gridDim.x * blockDim.x) {
    PrimaryJob job{F, cur, Q, trace, store_y0, 0u};
    Ray<0> ray; uint32_t seed;
    job.fetch(G, j, ray, seed);
    int st = RAY_SKIP;
    while (st != RAY_DONE) st = st == RAY_SKIP ? ray.cell_step(G) : ray.collide_step(G, seed);
    job.retire(G, ray, seed);
  }
}

__global__ void __launch_bounds__(128, 8) k_finish(const LightsDev L, const FrameParams* __restrict__ Fp, Planes cur, Planes prev, ResPlanes prevR,
                                                ResPlanes outR, Queues Q, uint32_t* __restrict__ trace, int store_y0, const PrevAccess PA,
                                                unsigned* __restrict__ out_of_halo) {
  const FrameParams& F = *Fp;
  const uint32_t nhit = Q.counters[Q_HIT];
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < nhit; s += gridDim.x * blockDim.x) {
    const uint32_t idx = Q.hit_pix[s];
    uint32_t seed = Q.hit_seed[s];
    Res res = unpackReservoir(outR.info[idx], outR.weight[idx]);
    GInfo gi = ginfo_from_planes(cur, idx, F);
    if ((F.flags & FLAG_VISIBILITY) != 0 && res.w > 0.0f) {                                            // :229-235 -> transmittance
      const float T = Q.hit_T[s];
      res.w = res.w * T;
      res.sumWeights = res.sumWeights * T;
    }
    if ((F.flags & FLAG_TEMPORAL) != 0) {                                                              // :237-284 (dormant block, intent)
      float q[4];
      mat_vec(F.prevVP, gi.worldPos.x, gi.worldPos.y, gi.worldPos.z, 1.0f, q);
      q[0] = q[0] / q[3]; q[1] = q[1] / q[3]; q[2] = q[2] / q[3];
      q[0] = (q[0] + 1.0f) * 0.5f * float(F.W);
      q[1] = (q[1] + 1.0f) * 0.5f * float(F.H);
      if (q[0] > 0.0f && q[1] > 0.0f && q[0] < float(F.W) && q[1] < float(F.H)) {
        int fx = int(q[0]), fy = int(q[1]);
        {   // how far the reprojection moved vertically (launchers size bands / halos from it: vrs_get_counters)
          const int dyr = abs(fy - ((int)(idx / F.W) + store_y0));
          const unsigned am = __activemask();
          const int mx = __reduce_max_sync(am, dyr);
          if ((threadIdx.x & 31) == __ffs(am) - 1) atomicMax(out_of_halo + 1, (unsigned)mx);
        }
        // Where the previous frame's pixel lives: in this context's own rows, or (several GPUs, peer memory) in the band of the
        // neighbour above / below, read in place over NVLink — only the few pixels whose reprojection crosses a band edge pay
        // that latency, and no halo rows have to be shipped for the temporal pass.
        const Planes* pp = &prev; const ResPlanes* prp = &prevR;
        int row0 = store_y0;
        bool have = fy >= PA.own_y0 && fy < PA.own_y1;
        if (!have && fy < PA.own_y0 && PA.up.worldPos && fy >= PA.up_y0) { pp = &PA.up; prp = &PA.upR; row0 = PA.up_row0; have = true; }
        if (!have && fy >= PA.own_y1 && PA.down.worldPos && fy < PA.down_y1) { pp = &PA.down; prp = &PA.downR; row0 = PA.down_row0; have = true; }
        if (!have) atomicAdd(out_of_halo, 1u);                                                         // beyond the rows anybody can supply (vrs_get_counters)
        else {
          const Planes& prev = *pp; const ResPlanes& prevR = *prp;
          size_t pidx = (size_t)(fy - row0) * F.W + (size_t)fx;
          if (!(prev.worldPos[pidx].w < 0.5f)) {                 // a previous miss holds no data (its zero normal fails :274 anyway)
            GInfo pg = ginfo_from_planes(prev, pidx, F);                                        // prevGInfo.camPos = gInfo.camPos (:259)
            V3 pd = sub(gi.worldPos, pg.worldPos);
            if (dot(pd, pd) < 0.01f) {
              V3 ad = v3(gi.albedo[0] - pg.albedo[0], gi.albedo[1] - pg.albedo[1], gi.albedo[2] - pg.albedo[2]);
              if (dot(ad, ad) < 0.01f) {
                if (dot(gi.normal, pg.normal) > 0.5f) {
                  Res pr = unpackReservoir(prevR.info[pidx], prevR.weight[pidx]);                      // at prevFrag
                  uint32_t cap = uint32_t(F.temporalMult) * res.M;
                  if (cap < pr.M) pr.M = cap;
                  if (pr.lightIndex < (uint32_t)L.nlights) combineReservoirsGeom(L, res, pr, gi, pg, seed);   // (a stale index can only come from a foreign history)
                }
              }
            }
          }
        }
      }
    }
    float4 a, b; packReservoir(res, a, b);                                                             // :286-289
    outR.info[idx] = a; outR.weight[idx] = b;
    if (trace) trace[(size_t)idx * 4 + 3] = seed;
  }
}

// -------------------------------------------------------------------------------------------------
// Spatial reuse — spatialReuse.comp main (:54-89) completed with the k-neighbour loop built on combineReservoirs
// (reservoir.glsl:56-76), normalisation deferred to the finally selected sample.  One thread per hit pixel (the
// `exist < 0.5` early-out of :76-79 is the hit list); neighbour G-buffer / reservoir reads are gathers through L2.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 8) k_spatial_thread(const LightsDev L, const FrameParams* __restrict__ Fp, Planes cur, ResPlanes inR, ResPlanes outR,
                                                           Queues Q, uint32_t iteration, int store_y0, int store_y1) {
  const FrameParams& F = *Fp;
  const uint32_t nhit = Q.counters[Q_HIT];
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < nhit; s += gridDim.x * blockDim.x) {
    const uint32_t idx = Q.hit_pix[s];
    const int x = (int)(idx % F.W), y = (int)(idx / F.W) + store_y0;
    uint32_t seed = pixel_seed((uint32_t)x, (uint32_t)y, F.clock, PASS_SPATIAL0 + iteration);   // :58-59
    Res res = unpackReservoir(inR.info[idx], inR.weight[idx]);
    GInfo gi = ginfo_from_planes(cur, idx, F);
    uint32_t Z = res.M;
    const float radius = F.spatialRadius;
    uint32_t k = F.spatialNeighbors; if (k > (uint32_t)MAX_NEIGHBORS) k = MAX_NEIGHBORS;
    uint32_t nb_idx[MAX_NEIGHBORS]; uint32_t nb_M[MAX_NEIGHBORS]; int nacc = 0;
    const ShadePre pre = shade_pre(gi);
    uint32_t sO = seed;                                                           // the 2k offset draws come first:
    for (uint32_t i = 0; i < 2u * k; ++i) lcg(seed);                              // `seed` continues behind them with the selection draws
    for (uint32_t i = 0; i < k; ++i) {
      float r1 = rnd(sO), r2 = rnd(sO);
      float dx = (r1 * 2.0f - 1.0f) * radius, dy = (r2 * 2.0f - 1.0f) * radius;
      int ox = int(dx), oy = int(dy);
      int nx = x + ox, ny = y + oy;
      const bool inside = !(dx * dx + dy * dy > radius * radius) && !(ox == 0 && oy == 0) && nx >= 0 && ny >= 0 && nx < (int)F.W && ny < (int)F.H &&
                          ny >= store_y0 && ny < store_y1;                          // (rows outside the ones held: halo too small)
      // all planes of the neighbour are requested at once (from this pixel's own slot when the offset is rejected outright)
      size_t nidx = inside ? (size_t)(ny - store_y0) * F.W + (size_t)nx : (size_t)idx;
      const float4 nwp = cur.worldPos[nidx], na = cur.albedo[nidx], nn = cur.normal[nidx], ri = inR.info[nidx], rw = inR.weight[nidx];
      if (!inside) continue;
      if (nwp.w < 0.5f) continue;
      V3 pd = sub(gi.worldPos, v3(nwp.x, nwp.y, nwp.z));
      if (!(dot(pd, pd) < 0.01f)) continue;
      V3 ad = v3(gi.albedo[0] - na.x, gi.albedo[1] - na.y, gi.albedo[2] - na.z);
      if (!(dot(ad, ad) < 0.01f)) continue;
      if (!(dot(gi.normal, v3(nn.x, nn.y, nn.z)) > 0.5f)) continue;
      Res nr = unpackReservoir(ri, rw);
      res.M += nr.M;                                                             // reservoir.glsl:61-68
      float pHat = evaluatePHat(L, nr.lightIndex, gi, pre);
      float weight = pHat * nr.w * float(nr.M);
      if (weight > 0.0f) updateReservoir(res, nr.lightIndex, nr.lightKind, weight, pHat, nr.w, seed, nr.sampleSeed);
      nb_idx[nacc] = (uint32_t)nidx; nb_M[nacc] = nr.M; ++nacc;
    }
    if (nacc > 0) {
      for (int j = 0; j < nacc; ++j) {                                           // reservoir.glsl:70-73
        GInfo ng = ginfo_from_planes(cur, nb_idx[j], F);
        float pHat = evaluatePHat(L, res.lightIndex, ng);
        if (pHat > 0.0f) Z += nb_M[j];
      }
      if (res.w > 0.0f) res.w = res.sumWeights / (float(Z) * res.pHat);          // :74-75
    }
    float4 a, b; packReservoir(res, a, b);
    outR.info[idx] = a; outR.weight[idx] = b;
  }
}


// -------------------------------------------------------------------------------------------------
// Final shade — restir_post.frag main (:57-105): shade, emissive override, firefly clamp, running mean.
// Every pixel; a miss pixel costs its worldPos read (16 B) and the accumulation update only.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 5) k_shade(const GridDev G, const LightsDev L, const FrameParams* __restrict__ Fp, Planes cur, ResPlanes rs,
                                               float4* __restrict__ accum, int y0, int y1, int store_y0) {
  const FrameParams& F = *Fp;
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = y0 + blockIdx.y * 8 + threadIdx.y;
  if (x >= (int)F.W || y >= y1) return;
  const size_t idx = (size_t)(y - store_y0) * F.W + x;
  const float4 wp = cur.worldPos[idx];
  V3 c;
  if (wp.w < 0.5f) {
    c = v3(F.clear[0], F.clear[1], F.clear[2]);
  } else {
    GInfo gi = ginfo_from_planes(cur, idx, F);
    Res res = unpackReservoir(rs.info[idx], rs.weight[idx]);
    gi.sampleSeed = res.sampleSeed;
    V3 pHat = evaluatePHatFull(L, res.lightIndex, gi);
    c = add(v3(0.0f, 0.0f, 0.0f), muls(pHat, res.w));                            // :80-81
    if ((F.flags & FLAG_FINAL_VISIBILITY) != 0 && res.w > 0.0f) {
      uint32_t seed = pixel_seed((uint32_t)x, (uint32_t)y, F.clock, PASS_SHADE);
      float4 lp = __ldg(&L.lights[2 * res.lightIndex]);
      float T = ratio_track(G, gi.worldPos, v3(lp.x, lp.y, lp.z), seed);
      c = muls(c, T);
    }
    if (gi.albedo[3] > 0.5f) c = v3(gi.albedo[0], gi.albedo[1], gi.albedo[2]);   // :82-84
    float lum = luminance_utils(c);                                              // :86-90
    if (lum > F.fireflyClamp) c = muls(c, F.fireflyClamp / lum);
    c = v3(0.0f < c.x ? c.x : 0.0f, 0.0f < c.y ? c.y : 0.0f, 0.0f < c.z ? c.z : 0.0f);   // :92
  }
  float4 out;
  if (F.frame < 1 || F.initialize == 1) {                                        // :94-102
    out = make_float4(c.x, c.y, c.z, 1.0f);
  } else {
    float4 old = accum[idx];
    float w = 1.0f / float(F.frame);
    out = make_float4(gmix(old.x, c.x, w), gmix(old.y, c.y, w), gmix(old.z, c.z, w), 1.0f);
  }
  accum[idx] = out;
}

// -------------------------------------------------------------------------------------------------
// Peer-memory halo exchange: boundary rows of up to 6 planes are stored straight into the neighbours' halo rows over
// NVLink (P2P stores), then the last block publishes the exchange serial with a system-scope release; the consumer side
// is k_halo_wait (one thread acquiring the flag the neighbour wrote).  No NCCL, capturable in the frame's CUDA graph.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_halo_push(HaloPush H) {
  const size_t stride = (size_t)gridDim.x * blockDim.x, tid0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = H.up_count + H.down_count;
  for (size_t i = tid0; i < total; i += stride) {           // one pixel per thread: consecutive lanes, consecutive 16-byte stores
    const bool up = i < H.up_count;
    const size_t k = up ? i : i - H.up_count;
    const size_t so = (up ? H.up_src_off : H.down_src_off) + k, dof = (up ? H.up_dst_off : H.down_dst_off) + k;
    const float4 wp = H.wp_src[so];
    if (H.copy_wp) (up ? H.wp_up : H.wp_down)[dof] = wp;
    if (!(wp.w < 0.5f)) {
#pragma unroll
      for (int p = 0; p < 4; ++p)
        if (p < H.nplanes) (up ? H.up_dst[p] : H.down_dst[p])[dof] = H.src[p][so];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicAdd(H.block_counter, 1u);
    if (ticket == gridDim.x - 1) {                       // every block's stores are fenced: publish
      *H.block_counter = 0u;
      const unsigned serial = *H.serial + 1u;
      *H.serial = serial;
      __threadfence_system();
      if (H.up_flag) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(H.up_flag), "r"(serial) : "memory");
      if (H.down_flag) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(H.down_flag), "r"(serial) : "memory");
    }
  }
}
__global__ void k_halo_wait(const unsigned* serial, const unsigned* flag_from_up, const unsigned* flag_from_down, unsigned* error) {
  const unsigned want = *serial;
  unsigned long long t0, now;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int side = 0; side < 2; ++side) {
    const unsigned* f = side == 0 ? flag_from_up : flag_from_down;
    if (!f) continue;
    unsigned v = 0;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if ((int)(v - want) >= 0) break;
      __nanosleep(200);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now - t0 > 4000000000ull) { *error = 1u; break; }   // 4 s: give up instead of hanging the GPU (vrs_synchronize reports VRS_ERR_COMM)
    }
  }
}

// Readback helper: planes in the reference layout for EVERY pixel (miss pixels hold stale data on the device; what the
// reference's images would contain there is the cleared G-buffer of restir.rgen:150-156,193-197 and an empty reservoir).
__global__ void __launch_bounds__(256) k_export(Planes cur, ResPlanes rs, float4* __restrict__ out6, size_t first_pix, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t idx = first_pix + i;
  const float4 wp = cur.worldPos[idx];
  const bool hit = !(wp.w < 0.5f);
  float4 a, b; packReservoir(newReservoir(), a, b);
  out6[0 * n + i] = hit ? wp : make_float4(0.f, 0.f, 0.f, 0.f);
  out6[1 * n + i] = hit ? cur.albedo[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
  out6[2 * n + i] = hit ? cur.normal[idx] : make_float4(0.f, 0.f, 0.f, 1.f);
  out6[3 * n + i] = hit ? cur.mat[idx] : make_float4(0.f, 0.f, 1.f, 1.f);
  out6[4 * n + i] = hit ? rs.info[idx] : a;
  out6[5 * n + i] = hit ? rs.weight[idx] : b;
}

// restir_post.frag:104: outColor = pow(outColor, vec3(1.0f / 0.8f)) -> 8-bit RGBA for the headless "swapchain"
__global__ void __launch_bounds__(256) k_display(const float4* __restrict__ accum, uchar4* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 c = accum[i];
  float r = powf(fmaxf(c.x, 0.0f), 1.0f / 0.8f), g = powf(fmaxf(c.y, 0.0f), 1.0f / 0.8f), b = powf(fmaxf(c.z, 0.0f), 1.0f / 0.8f);
  out[i] = make_uchar4((unsigned char)(fminf(r, 1.0f) * 255.0f + 0.5f), (unsigned char)(fminf(g, 1.0f) * 255.0f + 0.5f),
                       (unsigned char)(fminf(b, 1.0f) * 255.0f + 0.5f), 255);
}

__global__ void k_sample_density(const GridDev G, const int* __restrict__ ijk, uint32_t n, float* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = density_at(G, ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2]);
}

// Grid of a grid-stride kernel: exactly one resident wave (a second, partial wave of blocks would run at reduced occupancy
// for as long as the first).  VRS_WAVE_SCALE multiplies it (for measurements).
template <class K>
static int resident_grid(K kernel, int block_threads, int fallback_per_sm) {
  int d = 0, sms = 148, per_sm = 0;
  cudaGetDevice(&d); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block_threads, 0) != cudaSuccess || per_sm < 1) per_sm = fallback_per_sm;
  const int scale = getenv("VRS_WAVE_SCALE") ? atoi(getenv("VRS_WAVE_SCALE")) : 1;
  const int pct = getenv("VRS_GRID_PCT") ? atoi(getenv("VRS_GRID_PCT")) : 100;      // (measurements: leave room for the kernels of other stages)
  int g = sms * per_sm * (scale > 0 ? scale : 1);
  if (pct > 0 && pct < 100) g = g * pct / 100 < sms ? sms : g * pct / 100;
  return g;
}

// ------------------------------------------------------------------------------------------------- launchers
// `F` is the host copy (launch geometry, which kernels run); `dF` is the same struct in device memory, read by the
// kernels — so that a captured CUDA graph of the frame stays valid while the per-frame values change.
// Front of the initial pass (everything that does not depend on the previous frame), in two stages the host runtime may run
// on different streams and overlap with other frames (frames in flight):
//   stage A launch_front_trace  coverage mask, classification, primary volume event, hit list   -> work queues Q, hit scratch in cur.worldPos
//   stage B launch_front_ris    RIS candidates (G-buffer + tmp reservoir outR), shadow-ray queue -> cur, outR, Q.shadow
static int march_refill() {
  static const int sms = [] { int d = 0, n = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n; }();
  static const int target_warps = sms * (getenv("VRS_MIN_WARPS_PER_SM") ? atoi(getenv("VRS_MIN_WARPS_PER_SM")) : MIN_WARPS_PER_SM);
  // tuning knobs of the persistent raymarch kernels: idle lanes that trigger a refill | cell visits per scheduling decision << 8
  // | warps that a small launch is spread over (lanes_for) << 16
  static const int refill = (getenv("VRS_REFILL") ? atoi(getenv("VRS_REFILL")) : REFILL_MIN_IDLE) |
                            ((getenv("VRS_CELLS") ? atoi(getenv("VRS_CELLS")) : CELLS_PER_DECISION) << 8) | (target_warps << 16);
  return refill;
}
static bool simple_march() { static const bool v = getenv("VRS_MARCH") && getenv("VRS_MARCH")[0] == 's'; return v; }
static int march_waves() { static const int v = getenv("VRS_MARCH_WAVES") ? atoi(getenv("VRS_MARCH_WAVES")) : 1; return v; }

void launch_front_trace(cudaStream_t st, const GridDev& G, const FrameParams& F, const FrameParams* dF, Planes cur, const Queues& Q, uint32_t* trace,
                        int y0, int y1, int store_y0, int store_y1, int persistent_blocks, KTimer* kt) {
  const int refill = march_refill();
  cudaMemsetAsync(Q.counters, 0, 8 * sizeof(uint32_t), st);
  // screen-space coverage culling (k_cover); the RNG / cell trace of a culled ray would differ, so tracing turns it off
  static const bool no_cull = getenv("VRS_NO_CULL") != nullptr;
  int tiles_x = 0;
  const long long ncell = (long long)G.cdim[0] * G.cdim[1] * G.cdim[2];
  if (F.cull && !trace && !no_cull && Q.cover && ncell <= (1ll << 27)) {          // (8 threads per cell in a 32-bit grid index)
    tiles_x = ((int)F.W + COVER_TILE - 1) / COVER_TILE;
    const int band_ty0 = y0 / COVER_TILE, band_ty1 = (y1 - 1) / COVER_TILE;
    cudaMemsetAsync(Q.cover + (size_t)band_ty0 * tiles_x, 0, (size_t)tiles_x * (band_ty1 - band_ty0 + 1), st);
    const int tiles_y = ((int)F.H + COVER_TILE - 1) / COVER_TILE;
    k_cover<<<(unsigned)((ncell * 8 + 127) / 128), 128, 0, st>>>(G, dF, Q, tiles_x, tiles_y, band_ty0, band_ty1);
    ktick(kt, st, "k_cover");
  }
  dim3 block(32, 8), grid((F.W + 31) / 32, (y1 - y0 + 7) / 8);
  k_classify<<<grid, block, 0, st>>>(G, dF, cur, Q, trace, y0, y1, store_y0, tiles_x);
  ktick(kt, st, "k_classify");
  static const int g_ps = resident_grid(k_primary_simple, 128, 8) * march_waves();
  if (simple_march()) k_primary_simple<<<g_ps, 128, 0, st>>>(G, dF, cur, Q, trace, store_y0);
  else k_primary<<<persistent_blocks, 128, 0, st>>>(G, dF, cur, Q, trace, store_y0, refill);
  ktick(kt, st, "k_primary");
  // compaction runs over every stored row (8-byte aligned flag loads); flags outside the band rows stay 0
  const size_t npix = (size_t)(store_y1 - store_y0) * F.W;
  const uint32_t nblocks = (uint32_t)((npix + COMPACT_BLOCK - 1) / COMPACT_BLOCK);
  k_hit_compact<<<nblocks, 256, 0, st>>>(Q.flag, npix, Q.counters, Q.hit_pix);
  ktick(kt, st, "k_hit_compact");
}

void launch_front_ris(cudaStream_t st, const GridDev& G, const LightsDev& L, const FrameParams& F, const FrameParams* dF, Planes cur,
                      ResPlanes outR, const Queues& Q, uint32_t* trace, int store_y0, int persistent_blocks, KTimer* kt) {
  const int refill = march_refill();
  const int target_warps = refill >> 16;
  static const int sms = [] { int d = 0, n = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n; }();
  static const int ris_blocks = [] {   // one resident wave of the cooperative RIS kernel
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ris_coop, 128, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    if (getenv("VRS_RIS_BLOCKS_PER_SM")) per_sm = atoi(getenv("VRS_RIS_BLOCKS_PER_SM"));
    return sms * per_sm;
  }();
  // RIS stage: t(hread) | c(oop) | a(uto).  Measured on B200 (profiles/r01_summary.md): with light tables that stay
  // L1-resident the plain per-thread loop is fastest; once they spill to L2 (thousands of lights) the cooperative kernel,
  // which keeps the dependent table fetches of several pixels in flight, wins.
  static const char ris_env = getenv("VRS_RIS") ? getenv("VRS_RIS")[0] : 'a';
  static const uint32_t small_launch = getenv("VRS_RIS_SMALL") ? (uint32_t)atoi(getenv("VRS_RIS_SMALL")) : RIS_SMALL_LAUNCH;
  const bool big_tables = (size_t)L.nlights * 32 + (size_t)L.ntable * 16 > (size_t)64 * 1024;
  const bool vis = (F.flags & FLAG_VISIBILITY) != 0, temporal = (F.flags & FLAG_TEMPORAL) != 0;
  const int needs_finish = (vis || temporal) ? 1 : 0;
  static const int g_thread = resident_grid(k_ris_thread, 128, 8);
  if (ris_env == 't') k_ris_thread<<<g_thread, 128, 0, st>>>(G, L, dF, cur, outR, Q, trace, store_y0, needs_finish, 0u);
  else if (ris_env == 'c' || big_tables) k_ris_coop<<<ris_blocks, 128, 0, st>>>(G, L, dF, cur, outR, Q, trace, store_y0, needs_finish, target_warps, 0xFFFFFFFFu);
  else {   // auto, small tables: the hit count (known only on the device) picks the form; the other launch returns at once
    k_ris_coop<<<ris_blocks, 128, 0, st>>>(G, L, dF, cur, outR, Q, trace, store_y0, needs_finish, target_warps, small_launch);
    k_ris_thread<<<g_thread, 128, 0, st>>>(G, L, dF, cur, outR, Q, trace, store_y0, needs_finish, small_launch);
  }
  ktick(kt, st, "k_ris");
}
//   stage C launch_front_shadow  ratio-tracking transmittance toward the selected light          -> Q.hit_T / hit_seed
void launch_front_shadow(cudaStream_t st, const GridDev& G, const FrameParams& F, const Queues& Q, int persistent_blocks, KTimer* kt) {
  if ((F.flags & FLAG_VISIBILITY) == 0) return;
  static const int g_ss = resident_grid(k_shadow_simple, 128, 8) * march_waves();
  if (simple_march()) k_shadow_simple<<<g_ss, 128, 0, st>>>(G, Q);
  else k_shadow<<<persistent_blocks, 128, 0, st>>>(G, Q, march_refill());
  ktick(kt, st, "k_shadow");
}
// Back half of the initial pass: apply the shadow transmittance, temporal merge with the previous frame's G-buffer /
// final reservoirs (restir.rgen:229-284), final pack.  No-op when neither visibility nor temporal reuse is on.
void launch_initial_finish(cudaStream_t st, const LightsDev& L, const FrameParams& F, const FrameParams* dF, Planes cur, Planes prev, ResPlanes prevR,
                           ResPlanes outR, const Queues& Q, uint32_t* trace, int store_y0, const PrevAccess& PA, unsigned* out_of_halo, KTimer* kt) {
  if ((F.flags & (FLAG_VISIBILITY | FLAG_TEMPORAL)) == 0) return;
  static const int g_finish = resident_grid(k_finish, 128, 8);
  k_finish<<<g_finish, 128, 0, st>>>(L, dF, cur, prev, prevR, outR, Q, trace, store_y0, PA, out_of_halo);
  ktick(kt, st, "k_finish");
}
int front_trace_launches(bool culling) { return 3 /* classify, primary, compact */ + (culling ? 1 : 0); }
int front_ris_launches(int flags, const LightsDev& L) {
  const bool vis = (flags & FLAG_VISIBILITY) != 0;
  // auto mode with small light tables launches both RIS forms (the device-side hit count decides which one works)
  const char ris_env = getenv("VRS_RIS") ? getenv("VRS_RIS")[0] : 'a';
  const bool big_tables = (size_t)L.nlights * 32 + (size_t)L.ntable * 16 > (size_t)64 * 1024;
  const int ris = (ris_env == 'a' && !big_tables) ? 2 : 1;
  (void)vis;
  return ris;
}
int front_shadow_launches(int flags) { return (flags & FLAG_VISIBILITY) != 0 ? 1 : 0; }
void launch_spatial(cudaStream_t s, const LightsDev& L, const FrameParams* dF, Planes cur, ResPlanes inR, ResPlanes outR, const Queues& Q,
                    uint32_t iteration, int store_y0, int store_y1, KTimer* kt) {
  static const int g8 = resident_grid(k_spatial_thread, 128, 8);
  k_spatial_thread<<<g8, 128, 0, s>>>(L, dF, cur, inR, outR, Q, iteration, store_y0, store_y1);
  ktick(kt, s, "k_spatial");
}
void launch_shade(cudaStream_t s, const GridDev& G, const LightsDev& L, const FrameParams& F, const FrameParams* dF, Planes cur, ResPlanes rs,
                  float4* accum, int y0, int y1, int store_y0, KTimer* kt) {
  dim3 block(32, 8), grid((F.W + 31) / 32, (y1 - y0 + 7) / 8);
  k_shade<<<grid, block, 0, s>>>(G, L, dF, cur, rs, accum, y0, y1, store_y0);
  ktick(kt, s, "k_shade");
}
void launch_halo_push(cudaStream_t s, const HaloPush& H, int blocks, KTimer* kt) { k_halo_push<<<blocks, 256, 0, s>>>(H); ktick(kt, s, "k_halo_push"); }
void launch_halo_wait(cudaStream_t s, const unsigned* serial, const unsigned* from_up, const unsigned* from_down, unsigned* error, KTimer* kt) {
  k_halo_wait<<<1, 1, 0, s>>>(serial, from_up, from_down, error);
  ktick(kt, s, "k_halo_wait");
}
void launch_export(cudaStream_t s, Planes cur, ResPlanes rs, float4* out6, size_t first_pix, size_t n) {
  k_export<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(cur, rs, out6, first_pix, n);
}
void launch_display(cudaStream_t s, const float4* accum, uchar4* out, size_t n) {
  k_display<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(accum, out, n);
}
void launch_sample_density(cudaStream_t s, const GridDev& G, const int* ijk, uint32_t n, float* out) {
  k_sample_density<<<(n + 255) / 256, 256, 0, s>>>(G, ijk, n, out);
}

}  // namespace vrs
