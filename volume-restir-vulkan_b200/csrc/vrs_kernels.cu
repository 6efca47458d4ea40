// Hot-path kernels (sm_100a): initial (primary volume event + RIS + visibility + temporal), spatial reuse,
// final shade.  No tensor cores / RT cores: nothing here is a dense contraction (BASELINE.json north_star).
// Compile with -fmad=false (see vrs_device.cuh).
#include "vrs_device.cuh"
#include "vrs_kernels.h"

namespace vrs {

static constexpr uint32_t PASS_INITIAL = 0, PASS_SPATIAL0 = 1, PASS_SHADE = 5;
static constexpr int FLAG_VISIBILITY = 1 << 0, FLAG_TEMPORAL = 1 << 1;
static constexpr int FLAG_FINAL_VISIBILITY = 1 << 4, FLAG_FINALIZE_W = 1 << 5;
static constexpr int MAX_NEIGHBORS = 16;

// -------------------------------------------------------------------------------------------------
// Kernel A — restir.rgen main (:136-290) on a volume.
// Phase 1: one thread per pixel of a 16x16 tile: primary ray (:142-148), delta-tracking raymarch through the
//          sparse grid, G-buffer stores (:193-197).  Miss pixels store an empty reservoir and retire.
// Phase 2: the surviving (hit) pixels are compacted with warp ballots + a block scan into shared memory, so
//          that the M-candidate RIS loop, the shadow transmittance raymarch and the temporal merge run on
//          densely packed warps instead of on scattered lanes.
// -------------------------------------------------------------------------------------------------
struct HitRec {
  float P[3]; float n[3]; float albedo[4];
  uint32_t seed; uint32_t pix;      // pix = ly * 16 + lx inside the tile
};

template <bool TRACE>
__global__ void __launch_bounds__(256) k_initial(const GridDev G, const LightsDev L, const FrameParams F, Planes cur, Planes prev,
                                                 ResPlanes prevR, ResPlanes outR, uint32_t* __restrict__ trace, int y0, int y1,
                                                 int store_y0, int store_y1) {
  __shared__ HitRec s_hits[256];
  __shared__ int s_warp_count[8];
  __shared__ int s_total;

  const int tid = threadIdx.y * 16 + threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int x = blockIdx.x * 16 + threadIdx.x;
  const int y = y0 + blockIdx.y * 16 + threadIdx.y;
  const bool inside = x < (int)F.W && y < y1;

  bool hit = false;
  HitRec rec;
  uint32_t seed = 0;
  if (inside) {
    const size_t idx = (size_t)(y - store_y0) * F.W + x;
    seed = pixel_seed((uint32_t)x, (uint32_t)y, F.clock, PASS_INITIAL);          // :139-140
    // primary ray, :142-148
    float ux = float(x) / float(F.W), uy = float(y) / float(F.H);
    float dx = ux * 2.0f - 1.0f, dy = uy * 2.0f - 1.0f;
    float o4[4], t4[4], d4[4];
    mat_vec(F.viewInverse, 0.0f, 0.0f, 0.0f, 1.0f, o4);
    mat_vec(F.projInverse, dx, dy, 1.0f, 1.0f, t4);
    V3 tn = normalize(v3(t4[0], t4[1], t4[2]));
    mat_vec(F.viewInverse, tn.x, tn.y, tn.z, 0.0f, d4);
    V3 org = v3(o4[0], o4[1], o4[2]), dir = v3(d4[0], d4[1], d4[2]);

    TrackResult r = track<0>(G, org, dir, 0.0001f, 100000.0f, seed);            // :164-166 ray range
    float4 wp = make_float4(0.f, 0.f, 0.f, 0.f), al = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 nr = make_float4(0.f, 0.f, 0.f, 1.f), mt = make_float4(0.f, 0.f, 1.f, 1.f);
    uint32_t vcode = 0xFFFFFFFFu;
    if (r.hit) {
      hit = true;
      V3 P = add(org, muls(dir, r.t));
      int i = r.vox[0], j = r.vox[1], k = r.vox[2];
      vcode = uint32_t(i - G.vmin[0]) + uint32_t(G.vdim[0]) * (uint32_t(j - G.vmin[1]) + uint32_t(G.vdim[1]) * uint32_t(k - G.vmin[2]));
      float dens = density_at(G, i, j, k);
      V3 grad = v3(density_at(G, i + 1, j, k) - density_at(G, i - 1, j, k), density_at(G, i, j + 1, k) - density_at(G, i, j - 1, k),
                   density_at(G, i, j, k + 1) - density_at(G, i, j, k - 1));
      float gg = dot(grad, grad);
      V3 n;
      if (gg > 0.0f) { float l = sqrtf(gg); n = v3(-grad.x / l, -grad.y / l, -grad.z / l); }
      else n = v3(-dir.x, -dir.y, -dir.z);
      wp = make_float4(P.x, P.y, P.z, 1.0f);
      al = voxel_albedo(dens);
      nr = make_float4(n.x, n.y, n.z, 1.0f);
      mt = make_float4(G.roughness, G.metallic, 1.0f, 1.0f);
      rec.P[0] = P.x; rec.P[1] = P.y; rec.P[2] = P.z;
      rec.n[0] = n.x; rec.n[1] = n.y; rec.n[2] = n.z;
      rec.albedo[0] = al.x; rec.albedo[1] = al.y; rec.albedo[2] = al.z; rec.albedo[3] = al.w;
      rec.seed = seed; rec.pix = (uint32_t)tid;
    }
    cur.worldPos[idx] = wp; cur.albedo[idx] = al; cur.normal[idx] = nr; cur.mat[idx] = mt;   // :193-197
    if (TRACE) { trace[idx * 4 + 0] = vcode; trace[idx * 4 + 1] = r.ntent; trace[idx * 4 + 2] = r.ncells; }
    if (!hit) {                                                                   // miss: empty reservoir
      float4 a, b; packReservoir(newReservoir(), a, b);
      outR.info[idx] = a; outR.weight[idx] = b;
      if (TRACE) trace[idx * 4 + 3] = seed;
    }
  }

  // ---- block compaction of hit pixels (warp ballot + popc prefix, block scan over 8 warps)
  const unsigned ballot = __ballot_sync(0xffffffffu, hit);
  const int warp_prefix = __popc(ballot & ((1u << lane) - 1u));
  if (lane == 0) s_warp_count[warp] = __popc(ballot);
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { int c = s_warp_count[w]; s_warp_count[w] = acc; acc += c; }
    s_total = acc;
  }
  __syncthreads();
  if (hit) s_hits[s_warp_count[warp] + warp_prefix] = rec;
  __syncthreads();
  const int nhit = s_total;
  if (tid >= nhit) return;

  // ---- phase 2: RIS + visibility + temporal on packed lanes
  const HitRec h = s_hits[tid];
  const int px = blockIdx.x * 16 + (h.pix & 15), py = y0 + blockIdx.y * 16 + (h.pix >> 4);
  const size_t idx = (size_t)(py - store_y0) * F.W + px;
  seed = h.seed;
  GInfo gi;
  gi.albedo[0] = h.albedo[0]; gi.albedo[1] = h.albedo[1]; gi.albedo[2] = h.albedo[2]; gi.albedo[3] = h.albedo[3];
  gi.normal = v3(h.n[0], h.n[1], h.n[2]);
  gi.worldPos = v3(h.P[0], h.P[1], h.P[2]);
  gi.metallic = G.metallic; gi.roughness = G.roughness;
  gi.albedoLum = luminance_common(gi.albedo[0], gi.albedo[1], gi.albedo[2]);     // :182
  gi.camPos = v3(F.camPos[0], F.camPos[1], F.camPos[2]);                         // :183
  gi.sampleSeed = 0;
  Res res = newReservoir();
  if (dot(gi.normal, gi.normal) != 0.0f) {                                       // :205
    for (uint32_t i = 0; i < F.M; ++i) {                                         // :206-226
      gi.sampleSeed = seed;                                                      // :213
      float r1 = rnd(seed), r2 = rnd(seed);                                      // :116, GLSL left-to-right
      uint32_t sel; float pdf;
      aliasTableSample(L, r1, r2, sel, pdf);
      addSampleToReservoir(L, res, sel, 0, pdf, gi, seed);                       // :224-225
    }
  }
  if ((F.flags & FLAG_FINALIZE_W) != 0 && res.w > 0.0f) res.w = res.sumWeights / (float(res.M) * res.pHat);
  if ((F.flags & FLAG_VISIBILITY) != 0 && res.w > 0.0f) {                        // :229-235 -> transmittance
    float4 lp = __ldg(&L.lights[2 * res.lightIndex]);
    float T = ratio_track(G, gi.worldPos, v3(lp.x, lp.y, lp.z), seed);
    res.w = res.w * T;
    res.sumWeights = res.sumWeights * T;
  }
  if ((F.flags & FLAG_TEMPORAL) != 0) {                                          // :237-284 (dormant block, intent)
    float q[4];
    mat_vec(F.prevVP, gi.worldPos.x, gi.worldPos.y, gi.worldPos.z, 1.0f, q);
    q[0] = q[0] / q[3]; q[1] = q[1] / q[3]; q[2] = q[2] / q[3];
    q[0] = (q[0] + 1.0f) * 0.5f * float(F.W);
    q[1] = (q[1] + 1.0f) * 0.5f * float(F.H);
    if (q[0] > 0.0f && q[1] > 0.0f && q[0] < float(F.W) && q[1] < float(F.H)) {
      int fx = int(q[0]), fy = int(q[1]);
      if (fy >= store_y0 && fy < store_y1) {                                     // rows held by this context
        size_t pidx = (size_t)(fy - store_y0) * F.W + (size_t)fx;
        GInfo pg = ginfo_from_planes(prev, pidx, F.camPos);                      // prevGInfo.camPos = gInfo.camPos (:259)
        V3 pd = sub(gi.worldPos, pg.worldPos);
        if (dot(pd, pd) < 0.01f) {
          V3 ad = v3(gi.albedo[0] - pg.albedo[0], gi.albedo[1] - pg.albedo[1], gi.albedo[2] - pg.albedo[2]);
          if (dot(ad, ad) < 0.01f) {
            if (dot(gi.normal, pg.normal) > 0.5f) {
              Res pr = unpackReservoir(prevR.info[pidx], prevR.weight[pidx]);    // at prevFrag
              uint32_t cap = uint32_t(F.temporalMult) * res.M;
              if (cap < pr.M) pr.M = cap;
              combineReservoirsGeom(L, res, pr, gi, pg, seed);
            }
          }
        }
      }
    }
  }
  float4 a, b; packReservoir(res, a, b);                                         // :286-289
  outR.info[idx] = a; outR.weight[idx] = b;
  if (TRACE) trace[idx * 4 + 3] = seed;
}

// -------------------------------------------------------------------------------------------------
// Kernel B — spatial reuse: spatialReuse.comp main (:54-89) completed with the k-neighbour loop built on
// combineReservoirs (reservoir.glsl:56-76), normalisation deferred to the finally selected sample.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_spatial(const LightsDev L, const FrameParams F, Planes cur, ResPlanes inR, ResPlanes outR,
                                                 uint32_t iteration, int y0, int y1, int store_y0, int store_y1) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = y0 + blockIdx.y * 8 + threadIdx.y;
  if (x >= (int)F.W || y >= y1) return;
  const size_t idx = (size_t)(y - store_y0) * F.W + x;
  uint32_t seed = pixel_seed((uint32_t)x, (uint32_t)y, F.clock, PASS_SPATIAL0 + iteration);   // :58-59
  float4 ri = inR.info[idx], rw = inR.weight[idx];
  float exist = cur.worldPos[idx].w;
  if (!(exist < 0.5f)) {                                                         // :76-79
    Res res = unpackReservoir(ri, rw);
    GInfo gi = ginfo_from_planes(cur, idx, F.camPos);
    uint32_t Z = res.M;
    const float radius = F.spatialRadius;
    uint32_t k = F.spatialNeighbors; if (k > (uint32_t)MAX_NEIGHBORS) k = MAX_NEIGHBORS;
    uint32_t nb_idx[MAX_NEIGHBORS]; uint32_t nb_M[MAX_NEIGHBORS]; int nacc = 0;
    for (uint32_t i = 0; i < k; ++i) {
      float r1 = rnd(seed), r2 = rnd(seed);
      float dx = (r1 * 2.0f - 1.0f) * radius, dy = (r2 * 2.0f - 1.0f) * radius;
      if (dx * dx + dy * dy > radius * radius) continue;
      int ox = int(dx), oy = int(dy);
      if (ox == 0 && oy == 0) continue;
      int nx = x + ox, ny = y + oy;
      if (nx < 0 || ny < 0 || nx >= (int)F.W || ny >= (int)F.H) continue;
      if (ny < store_y0 || ny >= store_y1) continue;                             // outside the rows held (halo too small)
      size_t nidx = (size_t)(ny - store_y0) * F.W + (size_t)nx;
      if (cur.worldPos[nidx].w < 0.5f) continue;
      GInfo ng = ginfo_from_planes(cur, nidx, F.camPos);
      V3 pd = sub(gi.worldPos, ng.worldPos);
      if (!(dot(pd, pd) < 0.01f)) continue;
      V3 ad = v3(gi.albedo[0] - ng.albedo[0], gi.albedo[1] - ng.albedo[1], gi.albedo[2] - ng.albedo[2]);
      if (!(dot(ad, ad) < 0.01f)) continue;
      if (!(dot(gi.normal, ng.normal) > 0.5f)) continue;
      Res nr = unpackReservoir(inR.info[nidx], inR.weight[nidx]);
      res.M += nr.M;                                                             // reservoir.glsl:61-68
      float pHat = evaluatePHat(L, nr.lightIndex, gi);
      float weight = pHat * nr.w * float(nr.M);
      if (weight > 0.0f) updateReservoir(res, nr.lightIndex, nr.lightKind, weight, pHat, nr.w, seed, nr.sampleSeed);
      nb_idx[nacc] = (uint32_t)nidx; nb_M[nacc] = nr.M; ++nacc;
    }
    if (nacc > 0) {
      for (int j = 0; j < nacc; ++j) {                                           // reservoir.glsl:70-73
        GInfo ng = ginfo_from_planes(cur, nb_idx[j], F.camPos);
        float pHat = evaluatePHat(L, res.lightIndex, ng);
        if (pHat > 0.0f) Z += nb_M[j];
      }
      if (res.w > 0.0f) res.w = res.sumWeights / (float(Z) * res.pHat);          // :74-75
    }
    packReservoir(res, ri, rw);
  }
  outR.info[idx] = ri; outR.weight[idx] = rw;
}

// -------------------------------------------------------------------------------------------------
// Kernel C — restir_post.frag main (:57-105): shade, emissive override, firefly clamp, running mean.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_shade(const GridDev G, const LightsDev L, const FrameParams F, Planes cur, ResPlanes rs,
                                               float4* __restrict__ accum, int y0, int y1, int store_y0) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = y0 + blockIdx.y * 8 + threadIdx.y;
  if (x >= (int)F.W || y >= y1) return;
  const size_t idx = (size_t)(y - store_y0) * F.W + x;
  GInfo gi = ginfo_from_planes(cur, idx, F.camPos);
  Res res = unpackReservoir(rs.info[idx], rs.weight[idx]);
  gi.sampleSeed = res.sampleSeed;
  float exist = cur.worldPos[idx].w;
  V3 c;
  if (exist < 0.5f) {
    c = v3(F.clear[0], F.clear[1], F.clear[2]);
  } else {
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

__global__ void k_sample_density(const GridDev G, const int* __restrict__ ijk, uint32_t n, float* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = density_at(G, ijk[3 * i], ijk[3 * i + 1], ijk[3 * i + 2]);
}

// ------------------------------------------------------------------------------------------------- launchers
void launch_initial(cudaStream_t s, const GridDev& G, const LightsDev& L, const FrameParams& F, Planes cur, Planes prev, ResPlanes prevR,
                    ResPlanes outR, uint32_t* trace, int y0, int y1, int store_y0, int store_y1) {
  dim3 block(16, 16), grid((F.W + 15) / 16, (y1 - y0 + 15) / 16);
  if (trace) k_initial<true><<<grid, block, 0, s>>>(G, L, F, cur, prev, prevR, outR, trace, y0, y1, store_y0, store_y1);
  else k_initial<false><<<grid, block, 0, s>>>(G, L, F, cur, prev, prevR, outR, nullptr, y0, y1, store_y0, store_y1);
}
void launch_spatial(cudaStream_t s, const LightsDev& L, const FrameParams& F, Planes cur, ResPlanes inR, ResPlanes outR, uint32_t iteration,
                    int y0, int y1, int store_y0, int store_y1) {
  dim3 block(32, 8), grid((F.W + 31) / 32, (y1 - y0 + 7) / 8);
  k_spatial<<<grid, block, 0, s>>>(L, F, cur, inR, outR, iteration, y0, y1, store_y0, store_y1);
}
void launch_shade(cudaStream_t s, const GridDev& G, const LightsDev& L, const FrameParams& F, Planes cur, ResPlanes rs, float4* accum,
                  int y0, int y1, int store_y0) {
  dim3 block(32, 8), grid((F.W + 31) / 32, (y1 - y0 + 7) / 8);
  k_shade<<<grid, block, 0, s>>>(G, L, F, cur, rs, accum, y0, y1, store_y0);
}
void launch_sample_density(cudaStream_t s, const GridDev& G, const int* ijk, uint32_t n, float* out) {
  k_sample_density<<<(n + 255) / 256, 256, 0, s>>>(G, ijk, n, out);
}

}  // namespace vrs
