// pool_kernel.cuh — sample_poolkernel: the sample job as a warp-local wavefront.
//
// Replaces SampleBatchJob.Execute / Sample (Runtime/Jobs/SampleBatchJob.cs:58-401) like
// sample_megakernel does, with the same per-path arithmetic (kernel_common.cuh) and the same
// order-independent fixed-point sums, so the two kernels produce bit-identical buffers.  What
// changes is how a warp's 32 lanes are kept busy.  In the megakernel a lane owns one path from
// camera to sky: lanes that finish their BVH walk early idle until the slowest lane of the warp is
// done, and the shade / sky+refill branches run one after the other with part of the warp each.
// Here a warp owns a POOL of kPool (> 32) paths in shared memory and alternates three phases, each
// of which runs one kind of work with (nearly) all lanes:
//
//   shade    every pooled path whose ray hit a sphere, 32 at a time: hit record, Material.Scatter,
//            throughput, next ray (+ its reciprocal direction and the root box test)
//   finish   every pooled path whose ray left the world (or ran out of bounces), 32 at a time: sky,
//            per-pixel accumulation by warp reduction, then the slot is refilled with the tile's next
//            pixel-sample (camera ray)
//   walk     the closest-hit BVH walk over all queued rays with DYNAMIC FETCH: a lane that finishes
//            its ray takes the next one from the queue, so lanes stay busy until the queue is empty
//
// Everything is warp-synchronous (no block barrier after the world is staged): a warp claims tiles of
// <= 16 pixels from a global counter exactly like the megakernel.
#pragma once

#include "sample_kernels.cuh"

namespace rtbk {

#ifndef RTB_POOL_WARPS
#define RTB_POOL_WARPS 16
#endif
#ifndef RTB_POOL_SLOTS
#define RTB_POOL_SLOTS 64
#endif
constexpr int kPoolWarps = RTB_POOL_WARPS;
constexpr int kPoolBlock = kPoolWarps * 32;
constexpr int kPool = RTB_POOL_SLOTS;                 // paths per warp (<= 255: slot ids travel as bytes)
constexpr int kPoolRounds = (kPool + 31) / 32;
static_assert(kPool >= 32 && kPool <= 255, "pool size");

constexpr uint32_t kMetaLive = 1u << 8;               // slot holds a path
constexpr uint32_t kMetaNonSpecular = 1u << 9;        // first non-specular hit seen (SampleBatchJob.cs:313)
constexpr uint32_t kMetaFailed = 1u << 10;            // depth == TraceDepth (SampleBatchJob.cs:379-381)
// meta = pixel slot in the tile (bits 0-7) | flags | depth << 16

struct WarpPool {
  // ray: written by shade / finish, read by walk
  float ox[kPool], oy[kPool], oz[kPool], dx[kPool], dy[kPool], dz[kPool], ix[kPool], iy[kPool], iz[kPool], aa[kPool];
  // nearest hit: written by walk (or by the producer of a ray that misses the root box)
  float t[kPool];
  int idx[kPool];
  // path
  float thx[kPool], thy[kPool], thz[kPool];          // throughput
  float rx[kPool], ry[kPool], rz[kPool];             // radiance
  float nx[kPool], ny[kPool], nz[kPool];             // sampleNormal
  float ax[kPool], ay[kPool], az[kPool];             // sampleAlbedo
  float events[kPool];
  uint32_t pixel[kPool], sample[kPool], meta[kPool];
  unsigned char walk_q[kPool + 1], hit_q[kPool + 1], done_q[kPool + 1];
  // tile
  unsigned long long acc[kTilePixelsMax][kAccValues];  // 2^-32 fixed point
  uint32_t successes[kTilePixelsMax], rays[kTilePixelsMax], node_tests[kTilePixelsMax], sphere_tests[kTilePixelsMax];
  uint32_t non_finite[kTilePixelsMax];
  float fallback[kTilePixelsMax][6];
  uint32_t prefix[kTilePixelsMax + 1];
};

__host__ __device__ inline size_t pool_smem_bytes(uint32_t blob_bytes, bool scene_in_smem) {
  size_t s = 16;
  if (scene_in_smem) s += blob_bytes;
  s = (s + 15) & ~(size_t)15;
  return s + sizeof(WarpPool) * kPoolWarps;
}

// Sum of a 64-bit two's-complement value over the warp (all 32 lanes call; every lane gets the sum).
__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long q) {
  const uint32_t lo = (uint32_t)q, hi = (uint32_t)(q >> 32);
  const uint32_t s0 = __reduce_add_sync(0xffffffffu, lo & 0xffffu);
  const uint32_t s1 = __reduce_add_sync(0xffffffffu, lo >> 16);
  const uint32_t s2 = __reduce_add_sync(0xffffffffu, hi);
  return ((unsigned long long)s2 << 32) + ((unsigned long long)s1 << 16) + (unsigned long long)s0;
}

// Ray bookkeeping shared by the two producers of rays (shade, finish): reciprocal direction with the
// NaN fix (SampleBatchJob.cs:409-412), dot(d, d), and the root box test (HitTests.cs:9-21) of the next walk.
__device__ __forceinline__ bool store_ray(WarpPool& w, int s, const SceneDesc& sd, f3 o, f3 d) {
  f3 inv = um::rcp(d);
  inv = um::mk(um::isnan(inv.x) ? um::INF : inv.x, um::isnan(inv.y) ? um::INF : inv.y, um::isnan(inv.z) ? um::INF : inv.z);
  w.ox[s] = o.x; w.oy[s] = o.y; w.oz[s] = o.z;
  w.dx[s] = d.x; w.dy[s] = d.y; w.dz[s] = d.z;
  w.ix[s] = inv.x; w.iy[s] = inv.y; w.iz[s] = inv.z;
  w.aa[s] = um::dot(d, d);
  float t_enter;
  const bool root = sd.has_root && aabb_hit(v3(sd.root_min), v3(sd.root_max), o, inv, &t_enter);
  if (!root) { w.t[s] = um::INF; w.idx[s] = -1; }
  return root;
}

template <bool SMEM, bool COUNTERS>
__global__ void __launch_bounds__(kPoolBlock, 1) sample_poolkernel(const __grid_constant__ BatchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  unsigned char* blob_smem = smem + 16;
  const size_t pools_off = (16 + (SMEM ? a.scene.blob_bytes : 0) + 15) & ~(size_t)15;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpPool& w = reinterpret_cast<WarpPool*>(smem + pools_off)[warp];

  SceneView<SMEM> sv;
  if (SMEM) {
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      mbar_fence_init();
      mbar_expect_tx(bar, a.scene.blob_bytes);
      constexpr uint32_t kChunk = 32768;
      for (uint32_t off = 0; off < a.scene.blob_bytes; off += kChunk) {
        uint32_t n = a.scene.blob_bytes - off < kChunk ? a.scene.blob_bytes - off : kChunk;
        tma_bulk_g2s(blob_smem + off, a.scene.blob + off, n, bar);
      }
    }
    __syncthreads();
    mbar_wait(bar, 0);
    sv.bind(blob_smem, a.scene);
  } else {
    sv.bind(a.scene.blob, a.scene);
  }

  const rtb_batch_params& p = a.p;
  const SceneDesc& sd = a.scene;
  const uint32_t FULL = 0xffffffffu;
  const uint32_t lt_mask = (1u << lane) - 1u;

  for (int s = lane; s < kPool; s += 32) w.meta[s] = 0;
  uint32_t tile_base = 0, next_item = 0, total_items = 0;   // warp-uniform
  int tile_n = 0;
  int n_live = 0;
  __syncwarp();

  for (;;) {
    // ---- no path in flight and the tile's samples are used up: retire the tile, claim the next ----
    if (n_live == 0 && next_item >= total_items) {
      if (tile_n > 0 && lane < tile_n) {
        int cx, cy;
        uint32_t index;
        active_pixel(a, tile_base + (uint32_t)lane, &cx, &cy, &index);
        const float4 in_color = reinterpret_cast<const float4*>(a.b.in_color)[index];
        const float in_weight = a.b.in_sample_count_weight[index];
        const int sample_count = (int)in_color.w + (int)w.successes[lane];
        const bool bad = w.non_finite[lane] != 0;
        const float nan = um::asfloat(0x7fc00000u);
        float v[kAccValues];
#pragma unroll
        for (int k = 0; k < kAccValues; k++) v[k] = bad ? nan : __ll2float_rn((long long)w.acc[lane][k]) * kFixedInvScale;
        reinterpret_cast<float4*>(a.b.out_color)[index] =
            make_float4(in_color.x + v[0], in_color.y + v[1], in_color.z + v[2], (float)sample_count);
        const float* in_n = a.b.in_normal + 3 * (size_t)index;
        const float* in_a = a.b.in_albedo + 3 * (size_t)index;
        float* on = a.b.out_normal + 3 * (size_t)index;
        float* oa = a.b.out_albedo + 3 * (size_t)index;
        if (sample_count == 0) {
          on[0] = w.fallback[lane][0]; on[1] = w.fallback[lane][1]; on[2] = w.fallback[lane][2];
          oa[0] = w.fallback[lane][3]; oa[1] = w.fallback[lane][4]; oa[2] = w.fallback[lane][5];
        } else {
          on[0] = in_n[0] + v[3]; on[1] = in_n[1] + v[4]; on[2] = in_n[2] + v[5];
          oa[0] = in_a[0] + v[6]; oa[1] = in_a[1] + v[7]; oa[2] = in_a[2] + v[8];
        }
        a.b.out_sample_count_weight[index] = in_weight + v[9];
        if (a.b.out_diagnostics) {
          rtb_diagnostics dg;
          dg.ray_count = (float)w.rays[lane];
          dg.bounds_hit_count = (float)w.node_tests[lane];
          dg.candidate_count = (float)w.sphere_tests[lane];
          dg.sample_count_weight = um::div(in_weight, (float)(int)in_color.w);
          a.b.out_diagnostics[index] = dg;
        }
      }
      uint32_t tnext = 0;
      if (lane == 0) tnext = atomicAdd(a.tile_counter, 1u);
      tnext = __shfl_sync(FULL, tnext, 0);
      if (tnext >= a.n_tiles) break;
      tile_range(a, tnext, &tile_base, &tile_n);
      uint32_t n_samples = 0;
      if (lane < tile_n) {
        int cx, cy;
        uint32_t index;
        active_pixel(a, tile_base + (uint32_t)lane, &cx, &cy, &index);
        const float in_w = a.b.in_color[4 * (size_t)index + 3];
        const float in_weight = a.b.in_sample_count_weight[index];
        float scw;
        n_samples = samples_to_accumulate(p, in_w, in_weight, &scw);
#pragma unroll
        for (int k = 0; k < kAccValues; k++) w.acc[lane][k] = 0ull;
        w.successes[lane] = 0;
        w.rays[lane] = 0;
        w.node_tests[lane] = 0;
        w.sphere_tests[lane] = 0;
        w.non_finite[lane] = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) w.fallback[lane][k] = 0;
      }
      uint32_t incl = n_samples;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += y;
      }
      if (lane < tile_n) w.prefix[lane] = incl - n_samples;
      total_items = __shfl_sync(FULL, incl, 31);
      if (lane == 0) w.prefix[tile_n] = total_items;
      next_item = 0;
      __syncwarp();
    }

    // ---- classify the pool: paths whose ray hit a sphere / everything else (miss, or empty slot) ----
    int n_hit = 0, n_done = 0;
#pragma unroll
    for (int r = 0; r < kPoolRounds; r++) {
      const int s = r * 32 + lane;
      bool hit = false, other = false;
      if (s < kPool) {
        const bool live = (w.meta[s] & kMetaLive) != 0;
        hit = live && w.idx[s] >= 0;
        other = !hit;
      }
      const uint32_t bh = __ballot_sync(FULL, hit), bo = __ballot_sync(FULL, other);
      if (hit) w.hit_q[n_hit + __popc(bh & lt_mask)] = (unsigned char)s;
      if (other) w.done_q[n_done + __popc(bo & lt_mask)] = (unsigned char)s;
      n_hit += __popc(bh);
      n_done += __popc(bo);
    }
    int n_walk = 0;
    n_live = 0;
    __syncwarp();

    // ---- shade: one bounce for every path that hit (SampleBatchJob.cs:308-336) ----
    for (int c = 0; c < n_hit; c += 32) {
      const bool act = c + lane < n_hit;
      bool walk = false, failed = false, dielectric = false;
      int s = 0;
      if (act) {
        s = w.hit_q[c + lane];
        const f3 o = um::mk(w.ox[s], w.oy[s], w.oz[s]), d = um::mk(w.dx[s], w.dy[s], w.dz[s]);
        const float t_hit = w.t[s];
        const int hit_idx = w.idx[s];
        uint32_t meta = w.meta[s];
        const uint32_t depth = meta >> 16, pixel = w.pixel[s], sample = w.sample[s];
        const int pix = (int)(meta & 0xffu);
        const float4 sp = sv.sphere(hit_idx);
        const uint32_t mi = sv.material_of(hit_idx);
        const float4* mp = reinterpret_cast<const float4*>(sd.materials + mi);
        const float4 m0 = __ldg(mp), m1 = __ldg(mp + 1), m2 = __ldg(mp + 2), m3 = __ldg(mp + 3);
        // HitRecord (Entity.cs:57-72, HitTests.cs:41-45)
        const f3 N = hit_normal<SMEM, kFlavorGeneral>(sv, sp, o, d, t_hit, RayClock{});
        const f3 P = um::mad(d, t_hit, o);
        const ScatterResult sc = scatter(m0, m1, m2, m3, d, N, pixel, sample, depth, p.seed);
        dielectric = __float_as_uint(m0.w) == RTB_MATERIAL_DIELECTRIC;
        const f3 emission = um::mk(m1.x, m1.y, m1.z);
        if (depth == 0) {
          w.nx[s] = N.x; w.ny[s] = N.y; w.nz[s] = N.z;
          if (sample == 0) { w.fallback[pix][0] = N.x; w.fallback[pix][1] = N.y; w.fallback[pix][2] = N.z; }
        }
        if (!(meta & kMetaNonSpecular) && __float_as_uint(m2.z) == 0u) {
          const f3 s_albedo = emission + sc.reflectance;
          w.ax[s] = s_albedo.x; w.ay[s] = s_albedo.y; w.az[s] = s_albedo.z;
          w.nx[s] = N.x; w.ny[s] = N.y; w.nz[s] = N.z;
          meta |= kMetaNonSpecular;
          if (sample == 0) {
            w.fallback[pix][0] = N.x; w.fallback[pix][1] = N.y; w.fallback[pix][2] = N.z;
            w.fallback[pix][3] = s_albedo.x; w.fallback[pix][4] = s_albedo.y; w.fallback[pix][5] = s_albedo.z;
          }
        }
        // forward form of the emission/attenuation unstack (SampleBatchJob.cs:383-396)
        const f3 thr = um::mk(w.thx[s], w.thy[s], w.thz[s]);
        const f3 rad = um::mad(thr, emission, um::mk(w.rx[s], w.ry[s], w.rz[s]));
        const f3 thr2 = thr * sc.reflectance;
        w.rx[s] = rad.x; w.ry[s] = rad.y; w.rz[s] = rad.z;
        w.thx[s] = thr2.x; w.thy[s] = thr2.y; w.thz[s] = thr2.z;
        w.events[s] += sc.random_events * pow2_neg(depth);
        // next ray (SampleBatchJob.cs:335-336, Ray.cs:18)
        const f3 off_n = um::dot(sc.dir, N) >= 0 ? N : -N;
        const f3 o2 = um::mad(off_n, 0.001f, P);
        meta += 1u << 16;
        if ((int)(depth + 1) == p.trace_depth) {
          failed = true;
          meta |= kMetaFailed;
        } else {
          walk = store_ray(w, s, sd, o2, sc.dir);
        }
        w.meta[s] = meta;
      }
      const uint32_t bw = __ballot_sync(FULL, walk), bf = __ballot_sync(FULL, failed);
      if (walk) w.walk_q[n_walk + __popc(bw & lt_mask)] = (unsigned char)s;
      if (failed) w.done_q[n_done + __popc(bf & lt_mask)] = (unsigned char)s;
      n_walk += __popc(bw);
      n_done += __popc(bf);
      n_live += __popc(__ballot_sync(FULL, act && !failed));
      if (COUNTERS && a.counters) {
        const uint32_t n_d = __popc(__ballot_sync(FULL, act && dielectric)), n_s = __popc(__ballot_sync(FULL, act && !dielectric));
        if (lane == 0 && n_s) atomicAdd(&a.counters[4], (unsigned long long)n_s);
        if (lane == 0 && n_d) atomicAdd(&a.counters[5], (unsigned long long)n_d);
      }
    }
    __syncwarp();

    // ---- finish + refill: sky, accumulate, next pixel-sample of the tile into the freed slot ----
    for (int c = 0; c < n_done; c += 32) {
      const bool act = c + lane < n_done;
      int s = 0, pix = -1;
      bool finishing = false, success = false;
      uint32_t path_rays = 0;
      float vals[kAccValues];
#pragma unroll
      for (int k = 0; k < kAccValues; k++) vals[k] = 0.0f;
      if (act) {
        s = w.done_q[c + lane];
        const uint32_t meta = w.meta[s];
        if (meta & kMetaLive) {
          finishing = true;
          pix = (int)(meta & 0xffu);
          const uint32_t depth = meta >> 16;
          if (meta & kMetaFailed) {
            path_rays = depth;                       // every bounce-loop iteration ran (SampleBatchJob.cs:203)
          } else {
            success = true;
            path_rays = depth + 1;
            const f3 d = um::mk(w.dx[s], w.dy[s], w.dz[s]);
            const f3 sky = sky_color(p.environment, sd, d);
            const f3 rad = um::mad(um::mk(w.thx[s], w.thy[s], w.thz[s]), sky, um::mk(w.rx[s], w.ry[s], w.rz[s]));
            f3 s_normal = um::mk(w.nx[s], w.ny[s], w.nz[s]), s_albedo = um::mk(w.ax[s], w.ay[s], w.az[s]);
            if (!(meta & kMetaNonSpecular)) {         // SampleBatchJob.cs:366-370
              s_normal = -d;
              s_albedo = sky;
              if (w.sample[s] == 0) {
                w.fallback[pix][0] = s_normal.x; w.fallback[pix][1] = s_normal.y; w.fallback[pix][2] = s_normal.z;
                w.fallback[pix][3] = sky.x; w.fallback[pix][4] = sky.y; w.fallback[pix][5] = sky.z;
              }
            }
            vals[0] = rad.x; vals[1] = rad.y; vals[2] = rad.z;
            vals[3] = s_normal.x; vals[4] = s_normal.y; vals[5] = s_normal.z;
            vals[6] = s_albedo.x; vals[7] = s_albedo.y; vals[8] = s_albedo.z;
            vals[9] = w.events[s];
          }
        }
      }
      // per-pixel sums: one warp reduction per distinct pixel among the finishing paths (usually 1-2)
      bool finite = true;
#pragma unroll
      for (int k = 0; k < kAccValues; k++) finite = finite && (um::abs(vals[k]) < 1.0e9f);
      uint32_t todo = __ballot_sync(FULL, finishing);
      while (todo) {
        const int leader = __ffs(todo) - 1;
        const int lpix = __shfl_sync(FULL, pix, leader);
        const bool mine = finishing && pix == lpix;
        todo &= ~__ballot_sync(FULL, mine);
        const bool add = mine && success && finite;
#pragma unroll
        for (int k = 0; k < kAccValues; k++) {
          const unsigned long long q = add ? (unsigned long long)__float2ll_rn(vals[k] * kFixedScale) : 0ull;
          const unsigned long long sum = warp_sum64(q);
          if (lane == k) w.acc[lpix][k] += sum;
        }
        const uint32_t n_succ = __reduce_add_sync(FULL, (mine && success) ? 1u : 0u);
        const uint32_t n_rays = __reduce_add_sync(FULL, mine ? path_rays : 0u);
        const uint32_t n_bad = __reduce_add_sync(FULL, (mine && success && !finite) ? 1u : 0u);
        if (lane == 10) w.successes[lpix] += n_succ;
        if (lane == 11) w.rays[lpix] += n_rays;
        if (lane == 12 && n_bad) w.non_finite[lpix] = 1;
      }
      // refill
      const uint32_t need = __ballot_sync(FULL, act);
      const uint32_t my_item = next_item + __popc(need & lt_mask);
      bool walk = false, born = false;
      if (act) {
        if (my_item < total_items) {
          int lo = 0, hi = tile_n;     // last pixel slot with prefix[slot] <= my_item
          while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (w.prefix[mid] <= my_item) lo = mid; else hi = mid;
          }
          const uint32_t sample = my_item - w.prefix[lo];
          int cx, cy;
          uint32_t pixel;
          active_pixel(a, tile_base + (uint32_t)lo, &cx, &cy, &pixel);
          const PathRay ray = camera_ray(p, cx, cy, pixel, sample);
          w.thx[s] = 1.0f; w.thy[s] = 1.0f; w.thz[s] = 1.0f;
          w.rx[s] = 0.0f; w.ry[s] = 0.0f; w.rz[s] = 0.0f;
          w.nx[s] = 0.0f; w.ny[s] = 0.0f; w.nz[s] = 0.0f;
          w.ax[s] = 0.0f; w.ay[s] = 0.0f; w.az[s] = 0.0f;
          w.events[s] = 0.0f;
          w.pixel[s] = pixel;
          w.sample[s] = sample;
          w.meta[s] = (uint32_t)lo | kMetaLive;
          born = true;
          walk = store_ray(w, s, sd, ray.o, ray.d);
        } else {
          w.meta[s] = 0;
        }
      }
      next_item = min(next_item + (uint32_t)__popc(need), total_items);
      const uint32_t bw = __ballot_sync(FULL, walk);
      if (walk) w.walk_q[n_walk + __popc(bw & lt_mask)] = (unsigned char)s;
      n_walk += __popc(bw);
      n_live += __popc(__ballot_sync(FULL, born));
    }
    __syncwarp();

    // ---- walk: closest hit for every queued ray, lanes fetch the next ray as they finish ----
    // (FindHitCandidates + FindHits, SampleBatchJob.cs:403-475; see closest_hit in kernel_common.cuh
    //  for why the pruned, ordered walk returns the record the reference's collect-all + sort does)
    if (n_walk > 0) {
      int my = -1, cur = kTraversalDone, sp = 0, best_idx = -1, next = 0;
      float best_t = um::INF, aa = 0;
      f3 o = um::mk(0.0f), d = um::mk(0.0f), inv = um::mk(0.0f);
      int stack[kStackMax];
      WorkCounters wc;
      for (;;) {
        const uint32_t idle = __ballot_sync(FULL, my < 0);
        if (idle) {
          if (next < n_walk) {
            const int k = next + __popc(idle & lt_mask);
            if (my < 0 && k < n_walk) {
              my = w.walk_q[k];
              o = um::mk(w.ox[my], w.oy[my], w.oz[my]);
              d = um::mk(w.dx[my], w.dy[my], w.dz[my]);
              inv = um::mk(w.ix[my], w.iy[my], w.iz[my]);
              aa = w.aa[my];
              best_t = um::INF;
              best_idx = -1;
              cur = sd.root_ref;
              stack[0] = kTraversalDone;
              sp = 1;
              if (COUNTERS) wc.node_tests++;        // the root box test made by store_ray
            }
            next = min(next + __popc(idle), n_walk);
          } else if (idle == FULL) {
            break;
          }
        }
        if (my >= 0) {
          bool pop = true;
          if (cur >= 0) {
            const float4 q0 = sv.node(cur, 0), q1 = sv.node(cur, 1), q2 = sv.node(cur, 2), q3 = sv.node(cur, 3);
            float tl, tr;
            bool hl = aabb_hit(um::mk(q0.x, q0.y, q0.z), um::mk(q0.w, q1.x, q1.y), o, inv, &tl);
            bool hr = aabb_hit(um::mk(q1.z, q1.w, q2.x), um::mk(q2.y, q2.z, q2.w), o, inv, &tr);
            if (COUNTERS) wc.node_tests += 2;
            const float limit = best_t * kPruneMargin;
            hl = hl && tl < limit;
            hr = hr && tr < limit;
            const int left = __float_as_int(q3.x), right = __float_as_int(q3.y);
            if (hl && hr) {
              const bool left_first = tl <= tr;
              stack[sp++] = left_first ? right : left;
              cur = left_first ? left : right;
              pop = false;
            } else if (hl || hr) {
              cur = hl ? left : right;
              pop = false;
            }
          } else {
            const uint32_t code = (uint32_t)~cur;
            const int first = (int)(code & ~15u);
            int count = (int)(code & 15u) + 1;
            if (count == 16) count = (int)sv.leaf_count(first);
            for (int i = 0; i < count; i++) {
              const float4 prim = sv.sphere(first + 16 * i);
              if (prim.w != prim.w) triangle_hit(sv, __float_as_uint(prim.x), first + 16 * i, o, d, best_t, best_idx);
              else sphere_hit<true>(sd, prim, first + 16 * i, o, d, inv, aa, best_t, best_idx);
            }
            if (COUNTERS) wc.sphere_tests += count;
          }
          if (pop) {
            cur = stack[--sp];
            if (cur == kTraversalDone) {
              w.t[my] = best_t;
              w.idx[my] = best_idx;
              if (COUNTERS) {
                const int pix = (int)(w.meta[my] & 0xffu);
                atomicAdd(&w.node_tests[pix], wc.node_tests);
                atomicAdd(&w.sphere_tests[pix], wc.sphere_tests);
                wc = WorkCounters();
              }
              my = -1;
            }
          }
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace rtbk
