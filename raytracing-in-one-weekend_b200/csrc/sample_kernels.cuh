// sample_kernels.cuh — the two kernels that replace SampleBatchJob.Execute
// (Runtime/Jobs/SampleBatchJob.cs:58-164) and SampleBatchJob.Sample (:166-401).
//
//   sample_megakernel   the product: persistent warps, one lane per pixel-sample, dead lanes
//                       refilled from the warp's tile by ballot/popc compaction of the work
//                       stream, scene + flattened BVH staged into shared memory by TMA bulk
//                       copies, order-independent fixed-point accumulation in shared memory,
//                       coalesced float4 stores when a tile retires.
//   sample_simple       one thread per pixel looping over its samples in order; the
//                       accumulation order is the reference's (bit-exact against the CPU
//                       oracle up to TraceDepth 64).  Validation and tiny images only.
#pragma once

#include "kernel_common.cuh"
#include "media.cuh"

namespace rtbk {

#ifndef RTB_MEGA_BLOCK
#define RTB_MEGA_BLOCK 1024           // threads per CTA; one persistent CTA per SM shares one staged world (64 regs/thread, no spills)
#endif
#ifndef RTB_MEGA_MIN_BLOCKS
#define RTB_MEGA_MIN_BLOCKS 1
#endif
#ifndef RTB_MEGA_BLOCK_GENERAL
#define RTB_MEGA_BLOCK_GENERAL 896    // the general flavour (triangles) needs 72 registers to stay out of local memory
#endif
#ifndef RTB_MEGA_BLOCK_PLACED
#define RTB_MEGA_BLOCK_PLACED 768     // the placed-entity flavour (transforms, Rect, Box): 80 registers.  The flavour is instruction-fetch bound (ncu: 5.8 no-instruction stalls per issue at 640 threads): more warps hide it — round 2, Cornell world: 384 / 512 / 640 / 704 / 768 / 832 / 896 / 1024 threads = 213 / 188 / 168 / 147 / 133 / 144 / 140 / 144 ms
#endif
#ifndef RTB_MEGA_BLOCK_MEDIA
#define RTB_MEGA_BLOCK_MEDIA 512      // the media flavour (media.cuh; hit lists in local memory): 128 registers.  With the CTA's warps running in step
                                      // (kPhased): 256 / 384 / 448 / 512 / 576 / 640 / 768 / 896 / 1024 threads = 523 / 410 / 387 / 356 / 392 / 372 / 397 /
                                      // 392 / 427 ms on the fog Cornell box (free-running warps: 708 ... 619 ms for 384 ... 1024 threads)
#endif
// Threads per CTA by kernel flavour (measured on B200, profiles/README.md): 1024 x 64 registers for the lean sphere
// builds, 896 x 72 registers for the general build.
__host__ __device__ constexpr int mega_block(int flavor) {
  return flavor == kFlavorMedia ? RTB_MEGA_BLOCK_MEDIA : flavor >= kFlavorPlaced ? RTB_MEGA_BLOCK_PLACED : flavor >= kFlavorGeneral ? RTB_MEGA_BLOCK_GENERAL : RTB_MEGA_BLOCK;
}
constexpr float kFixedScale = 4294967296.0f;            // 2^32
constexpr float kFixedInvScale = 2.3283064365386963e-10f;  // 2^-32
constexpr int kAccValues = 10;           // color.xyz, normal.xyz, albedo.xyz, sampleCountWeight

// Shared-memory state of one warp's tile.  Sums are 64-bit fixed point (2^-32): integer addition
// is associative, so a pixel's sum does not depend on the order in which its paths retire nor on
// how the frame is tiled or sharded — the image is bit-reproducible for any GPU count.
//   lane_*  : each lane's private partial sums for the pixel slot it is currently feeding
//             (plain LDS/STS, conflict-free [value][lane] layout, no atomics in the bounce loop)
//   acc_*   : per-pixel totals; a lane flushes its partials here (native 32-bit shared atomics,
//             carry propagated by hand) only when it moves to another pixel or the tile retires
struct WarpTile {
  uint2 lane_acc[kAccValues][32];
  uint32_t lane_counts[32];              // successes << 20 | rays (per flush interval)
  float aov[6][32];                      // the live path's sampleNormal / sampleAlbedo (written <= 2x per path, read once)
  uint32_t acc_lo[kTilePixelsMax][kAccValues];
  uint32_t acc_hi[kTilePixelsMax][kAccValues];
  uint32_t successes[kTilePixelsMax];
  uint32_t rays[kTilePixelsMax];
  uint32_t node_tests[kTilePixelsMax];
  uint32_t sphere_tests[kTilePixelsMax];
  uint32_t non_finite[kTilePixelsMax];
  float fallback[kTilePixelsMax][6];     // first sample's normal/albedo (SampleBatchJob.cs:152-156)
  // Two tiles are in flight per warp: the one whose samples are being issued and the previous one, whose last paths are
  // still running (its lanes are refilled from the NEXT tile meanwhile, so no lane idles through a tile's drain).  A tile
  // occupies one half of the pixel slots: slots [0, kHalfPixels) or [kHalfPixels, kTilePixelsMax).
  uint32_t prefix[2][kTilePixelsMax / 2 + 1];   // per half: exclusive prefix of per-pixel sample counts
  uint32_t pix_xy[kTilePixelsMax];       // image coordinates of the slot's pixel: x | y << 16
  uint32_t half_base[2];                 // per half: first active-pixel index of its tile
};
constexpr int kHalfPixels = kTilePixelsMax / 2;

// lane-private += x (exact: x * 2^32 is an integer for |x| >= 2^-9, rounded to 2^-32 below that)
__device__ __forceinline__ void lane_add(WarpTile& t, int lane, int v, float x) {
  const unsigned long long q = (unsigned long long)__float2ll_rn(x * kFixedScale);
  uint2 cur = t.lane_acc[v][lane];
  const unsigned long long sum = (((unsigned long long)cur.y << 32) | cur.x) + q;
  t.lane_acc[v][lane] = make_uint2((uint32_t)sum, (uint32_t)(sum >> 32));
}
// tile totals += this lane's partials; partials := 0.
//
// Range (rtb.h "Accumulation range"): a pixel total must stay below 2^30 in magnitude.  A total whose two top bits differ
// (|sum| in [2^30, 2^31)) marks the pixel instead of ever wrapping.  The window cannot be jumped: ordinary samples are below
// 2^20 and a lane's partial holds at most 1024 of them (< 2^30 ... flushed long before: see finish_path), a sample of 2^20
// or more is added alone (and is below 2^25, or the pixel is marked without adding it), and the totals are read back AFTER
// this lane's own adds, so the lane whose add carries a total across 2^30 sees it there (the concurrent flushes of the 31
// other lanes move it by < 31 * 2^25 < 2^30 meanwhile).  The hi-word add itself stays a fire-and-forget shared atomic: waiting for its result
// (ten dependent round trips per flush) cost the 64-spp configs up to 10 % (measured on the Cornell world).
__device__ __forceinline__ void lane_flush_inline(WarpTile& t, int lane, int slot) {
#pragma unroll
  for (int v = 0; v < kAccValues; v++) {
    const uint2 cur = t.lane_acc[v][lane];
    if (cur.x | cur.y) {
      const uint32_t old = atomicAdd(&t.acc_lo[slot][v], cur.x);
      const uint32_t hi = cur.y + ((old + cur.x) < old ? 1u : 0u);
      if (hi) atomicAdd(&t.acc_hi[slot][v], hi);
      t.lane_acc[v][lane] = make_uint2(0u, 0u);
    }
  }
  uint32_t window = 0u;
#pragma unroll
  for (int v = 0; v < kAccValues; v++) {
    const uint32_t h = t.acc_hi[slot][v];
    window |= h ^ (h << 1);
  }
  if (window & 0x80000000u) t.non_finite[slot] = 1;
  const uint32_t c = t.lane_counts[lane];
  if (c) {
    if (c >> 20) atomicAdd(&t.successes[slot], c >> 20);
    atomicAdd(&t.rays[slot], c & 0xfffffu);
    t.lane_counts[lane] = 0;
  }
}
// The lean sphere builds call it out of line (their hot loop is instruction-fetch sensitive: three inlined copies raised
// ncu's no-instruction stalls from 0.35 to 0.98 per issue on config 3); the general and placed builds, which run 64-spp
// worlds with a flush every other path and have registers to spare, inline it (Cornell world: 179 -> 161 ms).
__device__ __noinline__ void lane_flush_call(WarpTile& t, int lane, int slot) { lane_flush_inline(t, lane, slot); }
template <int FLAVOR>
__device__ __forceinline__ void lane_flush(WarpTile& t, int lane, int slot) {
#ifdef RTB_FLUSH_CALL_ALWAYS
  lane_flush_call(t, lane, slot);
#else
  if (FLAVOR >= kFlavorGeneral) lane_flush_inline(t, lane, slot);
  else lane_flush_call(t, lane, slot);
#endif
}
__device__ __forceinline__ float fixed_read(const WarpTile& t, int slot, int v) {
  long long q = (long long)(((unsigned long long)t.acc_hi[slot][v] << 32) | t.acc_lo[slot][v]);
  return __ll2float_rn(q) * kFixedInvScale;
}

__host__ __device__ inline size_t mega_smem_bytes(uint32_t blob_bytes, bool scene_in_smem, int flavor) {
  size_t s = 16;  // mbarrier
  if (scene_in_smem) s += blob_bytes;
  s = (s + 15) & ~(size_t)15;
  return s + sizeof(WarpTile) * (size_t)(mega_block(flavor) / 32);
}

// Active pixel k (0 <= k < n_active_pixels) -> image coordinates and the reference's index.
__device__ __forceinline__ void active_pixel(const BatchArgs& a, uint32_t k, int* cx, int* cy, uint32_t* index) {
  uint32_t j = k / (uint32_t)a.width;
  *cx = (int)(k - j * (uint32_t)a.width);
  *cy = a.first_row + (int)j * a.row_step;
  *index = (uint32_t)*cy * (uint32_t)a.width + (uint32_t)*cx;
}

// Writes the finished tile of `half` (n pixels) to the output buffers: lanes 0..n-1, one pixel each, float4 colour store.
// Out of line, like claim_tile: once per tile, and the hot loop stays short.
__device__ __noinline__ void retire_tile(WarpTile& tile, const BatchArgs& a, int lane, int half, int n) {
  if (lane < n) {
    const int s = half * kHalfPixels + lane;
    int cx, cy;
    uint32_t index;
    active_pixel(a, tile.half_base[half] + (uint32_t)lane, &cx, &cy, &index);
    const float4 in_color = reinterpret_cast<const float4*>(a.b.in_color)[index];
    const float in_weight = a.b.in_sample_count_weight[index];
    const int succ = (int)tile.successes[s];
    const int sample_count = (int)in_color.w + succ;
    const bool bad = tile.non_finite[s] != 0;
    const float nan = um::asfloat(0x7fc00000u);
    float v[kAccValues];
#pragma unroll
    for (int k = 0; k < kAccValues; k++) v[k] = bad ? nan : fixed_read(tile, s, k);
    float4 oc = make_float4(in_color.x + v[0], in_color.y + v[1], in_color.z + v[2], (float)sample_count);
    reinterpret_cast<float4*>(a.b.out_color)[index] = oc;
    const float* in_n = a.b.in_normal + 3 * (size_t)index;
    const float* in_a = a.b.in_albedo + 3 * (size_t)index;
    float* on = a.b.out_normal + 3 * (size_t)index;
    float* oa = a.b.out_albedo + 3 * (size_t)index;
    if (sample_count == 0) {
      on[0] = tile.fallback[s][0]; on[1] = tile.fallback[s][1]; on[2] = tile.fallback[s][2];
      oa[0] = tile.fallback[s][3]; oa[1] = tile.fallback[s][4]; oa[2] = tile.fallback[s][5];
    } else {
      on[0] = in_n[0] + v[3]; on[1] = in_n[1] + v[4]; on[2] = in_n[2] + v[5];
      oa[0] = in_a[0] + v[6]; oa[1] = in_a[1] + v[7]; oa[2] = in_a[2] + v[8];
    }
    a.b.out_sample_count_weight[index] = in_weight + v[9];
    if (a.b.out_diagnostics) {
      rtb_diagnostics dg;
      dg.ray_count = (float)tile.rays[s];
      dg.bounds_hit_count = (float)tile.node_tests[s];
      dg.candidate_count = (float)tile.sphere_tests[s];
      dg.sample_count_weight = um::div(in_weight, (float)(int)in_color.w);
      a.b.out_diagnostics[index] = dg;
    }
  }
}
// Claims the launch's next tile into `half`: its pixels' slots are cleared and the (pixel slot, sample) stream is set up.
// Returns the tile's pixel count and its total sample count, or pixel count -1 when there is none (or the host's
// CancellationToken was set: SampleBatchJob.cs:61 polls it per pixel; here the warp that claims a tile reads the flag, and
// tells the CTA's other warps to stop issuing samples).
struct ClaimedTile { int n; uint32_t items; };
__device__ __noinline__ ClaimedTile claim_tile(WarpTile& tile, const BatchArgs& a, int lane, int half, volatile uint32_t* cta_cancelled) {
  const rtb_batch_params& p = a.p;
  int tile_n = 0;
  uint32_t total_items = 0;
  uint32_t t = 0;
  if (lane == 0) {
    const bool cancelled = cancel_requested(a.cancel_flag, a.cancel_epoch);
    t = atomicAdd(a.tile_counter, 1u);
    if (cancelled) { *cta_cancelled = 1u; t = 0xffffffffu; }
  }
  t = __shfl_sync(0xffffffffu, t, 0);
  if (t >= a.n_tiles) return ClaimedTile{-1, 0u};
  uint32_t base;
  tile_range(a, t, &base, &tile_n);
  uint32_t n_samples = 0;
  if (lane < tile_n) {
    const int s = half * kHalfPixels + lane;
    int cx, cy;
    uint32_t index;
    active_pixel(a, base + (uint32_t)lane, &cx, &cy, &index);
    tile.pix_xy[s] = (uint32_t)cx | ((uint32_t)cy << 16);
    const float in_w = a.b.in_color[4 * (size_t)index + 3];
    const float in_weight = a.b.in_sample_count_weight[index];
    float scw;
    n_samples = samples_to_accumulate(p, in_w, in_weight, &scw);
#pragma unroll
    for (int k = 0; k < kAccValues; k++) { tile.acc_lo[s][k] = 0; tile.acc_hi[s][k] = 0; }
    tile.successes[s] = 0;
    tile.rays[s] = 0;
    tile.node_tests[s] = 0;
    tile.sphere_tests[s] = 0;
    tile.non_finite[s] = 0;
#pragma unroll
    for (int k = 0; k < 6; k++) tile.fallback[s][k] = 0;
  }
  // exclusive prefix over the tile's pixels
  uint32_t incl = n_samples;
#pragma unroll
  for (int o = 1; o < kHalfPixels; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane < tile_n) tile.prefix[half][lane] = incl - n_samples;
  total_items = __shfl_sync(0xffffffffu, incl, kHalfPixels - 1);
  if (lane == 0) { tile.prefix[half][tile_n] = total_items; tile.half_base[half] = base; }
  __syncwarp();
  return ClaimedTile{tile_n, total_items};
}


// The media flavour's step (1): one bounce-loop iteration's hit search with the volume bookkeeping (media.cuh), out of line —
// its hit list lives in this call's frame, not in the kernel's.
template <bool SMEM, bool COUNTERS>
__device__ __noinline__ void media_step_call(const SceneView<SMEM>& sv, const SceneDesc& sd, f3 o, f3 d, const RayClock& clk, int current_volume,
                                             uint32_t pixel, uint32_t sample, uint32_t depth, uint32_t seed, WorkCounters& wc, MediaStep* out) {
  RayHits hits;
  WhiteNoise unused{};
  *out = media_step<SMEM, COUNTERS, true, false>(sv, sd, o, d, clk, current_volume, pixel, sample, depth, seed, unused, wc, hits);
}

template <bool SMEM, bool COUNTERS, int FLAVOR>
__global__ void __launch_bounds__(mega_block(FLAVOR), RTB_MEGA_MIN_BLOCKS) sample_megakernel(const __grid_constant__ BatchArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  volatile uint32_t* cta_cancelled = reinterpret_cast<volatile uint32_t*>(smem + 8);   // set by the first warp that sees the token
  unsigned char* blob_smem = smem + 16;
  const size_t tiles_off = (16 + (SMEM ? a.scene.blob_bytes : 0) + 15) & ~(size_t)15;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpTile& tile = reinterpret_cast<WarpTile*>(smem + tiles_off)[warp];

  // ---- stage the world into shared memory: TMA bulk copies signalled on one mbarrier ----
  SceneView<SMEM> sv;
  if (threadIdx.x == 0) *cta_cancelled = 0u;
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
    __syncthreads();          // the barrier init is visible before anyone polls it
    mbar_wait(bar, 0);
    sv.bind(blob_smem, a.scene);
    if (RTB_ABS_REFS) {
      sv.rebias_inner_refs(a.scene.n_inner, threadIdx.x, blockDim.x);
      __syncthreads();
    }
  } else {
    __syncthreads();
    sv.bind(a.scene.blob, a.scene);
  }

  const rtb_batch_params& p = a.p;
  const uint32_t lt_mask = (1u << lane) - 1u;

  // ---- per-lane path state ----
  bool alive = false;
  int slot = 0;               // pixel slot inside the warp tile
  uint32_t pixel = 0, sample = 0;
  int depth = 0;
  bool first_non_specular = false;
  PathRay ray{um::mk(0.0f), um::mk(0.0f)};
  f3 throughput = um::mk(1.0f), radiance = um::mk(0.0f);
  float events_acc = 0;
  // media flavour only: currentProbabilisticVolumeMaterial (material index, -1 = null; SampleBatchJob.cs:181) and "the ray's Time
  // is 0" (a ray scattered inside a medium is a new Ray(point, direction): Material.cs:163-168)
  int current_volume = -1;
  bool time_zero = false;
  auto set_normal = [&](f3 n) { tile.aov[0][lane] = n.x; tile.aov[1][lane] = n.y; tile.aov[2][lane] = n.z; };
  auto set_albedo = [&](f3 c) { tile.aov[3][lane] = c.x; tile.aov[4][lane] = c.y; tile.aov[5][lane] = c.z; };
  int acc_slot = -1;          // pixel slot this lane's private partial sums belong to
  WorkCounters wc;
#pragma unroll
  for (int k = 0; k < kAccValues; k++) tile.lane_acc[k][lane] = make_uint2(0u, 0u);
  tile.lane_counts[lane] = 0;

  // ---- warp-uniform tile state ----
  uint32_t next_item = 0, total_items = 0;   // work stream of the tile being issued
  int tile_n = 0;             // its pixel count (0: none)
  int drain_n = 0;            // pixel count of the previous tile while its last paths run (0: none)
  uint32_t tile_flags = 0;    // bit 0: half of the tile being issued; bit 1: the launch has no tiles left

  // A finished path (sky reached: success; TraceDepth exhausted: failed, SampleBatchJob.cs:379-381) is added to the
  // lane's private partial sums for its pixel; the lane is then free for the tile's next pixel-sample.
  auto finish_path = [&](bool success) {
    alive = false;
    if (acc_slot != slot) {
      if (acc_slot >= 0) lane_flush<FLAVOR>(tile, lane, acc_slot);
      acc_slot = slot;
    }
    if (COUNTERS) {
      atomicAdd(&tile.node_tests[slot], wc.node_tests);
      atomicAdd(&tile.sphere_tests[slot], wc.sphere_tests);
      if (a.counters) {
        if (wc.shade_standard) atomicAdd(&a.counters[4], (unsigned long long)wc.shade_standard);
        if (wc.shade_dielectric) atomicAdd(&a.counters[5], (unsigned long long)wc.shade_dielectric);
      }
      wc = WorkCounters();
    }
    // bounce-loop iterations of this path (SampleBatchJob.cs:203): every hit so far, plus the miss that ended it
    uint32_t counts = tile.lane_counts[lane] + (uint32_t)depth + (success ? 1u : 0u);
    bool flush_now = false;
    if (success) {
      const float vals[kAccValues] = {radiance.x, radiance.y, radiance.z, tile.aov[0][lane], tile.aov[1][lane], tile.aov[2][lane],
                                      tile.aov[3][lane], tile.aov[4][lane], tile.aov[5][lane], events_acc};
      // Range of the fixed-point sums (rtb.h "Accumulation range").  Ordinary samples (every component below 2^20): a lane's
      // partial holds at most 32 of them between flushes (see below), so it stays below 2^25.  A brighter sample goes alone:
      // partials out first, the sample in, out again — lane_flush sees the pixel total reach 2^30 if it does.
      bool ordinary = true;
#pragma unroll
      for (int k = 0; k < kAccValues; k++) ordinary = ordinary && (um::abs(vals[k]) < 1048576.0f);
      if (!ordinary) {
        bool finite = true;
#pragma unroll
        for (int k = 0; k < kAccValues; k++) finite = finite && (um::abs(vals[k]) < 33554432.0f);   // 2^25: 32 lanes of them < 2^30
        if (finite) {
          lane_flush<FLAVOR>(tile, lane, acc_slot);
          counts = (uint32_t)depth + 1u;
          flush_now = true;
        } else {
          tile.non_finite[slot] = 1;
        }
        ordinary = finite;
      }
      if (ordinary) {
#pragma unroll
        for (int k = 0; k < kAccValues; k++) lane_add(tile, lane, k, vals[k]);
      }
      counts += 1u << 20;
    }
    tile.lane_counts[lane] = counts;
    // the packed counters hold 2^12 - 1 successes / 2^20 - 1 rays: flush well before either wraps, and often enough (every
    // 32 successes: a partial then stays below 2^25) for lane_flush's overflow window to be airtight
    if (flush_now || (counts >> 20) >= 32u || (counts & 0xfffffu) >= 0x80000u) lane_flush<FLAVOR>(tile, lane, acc_slot);
  };

  // One trip of the loop, for every lane of the warp:
  //   (1) walk    closest hit for the lane's ray; a ray that leaves the world ends its path here (sky, accumulate)
  //   (2) refill  lanes without a path take the tile's next pixel-samples (ballot / popc compaction of the work stream)
  //   (3) draw    ONE converged step serves both kinds of lanes: the Philox block of the bounce (hit lanes) or of the
  //               camera ray (refilled lanes), and the sincos both need
  //   (4) shade   hit lanes finish Material.Scatter and form the next ray; refilled lanes form their camera ray
  // The placed and media flavours run the trip's two halves IN STEP across the CTA's warps: a barrier at the top of the trip
  // and one between step (1) — the walk with its entity tests, the media bookkeeping, the accumulation of finished paths —
  // and steps (2)-(4) — refill, draw, shade.  These flavours are bound by instruction fetch (ncu: 5.8 and 17 no-instruction
  // stalls per issue; 5 k and 6.6 k instructions of divergent code against a 32 KB L1.5 instruction cache, insensitive to the
  // CTA size): with every warp of the SM inside the same half, the code in flight fits.  Measured (bit-identical images):
  // Cornell box 132.8 -> 114.9 ms, as a linear list 99.3 -> 86.0, with media 593 -> 393; barrier only at the top 123.7 /
  // 513; a third one before the shade 115.7 / 393; barriers inside the media hit search (walk | tests) 500.  The general
  // flavour (mesh world 69 -> 135 ms: long walks of uneven length wait for each other) and the lean sphere builds (config 3
  // 128.5 -> 163) are not fetch-bound and keep free-running warps.
  constexpr bool kPhased = FLAVOR >= kFlavorPlaced;
  bool warp_done = false;     // kPhased: this warp has nothing left, but keeps meeting the CTA's barriers until every warp says so
  for (;;) {
    if (kPhased) { if (__syncthreads_and(warp_done ? 1 : 0)) break; }
    if (!(kPhased && warp_done)) {
    const uint32_t alive_mask = __ballot_sync(0xffffffffu, alive);
    // the previous tile retires as soon as its last path has finished
    if (drain_n > 0) {
      const int dh = (int)(tile_flags & 1u) ^ 1;
      if (__ballot_sync(0xffffffffu, alive && (slot >= kHalfPixels ? 1 : 0) == dh) == 0) {
        if (acc_slot >= 0 && (acc_slot >= kHalfPixels ? 1 : 0) == dh) { lane_flush<FLAVOR>(tile, lane, acc_slot); acc_slot = -1; }
        __syncwarp();
        retire_tile(tile, a, lane, dh, drain_n);
        drain_n = 0;
      }
    }
    // the tile being issued has no samples left: it becomes the draining one, the next tile goes to the other half
    // (if the previous tile is STILL draining — a path of it outlived a whole tile — the free lanes wait for it)
    if (next_item >= total_items && drain_n == 0) {
      if (!(tile_flags & 2u)) {
        drain_n = tile_n;
        tile_flags ^= 1u;
        const ClaimedTile c = claim_tile(tile, a, lane, (int)(tile_flags & 1u), cta_cancelled);
        next_item = 0;
        if (c.n < 0) { tile_flags |= 2u; tile_n = 0; total_items = 0; }
        else { tile_n = c.n; total_items = c.items; }
      } else if (alive_mask == 0) {
        if (kPhased) warp_done = true; else
        break;                // no tile left, nothing draining, nothing in flight
      }
    }
    }

    // (1) one walk for every live lane (SampleBatchJob.cs:184-206, 341-374)
    bool hit = false;
    float t_hit = 0;
    float4 m0 = make_float4(0, 0, 0, 0), m1 = m0, m2 = m0, m3 = m0;
    f3 N = um::mk(0.0f), P = um::mk(0.0f);
    float events0 = 0;          // media flavour: RandomEvents of this iteration's Material.ProbabilisticHit draws
    if (alive) {
      int hit_idx = -1;
      const RayClock clk{pixel, sample, p.seed, 0.0f, FLAVOR == kFlavorMedia && time_zero};
      const bool exhausted = depth == p.trace_depth;      // a failed sample (SampleBatchJob.cs:379-381), found at the end of the last trip
      if (FLAVOR == kFlavorMedia) {
        if (!exhausted) {
          // the iteration's hit search + the reference's volume bookkeeping (SampleBatchJob.cs:184-303; media.cuh)
          MediaStep st;
          media_step_call<SMEM, COUNTERS>(sv, a.scene, ray.o, ray.d, clk, current_volume, pixel, sample, (uint32_t)depth, p.seed, wc, &st);
          current_volume = st.current_volume;
          events0 = st.events;
          if (st.hit) {
            hit_idx = 0;
            t_hit = st.t;
            const float4* mp = reinterpret_cast<const float4*>(a.scene.materials + st.material);
            m0 = __ldg(mp); m1 = __ldg(mp + 1); m2 = __ldg(mp + 2); m3 = __ldg(mp + 3);
            N = st.n;                                     // the record's normal (-direction for a hit inside a medium)
            P = um::mad(ray.d, t_hit, ray.o);
          }
        }
      } else if (!exhausted) {
        closest_hit<SMEM, COUNTERS, FLAVOR>(sv, a.scene, ray.o, ray.d, t_hit, hit_idx, wc, clk);
      }
      if (hit_idx >= 0) {
        hit = true;
        if (FLAVOR != kFlavorMedia) {
        const float4 s = sv.sphere(hit_idx);
        const uint32_t mi = sv.material_of(hit_idx);
        const float4* mp = reinterpret_cast<const float4*>(a.scene.materials + mi);
        m0 = __ldg(mp); m1 = __ldg(mp + 1); m2 = __ldg(mp + 2); m3 = __ldg(mp + 3);
        if (flavor_has_textures(FLAVOR) && __float_as_uint(m3.w) != 0u) {
          // through copies: the out-of-line call takes addresses, and the material registers must not move to local memory for it
          float4 t0 = m0, t1 = m1, t2 = m2, t3 = m3;
          resolve_textures(sv, a.scene, mi, s, ray.o, ray.d, t0, t1, t2, t3);
          m0 = t0; m1 = t1; m2 = t2; m3 = t3;
        }
        // HitRecord (Entity.cs:57-72, HitTests.cs:41-45)
        N = hit_normal<SMEM, FLAVOR>(sv, s, ray.o, ray.d, t_hit, clk);
        P = um::mad(ray.d, t_hit, ray.o);
        }
      } else {
        if (!exhausted) {
          const f3 sky = sky_color(p.environment, a.scene, ray.d);
          radiance = um::mad(throughput, sky, radiance);
          if (FLAVOR == kFlavorMedia) events_acc += events0 * pow2_neg((uint32_t)depth);   // draws of media the ray passed through
          if (!first_non_specular) {
            const f3 s_normal = -ray.d;
            set_albedo(sky);
            set_normal(s_normal);
            if (sample == 0) {
              tile.fallback[slot][0] = s_normal.x; tile.fallback[slot][1] = s_normal.y; tile.fallback[slot][2] = s_normal.z;
              tile.fallback[slot][3] = sky.x; tile.fallback[slot][4] = sky.y; tile.fallback[slot][5] = sky.z;
            }
          }
        }
        finish_path(!exhausted);        // the kernel's ONE copy of the accumulation code
      }
    }

    if (kPhased) __syncthreads();
    // (2) refill dead lanes from the tile's work stream
    bool fresh = false;
    int cx = 0, cy = 0;
    const uint32_t need = __ballot_sync(0xffffffffu, !alive);
    // Refill once `refill_min` lanes are free (or none is alive): the refill and camera-ray code then runs with that many
    // lanes instead of one or two, and fewer trips pay for it; the free lanes idle through a walk or two meanwhile.
    // Measured on config 3 (BVH walk: lanes idle in it anyway): 1 / 4 / 8 / 12 / 16 / 24 lanes = 124.9 / 123.7 / 121.8 /
    // 122.7 / 123.7 / 132.3 ms, mesh world 70.2 -> 67.1 ms; on the linear list (converged walk: an idle lane is pure loss)
    // and on worlds of a dozen entities 1 is best — the plugin picks 8 for trees of >= 64 inner nodes.
    if (need && ((uint32_t)__popc(need) >= a.refill_min || need == 0xffffffffu)) {
      if (*cta_cancelled) next_item = total_items;      // cancelled: the tile's remaining samples are not started
      const uint32_t my_item = next_item + __popc(need & lt_mask);
      if (!alive && my_item < total_items) {
        // item -> (pixel slot, sample) through the tile's prefix table
        const uint32_t* prefix = tile.prefix[tile_flags & 1u];
        int lo = 0, hi = tile_n;   // find the last pixel with prefix[pixel] <= my_item
        while (hi - lo > 1) {
          int mid = (lo + hi) >> 1;
          if (prefix[mid] <= my_item) lo = mid; else hi = mid;
        }
        sample = my_item - prefix[lo];
        slot = (int)(tile_flags & 1u) * kHalfPixels + lo;
        const uint32_t xy = tile.pix_xy[slot];
        cx = (int)(xy & 0xffffu);
        cy = (int)(xy >> 16);
        pixel = (uint32_t)cy * (uint32_t)a.width + (uint32_t)cx;
        fresh = true;
        depth = 0;
        first_non_specular = false;
        throughput = um::mk(1.0f);
        radiance = um::mk(0.0f);
        set_normal(um::mk(0.0f));
        set_albedo(um::mk(0.0f));
        events_acc = 0;
        current_volume = -1;
        time_zero = false;
      }
      next_item = min(next_item + (uint32_t)__popc(need), total_items);
    }

    // (3) the shared draw: Philox(pixel, sample, bounce | CAMERA, block) and sincos(u * 2 PI)
    ScatterPlan pl{};
    bool want_rng = false, want_angle = false;
    uint32_t c2 = kBounceCamera, c3 = 0;
    const bool isotropic = FLAVOR == kFlavorMedia && hit && __float_as_uint(m0.w) == RTB_MATERIAL_PROBABILISTIC_VOLUME;
    if (hit) {
      pl = scatter_plan(m0, m1, m2);
      if (isotropic) { pl.block = 0u; pl.need_angle = true; }   // Material.cs:163-168: RandomSource.NextFloat3Direction, block 0's u, v
      c2 = (uint32_t)depth;
      c3 = pl.block;
      want_rng = true;
      want_angle = pl.need_angle;
    } else if (fresh) {
      want_angle = p.view.lens_radius != 0;
      want_rng = p.sub_pixel_jitter || want_angle;
    }
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    float sn = 0, cs = 1;
    if (want_rng) {
      r = philox4x32_10(pixel, sample, c2, c3, p.seed, kPhiloxKey1);
      if (want_angle) unit_angle_sincos(u2f(hit ? r.y : r.z), &sn, &cs);
    }

    // (4) shade / camera ray
    if (hit) {
      ScatterResult sc;
      if (isotropic) {
        sc.dir = random_direction(u2f(r.x), sn, cs);
        sc.reflectance = um::mk(m0.x, m0.y, m0.z);
        sc.random_events = events0 + 2.0f;
        time_zero = true;
      } else {
        sc = scatter_finish(m0, m1, m2, m3, ray.d, N, pl, r, sn, cs, pixel, sample, (uint32_t)depth, p.seed, FLAVOR == kFlavorMedia ? events0 : 0.0f);
        if (COUNTERS) { if (__float_as_uint(m0.w) == RTB_MATERIAL_DIELECTRIC) wc.shade_dielectric++; else wc.shade_standard++; }
      }
      const f3 emission = um::mk(m1.x, m1.y, m1.z);
      if (depth == 0) {
        set_normal(N);
        if (sample == 0) { tile.fallback[slot][0] = N.x; tile.fallback[slot][1] = N.y; tile.fallback[slot][2] = N.z; }
      }
      if (!first_non_specular && __float_as_uint(m2.z) == 0u) {
        const f3 s_albedo = emission + sc.reflectance;
        set_albedo(s_albedo);
        set_normal(N);
        first_non_specular = true;
        if (sample == 0) {
          tile.fallback[slot][0] = N.x; tile.fallback[slot][1] = N.y; tile.fallback[slot][2] = N.z;
          tile.fallback[slot][3] = s_albedo.x; tile.fallback[slot][4] = s_albedo.y; tile.fallback[slot][5] = s_albedo.z;
        }
      }
      // forward form of the emission/attenuation unstack (SampleBatchJob.cs:383-396)
      radiance = um::mad(throughput, emission, radiance);
      throughput = throughput * sc.reflectance;
      // RandomEvents / pow(2, depth) (SampleBatchJob.cs:332): dividing by a power of two == multiplying by its inverse, exactly
      events_acc += sc.random_events * pow2_neg((uint32_t)depth);
      // next ray (SampleBatchJob.cs:335-336, Ray.cs:18)
      const f3 off_n = um::dot(sc.dir, N) >= 0 ? N : -N;
      ray.o = um::mad(off_n, 0.001f, P);
      ray.d = sc.dir;
      depth++;                          // depth == TraceDepth: the sample failed (:379-381); the next trip's step (1) retires it
    } else if (fresh) {
      ray = camera_ray_finish(p, cx, cy, r, sn, cs);
      alive = true;
    }
    __syncwarp();
  }

}

// Sums the per-pixel diagnostics into rtb_counters (instrumented runs only).
__global__ void counters_from_diagnostics(const rtb_diagnostics* __restrict__ diag, const float* __restrict__ color4,
                                          const float* __restrict__ in_color4, BatchArgs a) {
  // grid-stride over active pixels
  unsigned long long rays = 0, nodes = 0, spheres = 0, succ = 0, attempted = 0;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < a.n_active_pixels; k += gridDim.x * blockDim.x) {
    int cx, cy;
    uint32_t index;
    active_pixel(a, k, &cx, &cy, &index);
    float scw;
    attempted += samples_to_accumulate(a.p, in_color4[4 * (size_t)index + 3], a.b.in_sample_count_weight[index], &scw);
    rays += (unsigned long long)diag[index].ray_count;
    nodes += (unsigned long long)diag[index].bounds_hit_count;
    spheres += (unsigned long long)diag[index].candidate_count;
    succ += (unsigned long long)((int)color4[4 * (size_t)index + 3] - (int)in_color4[4 * (size_t)index + 3]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    rays += __shfl_down_sync(0xffffffffu, rays, o);
    nodes += __shfl_down_sync(0xffffffffu, nodes, o);
    spheres += __shfl_down_sync(0xffffffffu, spheres, o);
    succ += __shfl_down_sync(0xffffffffu, succ, o);
    attempted += __shfl_down_sync(0xffffffffu, attempted, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&a.counters[0], attempted);
    atomicAdd(&a.counters[7], attempted - succ);   // failed samples (SampleBatchJob.cs:379-381)
    atomicAdd(&a.counters[1], rays);
    atomicAdd(&a.counters[2], nodes);
    atomicAdd(&a.counters[3], spheres);
    atomicAdd(&a.counters[6], succ);   // successes == sky terminations
  }
}

// ---------------------------------------------------------------------------------------
// sample_simple: thread per active pixel, samples in order.  Emission/attenuation stacks
// are kept (up to 64 entries) and unwound tail -> head exactly like SampleBatchJob.cs:383-396,
// and samples are added in index order, so colour sums are bit-identical to the oracle's.
// ---------------------------------------------------------------------------------------
constexpr int kSimpleStack = 64;

template <bool COUNTERS, bool WHITE>
__global__ void __launch_bounds__(128) sample_simple(const __grid_constant__ BatchArgs a) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.n_active_pixels) return;
  if (cancel_requested(a.cancel_flag, a.cancel_epoch)) return;          // CancellationToken (SampleBatchJob.cs:61): pixels not yet started are skipped
  const rtb_batch_params& p = a.p;
  SceneView<false> sv;
  sv.bind(a.scene.blob, a.scene);
  int cx, cy;
  uint32_t index;
  active_pixel(a, k, &cx, &cy, &index);

  const float4 in_color = reinterpret_cast<const float4*>(a.b.in_color)[index];
  f3 color_acc = um::mk(in_color.x, in_color.y, in_color.z);
  f3 normal_acc = v3(a.b.in_normal + 3 * (size_t)index);
  f3 albedo_acc = v3(a.b.in_albedo + 3 * (size_t)index);
  float weight_acc = a.b.in_sample_count_weight[index];
  int sample_count = (int)in_color.w;
  float scw;
  const uint32_t n = samples_to_accumulate(p, in_color.w, weight_acc, &scw);
  f3 fb_normal = um::mk(0.0f), fb_albedo = um::mk(0.0f);
  uint32_t rays = 0;
  WorkCounters wc;
  const bool exact = p.trace_depth <= kSimpleStack;
  f3 att[kSimpleStack], emi[kSimpleStack];
  WhiteNoise white{};
  if (WHITE) white.init((p.seed * 0x8C4CA03Fu) ^ (index * 0x7383ED49u));      // SampleBatchJob.cs:91

  for (uint32_t s = 0; s < n; s++) {
    RayClock clk{index, s, p.seed, 0.0f, true};
    PathRay ray = WHITE ? camera_ray_white(p, cx, cy, white, &clk.value) : camera_ray(p, cx, cy, index, s);
    if (!WHITE && a.scene.n_placed) { clk.known = false; clk.value = clk.time(); clk.known = true; }
    f3 throughput = um::mk(1.0f), radiance = um::mk(0.0f);
    f3 s_normal = um::mk(0.0f), s_albedo = um::mk(0.0f);
    bool first_non_specular = false;
    float events_acc = 0, pow2depth = 1;
    int depth = 0, entries = 0;
    for (; depth < p.trace_depth; depth++) {
      float t_hit;
      int hit_idx;
      closest_hit<false, COUNTERS, kFlavorPlaced, /*TIES*/ true>(sv, a.scene, ray.o, ray.d, t_hit, hit_idx, wc, clk);
      rays++;
      if (hit_idx >= 0) {
        const float4 sp = sv.sphere(hit_idx);
        const uint32_t mi = sv.material_of(hit_idx);
        const float4* mp = reinterpret_cast<const float4*>(a.scene.materials + mi);
        float4 m0 = __ldg(mp), m1 = __ldg(mp + 1), m2 = __ldg(mp + 2), m3 = __ldg(mp + 3);
        if (__float_as_uint(m3.w) != 0u) resolve_textures(sv, a.scene, mi, sp, ray.o, ray.d, m0, m1, m2, m3);
        const f3 N = hit_normal<false, kFlavorPlaced>(sv, sp, ray.o, ray.d, t_hit, clk);
        const f3 P = um::mad(ray.d, t_hit, ray.o);
        const ScatterResult sc = WHITE ? scatter_white(m0, m1, m2, m3, ray.d, N, white)
                                       : scatter(m0, m1, m2, m3, ray.d, N, index, s, (uint32_t)depth, p.seed);
        if (COUNTERS) { if (__float_as_uint(m0.w) == RTB_MATERIAL_DIELECTRIC) wc.shade_dielectric++; else wc.shade_standard++; }
        const f3 emission = um::mk(m1.x, m1.y, m1.z);
        if (depth == 0) s_normal = N;
        if (!first_non_specular && __float_as_uint(m2.z) == 0u) {
          s_albedo = emission + sc.reflectance;
          s_normal = N;
          first_non_specular = true;
        }
        if (exact) { emi[entries] = emission; att[entries] = sc.reflectance; entries++; }
        radiance = um::mad(throughput, emission, radiance);
        throughput = throughput * sc.reflectance;
        events_acc += um::div(sc.random_events, pow2depth);
        const f3 off_n = um::dot(sc.dir, N) >= 0 ? N : -N;
        ray.o = um::mad(off_n, 0.001f, P);
        ray.d = sc.dir;
        pow2depth *= 2.0f;
      } else {
        const f3 sky = sky_color(p.environment, a.scene, ray.d);
        if (exact) { emi[entries] = sky; att[entries] = um::mk(1.0f); entries++; }
        radiance = um::mad(throughput, sky, radiance);
        if (!first_non_specular) { s_albedo = sky; s_normal = -ray.d; }
        break;
      }
    }
    if (depth != p.trace_depth) {
      f3 c = radiance;
      if (exact) {
        c = um::mk(0.0f);
        for (int e = entries; e-- > 0;) { c = c * att[e]; c = c + emi[e]; }
      }
      color_acc = color_acc + c;
      normal_acc = normal_acc + s_normal;
      albedo_acc = albedo_acc + s_albedo;
      weight_acc += events_acc;
      sample_count++;
    }
    if (s == 0) { fb_normal = s_normal; fb_albedo = s_albedo; }
  }

  reinterpret_cast<float4*>(a.b.out_color)[index] = make_float4(color_acc.x, color_acc.y, color_acc.z, (float)sample_count);
  const f3 on = sample_count == 0 ? fb_normal : normal_acc;
  const f3 oa = sample_count == 0 ? fb_albedo : albedo_acc;
  float* pn = a.b.out_normal + 3 * (size_t)index;
  float* pa = a.b.out_albedo + 3 * (size_t)index;
  pn[0] = on.x; pn[1] = on.y; pn[2] = on.z;
  pa[0] = oa.x; pa[1] = oa.y; pa[2] = oa.z;
  a.b.out_sample_count_weight[index] = weight_acc;
  if (COUNTERS && a.counters) {
    if (wc.shade_standard) atomicAdd(&a.counters[4], (unsigned long long)wc.shade_standard);
    if (wc.shade_dielectric) atomicAdd(&a.counters[5], (unsigned long long)wc.shade_dielectric);
  }
  if (a.b.out_diagnostics) {
    rtb_diagnostics dg;
    dg.ray_count = (float)rays;
    dg.bounds_hit_count = (float)wc.node_tests;
    dg.candidate_count = (float)wc.sphere_tests;
    dg.sample_count_weight = scw;
    a.b.out_diagnostics[index] = dg;
  }
}

}  // namespace rtbk
