// sample_volumes: the sample job for worlds with MaterialType.ProbabilisticVolume materials.
//
// Participating media need what the closest-hit walk of the other kernels throws away: the reference decides entry /
// exit / containment from the SORTED LIST OF ALL HITS along the ray, with an extra exit hit injected for convex media
// and a backwards ray to find out whether the origin is inside a medium (SampleBatchJob.cs:194-303, 450-524).  This
// kernel therefore follows the reference's shape — one thread per pixel, samples in order, collect every hit, sort,
// run the volume bookkeeping — with the same arithmetic and the same accumulation order as sample_simple, so its
// outputs equal the CPU oracle's bit for bit.  It is the functional path for such worlds, not a tuned one.
#pragma once

#include "sample_kernels.cuh"
#include "media.cuh"

namespace rtbk {

constexpr uint32_t kLaneSamples = 8;  // samples a lane traces back to back before the warp accumulates (Philox mode)

struct VolumeSample {            // what one camera path hands to the pixel's accumulators (SampleBatchJob.cs:136-157)
  bool ok;                       // Sample() returned true (the path reached the sky within TraceDepth)
  f3 color, normal, albedo;
  float events;                  // randomEventsLocalAcc
  uint32_t rays;                 // bounce-loop iterations (Diagnostics.RayCount)
};

// SampleBatchJob.Sample (:166-401) with the volume bookkeeping, for the path that starts with `ray`.
template <bool COUNTERS, bool WHITE>
__device__ __noinline__ VolumeSample trace_volume_sample(const BatchArgs& a, const SceneView<false>& sv, PathRay ray, RayClock clk, uint32_t index,
                                                        uint32_t s, WhiteNoise& white, WorkCounters& wc) {
  const rtb_batch_params& p = a.p;
  const SceneDesc& sd = a.scene;
  const bool exact = p.trace_depth <= kSimpleStack;
  f3 att[kSimpleStack], emi[kSimpleStack];
  RayHits hits;
  VolumeSample out;
  out.rays = 0;
  f3 throughput = um::mk(1.0f), radiance = um::mk(0.0f);
  f3 s_normal = um::mk(0.0f), s_albedo = um::mk(0.0f);
  bool first_non_specular = false;
  float events_acc = 0, pow2depth = 1;
  int depth = 0, entries = 0;
  int current_volume = -1;                          // currentProbabilisticVolumeMaterial (material index, -1 = null)
  for (; depth < p.trace_depth; depth++) {
    // the iteration's hit search + volume bookkeeping (media.cuh), on the reference's own list: every hit, sorted
    const MediaStep st = media_step<false, COUNTERS, false, WHITE>(sv, sd, ray.o, ray.d, clk, current_volume, index, s, (uint32_t)depth,
                                                                    p.seed, white, wc, hits);
    current_volume = st.current_volume;
    float events = st.events;                       // rng.RandomEvents of this iteration
    out.rays++;

    bool scattered = false;
    if (st.hit) {
      const float rec_t = st.t;
      const f3 rec_n = st.n;
      const int rec_slot = st.slot;
      const uint32_t mi = st.material;
      const bool medium_hit = st.medium_hit;
      const float4* mp = reinterpret_cast<const float4*>(sd.materials + mi);
      float4 m0 = __ldg(mp), m1 = __ldg(mp + 1), m2 = __ldg(mp + 2), m3 = __ldg(mp + 3);
      if (__float_as_uint(m3.w) != 0u) {
        // HitRecord.TexCoords: the hit triangle's, or 0 (other entities, and the record made inside a medium)
        const float4 prim = medium_hit ? make_float4(0, 0, 0, 1) : sv.sphere(rec_slot);
        resolve_textures(sv, sd, mi, prim, ray.o, ray.d, m0, m1, m2, m3);
      }
      const f3 N = rec_n;
      const f3 P = um::mad(ray.d, rec_t, ray.o);
      ScatterResult sc;
      if (__float_as_uint(m0.w) == RTB_MATERIAL_PROBABILISTIC_VOLUME) {
        // Material.cs:163-168: isotropic; new Ray(rec.Point, direction) has Time = 0
        float rx, ry;
        if (WHITE) { rx = white.next_float(); ry = white.next_float(); }
        else { const uint4 r = philox4x32_10(index, s, (uint32_t)depth, 0u, p.seed, kPhiloxKey1); rx = u2f(r.x); ry = u2f(r.y); }
        float sn, cs;
        unit_angle_sincos(ry, &sn, &cs);
        sc.dir = random_direction(rx, sn, cs);
        sc.reflectance = um::mk(m0.x, m0.y, m0.z);
        sc.random_events = events + 2.0f;
        clk.value = 0.0f;
      } else {
        sc = WHITE ? scatter_white(m0, m1, m2, m3, ray.d, N, white, events)
                   : scatter(m0, m1, m2, m3, ray.d, N, index, s, (uint32_t)depth, p.seed, events);
        if (COUNTERS) { if (__float_as_uint(m0.w) == RTB_MATERIAL_DIELECTRIC) wc.shade_dielectric++; else wc.shade_standard++; }
      }
      events = sc.random_events;                   // RandomEvents after Scatter (it started from this iteration's count)
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
      events_acc += um::div(events, pow2depth);
      events = 0;
      const f3 off_n = um::dot(sc.dir, N) >= 0 ? N : -N;
      ray.o = um::mad(off_n, 0.001f, P);
      ray.d = sc.dir;
      scattered = true;
    }

    if (!scattered) {                               // no hit (or every hit passed through / dropped): the sky ends the path
      const f3 sky = sky_color(p.environment, sd, ray.d);
      if (exact) { emi[entries] = sky; att[entries] = um::mk(1.0f); entries++; }
      radiance = um::mad(throughput, sky, radiance);
      events_acc += um::div(events, pow2depth);
      if (!first_non_specular) { s_albedo = sky; s_normal = -ray.d; }
      break;
    }
    pow2depth *= 2.0f;
  }
  out.ok = depth != p.trace_depth;
  f3 c = radiance;
  if (out.ok && exact) {
    c = um::mk(0.0f);
    for (int e = entries; e-- > 0;) { c = c * att[e]; c = c + emi[e]; }
  }
  out.color = c;
  out.normal = s_normal;
  out.albedo = s_albedo;
  out.events = events_acc;
  return out;
}

// Philox mode: one WARP per pixel — the lanes trace 32 consecutive samples of the pixel (paths of one pixel cost about the
// same, paths of neighbouring pixels do not), then the warp adds the 32 results in sample order, every lane the same sums, so
// the accumulation order is the reference's.  White-noise mode (one sequential stream per pixel): one THREAD per pixel.
template <bool COUNTERS, bool WHITE>
#ifndef RTB_VOLUME_MIN_BLOCKS
#define RTB_VOLUME_MIN_BLOCKS 4   // <= 128 registers: the kernel is instruction-fetch bound, occupancy hides it (1462 -> 1249 ms on the fog Cornell box)
#endif
__global__ void __launch_bounds__(128, RTB_VOLUME_MIN_BLOCKS) sample_volumes(const __grid_constant__ BatchArgs a) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t k = WHITE ? blockIdx.x * blockDim.x + threadIdx.x : (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
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
  WhiteNoise white{};
  if (WHITE) white.init((p.seed * 0x8C4CA03Fu) ^ (index * 0x7383ED49u));      // SampleBatchJob.cs:91

  auto add = [&](bool ok, f3 c, f3 nrm, f3 alb, float ev, uint32_t s) {        // SampleBatchJob.cs:139-157
    if (ok) {
      color_acc = color_acc + c;
      normal_acc = normal_acc + nrm;
      albedo_acc = albedo_acc + alb;
      weight_acc += ev;
      sample_count++;
    }
    if (s == 0) { fb_normal = nrm; fb_albedo = alb; }
  };

  if (WHITE) {
    for (uint32_t s = 0; s < n; s++) {
      RayClock clk{index, s, p.seed, 0.0f, true};
      const PathRay ray = camera_ray_white(p, cx, cy, white, &clk.value);
      const VolumeSample r = trace_volume_sample<COUNTERS, WHITE>(a, sv, ray, clk, index, s, white, wc);
      rays += r.rays;
      add(r.ok, r.color, r.normal, r.albedo, r.events, s);
    }
  } else {
    // chunks of 32 * kLaneSamples samples: lane l traces samples base + l, base + l + 32, ... back to back (no lane waits for
    // another between them, so path-length differences average out), then the warp adds the chunk in sample order
    for (uint32_t base = 0; base < n; base += 32u * kLaneSamples) {
      VolumeSample r[kLaneSamples];
#pragma unroll 1
      for (uint32_t j = 0; j < kLaneSamples; j++) {
        const uint32_t s = base + 32u * j + lane;
        r[j].ok = false; r[j].rays = 0; r[j].events = 0;
        r[j].color = r[j].normal = r[j].albedo = um::mk(0.0f);
        if (s < n) {
          RayClock clk{index, s, p.seed, 0.0f, false};
          clk.value = clk.time();
          clk.known = true;
          const PathRay ray = camera_ray(p, cx, cy, index, s);
          r[j] = trace_volume_sample<COUNTERS, WHITE>(a, sv, ray, clk, index, s, white, wc);
          rays += r[j].rays;
        }
      }
      __syncwarp();
#pragma unroll 1
      for (uint32_t j = 0; j < kLaneSamples; j++) {
        if (base + 32u * j >= n) break;
        const uint32_t m = n - (base + 32u * j) < 32u ? n - (base + 32u * j) : 32u;
        const VolumeSample x = r[j];
        for (uint32_t l = 0; l < m; l++) {
          const bool ok = __shfl_sync(0xffffffffu, (int)x.ok, l) != 0;
          const f3 c = um::mk(__shfl_sync(0xffffffffu, x.color.x, l), __shfl_sync(0xffffffffu, x.color.y, l), __shfl_sync(0xffffffffu, x.color.z, l));
          const f3 nr = um::mk(__shfl_sync(0xffffffffu, x.normal.x, l), __shfl_sync(0xffffffffu, x.normal.y, l), __shfl_sync(0xffffffffu, x.normal.z, l));
          const f3 al = um::mk(__shfl_sync(0xffffffffu, x.albedo.x, l), __shfl_sync(0xffffffffu, x.albedo.y, l), __shfl_sync(0xffffffffu, x.albedo.z, l));
          const float ev = __shfl_sync(0xffffffffu, x.events, l);
          add(ok, c, nr, al, ev, base + 32u * j + l);
        }
      }
    }
    // per-lane tallies -> the pixel's
    for (int o = 16; o > 0; o >>= 1) {
      rays += __shfl_xor_sync(0xffffffffu, rays, o);
      wc.node_tests += __shfl_xor_sync(0xffffffffu, wc.node_tests, o);
      wc.sphere_tests += __shfl_xor_sync(0xffffffffu, wc.sphere_tests, o);
      wc.shade_standard += __shfl_xor_sync(0xffffffffu, wc.shade_standard, o);
      wc.shade_dielectric += __shfl_xor_sync(0xffffffffu, wc.shade_dielectric, o);
    }
    if (lane != 0) return;
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
