// aux_kernels.cuh — the two O(pixels) jobs that sit right after the sample job in the
// reference's per-batch pipeline (Raytracer.cs:745-754, :844), as HBM-bound device kernels:
//   combine_kernel         CombineJob.Execute        (Runtime/Jobs/CombineJob.cs:29-71)
//   reduce_metrics_kernel  ReduceMetricsJob.Execute  (Runtime/Jobs/ReduceMetricsJob.cs:22-45)
//   finalize_kernel        FinalizeTexturesJob.Execute (Runtime/Jobs/FinalizeTexturesJob.cs:23-55)
#pragma once

#include "kernel_common.cuh"

namespace rtbk {

// CombineJob: colour / sample count with the interlace look-around, NaN -> 0 (or debug
// colours), albedo / max(n,1) (optionally clamped to 1), normalizesafe(normal / max(n,1)).
__global__ void combine_kernel(int width, int height, int debug_mode, int ldr_albedo,
                               const float4* __restrict__ color, const float* __restrict__ normal,
                               const float* __restrict__ albedo, float* __restrict__ out_color,
                               float* __restrict__ out_normal, float* __restrict__ out_albedo) {
  const int n = width * height;
  for (int index = blockIdx.x * blockDim.x + threadIdx.x; index < n; index += gridDim.x * blockDim.x) {
    float4 c = color[index];
    int real = (int)c.w;
    f3 final_color;
    if (!debug_mode) {
      if (real == 0) {
        int tentative = index;
        while (real == 0 && (tentative -= width) >= 0) {   // look-around (interlaced buffer)
          c = color[tentative];
          real = (int)c.w;
        }
      }
    }
    const bool any_nan = um::isnan(c.x) || um::isnan(c.y) || um::isnan(c.z) || um::isnan(c.w);
    if (real == 0) final_color = debug_mode ? um::mk(1.0f, 0.0f, 1.0f) : um::mk(0.0f);
    else if (any_nan) final_color = debug_mode ? um::mk(0.0f, 1.0f, 1.0f) : um::mk(0.0f);
    else final_color = um::mk(c.x, c.y, c.z) / (float)real;
    const float denom = (float)max(real, 1);
    if (out_color) { out_color[3 * (size_t)index] = final_color.x; out_color[3 * (size_t)index + 1] = final_color.y; out_color[3 * (size_t)index + 2] = final_color.z; }
    if (out_albedo) {
      f3 al = v3(albedo + 3 * (size_t)index) / denom;
      if (ldr_albedo) al = um::min(al, um::mk(1.0f));
      out_albedo[3 * (size_t)index] = al.x; out_albedo[3 * (size_t)index + 1] = al.y; out_albedo[3 * (size_t)index + 2] = al.z;
    }
    if (out_normal) {
      f3 nn = v3(normal + 3 * (size_t)index) / denom;
      float len2 = um::dot(nn, nn);               // math.normalizesafe: 0 when |v|^2 <= FLT_MIN_NORMAL
      nn = len2 > 1.175494351e-38f ? nn * um::rsqrt(len2) : um::mk(0.0f);
      out_normal[3 * (size_t)index] = nn.x; out_normal[3 * (size_t)index + 1] = nn.y; out_normal[3 * (size_t)index + 2] = nn.z;
    }
  }
}

// MathExtensions.LinearToGamma (Util/MathExtensions.cs:17-21): max(1.055 * pow(max(v, 0), 0.416666667) - 0.055, 0),
// then FinalizeTexturesJob's saturate(...) * 255 truncated to a byte.  pow through the shared exp2/log of umath.h.
__device__ __noinline__ uint32_t gamma_byte_exact(float v) {      // v = max(v, 0) > 0
  const float p = um::pow_pos(v, 0.416666667f);
  const float g = um::max(1.055f * p - 0.055f, 0.0f);
  return (uint32_t)(um::saturate(g) * 255.0f);
}
// The same byte without the polynomial pow wherever that is safe: the job keeps only floor(255 g), so an estimate x of 255 g
// decides the byte whenever it lies further from an integer than it can be wrong.  MUFU lg2 / ex2 (2 ulp each) put x within
// 4e-4 of the true product for every v that does not round to byte 0 anyway, the polynomial pow is within 2e-4 of it:
// kGammaGuard = 1e-3.  (Measured: the dense sweep of `test_finalize_fast_path_*` — every float k / 2^24 of [0, 1] and a log
// sweep — still matches the exact path byte for byte with a guard of 5e-5: a factor of 20.)  0.2 % of the values (and white:
// 255 * 0.99999994 is byte 254) take the exact path.  Nine polynomial pows per pixel kept this kernel at 0.43 of the HBM
// roofline in round 1 (2790 GB/s); 4665 GB/s = 0.71 now (guard 2e-3: 4315; 5e-5: 4978).
#ifndef RTB_GAMMA_GUARD
#define RTB_GAMMA_GUARD 1.0e-3f
#endif
constexpr float kGammaGuard = RTB_GAMMA_GUARD;
__device__ __forceinline__ float mufu_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float mufu_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ uint32_t gamma_byte(float v) {
  v = um::max(v, 0.0f);                                            // math.max(v, 0): NaN -> 0
  if (!(v > 0.0f)) return 0u;                                      // pow is not evaluated: g = max(-0.055, 0) = 0
  if (!(v < 3.0e38f)) return gamma_byte_exact(v);
  // x ~ 255 * (1.055 pow(v, 1 / 2.4) - 0.055), clamped to [-1.5, 256.5] (everything outside is byte 0 or 255 anyway, and the
  // ends sit half-way between integers: never "too close to call").  Two MUFU operations per value and no conversion: the nearest integer n
  // comes from adding 1.5 * 2^23 (its low mantissa bits ARE n), the byte is floor(x) = n - (x < n).  lg2, ex2, float <-> int
  // conversions and floor all issue on the quarter-rate XU pipe: with four of them per value the kernel was XU-bound
  // (64 us per 4K frame against 61 us of HBM time).
  float x = __fmaf_rn(269.025f, mufu_ex2(mufu_lg2(v) * 0.416666667f), -14.025f);
  x = fminf(fmaxf(x, -1.5f), 256.5f);
  const float y = x + 12582912.0f;
  const int n = __float_as_int(y) - 0x4B400000;
  const float d = x - (y - 12582912.0f);
  if (fabsf(d) > kGammaGuard) return (uint32_t)min(max(n - (d < 0.0f ? 1 : 0), 0), 255);
  return gamma_byte_exact(v);
}
__device__ __forceinline__ uint32_t rgba32(f3 c) {      // PixelFormats.RGBA32: r, g, b, a = 255 in memory order
  return gamma_byte(c.x) | (gamma_byte(c.y) << 8) | (gamma_byte(c.z) << 16) | 0xff000000u;
}

// FinalizeTexturesJob: three float3 images -> three RGBA32 images (colour, normal * 0.5 + 0.5, albedo).
// HBM-bound: 36 B read + 12 B written per pixel.  Any in/out pair may be NULL.
__global__ void finalize_kernel(int n, const float* __restrict__ color, const float* __restrict__ normal,
                                const float* __restrict__ albedo, uint32_t* __restrict__ out_color,
                                uint32_t* __restrict__ out_normal, uint32_t* __restrict__ out_albedo, int first) {
  for (int i = first + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (out_color) out_color[i] = rgba32(v3(color + 3 * (size_t)i));
    if (out_normal) out_normal[i] = rgba32(v3(normal + 3 * (size_t)i) * 0.5f + um::mk(0.5f));
    if (out_albedo) out_albedo[i] = rgba32(v3(albedo + 3 * (size_t)i));
  }
}
// The same, four pixels per thread: three float4 loads and one uint4 store per image (16-byte aligned arrays; the last
// n % 4 pixels go through finalize_kernel).
__device__ __forceinline__ uint4 rgba32x4(const float4* __restrict__ src, size_t quad, bool as_normal) {
  const float4 a = __ldg(src + 3 * quad), b = __ldg(src + 3 * quad + 1), c = __ldg(src + 3 * quad + 2);
  const float v[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    f3 px = um::mk(v[3 * k], v[3 * k + 1], v[3 * k + 2]);
    if (as_normal) px = px * 0.5f + um::mk(0.5f);
    o[k] = rgba32(px);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}
__global__ void __launch_bounds__(256) finalize_kernel_x4(int n_quads, const float4* __restrict__ color, const float4* __restrict__ normal,
                                                         const float4* __restrict__ albedo, uint4* __restrict__ out_color,
                                                         uint4* __restrict__ out_normal, uint4* __restrict__ out_albedo) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += gridDim.x * blockDim.x) {
    if (out_color) out_color[q] = rgba32x4(color, (size_t)q, false);
    if (out_normal) out_normal[q] = rgba32x4(normal, (size_t)q, true);
    if (out_albedo) out_albedo[q] = rgba32x4(albedo, (size_t)q, false);
  }
}

struct MetricsAcc {
  long long rays, samples;
  float w_min, w_max;
  int s_min, s_max;
};

// ReduceMetricsJob.  The reference is a serial loop; min/max/integer sums are order-free, so
// a tree reduction returns the same values.  `partial` holds one MetricsAcc per block.
__global__ void reduce_metrics_kernel(int n, const rtb_diagnostics* __restrict__ diag, const float4* __restrict__ color,
                                      const float* __restrict__ weight, MetricsAcc* __restrict__ partial) {
  MetricsAcc m{0, 0, um::INF, -um::INF, 0x7fffffff, (int)0x80000000};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (diag) m.rays += (int)diag[i].ray_count;
    const int sc = (int)color[i].w;
    m.samples += sc;
    const float w = um::div(weight[i], (float)sc);
    m.w_min = um::min(m.w_min, w);
    m.w_max = um::max(m.w_max, w);
    m.s_min = min(m.s_min, sc);
    m.s_max = max(m.s_max, sc);
  }
  __shared__ MetricsAcc sh[32];
  for (int o = 16; o > 0; o >>= 1) {
    m.rays += __shfl_down_sync(0xffffffffu, m.rays, o);
    m.samples += __shfl_down_sync(0xffffffffu, m.samples, o);
    m.w_min = um::min(m.w_min, __shfl_down_sync(0xffffffffu, m.w_min, o));
    m.w_max = um::max(m.w_max, __shfl_down_sync(0xffffffffu, m.w_max, o));
    m.s_min = min(m.s_min, __shfl_down_sync(0xffffffffu, m.s_min, o));
    m.s_max = max(m.s_max, __shfl_down_sync(0xffffffffu, m.s_max, o));
  }
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    MetricsAcc t = sh[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) {
      t.rays += sh[w].rays;
      t.samples += sh[w].samples;
      t.w_min = um::min(t.w_min, sh[w].w_min);
      t.w_max = um::max(t.w_max, sh[w].w_max);
      t.s_min = min(t.s_min, sh[w].s_min);
      t.s_max = max(t.s_max, sh[w].s_max);
    }
    partial[blockIdx.x] = t;
  }
}

// Row costs for the multi-GPU tile balancer (plugin.cu: rtb_multi): per image row, the instruction-weighted work of an
// instrumented probe batch — rays, executed box tests and executed entity tests per pixel (the reference's FULL_DIAGNOSTICS
// fields, Raytracer.cs:56-60), weighted by what each costs the walk.  One block per row of [row_begin, row_end).
__global__ void __launch_bounds__(256) row_cost_kernel(const rtb_diagnostics* __restrict__ diag, int width, int row_begin,
                                                       float* __restrict__ row_cost) {
  const int row = row_begin + (int)blockIdx.x;
  float s = 0.0f;
  for (int x = threadIdx.x; x < width; x += blockDim.x) {
    const rtb_diagnostics d = diag[(size_t)row * (size_t)width + (size_t)x];
    s += 450.0f * d.ray_count + 35.0f * d.bounds_hit_count + 40.0f * d.candidate_count;
  }
  __shared__ float sh[8];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sh[w];
    row_cost[row] = t;
  }
}

// FP32-pipe roofline microbenchmark: 16 independent FMA chains per thread, register operands.
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters, float b, float c) {
  float a[16];
#pragma unroll
  for (int k = 0; k < 16; k++) a[k] = (float)(threadIdx.x + k) * 1e-3f;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = __fmaf_rn(a[k], b, c);
  }
  float s = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) s += a[k];
  if (s == 123456.789f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // keeps the chains alive
}

}  // namespace rtbk
