/*
 * umath.h — restatement of the slice of com.unity.mathematics 1.2.5 that the
 * reference's sample path calls (Packages/manifest.json:7; the package source is
 * NOT in /root/reference, see SURVEY.md §8c assumption A1).
 *
 * Why this header is shared by the CUDA kernel and the CPU oracle:
 * parity at 1e-4 per pixel is a bit-level problem (one flipped scatter decision
 * moves a pixel by ~4e-3), so both sides must evaluate the *library* functions
 * (dot/normalize/reflect/sincos/log/...) with identical roundings.  Everything
 * here is built from IEEE-754 single-precision +,-,*,/,sqrt and fused
 * multiply-add, all of which are correctly rounded on x86-64 (SSE/FMA3) and on
 * sm_100a (FADD/FMUL/FFMA, div.rn, sqrt.rn), so a function in this file returns
 * the same bits on both.  Build rules that make that true:
 *     host  : -ffp-contract=off   (gcc must not invent FMAs; explicit fmaf only)
 *     device: -fmad=false         (ptxas must not invent FMAs; --prec-div/sqrt default true)
 *
 * The reference compiles with Burst FloatMode.Fast / FloatPrecision.Medium
 * (SampleBatchJob.cs:16), i.e. contraction and 3.5-ULP transcendentals are
 * allowed and unspecified; we fix ONE legal evaluation (explicit FMA points,
 * polynomial sincos/log with <=2 ULP error) and use it everywhere.
 *
 * The reference's OWN code (SampleBatchJob, Material, HitTests, View, ...) is
 * NOT in here: the kernel and the oracle each restate it separately.
 */
#ifndef RTB_UMATH_H
#define RTB_UMATH_H

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define UM_HD __host__ __device__ __forceinline__
#else
#define UM_HD inline __attribute__((always_inline))
#endif

namespace um {

static constexpr float PI = 3.14159265f;           /* math.PI (float) */
static constexpr float TWO_PI = 6.28318531f;       /* 2 * PI, exact doubling */
static constexpr float INF = __builtin_huge_valf();

/* ---- bit casts (math.asfloat / math.asuint) ---- */
UM_HD float asfloat(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
UM_HD uint32_t asuint(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}

/* ---- scalar helpers ---- */
UM_HD float fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return __builtin_fmaf(a, b, c);
#endif
}
/* RTB_FAST_MATH (device only; csrc/fast_kernels.cu): the opt-in build in the spirit of the reference's own
 * [BurstCompile(FloatPrecision.Medium, FloatMode.Fast)] (SampleBatchJob.cs:16) — hardware approximations for sqrt,
 * division, sincos and log (MUFU.RSQ/RCP/SIN/COS/LG2, ~2 ulp), FMA contraction everywhere (-fmad=true).  Its images
 * equal the parity build's statistically, not bitwise; tools/fast_math_report.py measures by how much. */
UM_HD float sqrt(float x) {
#if defined(__CUDA_ARCH__) && defined(RTB_FAST_MATH)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#elif defined(__CUDA_ARCH__)
  return __fsqrt_rn(x);
#else
  return __builtin_sqrtf(x);
#endif
}
UM_HD float div(float a, float b) {
#if defined(__CUDA_ARCH__) && defined(RTB_FAST_MATH)
  return __fdividef(a, b);
#elif defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
#if defined(__CUDA_ARCH__) && defined(RTB_FAST_MATH)
UM_HD float rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
UM_HD float rsqrt(float x) { return rsqrtf(x); }
#else
UM_HD float rcp(float x) { return div(1.0f, x); }          /* math.rcp = 1/x */
UM_HD float rsqrt(float x) { return div(1.0f, um::sqrt(x)); } /* math.rsqrt = 1/sqrt(x) */
#endif
UM_HD bool isnan(float x) { return x != x; }
UM_HD bool isinf(float x) { return (asuint(x) & 0x7fffffffu) == 0x7f800000u; }
UM_HD float abs(float x) { return asfloat(asuint(x) & 0x7fffffffu); }
/* math.min/max: "isnan(y) || x < y ? x : y" — returns the non-NaN operand when exactly one is NaN */
UM_HD float min(float x, float y) { return (y != y || x < y) ? x : y; }
UM_HD float max(float x, float y) { return (y != y || x > y) ? x : y; }
UM_HD float clamp(float x, float a, float b) { return um::max(a, um::min(b, x)); }
UM_HD float saturate(float x) { return um::clamp(x, 0.0f, 1.0f); } /* saturate(NaN) == 1 */
UM_HD float lerp(float a, float b, float s) { return um::fma(s, b - a, a); }      /* a + s*(b-a) */
UM_HD float unlerp(float a, float b, float x) { return um::div(x - a, b - a); }
UM_HD float select(float a, float b, bool c) { return c ? b : a; }
/* math.round: MathF.Round = round-half-to-even */
UM_HD float round(float x) {
#if defined(__CUDA_ARCH__)
  return rintf(x);
#else
  return __builtin_rintf(x);
#endif
}

/* ---- float3 ---- */
struct f3 { float x, y, z; };
UM_HD f3 mk(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
UM_HD f3 mk(float s) { return mk(s, s, s); }
UM_HD f3 operator+(f3 a, f3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
UM_HD f3 operator-(f3 a, f3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
UM_HD f3 operator-(f3 a) { return mk(-a.x, -a.y, -a.z); }
UM_HD f3 operator*(f3 a, f3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
UM_HD f3 operator*(f3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
UM_HD f3 operator*(float s, f3 a) { return mk(a.x * s, a.y * s, a.z * s); }
UM_HD f3 operator/(f3 a, float s) { return mk(div(a.x, s), div(a.y, s), div(a.z, s)); }
/* math.mad(a, b, c) = a*b + c, fused */
UM_HD f3 mad(f3 a, float s, f3 c) { return mk(um::fma(a.x, s, c.x), um::fma(a.y, s, c.y), um::fma(a.z, s, c.z)); }
UM_HD f3 mad(f3 a, f3 b, f3 c) { return mk(um::fma(a.x, b.x, c.x), um::fma(a.y, b.y, c.y), um::fma(a.z, b.z, c.z)); }
/* dot = x0*y0 + x1*y1 + x2*y2, left to right, contracted */
UM_HD float dot(f3 a, f3 b) { return um::fma(a.z, b.z, um::fma(a.y, b.y, a.x * b.x)); }
UM_HD f3 cross(f3 a, f3 b) {
  return mk(um::fma(a.y, b.z, -(a.z * b.y)), um::fma(a.z, b.x, -(a.x * b.z)), um::fma(a.x, b.y, -(a.y * b.x)));
}
UM_HD f3 normalize(f3 v) { return v * um::rsqrt(um::dot(v, v)); } /* rsqrt(dot(v,v)) * v */
/* reflect(i, n) = i - 2*n*dot(i,n); the doubling is exact so (2n)*dt == n*(2dt) */
UM_HD f3 reflect(f3 i, f3 n) { float k = -2.0f * um::dot(i, n); return mad(n, k, i); }
UM_HD f3 lerp(f3 a, f3 b, float s) { return mad(b - a, s, a); }
UM_HD f3 min(f3 a, f3 b) { return mk(um::min(a.x, b.x), um::min(a.y, b.y), um::min(a.z, b.z)); }
UM_HD f3 max(f3 a, f3 b) { return mk(um::max(a.x, b.x), um::max(a.y, b.y), um::max(a.z, b.z)); }
UM_HD float cmax(f3 a) { return um::max(um::max(a.x, a.y), a.z); }
UM_HD float cmin(f3 a) { return um::min(um::min(a.x, a.y), a.z); }
UM_HD f3 rcp(f3 a) { return mk(um::rcp(a.x), um::rcp(a.y), um::rcp(a.z)); }
UM_HD float comp(f3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
/* mul(float3x3(c0, c1, c2), v) = c0*v.x + c1*v.y + c2*v.z (column-constructed matrix) */
UM_HD f3 mul_cols(f3 c0, f3 c1, f3 c2, f3 v) { return mad(c2, v.z, mad(c1, v.y, c0 * v.x)); }

/* ---- quaternion / RigidTransform (Entity.cs:52,65,95-97) ---- */
struct quat { float x, y, z, w; };
struct rigid { quat rot; f3 pos; };
UM_HD quat quat_identity() { quat q; q.x = 0; q.y = 0; q.z = 0; q.w = 1; return q; }
/* rotate(q, v): t = 2*cross(q.xyz, v); v + q.w*t + cross(q.xyz, t) */
UM_HD f3 rotate(quat q, f3 v) {
  f3 qv = mk(q.x, q.y, q.z);
  f3 t = 2.0f * cross(qv, v);
  return v + q.w * t + cross(qv, t);
}
UM_HD f3 transform(rigid a, f3 p) { return rotate(a.rot, p) + a.pos; }
UM_HD quat conjugate(quat q) { quat r; r.x = -q.x; r.y = -q.y; r.z = -q.z; r.w = q.w; return r; }
/* inverse(RigidTransform): invRot = inverse(rot) = conj(rot)/|rot|^2 ; pos = rotate(invRot, -pos) */
UM_HD rigid inverse(rigid a) {
  float n2 = um::fma(a.rot.w, a.rot.w, um::fma(a.rot.z, a.rot.z, um::fma(a.rot.y, a.rot.y, a.rot.x * a.rot.x)));
  float r = um::rcp(n2);
  quat c = conjugate(a.rot);
  rigid o; o.rot.x = r * c.x; o.rot.y = r * c.y; o.rot.z = r * c.z; o.rot.w = r * c.w;
  o.pos = rotate(o.rot, -a.pos);
  return o;
}

/* ---- transcendentals (deterministic, identical bits on host and device) ---- */

/* sincos(theta), theta in [0, ~8].  Quadrant reduction with a 2-term Cody-Waite
 * constant, then the Cephes single-precision minimax polynomials on [-pi/4, pi/4].
 * Max error vs correctly-rounded: < 2 ULP on [0, 2pi] (tests/test_umath.py pins this). */
UM_HD void sincos(float theta, float* s, float* c) {
#if defined(__CUDA_ARCH__) && defined(RTB_FAST_MATH)
  __sincosf(theta, s, c);
  return;
#endif
  const float TWO_OVER_PI = 0.636619772f;
  const float PIO2_HI = 1.57079637050628662109375f; /* float(pi/2) = 0x3FC90FDB */
  const float PIO2_LO = -4.371139000186241e-08f;    /* pi/2 - PIO2_HI */
  int k = (int)um::fma(theta, TWO_OVER_PI, 0.5f);
  float fk = (float)k;
  float r = um::fma(-fk, PIO2_HI, theta);
  r = um::fma(-fk, PIO2_LO, r);
  float z = r * r;
  /* sin(r) = r + r*z*(S1 + z*(S2 + z*S3)) */
  float ps = um::fma(z, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = um::fma(z, ps, -1.6666654611e-1f);
  float sr = um::fma(r * z, ps, r);
  /* cos(r) = 1 - z/2 + z*z*(C1 + z*(C2 + z*C3)) */
  float pc = um::fma(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = um::fma(z, pc, 4.166664568298827e-2f);
  float cr = um::fma(z * z, pc, um::fma(z, -0.5f, 1.0f));
  switch (k & 3) {
    case 0: *s = sr; *c = cr; break;
    case 1: *s = cr; *c = -sr; break;
    case 2: *s = -sr; *c = -cr; break;
    default: *s = -cr; *c = sr; break;
  }
}
UM_HD float tan(float x) { float s, c; um::sincos(x, &s, &c); return um::div(s, c); }

/* log(x) for normal positive x (Cephes logf layout).  Used by RoughnessToAlpha
 * (Microfacet.cs:71-80) on [1e-3, 1].  < 2 ULP. */
UM_HD float log(float x) {
#if defined(__CUDA_ARCH__) && defined(RTB_FAST_MATH)
  return __logf(x);
#endif
  uint32_t u = asuint(x);
  int e = (int)(u >> 23) - 126;                     /* x = m * 2^e, m in [0.5, 1) */
  float m = asfloat((u & 0x007fffffu) | 0x3f000000u);
  if (m < 0.707106781186547524f) { e -= 1; m = m + m - 1.0f; } else { m = m - 1.0f; }
  float z = m * m;
  float y = 7.0376836292e-2f;
  y = um::fma(y, m, -1.1514610310e-1f);
  y = um::fma(y, m, 1.1676998740e-1f);
  y = um::fma(y, m, -1.2420140846e-1f);
  y = um::fma(y, m, 1.4249322787e-1f);
  y = um::fma(y, m, -1.6668057665e-1f);
  y = um::fma(y, m, 2.0000714765e-1f);
  y = um::fma(y, m, -2.4999993993e-1f);
  y = um::fma(y, m, 3.3333331174e-1f);
  y = y * m * z;
  float fe = (float)e;
  y = um::fma(fe, -2.12194440e-4f, y);
  y = um::fma(z, -0.5f, y);
  return um::fma(fe, 0.693359375f, m + y);
}

/* log of a uniform draw in [0, 1): log(0) = -inf like math.log (Material.ProbabilisticHit, Material.cs:56) */
UM_HD float log_unit(float x) { return x == 0.0f ? -INF : um::log(x); }

/* pow(x, 2) and pow(x, 5) as the reference calls them (Material.cs:82,216).  A
 * correctly-rounded powf(x,2) IS x*x; x^5 is evaluated as (x^2)^2 * x (<= 1.5 ULP,
 * sign-correct for negative x, which happens for cosine > 1 in Dielectric). */
UM_HD float pow2(float x) { return x * x; }
UM_HD float pow5(float x) { float x2 = x * x; return (x2 * x2) * x; }

/* exp2/log2 pair for LinearToGamma's pow(v, 0.41666) (MathExtensions.cs:17-21). */
UM_HD float exp2_poly(float x) {
  /* x in [-126, 127]: split integer/fraction, 2^f with a degree-6 polynomial on [-0.5, 0.5] */
  float fi = um::round(x);
  float f = x - fi;
  float p = 1.535336188319500e-4f;
  p = um::fma(p, f, 1.339887440266574e-3f);
  p = um::fma(p, f, 9.618437357674640e-3f);
  p = um::fma(p, f, 5.550332471162809e-2f);
  p = um::fma(p, f, 2.402264791363012e-1f);
  p = um::fma(p, f, 6.931472028550421e-1f);
  p = um::fma(p, f, 1.0f);
  int i = (int)fi;
  return asfloat((uint32_t)(i + 127) << 23) * p;
}
UM_HD float pow_pos(float x, float y) {           /* x > 0 */
  const float LOG2E = 1.44269504088896341f;
  return exp2_poly(um::log(x) * LOG2E * y);
}

/* half -> float (IEEE binary16, exact), as (float)half does in Unity.Mathematics */
UM_HD float half_to_float(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
  if (e == 0) {
    if (m == 0) return asfloat(sign);
    int shift = 0;                                   /* subnormal: normalise */
    while (!(m & 0x400u)) { m <<= 1; shift++; }
    m &= 0x3ffu;
    return asfloat(sign | ((uint32_t)(113 - shift) << 23) | (m << 13));
  }
  if (e == 31) return asfloat(sign | 0x7f800000u | (m << 13));
  return asfloat(sign | ((e + 112u) << 23) | (m << 13));
}

} /* namespace um */

#endif /* RTB_UMATH_H */
