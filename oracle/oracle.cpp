// oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT.
//
// CPU restatement of the reference's per-pixel sample job, structured like the C# it
// follows (collect all BVH candidates -> intersect all -> sort -> take the nearest;
// emission/attenuation stacks unwound tail->head; failed samples dropped).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// it; nothing under raytracing-in-one-weekend_b200/ includes, links or calls it.
//
// PARITY UNPINNED: the reference has no tests, golden vectors or fixtures for this path
//  and its Unity/Burst C# cannot be compiled or run in this image, so
// this restatement is pinned only by our own known-answer tests (tests/test_oracle_kat.py:
// Unity xorshift32 stream, Philox4x32-10 Random123 vectors, closed-form intersection and
// shading cases, scene/BVH counts) — not by outputs of the reference itself.
//
// Paths below are relative to /root/reference/RaytracingInOneWeekend/Assets/Scripts.
// Library math (Unity.Mathematics) comes from include/rtb/umath.h; everything else in
// this file restates the reference's own code.  Float contraction: FMAs appear only
// inside um:: functions and at the three places marked [FMA] (Burst FloatMode.Fast,
// SampleBatchJob.cs:16, permits any contraction; we fix one).
//
// Build: oracle/Makefile (strict: -O2 -ffp-contract=off; fast: -O3 unsafe-math for timing).

#include "rtb.h"
#include "rtb/umath.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

using um::f3;

namespace {

// ======================================================================================
// RNG
// ======================================================================================

// Unity.Mathematics.Random (package not vendored).
struct UnityRandom {
  uint32_t state;
  void init(uint32_t seed) { state = seed; next_state(); }
  uint32_t next_state() {
    uint32_t t = state;
    state ^= state << 13;
    state ^= state >> 17;
    state ^= state << 5;
    return t;
  }
};

// Philox4x32-10 (Salmon et al., SC'11), written from the paper's round function.
void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Random-draw slots.  The reference's white-noise stream is sequential per pixel
// (SampleBatchJob.cs:91), so data-dependent draws shift every later draw; the Philox
// replacement (BASELINE.json north_star) gives each draw site a fixed slot so a skipped
// draw does not move the others.
//   camera  (bounce word 0xFFFFFFFF): block 0 = {jitter.x, jitter.y, lens theta, lens radius}, block 1 = {time}
//   bounce d (bounce word d):         block 0 = {rough-normal / random-direction u, v, reflect-or-Schlick test, metal test}
//                                     block 1 = {diffuse direction u, v}
enum Slot {
  SLOT_JITTER = 0,       // 2 floats: block 0 [0,1]
  SLOT_LENS = 2,         // 2 floats: block 0 [2,3]
  SLOT_TIME = 4,         // 1 float : block 1 [0]
  SLOT_ROUGH = 0,        // 2 floats: block 0 [0,1]   (Material.cs:83 / :124)
  SLOT_REFLECT = 2,      // 1 float : block 0 [2]     (Material.cs:91 / :143)
  SLOT_METAL = 3,        // 1 float : block 0 [3]     (Material.cs:99)
  SLOT_DIFFUSE = 4,      // 2 floats: block 1 [0,1]   (Material.cs:107)
  SLOT_VOLUME = 8        // 1 float per Material.ProbabilisticHit call of the bounce: block 2 + k / 4 [k % 4] (Material.cs:56)
};
constexpr uint32_t BOUNCE_CAMERA = 0xFFFFFFFFu;
constexpr uint32_t PHILOX_KEY1 = 0x52544232u;  // "RTB2"

enum NoiseMode { NOISE_WHITE_XORSHIFT = 0, NOISE_PHILOX = 1 };

// RandomSource.cs:15-151 (White case).  Every draw site passes its slot; the xorshift
// mode ignores it and consumes the stream in call order exactly like the reference.
struct RandomSource {
  int mode = NOISE_WHITE_XORSHIFT;
  UnityRandom white{};
  uint32_t key[2] = {0, 0};
  uint32_t pixel = 0, sample = 0, bounce = 0;
  uint32_t cached_block_id = 0xFFFFFFFFu, cached_bounce = 0, cached_sample = 0;
  uint32_t cached[4] = {0, 0, 0, 0};
  float random_events = 0;  // RandomSource.cs:33-37

  static float to_float(uint32_t u) { return um::asfloat(0x3f800000u | (u >> 9)) - 1.0f; }

  uint32_t raw(int slot) {
    if (mode == NOISE_WHITE_XORSHIFT) return white.next_state();
    uint32_t block = (uint32_t)slot >> 2;
    if (block != cached_block_id || bounce != cached_bounce || sample != cached_sample) {
      uint32_t ctr[4] = {pixel, sample, bounce, block};
      philox4x32_10(ctr, key, cached);
      cached_block_id = block; cached_bounce = bounce; cached_sample = sample;
    }
    return cached[slot & 3];
  }
  float NextFloat(int slot) { return to_float(raw(slot)); }
  void NextFloat2(int slot, float* x, float* y) { *x = to_float(raw(slot)); *y = to_float(raw(slot + 1)); }

  // RandomSource.cs:40-61
  void InUnitDisk(float* x, float* y) {
    float theta = NextFloat(SLOT_LENS) * (2 * um::PI - 0) + 0;   // whiteNoise.NextFloat(0, 2 * PI)
    float radius = um::sqrt(NextFloat(SLOT_LENS + 1));
    float s, c;
    um::sincos(theta, &s, &c);
    *x = radius * c;
    *y = radius * s;
  }
  f3 OnCosineWeightedHemisphere(f3 normal, int slot);
  // RandomSource.cs:113-128
  f3 NextFloat3Direction(int slot) {
    float rx, ry;
    NextFloat2(slot, &rx, &ry);
    float z = rx * 2.0f - 1.0f;
    float r = um::sqrt(um::max(1.0f - z * z, 0.0f));
    float angle = ry * um::PI * 2.0f;
    float s, c;
    um::sincos(angle, &s, &c);
    return um::mk(c * r, s * r, z);
  }
};

// Tools.GetOrthonormalBasis (Util/Tools.cs:19-28)
void GetOrthonormalBasis(f3 normal, f3* tangent, f3* bitangent) {
  float s = normal.z >= 0 ? 1.0f : -1.0f;
  float a = um::div(-1.0f, s + normal.z);
  float b = normal.x * normal.y * a;
  *tangent = um::mk(1 + s * normal.x * normal.x * a, s * b, -s * normal.x);
  *bitangent = um::mk(b, s + normal.y * normal.y * a, -normal.y);
}
// Tools.TangentToWorldSpace (Util/Tools.cs:30-37)
f3 TangentToWorldSpace(f3 v, f3 normal) {
  f3 tangent, bitangent;
  GetOrthonormalBasis(normal, &tangent, &bitangent);
  return um::normalize(um::mul_cols(tangent, normal, bitangent, v));
}
// RandomSource.cs:63-89
f3 RandomSource::OnCosineWeightedHemisphere(f3 normal, int slot) {
  float ux, uy;
  NextFloat2(slot, &ux, &uy);
  float u = ux;
  float radius = um::sqrt(u);
  float theta = uy * 2 * um::PI;
  float s, c;
  um::sincos(theta, &s, &c);
  f3 tangentSpaceDirection = um::mk(radius * c, um::sqrt(1 - u), radius * s);
  return TangentToWorldSpace(tangentSpaceDirection, normal);
}

// ======================================================================================
// Runtime structs
// ======================================================================================

f3 v3(const float* p) { return um::mk(p[0], p[1], p[2]); }

struct Ray {  // Ray.cs
  f3 Origin, Direction;
  float Time;
  Ray OffsetTowards(f3 normal) const { return Ray{um::mad(normal, 0.001f, Origin), Direction, Time}; }  // [FMA] Ray.cs:18
  f3 GetPoint(float t) const { return um::mad(Direction, t, Origin); }                                     // [FMA] Ray.cs:20
};

struct HitRecord {  // HitRecord.cs
  float Distance;
  f3 Point, Normal;
  int Entity;  // index instead of Entity*
  float TexU = 0, TexV = 0;  // HitRecord.TexCoords: triangles only ("TODO: Texcoord support for primitives", Entity.cs:107)
};

// Image textures (TextureType.Image, Texture.cs:80-89,128-137): set by oracle_set_textures (test infrastructure: one set
// for the process, like the sky).  Images are borrowed; texels are clamped to the image like the plugin does (the
// reference reads past it for a coordinate of exactly 1).
const rtb_image* g_images = nullptr; size_t g_image_count = 0;
const rtb_material_textures* g_mat_tex = nullptr; size_t g_mat_tex_count = 0;
const float* g_tri_uv = nullptr; size_t g_tri_uv_count = 0;
const uint8_t* TexelOf(int image, float tu, float tv) {
  const rtb_image& im = g_images[image];
  int cx = (int)(tu * (float)im.width), cy = (int)(tv * (float)im.height);   // (int2)(textureCoordinates * ImageSize)
  cx = cx < 0 ? 0 : (cx > im.width - 1 ? im.width - 1 : cx);
  cy = cy < 0 ? 0 : (cy > im.height - 1 ? im.height - 1 : cy);
  return im.pixels + ((size_t)cy * im.width + cx) * im.pixel_stride;
}

struct Scene {
  const rtb_sphere* spheres; size_t sphere_count;
  const rtb_material* materials; size_t material_count;
  const rtb_bvh_node* nodes; size_t node_count;
  std::vector<um::rigid> origin_transform, inverse_transform;  // Entity.OriginTransform / InverseTransform (per sphere)
  // bvhEntities: what the leaves index (nullptr: entity i is sphere i); Entity.Type / Entity.Content
  const rtb_entity* entities = nullptr; size_t entity_count = 0;
  const rtb_triangle* triangles = nullptr; size_t triangle_count = 0;
  // entities with the reference's full Entity record (rotation, motion, Rect / Box content)
  const rtb_placed_entity* placed = nullptr; size_t placed_count = 0;
  bool has_volumes = false;  // some material is a ProbabilisticVolume (DetermineVolumeContainment is a no-op otherwise)
  std::vector<um::rigid> placed_inverse;  // Entity.InverseTransform of the static ones (Entity ctor, Entity.cs:51-52)
  static bool is_placed(const rtb_entity& e) {
    return (e.type & RTB_ENTITY_PLACED) || e.type == RTB_ENTITY_RECT || e.type == RTB_ENTITY_BOX;
  }
  bool is_convex_hull(int entity) const {  // EntityType.IsConvexHull (Entity.cs:22-25): Box or Sphere
    if (!entities) return true;
    const uint32_t base = entities[entity].type & ~(uint32_t)RTB_ENTITY_PLACED;
    return base == RTB_ENTITY_SPHERE || base == RTB_ENTITY_BOX;
  }
  uint32_t material_of(int entity) const {
    if (!entities) return spheres[entity].material;
    const rtb_entity& e = entities[entity];
    if (is_placed(e)) return placed[e.index].material;
    return e.type == RTB_ENTITY_TRIANGLE ? triangles[e.index].material : spheres[e.index].material;
  }
};

// HitTests.Hit(this AxisAlignedBoundingBox) (HitTests.cs:9-21)
bool AabbHit(const rtb_bvh_node& n, f3 rayOrigin, f3 rayInvDirection) {
  f3 mn = um::mk(n.bounds_min[0], n.bounds_min[1], n.bounds_min[2]);
  f3 mx = um::mk(n.bounds_max[0], n.bounds_max[1], n.bounds_max[2]);
  f3 t0 = (mn - rayOrigin) * rayInvDirection;
  f3 t1 = (mx - rayOrigin) * rayInvDirection;
  float tMin = um::max(0.0f, um::cmax(um::min(t0, t1)));
  float tMax = um::cmin(um::max(t0, t1));
  return tMin < tMax;
}

// HitTests.Hit(this Sphere) (HitTests.cs:23-60), in entity space
bool SphereHit(float radius, const Ray& r, float tMin, float tMax, float* distance, f3* normal) {
  float squaredRadius = radius * radius;  // Sphere ctor, Sphere.cs:10-14
  f3 oc = r.Origin;
  float a = um::dot(r.Direction, r.Direction);
  float b = um::dot(oc, r.Direction);
  float c = um::dot(oc, oc) - squaredRadius;
  float discriminant = um::fma(b, b, -(a * c));  // [FMA] b*b - a*c
  if (discriminant > 0) {
    float sqrtDiscriminant = um::sqrt(discriminant);
    float t = um::div(-b - sqrtDiscriminant, a);
    if (t < tMax && t > tMin) {
      *distance = t;
      *normal = r.GetPoint(t) / radius;
      return true;
    }
    t = um::div(-b + sqrtDiscriminant, a);
    if (t < tMax && t > tMin) {
      *distance = t;
      *normal = r.GetPoint(t) / radius;
      return true;
    }
  }
  *distance = 0;
  *normal = um::mk(0.0f);
  return false;
}

// HitTests.Hit(this Triangle) (HitTests.cs:113-150): Moeller-Trumbore, both faces, vertex normals interpolated
bool TriangleHit(const rtb_triangle& tri, const Ray& r, float tMin, float tMax, float* distance, f3* normal,
                 const float* uv6 = nullptr, float* texU = nullptr, float* texV = nullptr) {
  *distance = 0;
  *normal = um::mk(0.0f);
  const f3 data0 = v3(tri.edge2), data1 = v3(tri.edge1), data2 = v3(tri.v0);
  f3 pvec = um::cross(r.Direction, data0);
  float det = um::dot(data1, pvec);
  if (det == 0) return false;
  float invDet = um::div(1.0f, det);
  f3 tvec = r.Origin - data2;
  float u = um::dot(tvec, pvec) * invDet;
  if (u < 0 || u > 1) return false;
  f3 qvec = um::cross(tvec, data1);
  float v = um::dot(r.Direction, qvec) * invDet;
  if (v < 0 || u + v > 1) return false;
  float d = um::dot(data0, qvec) * invDet;
  if (d < tMin || d > tMax) return false;
  *distance = d;
  f3 barycentricCoords = um::mk(1 - u - v, u, v);
  *normal = um::mul_cols(v3(tri.normals[0]), v3(tri.normals[1]), v3(tri.normals[2]), barycentricCoords);
  if (uv6 && texU && texV) {  // texCoord = mul(tri.TextureCoordinates, barycentricCoords) (HitTests.cs:147)
    *texU = um::fma(uv6[4], barycentricCoords.z, um::fma(uv6[2], barycentricCoords.y, uv6[0] * barycentricCoords.x));
    *texV = um::fma(uv6[5], barycentricCoords.z, um::fma(uv6[3], barycentricCoords.y, uv6[1] * barycentricCoords.x));
  }
  return true;
}

// math.sign: (x > 0 ? 1 : 0) - (x < 0 ? 1 : 0)
float Sign(float x) { return (x > 0 ? 1.0f : 0.0f) - (x < 0 ? 1.0f : 0.0f); }

// HitTests.Hit(this Rect) (HitTests.cs:62-78); Rect ctor (Rect.cs:11-15): From = -size / 2, To = size / 2
bool RectHit(const float size[3], const Ray& r, float tMin, float tMax, float* distance, f3* normal) {
  *distance = 0;
  *normal = um::mk(0.0f);
  const float fromX = um::div(-size[0], 2.0f), fromY = um::div(-size[1], 2.0f);
  const float toX = um::div(size[0], 2.0f), toY = um::div(size[1], 2.0f);
  if (r.Direction.z >= 0) return false;
  float t = um::div(-r.Origin.z, r.Direction.z);
  if (t < tMin || t > tMax) return false;
  float x = r.Origin.x + t * r.Direction.x, y = r.Origin.y + t * r.Direction.y;
  if (x < fromX || y < fromY || x > toX || y > toY) return false;
  *distance = t;
  *normal = um::mk(0.0f, 0.0f, 1.0f);
  return true;
}

// HitTests.Hit(this Box) (HitTests.cs:80-111, Majercik et al.); Box ctor (Box.cs:11-15): Extents = size / 2,
// InverseExtents = 1 / Extents
bool BoxHit(const float size[3], const Ray& ray, float tMin, float tMax, float* distance, f3* normal) {
  *distance = 0;
  *normal = um::mk(0.0f);
  const f3 extents = um::mk(um::div(size[0], 2.0f), um::div(size[1], 2.0f), um::div(size[2], 2.0f));
  const f3 inverseExtents = um::rcp(extents);
  const f3 origin = ray.Origin + ray.Direction * tMin;   // "offset origin by tMin"
  const f3 rayDirection = ray.Direction;
  const f3 ao = um::mk(um::abs(origin.x), um::abs(origin.y), um::abs(origin.z)) * inverseExtents;
  const float winding = um::cmax(ao) < 1 ? -1.0f : 1.0f;
  f3 sgn = um::mk(-Sign(rayDirection.x), -Sign(rayDirection.y), -Sign(rayDirection.z));
  const f3 num = extents * winding * sgn - origin;
  const f3 distanceToPlane = um::mk(um::div(num.x, rayDirection.x), um::div(num.y, rayDirection.y), um::div(num.z, rayDirection.z));
  auto inside = [](float o1, float d1, float o2, float d2, float t, float e1, float e2) {
    return um::abs(o1 + d1 * t) < e1 && um::abs(o2 + d2 * t) < e2;
  };
  const bool testX = distanceToPlane.x >= 0 &&
                     inside(origin.y, rayDirection.y, origin.z, rayDirection.z, distanceToPlane.x, extents.y, extents.z);
  const bool testY = distanceToPlane.y >= 0 &&
                     inside(origin.z, rayDirection.z, origin.x, rayDirection.x, distanceToPlane.y, extents.z, extents.x);
  const bool testZ = distanceToPlane.z >= 0 &&
                     inside(origin.x, rayDirection.x, origin.y, rayDirection.y, distanceToPlane.z, extents.x, extents.y);
  sgn = testX ? um::mk(sgn.x, 0, 0) : testY ? um::mk(0, sgn.y, 0) : um::mk(0, 0, testZ ? sgn.z : 0.0f);
  const bool nzX = sgn.x != 0, nzY = sgn.y != 0, nzZ = sgn.z != 0;
  if (!(nzX || nzY || nzZ)) return false;
  float d = nzX ? distanceToPlane.x : nzY ? distanceToPlane.y : distanceToPlane.z;
  d += tMin;
  if (d > tMax) return false;
  *distance = d;
  *normal = sgn;
  return true;
}

// Entity.TransformAtTime (Entity.cs:124-127)
um::rigid TransformAtTime(const rtb_placed_entity& e, float t) {
  const float s = um::clamp(um::unlerp(e.time_range[0], e.time_range[1], t), 0.0f, 1.0f);
  um::rigid r;
  r.rot = um::quat{e.rotation[0], e.rotation[1], e.rotation[2], e.rotation[3]};
  r.pos = v3(e.position) + v3(e.destination_offset) * s;
  return r;
}

// Entity.Hit -> HitInternal -> HitContent (Entity.cs:57-122) for an entity with the full record
bool PlacedHit(const Scene& sc, uint32_t index, int entity, const Ray& ray, float tMin, float tMax, HitRecord* rec) {
  const rtb_placed_entity& e = sc.placed[index];
  um::rigid transformAtTime, inverseTransform;
  if (!e.moving) {
    transformAtTime = um::rigid{um::quat{e.rotation[0], e.rotation[1], e.rotation[2], e.rotation[3]}, v3(e.position)};
    inverseTransform = sc.placed_inverse[index];
  } else {
    transformAtTime = TransformAtTime(e, ray.Time);
    inverseTransform = um::inverse(transformAtTime);
  }
  Ray entitySpaceRay{um::transform(inverseTransform, ray.Origin), um::rotate(inverseTransform.rot, ray.Direction), ray.Time};
  float distance;
  f3 entityLocalNormal;
  bool hit;
  switch (e.type) {
    case RTB_ENTITY_SPHERE: hit = SphereHit(e.size[0], entitySpaceRay, tMin, tMax, &distance, &entityLocalNormal); break;
    case RTB_ENTITY_RECT: hit = RectHit(e.size, entitySpaceRay, tMin, tMax, &distance, &entityLocalNormal); break;
    case RTB_ENTITY_BOX: hit = BoxHit(e.size, entitySpaceRay, tMin, tMax, &distance, &entityLocalNormal); break;
    default: hit = false;
  }
  if (!hit) return false;
  rec->Distance = distance;
  rec->Point = ray.GetPoint(distance);
  rec->Normal = um::normalize(um::rotate(transformAtTime.rot, entityLocalNormal));
  rec->Entity = entity;
  return true;
}

// Entity.Hit -> HitInternal -> HitContent (Entity.cs:57-122): static sphere entity, or world-space triangle
bool EntityHit(const Scene& sc, int entity, const Ray& ray, float tMin, float tMax, HitRecord* rec) {
  int sphere = entity;
  if (sc.entities) {
    const rtb_entity& e = sc.entities[entity];
    if (Scene::is_placed(e)) return PlacedHit(sc, e.index, entity, ray, tMin, tMax, rec);
    if (e.type == RTB_ENTITY_TRIANGLE) {
      // "Triangles are always world-space" (Entity.cs:92-93); OriginTransform is the identity the mesh job gives
      // every triangle entity (AddMeshRuntimeEntitiesJob.cs), so rotate(transformAtTime, n) == n
      float distance;
      f3 entityLocalNormal;
      const float* uv6 = (g_tri_uv && e.index < g_tri_uv_count) ? g_tri_uv + 6 * (size_t)e.index : nullptr;
      rec->TexU = rec->TexV = 0;
      if (!TriangleHit(sc.triangles[e.index], ray, tMin, tMax, &distance, &entityLocalNormal, uv6, &rec->TexU, &rec->TexV)) return false;
      rec->Distance = distance;
      rec->Point = ray.GetPoint(distance);
      rec->Normal = um::normalize(um::rotate(um::quat_identity(), entityLocalNormal));
      rec->Entity = entity;
      return true;
    }
    sphere = (int)e.index;
  }
  const um::rigid& transformAtTime = sc.origin_transform[sphere];
  const um::rigid& inverseTransform = sc.inverse_transform[sphere];
  Ray entitySpaceRay{um::transform(inverseTransform, ray.Origin), um::rotate(inverseTransform.rot, ray.Direction), 0};
  float distance;
  f3 entityLocalNormal;
  if (!SphereHit(sc.spheres[sphere].radius, entitySpaceRay, tMin, tMax, &distance, &entityLocalNormal)) return false;
  rec->Distance = distance;
  rec->Point = ray.GetPoint(distance);
  rec->Normal = um::normalize(um::rotate(transformAtTime.rot, entityLocalNormal));
  rec->Entity = entity;
  return true;
}

// ======================================================================================
// Material (Material.cs, Microfacet.cs)
// ======================================================================================

float Schlick(float cosine, float refractiveIndex) {  // Material.cs:212-217
  float r0 = um::div(1 - refractiveIndex, 1 + refractiveIndex);
  r0 *= r0;
  return r0 + (1 - r0) * um::pow5(1 - cosine);
}
bool Refract(f3 v, f3 n, float niOverNt, f3* refracted) {  // Material.cs:198-210
  float dt = um::dot(v, n);
  float discriminant = 1 - niOverNt * niOverNt * (1 - dt * dt);
  if (discriminant > 0) {
    *refracted = niOverNt * (v - n * dt) - n * um::sqrt(discriminant);
    return true;
  }
  *refracted = um::mk(0.0f);
  return false;
}
float RoughnessToAlpha(float roughness) {  // Microfacet.cs:71-80
  roughness = um::max(roughness, 1e-3f);
  float x = um::log(roughness);
  return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}
float Lambda(f3 w, f3 normal, float roughness) {  // Microfacet.cs:53-69
  float cosTheta = um::dot(normal, w);
  float sqCosTheta = cosTheta * cosTheta;
  float sqSinTheta = um::max(0.0f, 1 - sqCosTheta);
  float sinTheta = um::sqrt(sqSinTheta);
  float tanTheta = um::div(sinTheta, cosTheta);
  float absTanTheta = um::abs(tanTheta);
  if (um::isinf(absTanTheta)) return 0;
  float alpha = RoughnessToAlpha(roughness);
  float alpha2Tan2Theta = (alpha * absTanTheta) * (alpha * absTanTheta);
  return um::div(-1 + um::sqrt(1 + alpha2Tan2Theta), 2.0f);
}
float SmithMaskingShadowing(f3 w, f3 normal, float roughness) {  // Microfacet.cs:9-12
  return um::div(1.0f, 1 + Lambda(w, normal, roughness));
}

bool AlmostEquals1(float v) { return um::abs(1.0f - v) < 1e-6f; }  // MathExtensions.cs:23-27
bool IsPerfectSpecular(const rtb_material& m, uint32_t materialIndex) {  // Material.cs:181-196
  switch (m.type) {
    case RTB_MATERIAL_DIELECTRIC: return true;
    case RTB_MATERIAL_STANDARD: {
      // Metallic.Type == Constant && ... && Glossiness.Type == Constant && ...
      if (g_mat_tex && materialIndex < g_mat_tex_count &&
          (g_mat_tex[materialIndex].metallic_image >= 0 || g_mat_tex[materialIndex].glossiness_image >= 0))
        return false;
      return AlmostEquals1(m.metallic) && AlmostEquals1(m.glossiness);
    }
  }
  return false;
}

// What Albedo / Emission .SampleColor and Glossiness / Metallic .SampleScalar return at the hit's TexCoords
// (Texture.cs:50-138): the material with its image textures evaluated, so that Scatter / Emit below read plain values.
rtb_material SampleTextures(const rtb_material& m, uint32_t materialIndex, const HitRecord& rec) {
  rtb_material r = m;
  if (!g_mat_tex || materialIndex >= g_mat_tex_count) return r;
  const rtb_material_textures& t = g_mat_tex[materialIndex];
  if (t.albedo_image >= 0) {   // float3(p[0], p[1], p[2]) / 255 * MainColor
    const uint8_t* p = TexelOf(t.albedo_image, rec.TexU, rec.TexV);
    for (int k = 0; k < 3; k++) r.albedo[k] = um::div((float)p[k], 255.0f) * m.albedo[k];
  }
  if (t.emission_image >= 0) {
    const uint8_t* p = TexelOf(t.emission_image, rec.TexU, rec.TexV);
    for (int k = 0; k < 3; k++) r.emission[k] = um::div((float)p[k], 255.0f) * m.emission[k];
  }
  if (t.glossiness_image >= 0)  // p[ScalarValueChannel] / 255.0f * MainColor[ScalarValueChannel]
    r.glossiness = um::div((float)TexelOf(t.glossiness_image, rec.TexU, rec.TexV)[t.glossiness_channel], 255.0f) * m.glossiness;
  if (t.metallic_image >= 0)
    r.metallic = um::div((float)TexelOf(t.metallic_image, rec.TexU, rec.TexV)[t.metallic_channel], 255.0f) * m.metallic;
  return r;
}

// Material.ProbabilisticHit (Material.cs:48-65); Density = Material.parameter = rtb_material.index_of_refraction
bool ProbabilisticHit(const rtb_material& m, float* hitDistance, RandomSource& rng, int* volumeDraws) {
  if (m.type != RTB_MATERIAL_PROBABILISTIC_VOLUME) return false;
  rng.random_events++;
  const float EPSILON = 1.1920928955078125e-7f;  // math.EPSILON
  float volumeHitDistance = -um::div(1.0f, um::max(m.index_of_refraction, EPSILON)) * um::log_unit(rng.NextFloat(SLOT_VOLUME + (*volumeDraws)++));
  if (volumeHitDistance < *hitDistance) {
    *hitDistance = volumeHitDistance;
    return true;
  }
  return false;
}

// Material.Scatter (Material.cs:67-173), Standard and Dielectric; constant textures
// (Texture.cs:50-59,101-108).
void Scatter(const rtb_material& m, const Ray& ray, const HitRecord& rec, RandomSource& rng,
             f3* reflectance, Ray* scattered) {
  if (m.type == RTB_MATERIAL_PROBABILISTIC_VOLUME) {  // Material.cs:163-168: isotropic; the ray's time is NOT carried over
    *reflectance = um::mk(m.albedo[0], m.albedo[1], m.albedo[2]);
    *scattered = Ray{rec.Point, rng.NextFloat3Direction(SLOT_ROUGH), 0.0f};
    rng.random_events += 2;
    return;
  }
  *reflectance = um::mk(m.albedo[0], m.albedo[1], m.albedo[2]);
  switch (m.type) {
    case RTB_MATERIAL_STANDARD: {
      float metallic = m.metallic;
      float glossiness = m.glossiness;
      float roughness = um::pow2(1 - glossiness);
      f3 roughNormal = roughness > 0
          ? um::normalize(um::lerp(rec.Normal, rng.OnCosineWeightedHemisphere(rec.Normal, SLOT_ROUGH), roughness))
          : rec.Normal;
      float incidentCosine = -um::dot(ray.Direction, roughNormal);
      float ior = um::lerp(1.5f, 1.1f, metallic);  // PlasticIor, MetalIor (Material.cs:18-19)
      float fresnel = Schlick(incidentCosine, ior);
      float maskingShadowing = SmithMaskingShadowing(ray.Direction, rec.Normal, roughness);
      float reflectionChance = um::saturate(fresnel * glossiness * maskingShadowing);

      if (reflectionChance > 0 && rng.NextFloat(SLOT_REFLECT) < reflectionChance) {
        *scattered = Ray{rec.Point, um::reflect(ray.Direction, roughNormal), ray.Time};
        *reflectance = um::mk(1.0f);
      } else {
        if (metallic > 0 && rng.NextFloat(SLOT_METAL) < metallic) {
          *scattered = Ray{rec.Point, um::reflect(ray.Direction, roughNormal), ray.Time};
        } else {
          *scattered = Ray{rec.Point, rng.OnCosineWeightedHemisphere(rec.Normal, SLOT_DIFFUSE), ray.Time};
        }
      }
      if (reflectionChance > 0 && reflectionChance < 1) rng.random_events++;
      if (metallic > 0 && metallic < 1) rng.random_events++;
      rng.random_events += roughness * (reflectionChance + (1 - reflectionChance) * metallic);
      rng.random_events += (1 - reflectionChance) * (1 - metallic);
      break;
    }
    case RTB_MATERIAL_DIELECTRIC: {
      float roughness = 1 - m.glossiness;
      f3 roughNormal = um::normalize(rec.Normal + roughness * rng.NextFloat3Direction(SLOT_ROUGH));
      float niOverNt, cosine;
      f3 outwardRoughNormal;
      float ddn = um::dot(ray.Direction, roughNormal);
      if (ddn > 0) {
        outwardRoughNormal = -roughNormal;
        niOverNt = m.index_of_refraction;
        cosine = m.index_of_refraction * ddn;
      } else {
        outwardRoughNormal = roughNormal;
        niOverNt = um::div(1.0f, m.index_of_refraction);
        cosine = -ddn;
      }
      f3 scatterDirection, refracted;
      if (Refract(ray.Direction, outwardRoughNormal, niOverNt, &refracted) &&
          rng.NextFloat(SLOT_REFLECT) > Schlick(cosine, m.index_of_refraction)) {
        scatterDirection = refracted;
      } else {
        scatterDirection = um::reflect(ray.Direction, roughNormal);
        *reflectance = um::mk(1.0f);
      }
      *scattered = Ray{rec.Point, scatterDirection, ray.Time};
      rng.random_events++;
      rng.random_events += roughness;
      break;
    }
    default:
      *scattered = ray;
      break;
  }
}

// ======================================================================================
// View (View.cs:38-48)
// ======================================================================================

Ray GetRay(const rtb_view& v, float nx, float ny, RandomSource& rng) {
  float rdx = 0, rdy = 0;
  if (v.lens_radius != 0) {
    rng.InUnitDisk(&rdx, &rdy);
    rdx = v.lens_radius * rdx;
    rdy = v.lens_radius * rdy;
  }
  f3 offset = v3(v.right) * rdx + v3(v.up) * rdy;
  f3 dir = um::normalize(v3(v.lower_left_corner) - offset + nx * v3(v.horizontal) + ny * v3(v.vertical));
  return Ray{v3(v.origin) + offset, dir, rng.NextFloat(SLOT_TIME)};
}

// ======================================================================================
// SampleBatchJob
// ======================================================================================

struct Job {
  const rtb_batch_params* p;
  const Scene* scene;
  const rtb_batch_buffers* buf;
  int noise;
};

struct Work {  // the stackalloc'd buffers of Execute (SampleBatchJob.cs:103-113); vectors grow like Hybrid* collections
  std::vector<f3> emission, attenuation;
  std::vector<int> node_stack, candidates;
  std::vector<HitRecord> hits;
};

// SampleBatchJob.FindHitCandidates (:403-448)
void FindHitCandidates(const Scene& sc, const Ray& ray, Work& w, rtb_diagnostics& diag) {
  f3 rayInvDirection = um::rcp(ray.Direction);
  rayInvDirection = um::mk(um::isnan(rayInvDirection.x) ? um::INF : rayInvDirection.x,
                           um::isnan(rayInvDirection.y) ? um::INF : rayInvDirection.y,
                           um::isnan(rayInvDirection.z) ? um::INF : rayInvDirection.z);
  w.node_stack.clear();
  w.candidates.clear();
  if (sc.node_count == 0) return;
  w.node_stack.push_back(0);
  while (!w.node_stack.empty()) {
    int ni = w.node_stack.back();
    w.node_stack.pop_back();
    const rtb_bvh_node& node = sc.nodes[ni];
    if (!AabbHit(node, ray.Origin, rayInvDirection)) continue;
    diag.bounds_hit_count++;
    if (node.first_entity >= 0) {
      for (int i = 0; i < node.entity_count; i++) w.candidates.push_back(node.first_entity + i);
      diag.candidate_count += node.entity_count;
    } else {
      w.node_stack.push_back(node.left);
      w.node_stack.push_back(node.right);
    }
  }
}

// SampleBatchJob.FindHits (:450-475)
void FindHits(const Scene& sc, const Ray& ray, Work& w) {
  w.hits.clear();
  while (!w.candidates.empty()) {
    int e = w.candidates.back();
    w.candidates.pop_back();
    HitRecord rec;
    if (EntityHit(sc, e, ray, 0, um::INF, &rec)) {
      w.hits.push_back(rec);
      // Inject exit hits for probabilistic convex hulls (:462-469)
      HitRecord exitRec;
      if (sc.materials[sc.material_of(e)].type == RTB_MATERIAL_PROBABILISTIC_VOLUME && sc.is_convex_hull(e) &&
          EntityHit(sc, e, ray, rec.Distance + 0.001f, um::INF, &exitRec))
        w.hits.push_back(exitRec);
    }
  }
  if (!w.hits.empty())
    std::stable_sort(w.hits.begin(), w.hits.end(),
                     [](const HitRecord& a, const HitRecord& b) { return a.Distance < b.Distance; });
}

// Cubemap (Texture.cs:141-211) over R16G16B16A16_SFloat faces; set by oracle_set_sky_cubemap (test infrastructure: one
// global sky for the process, like Environment.SkyCubemap is one per job)
const uint16_t* g_sky_faces = nullptr;
int g_sky_w = 0, g_sky_h = 0;
f3 CubemapSample(f3 vector) {
  if (!g_sky_faces) return um::mk(0.0f);
  const float absVector[4] = {um::abs(vector.x), um::abs(vector.y), um::abs(vector.z), 0.0f};
  float maxDistance = um::max(um::max(um::max(absVector[0], absVector[1]), absVector[2]), absVector[3]);
  int firstLane = 0;
  while (firstLane < 3 && !(maxDistance == absVector[firstLane])) firstLane++;     // tzcnt(bitmask(maxDistance == absVector))
  const float comp = firstLane == 0 ? vector.x : (firstLane == 1 ? vector.y : vector.z);
  const bool positive = comp >= 0;
  float u, v;
  switch (firstLane) {
    case 0: u = positive ? -vector.z : vector.z; v = -vector.y; break;
    case 1: u = vector.x; v = positive ? vector.z : -vector.z; break;
    default: u = positive ? vector.x : -vector.x; v = -vector.y; break;
  }
  u = um::div(u, absVector[firstLane > 2 ? 2 : firstLane]);
  v = um::div(v, absVector[firstLane > 2 ? 2 : firstLane]);
  const int halfW = g_sky_w / 2, halfH = g_sky_h / 2;
  int cx = (int)((u + 1) * (float)halfW), cy = (int)((v + 1) * (float)halfH);
  if (cx > g_sky_w - 1) cx = g_sky_w - 1;
  if (cy > g_sky_h - 1) cy = g_sky_h - 1;
  const int face = (firstLane > 2 ? 2 : firstLane) * 2 + (positive ? 0 : 1);
  const uint16_t* px = g_sky_faces + ((size_t)face * g_sky_w * g_sky_h + (size_t)cy * g_sky_w + cx) * 4;
  return um::mk(um::half_to_float(px[0]), um::half_to_float(px[1]), um::half_to_float(px[2]));
}

// SampleBatchJob.AnyBackwardsVolumeEntryHit (:508-524): scans the candidates of the backwards traversal (not popped)
bool AnyBackwardsVolumeEntryHit(const Scene& sc, const Ray& backwardsRay, Work& w) {
  for (size_t i = 0; i < w.candidates.size(); i++) {
    const int hitCandidate = w.candidates[i];
    HitRecord hitRecord;
    if (sc.materials[sc.material_of(hitCandidate)].type == RTB_MATERIAL_PROBABILISTIC_VOLUME &&
        EntityHit(sc, hitCandidate, backwardsRay, 0, um::INF, &hitRecord) &&
        um::dot(hitRecord.Normal, backwardsRay.Direction) > 0)
      return true;
  }
  return false;
}

// SampleBatchJob.DetermineVolumeContainment (:477-506): -1 = not inside a volume, else its material index
int DetermineVolumeContainment(const Scene& sc, const Ray& ray, Work& w, rtb_diagnostics& diag) {
  for (size_t i = 0; i < w.hits.size(); i++) {
    const HitRecord hit = w.hits[i];
    const uint32_t hitMaterial = sc.material_of(hit.Entity);
    if (sc.materials[hitMaterial].type == RTB_MATERIAL_PROBABILISTIC_VOLUME) {
      // Entry hit, early out
      if (um::dot(hit.Normal, ray.Direction) < 0) break;
      // Exit hit before an entry hit, we are likely inside this volume; throw a ray backwards to make sure
      Ray backwardsRay{ray.Origin, -ray.Direction, ray.Time};
      FindHitCandidates(sc, backwardsRay, w, diag);
      if (AnyBackwardsVolumeEntryHit(sc, backwardsRay, w)) return (int)hitMaterial;
    }
  }
  return -1;
}

// SampleBatchJob.Sample (:166-401)
bool Sample(const Job& job, Ray eyeRay, RandomSource& rng, Work& w, f3* sampleColor, f3* sampleNormal,
            f3* sampleAlbedo, rtb_diagnostics& diag, float* randomEventsAcc) {
  const Scene& sc = *job.scene;
  const rtb_batch_params& p = *job.p;
  w.emission.clear();
  w.attenuation.clear();
  float randomEventsLocalAcc = 0;
  int depth = 0;
  bool firstNonSpecularHit = false;
  *sampleColor = *sampleNormal = *sampleAlbedo = um::mk(0.0f);
  int currentProbabilisticVolumeMaterial = -1;  // Material* -> material index, null -> -1
  Ray ray = eyeRay;
  float pow2depth = 1.0f;  // pow(2, depth), exact (inf from depth 128 on, like powf)

  for (; depth < p.trace_depth; depth++) {
    rng.bounce = (uint32_t)depth;
    int volumeDraws = 0;  // Philox slot of the k-th ProbabilisticHit draw of this bounce
    FindHitCandidates(sc, ray, w, diag);
    FindHits(sc, ray, w);
    if (sc.has_volumes && currentProbabilisticVolumeMaterial < 0)
      currentProbabilisticVolumeMaterial = DetermineVolumeContainment(sc, ray, w, diag);
    diag.ray_count++;

    size_t hitIndex = 0;
    while (hitIndex < w.hits.size()) {
      HitRecord rec = w.hits[hitIndex];
      uint32_t materialIndex = sc.material_of(rec.Entity);

      if (currentProbabilisticVolumeMaterial >= 0 ||                                        // Inside a volume
          sc.materials[materialIndex].type == RTB_MATERIAL_PROBABILISTIC_VOLUME) {          // Entering a volume
        const bool isEntryHit = currentProbabilisticVolumeMaterial < 0;
        if (currentProbabilisticVolumeMaterial < 0) currentProbabilisticVolumeMaterial = (int)materialIndex;

        // Look for an obstacle or an exit hit
        size_t exitHitIndex = hitIndex;
        long lastExitIndex = -1;
        int sameMaterialEntries = 0;
        while (exitHitIndex < w.hits.size()) {
          const HitRecord& hit = w.hits[exitHitIndex];
          if ((int)sc.material_of(hit.Entity) == currentProbabilisticVolumeMaterial) {
            if (um::dot(hit.Normal, ray.Direction) < 0) {
              sameMaterialEntries++;
            } else {
              sameMaterialEntries--;
              lastExitIndex = (long)exitHitIndex;
            }
            if (sameMaterialEntries <= 0) break;
          } else {
            break;
          }
          exitHitIndex++;
        }
        if (sameMaterialEntries > 0 && lastExitIndex != -1) exitHitIndex = (size_t)lastExitIndex;

        if (exitHitIndex < w.hits.size()) {
          const HitRecord exitHitRecord = w.hits[exitHitIndex];
          float distanceInProbabilisticVolume = exitHitRecord.Distance;
          float probabilisticVolumeEntryDistance = 0;
          if (isEntryHit) {  // Factor in entry distance
            probabilisticVolumeEntryDistance = rec.Distance;
            distanceInProbabilisticVolume -= rec.Distance;
          }
          if (ProbabilisticHit(sc.materials[currentProbabilisticVolumeMaterial], &distanceInProbabilisticVolume, rng, &volumeDraws)) {
            // We hit inside the volume; hijack the current hit record's distance and material
            const float totalDistance = probabilisticVolumeEntryDistance + distanceInProbabilisticVolume;
            rec = HitRecord{totalDistance, ray.GetPoint(totalDistance), -ray.Direction, -1};
            materialIndex = (uint32_t)currentProbabilisticVolumeMaterial;
          } else {
            // No hit inside the volume, exit it
            currentProbabilisticVolumeMaterial = -1;
            if (sc.materials[sc.material_of(exitHitRecord.Entity)].type == RTB_MATERIAL_PROBABILISTIC_VOLUME &&
                um::dot(exitHitRecord.Normal, ray.Direction) > 0) {
              hitIndex = exitHitIndex + 1;  // Volume exit, move to next hit
              continue;
            }
            rec = exitHitRecord;  // Obstacle, scatter on the exit hit
            materialIndex = sc.material_of(rec.Entity);
          }
        } else {
          // No more surfaces to hit (probabilistic volume has holes)
          w.hits.clear();
          break;
        }
      }

      const rtb_material material = SampleTextures(sc.materials[materialIndex], materialIndex, rec);
      f3 albedo;
      Ray scatteredRay;
      Scatter(material, ray, rec, rng, &albedo, &scatteredRay);
      f3 emission = v3(material.emission);  // Material.Emit (Material.cs:175-179)
      w.emission.push_back(emission);
      if (depth == 0) *sampleNormal = rec.Normal;
      if (!firstNonSpecularHit) {
        if (!IsPerfectSpecular(material, materialIndex)) {
          *sampleAlbedo = emission + albedo;
          *sampleNormal = rec.Normal;
          firstNonSpecularHit = true;
        }
      }
      w.attenuation.push_back(albedo);
      randomEventsLocalAcc += um::div(rng.random_events, pow2depth);
      rng.random_events = 0;
      ray = scatteredRay;
      ray = ray.OffsetTowards(um::dot(scatteredRay.Direction, rec.Normal) >= 0 ? rec.Normal : -rec.Normal);
      break;
    }

    // No hit?
    if (hitIndex >= w.hits.size()) {
      f3 hitSkyColor = um::mk(0.0f);
      if (p.environment.sky_type == RTB_SKY_CUBEMAP)
        hitSkyColor = CubemapSample(ray.Direction);
      else if (p.environment.sky_type == RTB_SKY_GRADIENT)
        hitSkyColor = um::lerp(v3(p.environment.sky_bottom_color), v3(p.environment.sky_top_color),
                               0.5f * (ray.Direction.y + 1));
      w.emission.push_back(hitSkyColor);
      w.attenuation.push_back(um::mk(1.0f));
      randomEventsLocalAcc += um::div(rng.random_events, pow2depth);
      rng.random_events = 0;
      if (!firstNonSpecularHit) {
        *sampleAlbedo = hitSkyColor;
        *sampleNormal = -ray.Direction;
      }
      break;
    }
    pow2depth *= 2.0f;
  }

  *sampleColor = um::mk(0.0f);
  if (depth == p.trace_depth) return false;  // :379-381

  for (size_t k = w.emission.size(); k-- > 0;) {  // :383-396
    *sampleColor = *sampleColor * w.attenuation[k];
    *sampleColor = *sampleColor + w.emission[k];
  }
  *randomEventsAcc += randomEventsLocalAcc;
  return true;
}

// SampleBatchJob.Execute (:58-164)
void Execute(const Job& job, int index, Work& w) {
  const rtb_batch_params& p = *job.p;
  const rtb_batch_buffers& b = *job.buf;
  int width = (int)p.size[0];
  int cx = index % width, cy = index / width;
  if (cy % p.slice_divider != p.slice_offset) return;
  if (p.row_end > p.row_begin && (cy < p.row_begin || cy >= p.row_end)) return;  // rtb extension

  f3 colorAcc = v3(b.in_color + 4 * (size_t)index);
  float lastW = b.in_color[4 * (size_t)index + 3];
  f3 normalAcc = v3(b.in_normal + 3 * (size_t)index);
  f3 albedoAcc = v3(b.in_albedo + 3 * (size_t)index);
  float sampleCountWeightAcc = b.in_sample_count_weight[index];
  int sampleCount = (int)lastW;

  RandomSource rng;
  rng.mode = job.noise;
  if (job.noise == NOISE_WHITE_XORSHIFT) {
    rng.white.init((p.seed * 0x8C4CA03Fu) ^ ((uint32_t)index * 0x7383ED49u));  // :91
  } else {
    rng.key[0] = p.seed;
    rng.key[1] = PHILOX_KEY1;
    rng.pixel = (uint32_t)index;
  }

  f3 fallbackAlbedo = um::mk(0.0f), fallbackNormal = um::mk(0.0f);
  rtb_diagnostics diag{};

  uint32_t samplesToAccumulate;
  float sampleCountWeight = um::div(sampleCountWeightAcc, (float)sampleCount);
  if (sampleCountWeight == 0) {
    samplesToAccumulate = p.sample_count_range[0];
  } else {
    float normalized = um::saturate(um::unlerp(p.sample_count_weight_extrema[0], p.sample_count_weight_extrema[1], sampleCountWeight));
    samplesToAccumulate = (uint32_t)um::round(um::lerp((float)p.sample_count_range[0], (float)p.sample_count_range[1], normalized));
  }
  diag.sample_count_weight = sampleCountWeight;

  for (uint32_t s = 0; s < samplesToAccumulate; s++) {
    rng.sample = s;
    rng.bounce = BOUNCE_CAMERA;
    float jx = 0.5f, jy = 0.5f;
    if (p.sub_pixel_jitter) rng.NextFloat2(SLOT_JITTER, &jx, &jy);
    float nx = um::div((float)cx + jx, p.size[0]);
    float ny = um::div((float)cy + jy, p.size[1]);
    Ray eyeRay = GetRay(p.view, nx, ny, rng);

    f3 sampleColor, sampleNormal, sampleAlbedo;
    if (Sample(job, eyeRay, rng, w, &sampleColor, &sampleNormal, &sampleAlbedo, diag, &sampleCountWeightAcc)) {
      colorAcc = colorAcc + sampleColor;
      normalAcc = normalAcc + sampleNormal;
      albedoAcc = albedoAcc + sampleAlbedo;
      sampleCount++;
    }
    if (s == 0) {
      fallbackNormal = sampleNormal;
      fallbackAlbedo = sampleAlbedo;
    }
  }

  float* oc = b.out_color + 4 * (size_t)index;
  oc[0] = colorAcc.x; oc[1] = colorAcc.y; oc[2] = colorAcc.z; oc[3] = (float)sampleCount;
  f3 on = sampleCount == 0 ? fallbackNormal : normalAcc;
  f3 oa = sampleCount == 0 ? fallbackAlbedo : albedoAcc;
  float* pn = b.out_normal + 3 * (size_t)index;
  float* pa = b.out_albedo + 3 * (size_t)index;
  pn[0] = on.x; pn[1] = on.y; pn[2] = on.z;
  pa[0] = oa.x; pa[1] = oa.y; pa[2] = oa.z;
  b.out_sample_count_weight[index] = sampleCountWeightAcc;
  if (b.out_diagnostics) b.out_diagnostics[index] = diag;
}

}  // namespace

extern "C" {

#define ORACLE_API __attribute__((visibility("default")))

// Runs one SampleBatchJob over pixel indices [index_begin, index_end) (0,0 = all) with
// `threads` worker threads pulling one pixel at a time (== Schedule(W*H, 1, ...),
// Raytracer.cs:730).  noise: 0 = the reference's xorshift32 white noise, 1 = Philox slots.
ORACLE_API int oracle_sample_batch_world(const rtb_batch_params* params,
                                         const rtb_entity* entities, size_t entity_count,
                                         const rtb_sphere* spheres, size_t sphere_count,
                                         const rtb_triangle* triangles, size_t triangle_count,
                                         const rtb_material* materials, size_t material_count,
                                         const rtb_bvh_node* nodes, size_t node_count,
                                         const rtb_batch_buffers* buffers, int noise, int threads,
                                         int64_t index_begin, int64_t index_end);

ORACLE_API int oracle_sample_batch_placed(const rtb_batch_params* params,
                                          const rtb_entity* entities, size_t entity_count,
                                          const rtb_sphere* spheres, size_t sphere_count,
                                          const rtb_triangle* triangles, size_t triangle_count,
                                          const rtb_placed_entity* placed, size_t placed_count,
                                          const rtb_material* materials, size_t material_count,
                                          const rtb_bvh_node* nodes, size_t node_count,
                                          const rtb_batch_buffers* buffers, int noise, int threads,
                                          int64_t index_begin, int64_t index_end);

ORACLE_API int oracle_sample_batch(const rtb_batch_params* params,
                                   const rtb_sphere* spheres, size_t sphere_count,
                                   const rtb_material* materials, size_t material_count,
                                   const rtb_bvh_node* nodes, size_t node_count,
                                   const rtb_batch_buffers* buffers, int noise, int threads,
                                   int64_t index_begin, int64_t index_end) {
  return oracle_sample_batch_world(params, nullptr, 0, spheres, sphere_count, nullptr, 0, materials, material_count, nodes,
                                   node_count, buffers, noise, threads, index_begin, index_end);
}

ORACLE_API int oracle_sample_batch_world(const rtb_batch_params* params,
                                         const rtb_entity* entities, size_t entity_count,
                                         const rtb_sphere* spheres, size_t sphere_count,
                                         const rtb_triangle* triangles, size_t triangle_count,
                                         const rtb_material* materials, size_t material_count,
                                         const rtb_bvh_node* nodes, size_t node_count,
                                         const rtb_batch_buffers* buffers, int noise, int threads,
                                         int64_t index_begin, int64_t index_end) {
  return oracle_sample_batch_placed(params, entities, entity_count, spheres, sphere_count, triangles, triangle_count, nullptr, 0,
                                    materials, material_count, nodes, node_count, buffers, noise, threads, index_begin, index_end);
}

ORACLE_API int oracle_sample_batch_placed(const rtb_batch_params* params,
                                          const rtb_entity* entities, size_t entity_count,
                                          const rtb_sphere* spheres, size_t sphere_count,
                                          const rtb_triangle* triangles, size_t triangle_count,
                                          const rtb_placed_entity* placed, size_t placed_count,
                                          const rtb_material* materials, size_t material_count,
                                          const rtb_bvh_node* nodes, size_t node_count,
                                          const rtb_batch_buffers* buffers, int noise, int threads,
                                          int64_t index_begin, int64_t index_end) {
  if (!params || !buffers || params->slice_divider < 1 || params->trace_depth < 0) return RTB_ERR_INVALID_ARGUMENT;
  for (size_t i = 0; i < entity_count; i++) {
    const uint32_t base = entities[i].type & ~(uint32_t)RTB_ENTITY_PLACED;
    if (base < RTB_ENTITY_SPHERE || base > RTB_ENTITY_TRIANGLE) return RTB_ERR_UNSUPPORTED;
    if (Scene::is_placed(entities[i])) {
      if (base == RTB_ENTITY_TRIANGLE || entities[i].index >= placed_count || placed[entities[i].index].type != base)
        return RTB_ERR_INVALID_ARGUMENT;
    } else if (entities[i].index >= (base == RTB_ENTITY_SPHERE ? sphere_count : triangle_count)) {
      return RTB_ERR_INVALID_ARGUMENT;
    }
  }
  bool has_volumes = false;
  for (size_t i = 0; i < material_count; i++) {
    if (materials[i].type > RTB_MATERIAL_PROBABILISTIC_VOLUME) return RTB_ERR_UNSUPPORTED;
    has_volumes = has_volumes || materials[i].type == RTB_MATERIAL_PROBABILISTIC_VOLUME;
  }
  if (params->environment.sky_type == RTB_SKY_CUBEMAP && !g_sky_faces) return RTB_ERR_NO_SCENE;
  Scene sc{spheres, sphere_count, materials, material_count, nodes, node_count, {}, {}};
  sc.entities = entity_count ? entities : nullptr;
  sc.entity_count = entity_count;
  sc.triangles = triangles;
  sc.triangle_count = triangle_count;
  sc.placed = placed;
  sc.placed_count = placed_count;
  sc.has_volumes = has_volumes;
  sc.placed_inverse.resize(placed_count);
  for (size_t i = 0; i < placed_count; i++) {
    const rtb_placed_entity& e = placed[i];
    sc.placed_inverse[i] = um::inverse(um::rigid{um::quat{e.rotation[0], e.rotation[1], e.rotation[2], e.rotation[3]}, v3(e.position)});
  }
  sc.origin_transform.resize(sphere_count);
  sc.inverse_transform.resize(sphere_count);
  for (size_t i = 0; i < sphere_count; i++) {
    um::rigid t{um::quat_identity(), um::mk(spheres[i].center[0], spheres[i].center[1], spheres[i].center[2])};
    sc.origin_transform[i] = t;
    sc.inverse_transform[i] = um::inverse(t);  // Entity ctor, Entity.cs:52
  }
  Job job{params, &sc, buffers, noise};
  int64_t total = (int64_t)params->size[0] * (int64_t)params->size[1];
  if (index_end <= index_begin) { index_begin = 0; index_end = total; }
  if (index_end > total) index_end = total;
  if (threads < 1) threads = 1;
  std::atomic<int64_t> next(index_begin);
  auto worker = [&]() {
    Work w;
    for (;;) {
      int64_t i = next.fetch_add(1, std::memory_order_relaxed);
      if (i >= index_end) break;
      Execute(job, (int)i, w);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; t++) pool.emplace_back(worker);
  worker();
  for (auto& t : pool) t.join();
  return RTB_OK;
}

ORACLE_API void oracle_set_textures(const rtb_image* images, size_t image_count, const rtb_material_textures* material_textures,
                                    size_t material_count, const float* triangle_uvs, size_t triangle_count) {
  g_images = images; g_image_count = image_count;            // borrowed: the caller keeps the arrays alive while it renders
  g_mat_tex = material_textures; g_mat_tex_count = material_count;
  g_tri_uv = triangle_uvs; g_tri_uv_count = triangle_count;
}
ORACLE_API void oracle_set_sky_cubemap(const uint16_t* half_rgba, int face_width, int face_height) {
  g_sky_faces = half_rgba;     // borrowed: the caller keeps the array alive while it renders
  g_sky_w = face_width;
  g_sky_h = face_height;
}
ORACLE_API void oracle_cubemap_sample(const float dir[3], float out[3]) {
  const f3 c = CubemapSample(um::mk(dir[0], dir[1], dir[2]));
  out[0] = c.x; out[1] = c.y; out[2] = c.z;
}

// ---- known-answer-test hooks (tests/test_oracle_kat.py) ---------------------------------
ORACLE_API void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  philox4x32_10(ctr, key, out);
}
ORACLE_API void oracle_unity_random(uint32_t seed, uint32_t* out_states, float* out_floats, int n) {
  UnityRandom r; r.init(seed);
  for (int i = 0; i < n; i++) {
    uint32_t s = r.next_state();
    out_states[i] = s;
    out_floats[i] = RandomSource::to_float(s);
  }
}
ORACLE_API int oracle_sphere_hit(const float center[3], float radius, const float origin[3], const float dir[3],
                                 float* distance, float point[3], float normal[3]) {
  rtb_sphere s{};
  s.center[0] = center[0]; s.center[1] = center[1]; s.center[2] = center[2]; s.radius = radius;
  Scene sc{&s, 1, nullptr, 0, nullptr, 0, {}, {}};
  um::rigid t{um::quat_identity(), v3(center)};
  sc.origin_transform.push_back(t);
  sc.inverse_transform.push_back(um::inverse(t));
  HitRecord rec;
  Ray ray{v3(origin), v3(dir), 0};
  if (!EntityHit(sc, 0, ray, 0, um::INF, &rec)) return 0;
  *distance = rec.Distance;
  point[0] = rec.Point.x; point[1] = rec.Point.y; point[2] = rec.Point.z;
  normal[0] = rec.Normal.x; normal[1] = rec.Normal.y; normal[2] = rec.Normal.z;
  return 1;
}
ORACLE_API int oracle_triangle_hit(const rtb_triangle* tri, const float origin[3], const float dir[3], float* distance,
                                   float point[3], float normal[3]) {
  rtb_entity e{RTB_ENTITY_TRIANGLE, 0};
  Scene sc{nullptr, 0, nullptr, 0, nullptr, 0, {}, {}};
  sc.entities = &e; sc.entity_count = 1; sc.triangles = tri; sc.triangle_count = 1;
  HitRecord rec;
  Ray ray{v3(origin), v3(dir), 0};
  if (!EntityHit(sc, 0, ray, 0, um::INF, &rec)) return 0;
  *distance = rec.Distance;
  point[0] = rec.Point.x; point[1] = rec.Point.y; point[2] = rec.Point.z;
  normal[0] = rec.Normal.x; normal[1] = rec.Normal.y; normal[2] = rec.Normal.z;
  return 1;
}
ORACLE_API int oracle_placed_hit(const rtb_placed_entity* e, const float origin[3], const float dir[3], float time,
                                 float* distance, float point[3], float normal[3]) {
  rtb_entity ent{e->type | RTB_ENTITY_PLACED, 0};
  Scene sc{nullptr, 0, nullptr, 0, nullptr, 0, {}, {}};
  sc.entities = &ent; sc.entity_count = 1; sc.placed = e; sc.placed_count = 1;
  sc.placed_inverse.push_back(um::inverse(um::rigid{um::quat{e->rotation[0], e->rotation[1], e->rotation[2], e->rotation[3]}, v3(e->position)}));
  HitRecord rec;
  Ray ray{v3(origin), v3(dir), time};
  if (!EntityHit(sc, 0, ray, 0, um::INF, &rec)) return 0;
  *distance = rec.Distance;
  point[0] = rec.Point.x; point[1] = rec.Point.y; point[2] = rec.Point.z;
  normal[0] = rec.Normal.x; normal[1] = rec.Normal.y; normal[2] = rec.Normal.z;
  return 1;
}
ORACLE_API int oracle_aabb_hit(const float bmin[3], const float bmax[3], const float origin[3], const float dir[3]) {
  rtb_bvh_node n{};
  for (int i = 0; i < 3; i++) { n.bounds_min[i] = bmin[i]; n.bounds_max[i] = bmax[i]; }
  f3 inv = um::rcp(v3(dir));
  inv = um::mk(um::isnan(inv.x) ? um::INF : inv.x, um::isnan(inv.y) ? um::INF : inv.y, um::isnan(inv.z) ? um::INF : inv.z);
  return AabbHit(n, v3(origin), inv) ? 1 : 0;
}
ORACLE_API float oracle_schlick(float cosine, float ior) { return Schlick(cosine, ior); }
ORACLE_API float oracle_roughness_to_alpha(float r) { return RoughnessToAlpha(r); }
ORACLE_API float oracle_lambda(const float w[3], const float n[3], float roughness) { return Lambda(v3(w), v3(n), roughness); }
ORACLE_API int oracle_refract(const float v[3], const float n[3], float ni_over_nt, float out[3]) {
  f3 r;
  bool ok = Refract(v3(v), v3(n), ni_over_nt, &r);
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
  return ok ? 1 : 0;
}
ORACLE_API void oracle_get_ray(const rtb_view* view, float nx, float ny, uint32_t seed, uint32_t pixel, uint32_t sample,
                               int noise, float origin[3], float dir[3]) {
  RandomSource rng;
  rng.mode = noise;
  if (noise == NOISE_WHITE_XORSHIFT) rng.white.init(seed);
  rng.key[0] = seed; rng.key[1] = PHILOX_KEY1; rng.pixel = pixel; rng.sample = sample; rng.bounce = BOUNCE_CAMERA;
  Ray r = GetRay(*view, nx, ny, rng);
  origin[0] = r.Origin.x; origin[1] = r.Origin.y; origin[2] = r.Origin.z;
  dir[0] = r.Direction.x; dir[1] = r.Direction.y; dir[2] = r.Direction.z;
}
ORACLE_API void oracle_umath_sincos(const float* theta, float* s, float* c, int n) {
  for (int i = 0; i < n; i++) um::sincos(theta[i], s + i, c + i);
}
ORACLE_API void oracle_umath_log(const float* x, float* y, int n) {
  for (int i = 0; i < n; i++) y[i] = um::log(x[i]);
}
ORACLE_API void oracle_umath_pow(const float* x, float e, float* y, int n) {
  for (int i = 0; i < n; i++) y[i] = um::pow_pos(x[i], e);
}

// The vector / quaternion functions of include/rtb/umath.h, one record per call, so that tests can pin the header both
// consumers share (this oracle and the CUDA kernels) against an INDEPENDENT float64 restatement of the
// Unity.Mathematics formulas instead of against each other.  `in` / `out` record sizes per op:
//   0 reflect(i, n)            6 -> 3      1 rotate(q, v)             7 -> 3      2 inverse(RigidTransform(q, p))  7 -> 7
//   3 transform(rt(q, p), x)  10 -> 3      4 lerp(a, b, s) (float3)   7 -> 3      5 normalize(v)                   3 -> 3
//   6 cross(a, b)              6 -> 3      7 mul(float3x3(c0,c1,c2), v) 12 -> 3   8 dot(a, b)                      6 -> 1
//   9 scalars (x, y, z) -> min(x,y), max(x,y), saturate(x), round(x), unlerp(x,y,z), lerp(x,y,z), rcp(x), rsqrt(|x|)   3 -> 8
ORACLE_API int oracle_umath_vec(int op, const float* in, float* out, int n) {
  static const int in_size[10] = {6, 7, 7, 10, 7, 3, 6, 12, 6, 3}, out_size[10] = {3, 3, 7, 3, 3, 3, 3, 3, 1, 8};
  if (op < 0 || op > 9) return -1;
  auto put = [](float* o, f3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; };
  for (int r = 0; r < n; r++) {
    const float* a = in + (size_t)r * in_size[op];
    float* o = out + (size_t)r * out_size[op];
    switch (op) {
      case 0: put(o, um::reflect(um::mk(a[0], a[1], a[2]), um::mk(a[3], a[4], a[5]))); break;
      case 1: put(o, um::rotate(um::quat{a[0], a[1], a[2], a[3]}, um::mk(a[4], a[5], a[6]))); break;
      case 2: {
        const um::rigid inv = um::inverse(um::rigid{um::quat{a[0], a[1], a[2], a[3]}, um::mk(a[4], a[5], a[6])});
        o[0] = inv.rot.x; o[1] = inv.rot.y; o[2] = inv.rot.z; o[3] = inv.rot.w; put(o + 4, inv.pos);
        break;
      }
      case 3: put(o, um::transform(um::rigid{um::quat{a[0], a[1], a[2], a[3]}, um::mk(a[4], a[5], a[6])}, um::mk(a[7], a[8], a[9]))); break;
      case 4: put(o, um::lerp(um::mk(a[0], a[1], a[2]), um::mk(a[3], a[4], a[5]), a[6])); break;
      case 5: put(o, um::normalize(um::mk(a[0], a[1], a[2]))); break;
      case 6: put(o, um::cross(um::mk(a[0], a[1], a[2]), um::mk(a[3], a[4], a[5]))); break;
      case 7: put(o, um::mul_cols(um::mk(a[0], a[1], a[2]), um::mk(a[3], a[4], a[5]), um::mk(a[6], a[7], a[8]), um::mk(a[9], a[10], a[11]))); break;
      case 8: o[0] = um::dot(um::mk(a[0], a[1], a[2]), um::mk(a[3], a[4], a[5])); break;
      default:
        o[0] = um::min(a[0], a[1]); o[1] = um::max(a[0], a[1]); o[2] = um::saturate(a[0]); o[3] = um::round(a[0]);
        o[4] = um::unlerp(a[0], a[1], a[2]); o[5] = um::lerp(a[0], a[1], a[2]); o[6] = um::rcp(a[0]); o[7] = um::rsqrt(um::abs(a[0]));
        break;
    }
  }
  return 0;
}

}  // extern "C"
