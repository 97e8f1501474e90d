// librtb_host.so — host-side producers of the sample job's inputs (see include/rtb_host.h).
// Plain C++17, no CUDA.  Compiled with -ffp-contract=off so float results are the same on
// every x86-64 box (fixtures under tests/golden/ depend on it).
//
// Reference paths are relative to /root/reference/RaytracingInOneWeekend/Assets/Scripts.

#include "rtb_host.h"
#include "rtb/umath.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <set>
#include <thread>
#include <vector>

namespace {

using um::f3;

// ---- Unity.Mathematics.Random -----------------------------------------------------------
// state = seed; NextState(): t = state; state ^= state << 13; state ^= state >> 17; state ^= state << 5; return t
struct URandom {
  uint32_t state;
  explicit URandom(uint32_t seed) : state(seed) { next_state(); }
  uint32_t next_state() {
    uint32_t t = state;
    state ^= state << 13;
    state ^= state >> 17;
    state ^= state << 5;
    return t;
  }
  float next_float() { return um::asfloat(0x3f800000u | (next_state() >> 9)) - 1.0f; }
  float next_float(float lo, float hi) { return next_float() * (hi - lo) + lo; }
  f3 next_float3() {
    float x = next_float(), y = next_float(), z = next_float();
    return um::mk(x, y, z);
  }
  f3 next_float3(f3 lo, f3 hi) {
    f3 u = next_float3();
    return um::mk(u.x * (hi.x - lo.x) + lo.x, u.y * (hi.y - lo.y) + lo.y, u.z * (hi.z - lo.z) + lo.z);
  }
};

// ---- materials (legacy -> HEAD mapping) ----------------------------------
rtb_material lambertian(f3 albedo) {
  rtb_material m{};
  m.type = RTB_MATERIAL_STANDARD;
  m.albedo[0] = albedo.x; m.albedo[1] = albedo.y; m.albedo[2] = albedo.z;
  m.glossiness = 0.0f; m.metallic = 0.0f; m.index_of_refraction = 0.0f;
  return m;
}
rtb_material metal(f3 albedo, float fuzz) {
  rtb_material m{};
  m.type = RTB_MATERIAL_STANDARD;
  m.albedo[0] = albedo.x; m.albedo[1] = albedo.y; m.albedo[2] = albedo.z;
  m.glossiness = 1.0f - fuzz; m.metallic = 1.0f; m.index_of_refraction = 0.0f;
  return m;
}
rtb_material dielectric(float ior) {
  rtb_material m{};
  m.type = RTB_MATERIAL_DIELECTRIC;
  m.albedo[0] = m.albedo[1] = m.albedo[2] = 1.0f;
  m.glossiness = 1.0f; m.metallic = 0.0f; m.index_of_refraction = ior;
  return m;
}
rtb_sphere sphere(f3 c, float r, uint32_t material) {
  rtb_sphere s{};
  s.center[0] = c.x; s.center[1] = c.y; s.center[2] = c.z;
  s.radius = r; s.material = material;
  return s;
}

struct SceneBuild {
  std::vector<rtb_sphere> spheres;
  std::vector<rtb_material> materials;
  std::vector<uint8_t> exclude_from_overlap;
  rtbh_scene_info info{};
  void add(f3 c, float r, const rtb_material& m, bool exclude = false) {
    spheres.push_back(sphere(c, r, (uint32_t)materials.size()));
    materials.push_back(m);
    exclude_from_overlap.push_back(exclude ? 1 : 0);
    if (m.type == RTB_MATERIAL_DIELECTRIC) info.dielectric_count++;
    else if (m.metallic > 0.0f) info.metal_count++;
    else info.lambertian_count++;
  }
};

void set_sky(rtb_environment& e) {
  // skyBottomColor (1,1,1), skyTopColor (0.5,0.7,1): both legacy assets (:204-205 / :82-83)
  e.sky_type = RTB_SKY_GRADIENT;
  e.sky_bottom_color[0] = e.sky_bottom_color[1] = e.sky_bottom_color[2] = 1.0f;
  e.sky_top_color[0] = 0.5f; e.sky_top_color[1] = 0.7f; e.sky_top_color[2] = 1.0f;
}

// Three Spheres (Book 1).asset:15-83.  Material GUIDs dangle; book-1 values.
void scene_three_spheres(SceneBuild& sb) {
  sb.info.camera = rtbh_camera{{0.0f, 0.0f, -2.25f}, {0.0f, 0.0f, 0.0f}, 0.0f, 60.0f};
  set_sky(sb.info.environment);
  sb.add(um::mk(0.0f, -100.5f, 0.0f), 100.0f, lambertian(um::mk(0.8f, 0.8f, 0.0f)));
  sb.add(um::mk(1.0f, 0.0f, 0.0f), 0.5f, metal(um::mk(0.8f, 0.6f, 0.2f), 0.0f));
  sb.add(um::mk(0.0f, 0.0f, 0.0f), 0.5f, lambertian(um::mk(0.1f, 0.2f, 0.5f)));
  sb.add(um::mk(-1.0f, 0.0f, 0.0f), 0.5f, dielectric(1.5f));
  sb.add(um::mk(-1.0f, 0.0f, 0.0f), -0.45f, dielectric(1.5f));
}

// Parameters of the one RandomEntityGroup in Final Scene (Book 1).asset:85-202.
struct RandomGroup {
  uint32_t tentative_count = 1000;
  float spread_x = 22.0f, spread_y = 0.0f, spread_z = 22.0f;
  f3 offset = um::mk(0.0f, 0.2f, 0.0f);
  float radius_lo = 0.2f, radius_hi = 0.2f;
  float min_distance = 0.15f;
  float movement_chance = 0.0f;
  float lambert_chance = 0.8f, metal_chance = 0.15f, dielectric_chance = 0.05f, light_chance = 0.0f;
  f3 diffuse_lo = um::mk(0.0f), diffuse_hi = um::mk(1.0f);
  bool double_sample_diffuse = true;
  f3 metal_lo = um::mk(0.5f), metal_hi = um::mk(1.0f);
  float fuzz_lo = 0.0f, fuzz_hi = 0.5f;
  float ior_lo = 1.5f, ior_hi = 1.5f;
  uint32_t stop_after_accepted = 0;  // 0 = run tentative_count iterations (the reference); else until N accepted
};

// RandomDistribution.DartThrowing branch of CollectActiveEntities (Raytracer.cs:1453-1474)
// with GetEntity (:1420-1450), GetMaterial (:1364-1413), AnyOverlap (:1415-1418).
void dart_throw(SceneBuild& sb, URandom& rng, const RandomGroup& g) {
  // GetMaterial thresholds (:1366-1379), float arithmetic in the order written there
  float p_l = g.lambert_chance, p_m = g.metal_chance, p_d = g.dielectric_chance, p_e = g.light_chance;
  float sum = p_l + p_m + p_d + p_e;
  p_m += p_l; p_d += p_m; p_e += p_d;
  p_l /= sum; p_m /= sum; p_d /= sum; p_e /= sum;

  uint32_t accepted = 0, draws = 0;
  for (uint32_t i = 0; g.stop_after_accepted ? (accepted < g.stop_after_accepted) : (i < g.tentative_count); i++) {
    draws++;
    f3 center = rng.next_float3(um::mk(-g.spread_x / 2, -g.spread_y / 2, -g.spread_z / 2),
                                um::mk(g.spread_x / 2, g.spread_y / 2, g.spread_z / 2));
    center = um::mk(center.x + g.offset.x, center.y + g.offset.y, center.z + g.offset.z);
    float radius = rng.next_float(g.radius_lo, g.radius_hi);

    bool overlap = false;
    for (size_t k = 0; k < sb.spheres.size(); k++) {
      if (sb.exclude_from_overlap[k]) continue;
      const rtb_sphere& s = sb.spheres[k];
      float dx = center.x - s.center[0], dy = center.y - s.center[1], dz = center.z - s.center[2];
      float dist = std::sqrt(dx * dx + dy * dy + dz * dz);   // math.distance = length(b - a)
      if (dist < s.radius + radius + g.min_distance) { overlap = true; break; }
    }
    if (overlap) continue;

    // GetEntity: movement draw first, then the material draws
    bool moving = rng.next_float() < g.movement_chance;
    (void)moving;  // MovementChance is 0 in the asset; a moving entity would draw an offset here
    rtb_material material{};
    float u = rng.next_float();
    if (u < p_l) {
      f3 color = rng.next_float3(g.diffuse_lo, g.diffuse_hi);
      if (g.double_sample_diffuse) {
        f3 c2 = rng.next_float3(g.diffuse_lo, g.diffuse_hi);
        color = um::mk(color.x * c2.x, color.y * c2.y, color.z * c2.z);
      }
      material = lambertian(color);
    } else if (u < p_m) {
      f3 color = rng.next_float3(g.metal_lo, g.metal_hi);
      float fuzz = rng.next_float(g.fuzz_lo, g.fuzz_hi);
      material = metal(color, fuzz);
    } else if (u < p_d) {
      material = dielectric(rng.next_float(g.ior_lo, g.ior_hi));
    } else {
      // LightChance is 0; an emissive draw would go here (rtb does not need it for the BASELINE scenes)
      material = lambertian(um::mk(0.0f));
    }
    // Rotation = 0: rotate(identity, center - Offset) + Offset.  (c - o) + o can differ from c
    // by an ulp in float; the reference evaluates it, so do we.
    f3 pos = um::mk((center.x - g.offset.x) + g.offset.x, (center.y - g.offset.y) + g.offset.y,
                    (center.z - g.offset.z) + g.offset.z);
    sb.add(pos, radius, material);
    accepted++;
  }
  sb.info.tentative_draws = draws;
}

// Final Scene (Book 1).asset:15-84 — 4 named spheres, then the dart-thrown group.
void scene_final_named(SceneBuild& sb) {
  set_sky(sb.info.environment);
  sb.add(um::mk(0.0f, -1000.0f, 0.0f), 1000.0f, lambertian(um::mk(0.5f, 0.5f, 0.5f)), /*exclude*/ true);
  sb.add(um::mk(0.0f, 1.0f, 0.0f), 1.0f, dielectric(1.5f));
  sb.add(um::mk(-4.0f, 1.0f, 0.0f), 1.0f, lambertian(um::mk(0.4f, 0.2f, 0.1f)));
  sb.add(um::mk(4.0f, 1.0f, 0.0f), 1.0f, metal(um::mk(0.7f, 0.6f, 0.5f), 0.0f));
}

// ---- BVH build (BvhNodeData.cs) ---------------------------------------------------------
struct Aabb { f3 mn, mx; };
Aabb enclose(const Aabb& a, const Aabb& b) { return Aabb{um::min(a.mn, b.mn), um::max(a.mx, b.mx)}; }

struct BuildEntity {  // BvhBuildingEntity (:23-80)
  uint32_t index;
  Aabb bounds;
};

Aabb sphere_world_bounds(const rtb_sphere& s) {
  // Sphere.Bounds = [-|r|, |r|] (Sphere.cs:16-23), 8 corners through the rigid transform
  // (identity rotation: rotate() returns the corner unchanged), min/max (:41-78).
  float ar = std::fabs(s.radius);
  f3 pos = um::mk(s.center[0], s.center[1], s.center[2]);
  const float lo = -ar, hi = ar;
  f3 mn = um::mk(std::numeric_limits<float>::infinity());
  f3 mx = um::mk(-std::numeric_limits<float>::infinity());
  for (int i = 0; i < 8; i++) {
    f3 corner = um::mk((i & 1) ? hi : lo, (i & 2) ? hi : lo, (i & 4) ? hi : lo);
    f3 t = um::mk(corner.x + pos.x, corner.y + pos.y, corner.z + pos.z);
    mn = um::min(mn, t);
    mx = um::max(mx, t);
  }
  return Aabb{mn, mx};
}

// Threads for the BVH build: RTB_BUILD_THREADS, else the hardware's (capped at 32).
int build_threads() {
  if (const char* e = getenv("RTB_BUILD_THREADS")) return std::max(1, atoi(e));
  const unsigned hw = std::thread::hardware_concurrency();
  return (int)std::min(32u, std::max(1u, hw));
}

struct NodeData {  // BvhNodeData
  Aabb bounds{};
  int first_entity = -1, entity_count = 0, depth = 0;
  int left = -1, right = -1;
};

struct Builder {
  const rtb_sphere* spheres;
  int max_depth;
  std::vector<NodeData> nodes;          // bvhNodes (root at 0, children appended as created)
  std::vector<uint32_t> bvh_entities;   // bvhEntities order (indices into the input array)

  // BvhNodeData ctor (:122-213).  `ents` is the NativeSlice; sorting it in place is visible to
  // the caller exactly as in the reference.
  void build(int self, BuildEntity* ents, int count, int depth, int sort_axis) {
    NodeData nd;
    nd.depth = depth;
    Aabb entire{um::mk(std::numeric_limits<float>::max()), um::mk(std::numeric_limits<float>::lowest())};
    for (int i = 0; i < count; i++) entire = enclose(entire, ents[i].bounds);

    int biggest = -1;
    float biggest_size = std::numeric_limits<float>::lowest();
    f3 size = entire.mx - entire.mn;
    for (int i = 0; i < 3; i++) {
      float s = um::comp(size, i);
      if (s > biggest_size) { biggest = i; biggest_size = s; }
    }
    if (sort_axis != biggest && biggest >= 0) {
      // NativeSlice.Sort with comparer (int)sign(lhs.Min[axis] - rhs.Min[axis]) (:240-250).
      // Unity's sort is not stable; ties are measure-zero for these scenes and radiance does
      // not depend on tie order.  We use a stable sort so the flattened tree is reproducible.
      parallel_stable_sort(ents, count, biggest, threads_for(count));
    }

    if (depth == max_depth || count <= 1) {
      nd.first_entity = (int)bvh_entities.size();
      for (int i = 0; i < count; i++) bvh_entities.push_back(ents[i].index);
      if (count > 0) {
        nd.bounds = ents[0].bounds;
        for (int i = 1; i < count; i++) nd.bounds = enclose(nd.bounds, ents[i].bounds);
      } else {
        nd.bounds = Aabb{um::mk(0.0f), um::mk(0.0f)};
      }
      nd.entity_count = count;
      nodes[self] = nd;
      return;
    }

    int partition_length = 0;
    float partition_start = um::comp(ents[0].bounds.mn, biggest);
    for (int i = 0; i < count; i++) {
      partition_length++;
      const Aabb& b = ents[i].bounds;
      float bsize = um::comp(b.mx, biggest) - um::comp(b.mn, biggest);
      if (um::comp(b.mn, biggest) - partition_start > biggest_size / 2 || bsize > biggest_size / 2) break;
    }
    if (partition_length == count) partition_length--;

    int l, r;
    if (count >= kParallelEntities && worker_budget > 1) {
      // Large subtrees are built concurrently, each into its own builder, and spliced in the order the serial recursion
      // creates them (left subtree's nodes and leaf entities first, then the right one's): same node array, same
      // bvhEntities order, whatever the thread count.
      Builder lb, rb;
      lb.spheres = rb.spheres = spheres;
      lb.max_depth = rb.max_depth = max_depth;
      lb.worker_budget = worker_budget / 2;
      rb.worker_budget = worker_budget - lb.worker_budget;
      lb.nodes.emplace_back();
      rb.nodes.emplace_back();
      std::thread left_thread([&] { lb.build(0, ents, partition_length, depth + 1, biggest); });
      rb.build(0, ents + partition_length, count - partition_length, depth + 1, biggest);
      left_thread.join();
      l = splice(lb);
      r = splice(rb);
    } else {
      l = (int)nodes.size();
      nodes.emplace_back();
      build(l, ents, partition_length, depth + 1, biggest);
      r = (int)nodes.size();
      nodes.emplace_back();
      build(r, ents + partition_length, count - partition_length, depth + 1, biggest);
    }
    nd.left = l;
    nd.right = r;
    nd.bounds = enclose(nodes[l].bounds, nodes[r].bounds);
    nodes[self] = nd;
  }

  // ---- multithreaded build (the reference's BvhNodeData ctor is one serial job: 2.2 s for a million triangles) ----
  static constexpr int kParallelEntities = 16384;
  int worker_budget = 1;                // threads this subtree may use
  int threads_for(int count) const { return count >= kParallelEntities ? worker_budget : 1; }

  // appends another builder's subtree (root at its index 0): returns the index its root got
  int splice(const Builder& b) {
    const int node_base = (int)nodes.size(), entity_base = (int)bvh_entities.size();
    for (NodeData n : b.nodes) {
      if (n.left >= 0) n.left += node_base;
      if (n.right >= 0) n.right += node_base;
      if (n.first_entity >= 0) n.first_entity += entity_base;
      nodes.push_back(n);
    }
    bvh_entities.insert(bvh_entities.end(), b.bvh_entities.begin(), b.bvh_entities.end());
    return node_base;
  }

  // stable sort by bounds.min[axis]: chunks sorted concurrently, then merged pairwise (std::inplace_merge is stable)
  static void parallel_stable_sort(BuildEntity* ents, int count, int ax, int threads) {
    auto less = [ax](const BuildEntity& a, const BuildEntity& b) { return um::comp(a.bounds.mn, ax) < um::comp(b.bounds.mn, ax); };
    if (threads <= 1 || count < kParallelEntities) {
      std::stable_sort(ents, ents + count, less);
      return;
    }
    int chunks = 1;
    while (chunks * 2 <= threads && count / (chunks * 2) >= kParallelEntities / 4) chunks *= 2;
    std::vector<int> cut(chunks + 1);
    for (int i = 0; i <= chunks; i++) cut[i] = (int)((long long)count * i / chunks);
    {
      std::vector<std::thread> pool;
      for (int i = 1; i < chunks; i++) pool.emplace_back([&, i] { std::stable_sort(ents + cut[i], ents + cut[i + 1], less); });
      std::stable_sort(ents + cut[0], ents + cut[1], less);
      for (auto& t : pool) t.join();
    }
    for (int width = 1; width < chunks; width *= 2) {
      std::vector<std::thread> pool;
      for (int i = 0; i + width < chunks; i += 2 * width) {
        const int lo = cut[i], mid = cut[i + width], hi = cut[std::min(i + 2 * width, chunks)];
        pool.emplace_back([=] { std::inplace_merge(ents + lo, ents + mid, ents + hi, less); });
      }
      for (auto& t : pool) t.join();
    }
  }

  // BuildRuntimeBvhJob.WalkBvh (:18-33): post-order, written backwards, root lands on index 0.
  int next_index = 0;
  int flatten(int nd, rtb_bvh_node* out) {
    int l = -1, r = -1;
    if (nodes[nd].first_entity < 0) {
      if (nodes[nd].left >= 0) l = flatten(nodes[nd].left, out);
      if (nodes[nd].right >= 0) r = flatten(nodes[nd].right, out);
    }
    rtb_bvh_node& o = out[next_index];
    const NodeData& n = nodes[nd];
    o.bounds_min[0] = n.bounds.mn.x; o.bounds_min[1] = n.bounds.mn.y; o.bounds_min[2] = n.bounds.mn.z;
    o.bounds_max[0] = n.bounds.mx.x; o.bounds_max[1] = n.bounds.mx.y; o.bounds_max[2] = n.bounds.mx.z;
    o.left = l; o.right = r;
    o.first_entity = n.first_entity;
    o.entity_count = n.entity_count;
    return next_index--;
  }
};

// ---- auto-focus: HitTests.Hit(this BvhNode) (HitTests.cs:152-196) -----------------------
bool aabb_hit(const rtb_bvh_node& n, f3 o, f3 inv) {  // HitTests.cs:9-21
  f3 mn = um::mk(n.bounds_min[0], n.bounds_min[1], n.bounds_min[2]);
  f3 mx = um::mk(n.bounds_max[0], n.bounds_max[1], n.bounds_max[2]);
  f3 t0 = (mn - o) * inv, t1 = (mx - o) * inv;
  float tmin = um::max(0.0f, um::cmax(um::min(t0, t1)));
  float tmax = um::cmin(um::max(t0, t1));
  return tmin < tmax;
}
bool sphere_hit(const rtb_sphere& s, f3 o, f3 d, float t_min, float t_max, float* dist) {  // HitTests.cs:23-60
  f3 oc = o + um::mk(-s.center[0], -s.center[1], -s.center[2]);
  float a = d.x * d.x + d.y * d.y + d.z * d.z;
  float b = oc.x * d.x + oc.y * d.y + oc.z * d.z;
  float c = (oc.x * oc.x + oc.y * oc.y + oc.z * oc.z) - s.radius * s.radius;
  float disc = b * b - a * c;
  if (disc > 0) {
    float sq = std::sqrt(disc);
    float t = (-b - sq) / a;
    if (t < t_max && t > t_min) { *dist = t; return true; }
    t = (-b + sq) / a;
    if (t < t_max && t > t_min) { *dist = t; return true; }
  }
  return false;
}
bool node_hit(const rtb_bvh_node* nodes, const rtb_sphere* spheres, int idx, f3 o, f3 d, float* dist) {
  const rtb_bvh_node& n = nodes[idx];
  f3 inv = um::rcp(d);
  if (!aabb_hit(n, o, inv)) return false;
  if (n.first_entity >= 0) {
    bool any = false;
    for (int i = 0; i < n.entity_count; i++) {
      float t;
      if (sphere_hit(spheres[n.first_entity + i], o, d, 0.0f, std::numeric_limits<float>::infinity(), &t) &&
          (!any || t < *dist)) {
        any = true;
        *dist = t;
      }
    }
    return any;
  }
  float tl = 0, tr = 0;
  bool hl = n.left >= 0 && node_hit(nodes, spheres, n.left, o, d, &tl);
  bool hr = n.right >= 0 && node_hit(nodes, spheres, n.right, o, d, &tr);
  if (!hl && !hr) return false;
  if (hl && hr) { *dist = tl < tr ? tl : tr; return true; }
  *dist = hl ? tl : tr;
  return true;
}

}  // namespace

extern "C" {

void rtbh_random_init(rtbh_random* r, uint32_t seed) { URandom u(seed); r->state = u.state; }
uint32_t rtbh_random_next_state(rtbh_random* r) {
  URandom u(1); u.state = r->state;
  uint32_t t = u.next_state();
  r->state = u.state;
  return t;
}
float rtbh_random_next_float(rtbh_random* r) {
  return um::asfloat(0x3f800000u | (rtbh_random_next_state(r) >> 9)) - 1.0f;
}

int rtbh_scene_generate(int scene_id, uint32_t seed, uint32_t target_count,
                        rtb_sphere* spheres, size_t sphere_capacity,
                        rtb_material* materials, size_t material_capacity,
                        rtbh_scene_info* info) {
  SceneBuild sb;
  switch (scene_id) {
    case RTBH_SCENE_THREE_SPHERES:
      scene_three_spheres(sb);
      break;
    case RTBH_SCENE_FINAL: {
      sb.info.camera = rtbh_camera{{12.3f, 1.98f, -2.99f}, {11.342338f, 1.8494737f, -2.7333953f}, 0.0f, 20.0f};
      scene_final_named(sb);
      URandom rng(seed);
      dart_throw(sb, rng, RandomGroup{});
      break;
    }
    case RTBH_SCENE_STRESS: {
      // BASELINE config 5: same generator, 110x110 spread, draw until target_count accepted,
      // camera pulled back (SURVEY.md §8d "Scene C")
      sb.info.camera = rtbh_camera{{60.0f, 10.0f, -15.0f}, {0.0f, 0.0f, 0.0f}, 0.1f, 20.0f};
      scene_final_named(sb);
      URandom rng(seed);
      RandomGroup g;
      g.spread_x = 110.0f; g.spread_z = 110.0f;
      g.stop_after_accepted = target_count ? target_count : 10000;
      dart_throw(sb, rng, g);
      break;
    }
    default:
      return RTB_ERR_INVALID_ARGUMENT;
  }
  sb.info.sphere_count = (uint32_t)sb.spheres.size();
  sb.info.material_count = (uint32_t)sb.materials.size();
  if (info) *info = sb.info;
  if (spheres) {
    if (sphere_capacity < sb.spheres.size()) return RTB_ERR_INVALID_ARGUMENT;
    std::memcpy(spheres, sb.spheres.data(), sb.spheres.size() * sizeof(rtb_sphere));
  }
  if (materials) {
    if (material_capacity < sb.materials.size()) return RTB_ERR_INVALID_ARGUMENT;
    std::memcpy(materials, sb.materials.data(), sb.materials.size() * sizeof(rtb_material));
  }
  return RTB_OK;
}

int rtbh_build_bvh(const rtb_sphere* spheres, size_t sphere_count, int max_depth,
                   rtb_sphere* out_spheres, size_t out_sphere_capacity,
                   rtb_bvh_node* out_nodes, size_t node_capacity, size_t* out_node_count) {
  if ((!spheres && sphere_count) || !out_nodes || !out_node_count || max_depth < 0) return RTB_ERR_INVALID_ARGUMENT;
  if (out_sphere_capacity < sphere_count) return RTB_ERR_INVALID_ARGUMENT;
  std::vector<BuildEntity> ents(sphere_count);
  for (size_t i = 0; i < sphere_count; i++) ents[i] = BuildEntity{(uint32_t)i, sphere_world_bounds(spheres[i])};
  Builder b;
  b.spheres = spheres;
  b.max_depth = max_depth;
  b.nodes.reserve(sphere_count * 2 + 1);
  b.nodes.emplace_back();                       // BvhNodes.AddNoResize(default); BvhNodes[0] = new BvhNodeData(...)
  b.build(0, ents.data(), (int)sphere_count, 0, -1);
  if (node_capacity < b.nodes.size()) return RTB_ERR_INVALID_ARGUMENT;
  b.next_index = (int)b.nodes.size() - 1;
  b.flatten(0, out_nodes);
  for (size_t i = 0; i < b.bvh_entities.size(); i++) out_spheres[i] = spheres[b.bvh_entities[i]];
  *out_node_count = b.nodes.size();
  return RTB_OK;
}

int rtbh_build_bvh_from_bounds(const float* bounds, size_t entity_count, int max_depth, uint32_t* out_order,
                               size_t order_capacity, rtb_bvh_node* out_nodes, size_t node_capacity, size_t* out_node_count) {
  if ((!bounds && entity_count) || !out_nodes || !out_node_count || max_depth < 0) return RTB_ERR_INVALID_ARGUMENT;
  if (order_capacity < entity_count || (entity_count && !out_order)) return RTB_ERR_INVALID_ARGUMENT;
  std::vector<BuildEntity> ents(entity_count);
  for (size_t i = 0; i < entity_count; i++) {
    const float* b = bounds + 6 * i;
    ents[i] = BuildEntity{(uint32_t)i, Aabb{um::mk(b[0], b[1], b[2]), um::mk(b[3], b[4], b[5])}};
  }
  Builder b;
  b.spheres = nullptr;
  b.max_depth = max_depth;
  b.worker_budget = build_threads();
  b.nodes.reserve(entity_count * 2 + 1);
  b.nodes.emplace_back();
  b.build(0, ents.data(), (int)entity_count, 0, -1);
  if (node_capacity < b.nodes.size()) return RTB_ERR_INVALID_ARGUMENT;
  b.next_index = (int)b.nodes.size() - 1;
  b.flatten(0, out_nodes);
  for (size_t i = 0; i < b.bvh_entities.size(); i++) out_order[i] = b.bvh_entities[i];
  *out_node_count = b.nodes.size();
  return RTB_OK;
}

void rtbh_sphere_bounds(const rtb_sphere* sphere, float out_bounds[6]) {
  const Aabb a = sphere_world_bounds(*sphere);
  out_bounds[0] = a.mn.x; out_bounds[1] = a.mn.y; out_bounds[2] = a.mn.z;
  out_bounds[3] = a.mx.x; out_bounds[4] = a.mx.y; out_bounds[5] = a.mx.z;
}

void rtbh_triangle_bounds(const rtb_triangle* t, float out_bounds[6]) {
  // Triangle.Bounds (Triangle.cs:38-49): vertices -/+ |vertex normal| * 0.001, min / max over the three
  const f3 d0 = um::mk(t->edge2[0], t->edge2[1], t->edge2[2]), d1 = um::mk(t->edge1[0], t->edge1[1], t->edge1[2]);
  const f3 d2 = um::mk(t->v0[0], t->v0[1], t->v0[2]);
  const f3 vertices[3] = {d2, d1 + d2, d0 + d2};
  f3 mn = um::mk(0.0f), mx = um::mk(0.0f);
  for (int i = 0; i < 3; i++) {
    const f3 n = um::mk(std::fabs(t->normals[i][0]), std::fabs(t->normals[i][1]), std::fabs(t->normals[i][2]));
    const f3 neg = vertices[i] - n * 0.001f, pos = vertices[i] + n * 0.001f;
    mn = i == 0 ? neg : um::min(mn, neg);
    mx = i == 0 ? pos : um::max(mx, pos);
  }
  out_bounds[0] = mn.x; out_bounds[1] = mn.y; out_bounds[2] = mn.z;
  out_bounds[3] = mx.x; out_bounds[4] = mx.y; out_bounds[5] = mx.z;
}

void rtbh_placed_bounds(const rtb_placed_entity* e, float out_bounds[6]) {
  // BvhBuildingEntity ctor (BvhNodeData.cs:28-80): the content's local bounds (Sphere.cs:16-23, Rect.cs:17-19,
  // Box.cs:17), their 8 corners through OriginTransform; a moving entity takes the minimum through the transform
  // at min(origin, destination) and the maximum through the one at max(origin, destination) (:58-70)
  f3 lo, hi;
  if (e->type == RTB_ENTITY_SPHERE) {
    const float ar = std::fabs(e->size[0]);
    lo = um::mk(-ar); hi = um::mk(ar);
  } else if (e->type == RTB_ENTITY_RECT) {
    lo = um::mk(um::div(-e->size[0], 2.0f), um::div(-e->size[1], 2.0f), -0.001f);
    hi = um::mk(um::div(e->size[0], 2.0f), um::div(e->size[1], 2.0f), 0.001f);
  } else {
    hi = um::mk(um::div(e->size[0], 2.0f), um::div(e->size[1], 2.0f), um::div(e->size[2], 2.0f));
    lo = -hi;
  }
  um::rigid a, b;
  a.rot.x = e->rotation[0]; a.rot.y = e->rotation[1]; a.rot.z = e->rotation[2]; a.rot.w = e->rotation[3];
  a.pos = um::mk(e->position[0], e->position[1], e->position[2]);
  b = a;
  if (e->moving) {
    const f3 dest = a.pos + um::mk(e->destination_offset[0], e->destination_offset[1], e->destination_offset[2]);
    const f3 origin = a.pos;
    a.pos = um::min(origin, dest);
    b.pos = um::max(origin, dest);
  }
  f3 mn = um::mk(std::numeric_limits<float>::infinity());
  f3 mx = um::mk(-std::numeric_limits<float>::infinity());
  // corner order of BvhNodeData.cs:43-53 (min / max are order-independent; kept for readability)
  const int order[8][3] = {{0, 0, 0}, {0, 0, 1}, {0, 1, 0}, {1, 0, 0}, {0, 1, 1}, {1, 1, 0}, {1, 0, 1}, {1, 1, 1}};
  for (int i = 0; i < 8; i++) {
    const f3 c = um::mk(order[i][0] ? hi.x : lo.x, order[i][1] ? hi.y : lo.y, order[i][2] ? hi.z : lo.z);
    mn = um::min(mn, um::transform(a, c));
    mx = um::max(mx, um::transform(b, c));
  }
  out_bounds[0] = mn.x; out_bounds[1] = mn.y; out_bounds[2] = mn.z;
  out_bounds[3] = mx.x; out_bounds[4] = mx.y; out_bounds[5] = mx.z;
}

void rtbh_make_triangle(const float v1[3], const float v2[3], const float v3[3], const float* n1, const float* n2,
                        const float* n3, uint32_t material, rtb_triangle* out) {
  // Data = float3x3(v3 - v1, v2 - v1, v1) (Triangle.cs:16,25)
  const f3 a = um::mk(v1[0], v1[1], v1[2]), b = um::mk(v2[0], v2[1], v2[2]), c = um::mk(v3[0], v3[1], v3[2]);
  const f3 d0 = c - a, d1 = b - a;
  f3 n[3];
  if (n1 && n2 && n3) {
    n[0] = um::normalize(um::mk(n1[0], n1[1], n1[2]));
    n[1] = um::normalize(um::mk(n2[0], n2[1], n2[2]));
    n[2] = um::normalize(um::mk(n3[0], n3[1], n3[2]));
  } else {
    n[0] = n[1] = n[2] = um::normalize(um::cross(d1, d0));   // faceNormal = normalize(cross(Data[1], Data[0]))
  }
  *out = rtb_triangle{};
  out->edge2[0] = d0.x; out->edge2[1] = d0.y; out->edge2[2] = d0.z;
  out->edge1[0] = d1.x; out->edge1[1] = d1.y; out->edge1[2] = d1.z;
  out->v0[0] = a.x; out->v0[1] = a.y; out->v0[2] = a.z;
  for (int i = 0; i < 3; i++) { out->normals[i][0] = n[i].x; out->normals[i][1] = n[i].y; out->normals[i][2] = n[i].z; }
  out->material = material;
}

int rtbh_add_mesh(const float* vertices, const float* normals, const float* uvs, size_t vertex_count, const uint16_t* indices,
                  size_t index_count, const float rotation[4], const float position[3], float scale, uint32_t material,
                  rtb_triangle* out_triangles, float* out_uvs, size_t triangle_capacity, size_t* out_triangle_count) {
  // AddMeshRuntimeEntitiesJob.Execute (AddMeshRuntimeEntitiesJob.cs:30-96): per index triple, bake the transform into the
  // vertices (transform(RigidTransform, vertex * Scale)) and rotate the vertex normals (mul(RigidTransform.rot, normal)),
  // then the face-normal or vertex-normal Triangle ctor; uvs pass through (default = 0 without TexCoord0)
  if (!vertices || !indices || !rotation || !position || !out_triangles || !out_triangle_count) return RTB_ERR_INVALID_ARGUMENT;
  if (index_count % 3 != 0 || triangle_capacity < index_count / 3) return RTB_ERR_INVALID_ARGUMENT;
  um::rigid xf;
  xf.rot.x = rotation[0]; xf.rot.y = rotation[1]; xf.rot.z = rotation[2]; xf.rot.w = rotation[3];
  xf.pos = um::mk(position[0], position[1], position[2]);
  size_t n = 0;
  for (size_t i = 0; i < index_count; i += 3) {
    float wv[3][3], wn[3][3];
    for (int j = 0; j < 3; j++) {
      const size_t k = indices[i + j];
      if (k >= vertex_count) return RTB_ERR_INVALID_ARGUMENT;
      const f3 v = um::transform(xf, um::mk(vertices[3 * k], vertices[3 * k + 1], vertices[3 * k + 2]) * scale);
      wv[j][0] = v.x; wv[j][1] = v.y; wv[j][2] = v.z;
      if (normals) {
        const f3 nn = um::rotate(xf.rot, um::mk(normals[3 * k], normals[3 * k + 1], normals[3 * k + 2]));
        wn[j][0] = nn.x; wn[j][1] = nn.y; wn[j][2] = nn.z;
      }
      if (out_uvs) {
        out_uvs[6 * n + 2 * j] = uvs ? uvs[2 * k] : 0.0f;
        out_uvs[6 * n + 2 * j + 1] = uvs ? uvs[2 * k + 1] : 0.0f;
      }
    }
    rtbh_make_triangle(wv[0], wv[1], wv[2], normals ? wn[0] : nullptr, normals ? wn[1] : nullptr, normals ? wn[2] : nullptr, material,
                       &out_triangles[n]);
    n++;
  }
  *out_triangle_count = n;
  return RTB_OK;
}

void rtbh_make_view(const float origin[3], const float look_at[3], const float up_in[3],
                    float vertical_fov_degrees, float aspect, float aperture,
                    float focus_distance, rtb_view* out) {
  // View.cs:16-36.  Runs in managed C# in the reference (plain float ops, MathF.Tan); we use
  // um::tan so the result does not depend on the box's libm.
  f3 o = um::mk(origin[0], origin[1], origin[2]);
  f3 la = um::mk(look_at[0], look_at[1], look_at[2]);
  f3 up = um::mk(up_in[0], up_in[1], up_in[2]);
  float lens_radius = aperture / 2;
  float theta = vertical_fov_degrees * um::PI / 180;
  float half_height = um::tan(theta / 2);
  float half_width = aspect * half_height;
  f3 forward = um::normalize(o - la);
  f3 right = um::normalize(um::cross(forward, up));
  f3 upv = um::cross(right, forward);
  f3 llc = (half_width * focus_distance) * (-right) + (half_height * focus_distance) * (-upv) + focus_distance * (-forward);
  f3 horizontal = (2 * half_width * focus_distance) * right;
  f3 vertical = (2 * half_height * focus_distance) * upv;
  auto put = [](float* d, f3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; };
  put(out->origin, o);
  put(out->lower_left_corner, llc);
  put(out->horizontal, horizontal);
  put(out->vertical, vertical);
  put(out->forward, forward);
  put(out->up, upv);
  put(out->right, right);
  out->lens_radius = lens_radius;
}

int rtbh_hit_world(const rtb_bvh_node* nodes, size_t node_count, const rtb_sphere* spheres, size_t sphere_count,
                   const float origin[3], const float direction[3], float* out_distance) {
  (void)sphere_count;
  if (!nodes || node_count == 0) return 0;
  float t = 0;
  bool hit = node_hit(nodes, spheres, 0, um::mk(origin[0], origin[1], origin[2]),
                      um::mk(direction[0], direction[1], direction[2]), &t);
  if (hit && out_distance) *out_distance = t;
  return hit ? 1 : 0;
}

void rtbh_view_from_camera(const rtbh_camera* camera, float aspect,
                           const rtb_bvh_node* nodes, size_t node_count,
                           const rtb_sphere* spheres, size_t sphere_count,
                           float fallback_focus, rtb_view* out, float* out_focus_distance) {
  // Raytracer.cs:604-612: origin = camera position, forward = transform.forward,
  // lookAt = origin + forward, focusDistance = first hit along forward (else previous value).
  f3 pos = um::mk(camera->position[0], camera->position[1], camera->position[2]);
  f3 tgt = um::mk(camera->target[0], camera->target[1], camera->target[2]);
  f3 fwd = um::normalize(tgt - pos);
  f3 look_at = pos + fwd;
  float focus = fallback_focus;
  float fo[3] = {pos.x, pos.y, pos.z}, fd[3] = {fwd.x, fwd.y, fwd.z};
  float t;
  if (rtbh_hit_world(nodes, node_count, spheres, sphere_count, fo, fd, &t)) focus = t;
  float la[3] = {look_at.x, look_at.y, look_at.z};
  float up[3] = {0.0f, 1.0f, 0.0f};
  rtbh_make_view(fo, la, up, camera->vertical_fov, aspect, camera->aperture, focus, out);
  if (out_focus_distance) *out_focus_distance = focus;
}

int rtbh_space_filling_series(int length, int32_t* out, size_t capacity) {
  // Tools.SpaceFillingSeries (Tools.cs:101-124)
  if (length <= 0 || !out || capacity < (size_t)length) return RTB_ERR_INVALID_ARGUMENT;
  int current = 0, n = 0;
  std::set<int> seen;
  do {
    int divider = 2;
    do {
      int increment = (int)std::ceil((float)length / divider);
      for (int i = 0; i < divider; i++) {
        current = i * increment;
        if (!seen.count(current)) break;
      }
      divider *= 2;
    } while (seen.count(current));
    out[n++] = current;
    seen.insert(current);
  } while ((int)seen.size() < length);
  return RTB_OK;
}

}  // extern "C"
