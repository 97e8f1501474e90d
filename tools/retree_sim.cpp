// retree_sim.cpp — CPU estimate of what another BVH topology is worth to the device walk (no GPU needed).
//   g++ -O2 -std=c++17 -Iinclude -Iraytracing-in-one-weekend_b200/csrc tools/retree_sim.cpp \
//       -Lraytracing-in-one-weekend_b200/lib -lrtb_host -Wl,-rpath,$PWD/raytracing-in-one-weekend_b200/lib -o /tmp/retree_sim
// Traces pseudo-random paths (simplified materials: the ray distribution is what matters) through the book-1 final scene with
// the kernel's walk order (near child first, a box beyond the best hit is skipped) on the reference's tree and on the
// re-built one, and prints box visits and leaf tests per ray, and per 32 rays the longest walk (what a warp waits for).
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "retree.hpp"
#include "rtb_host.h"

struct V { float x, y, z; };
static V operator+(V a, V b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static V operator-(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static V operator*(V a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static float dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static V norm(V a) { return a * (1.0f / std::sqrt(dot(a, a))); }

struct Stats { double visits = 0, leaves = 0, rays = 0, trips = 0; };

static void slab(const rtb_bvh_node& b, V o, V inv, float* tn, float* tx) {
  const float ox[3] = {o.x, o.y, o.z}, iv[3] = {inv.x, inv.y, inv.z};
  float t0 = 0.0f, t1 = INFINITY;
  for (int k = 0; k < 3; k++) {
    const float a = (b.bounds_min[k] - ox[k]) * iv[k], c = (b.bounds_max[k] - ox[k]) * iv[k];
    t0 = std::fmax(t0, std::fmin(a, c));
    t1 = std::fmin(t1, std::fmax(a, c));
  }
  *tn = t0;
  *tx = t1;
}

static int walk(const std::vector<rtb_bvh_node>& nodes, const std::vector<rtb_sphere>& sph, V o, V d, float* best_t, int* trips, Stats& st) {
  const V inv{1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
  int best = -1;
  *best_t = INFINITY;
  int stack[128], top = 0, cur = 0, n_trips = 0;
  const float a = dot(d, d);
  for (;;) {
    n_trips++;
    bool pop = false;
    const rtb_bvh_node& nd = nodes[cur];
    if (nd.first_entity < 0) {
      st.visits++;
      float tl, xl, tr, xr;
      slab(nodes[nd.left], o, inv, &tl, &xl);
      slab(nodes[nd.right], o, inv, &tr, &xr);
      const float limit = *best_t * 1.0005f;
      const bool hl = tl < std::fmin(xl, limit), hr = tr < std::fmin(xr, limit);
      if (hl && hr) {
        const bool lf = tl <= tr;
        stack[top++] = lf ? nd.right : nd.left;
        cur = lf ? nd.left : nd.right;
      } else if (hl || hr) cur = hl ? nd.left : nd.right;
      else pop = true;
    }
    if (!pop && nodes[cur].first_entity >= 0) {
      const rtb_bvh_node& lf = nodes[cur];
      for (int i = 0; i < lf.entity_count; i++) {
        st.leaves++;
        const rtb_sphere& s = sph[lf.first_entity + i];
        const V oc = o - V{s.center[0], s.center[1], s.center[2]};
        const float b = dot(oc, d), c = dot(oc, oc) - s.radius * s.radius, disc = b * b - a * c;
        if (disc > 0) {
          const float sq = std::sqrt(disc);
          float t = (-b - sq) / a;
          if (!(t > 1e-4f)) t = (-b + sq) / a;
          if (t > 1e-4f && t < *best_t) { *best_t = t; best = lf.first_entity + i; }
        }
      }
      pop = true;
    }
    if (pop) {
      if (top == 0) break;
      cur = stack[--top];
    }
  }
  *trips = n_trips;
  st.rays++;
  st.trips += n_trips;
  return best;
}

int main(int argc, char** argv) {
  const int scene_id = argc > 1 ? atoi(argv[1]) : RTBH_SCENE_FINAL;
  const uint32_t target = argc > 2 ? (uint32_t)atoi(argv[2]) : 0;
  const int n_paths = argc > 3 ? atoi(argv[3]) : 40000;
  rtbh_scene_info info;
  rtbh_scene_generate(scene_id, 700, target, nullptr, 0, nullptr, 0, &info);
  std::vector<rtb_sphere> spheres(info.sphere_count), ordered(info.sphere_count);
  std::vector<rtb_material> mats(info.material_count);
  rtbh_scene_generate(scene_id, 700, target, spheres.data(), spheres.size(), mats.data(), mats.size(), &info);
  std::vector<rtb_bvh_node> ref(2 * spheres.size() + 1);
  size_t nn = 0;
  rtbh_build_bvh(spheres.data(), spheres.size(), 16, ordered.data(), ordered.size(), ref.data(), ref.size(), &nn);
  ref.resize(nn);
  std::vector<rtb_bvh_node> sah;
  const int passes = argc > 4 ? atoi(argv[4]) : RTB_RETREE_PASSES;
  const bool ok = rtb_retree::retree(ref.data(), ref.size(), 60, sah, passes);
  auto inner_area = [](const std::vector<rtb_bvh_node>& t) {
    double a = 0;
    for (const auto& n : t) if (n.first_entity < 0) a += ((double)n.bounds_max[0] - n.bounds_min[0]) * ((double)n.bounds_max[1] - n.bounds_min[1]) + ((double)n.bounds_max[1] - n.bounds_min[1]) * ((double)n.bounds_max[2] - n.bounds_min[2]) + ((double)n.bounds_max[2] - n.bounds_min[2]) * ((double)n.bounds_max[0] - n.bounds_min[0]);
    return a;
  };
  printf("sum of inner areas: reference %.4g, re-built %.4g (passes %d)\n", inner_area(ref), inner_area(sah), passes);
  printf("spheres %zu, reference nodes %zu, retree %s, nodes %zu\n", spheres.size(), ref.size(), ok ? "applied" : "NOT applied", sah.size());
  if (!ok) return 1;
  rtb_view view;
  float focus;
  rtbh_view_from_camera(&info.camera, 16.0f / 9.0f, ref.data(), ref.size(), ordered.data(), ordered.size(), 1.0f, &view, &focus);
  std::mt19937 rng(1);
  std::uniform_real_distribution<float> U(0.0f, 1.0f);
  auto unit = [&]() {
    for (;;) {
      V v{2 * U(rng) - 1, 2 * U(rng) - 1, 2 * U(rng) - 1};
      const float l = dot(v, v);
      if (l > 1e-4f && l <= 1.0f) return norm(v);
    }
  };
  Stats a, b;
  double warp_max_a = 0, warp_max_b = 0, warps = 0;
  int lane = 0, wa = 0, wb = 0;
  long mismatches = 0;
  for (int p = 0; p < n_paths; p++) {
    const float u = U(rng), v = U(rng);
    V o{view.origin[0], view.origin[1], view.origin[2]};
    V target_pt{view.lower_left_corner[0] + u * view.horizontal[0] + v * view.vertical[0],
                view.lower_left_corner[1] + u * view.horizontal[1] + v * view.vertical[1],
                view.lower_left_corner[2] + u * view.horizontal[2] + v * view.vertical[2]};
    V d = norm(target_pt - o);
    for (int depth = 0; depth < 50; depth++) {
      float ta, tb;
      int trips_a, trips_b;
      const int ha = walk(ref, ordered, o, d, &ta, &trips_a, a);
      const int hb = walk(sah, ordered, o, d, &tb, &trips_b, b);
      if (ha != hb) mismatches++;
      wa = std::max(wa, trips_a);
      wb = std::max(wb, trips_b);
      if (++lane == 32) { warp_max_a += wa; warp_max_b += wb; warps++; lane = 0; wa = wb = 0; }
      if (ha < 0) break;
      const rtb_sphere& s = ordered[ha];
      const V P = o + d * ta;
      V N = (P - V{s.center[0], s.center[1], s.center[2]}) * (1.0f / s.radius);
      const rtb_material& m = mats[s.material];
      V nd;
      if (m.type == RTB_MATERIAL_DIELECTRIC) {
        const bool entering = dot(d, N) < 0;
        const V n = entering ? N : N * -1.0f;
        const float eta = entering ? 1.0f / m.index_of_refraction : m.index_of_refraction;
        const float cosi = -dot(d, n), k = 1 - eta * eta * (1 - cosi * cosi);
        if (k < 0 || U(rng) < 0.1f) nd = d + n * (2 * cosi);
        else nd = d * eta + n * (eta * cosi - std::sqrt(k));
        N = dot(nd, N) >= 0 ? N : N * -1.0f;
      } else if (m.metallic > 0.5f) {
        nd = d - N * (2 * dot(d, N)) + unit() * (1.0f - m.glossiness);
        if (dot(nd, N) <= 0) break;
      } else {
        nd = N + unit();
      }
      o = P + N * 0.001f;
      d = norm(nd);
    }
  }
  printf("rays %.0f  hit mismatches %ld\n", a.rays, mismatches);
  printf("reference tree: box visits / ray %.2f  leaf tests / ray %.2f  trips / ray %.2f  longest of 32: %.2f\n", a.visits / a.rays, a.leaves / a.rays, a.trips / a.rays, warp_max_a / warps);
  printf("re-built tree : box visits / ray %.2f  leaf tests / ray %.2f  trips / ray %.2f  longest of 32: %.2f\n", b.visits / b.rays, b.leaves / b.rays, b.trips / b.rays, warp_max_b / warps);
  return 0;
}
