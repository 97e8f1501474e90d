// bvh4_sim.cpp — CPU estimate: visits / box tests / trips per ray for a 4-wide collapse of the shipped re-built tree against the binary walk.
//   g++ -O2 -std=c++17 -Iinclude -Iraytracing-in-one-weekend_b200/csrc tools/bvh4_sim.cpp -Lraytracing-in-one-weekend_b200/lib -lrtb_host -Wl,-rpath,$PWD/raytracing-in-one-weekend_b200/lib -o /tmp/bvh4_sim && /tmp/bvh4_sim 1 0 30000
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include <algorithm>
#include "retree.hpp"
#include "rtb_host.h"
struct V { float x, y, z; };
static V operator+(V a, V b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static V operator-(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static V operator*(V a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static float dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static V norm(V a) { return a * (1.0f / std::sqrt(dot(a, a))); }
static void slab(const rtb_bvh_node& b, V o, V inv, float* tn, float* tx) {
  const float ox[3] = {o.x, o.y, o.z}, iv[3] = {inv.x, inv.y, inv.z};
  float t0 = 0.0f, t1 = INFINITY;
  for (int k = 0; k < 3; k++) {
    const float a = (b.bounds_min[k] - ox[k]) * iv[k], c = (b.bounds_max[k] - ox[k]) * iv[k];
    t0 = std::fmax(t0, std::fmin(a, c)); t1 = std::fmin(t1, std::fmax(a, c));
  }
  *tn = t0; *tx = t1;
}
struct Wide { int child[4]; int n; };
std::vector<Wide> wide;   // per binary inner node index (only those that are wide roots are used)
static double area(const rtb_bvh_node& n){double dx=n.bounds_max[0]-n.bounds_min[0],dy=n.bounds_max[1]-n.bounds_min[1],dz=n.bounds_max[2]-n.bounds_min[2];return dx*dy+dy*dz+dz*dx;}
static void collapse(const std::vector<rtb_bvh_node>& t, int i) {
  Wide w; w.n = 2; w.child[0] = t[i].left; w.child[1] = t[i].right;
  while (w.n < 4) {
    int best = -1; double ba = -1;
    for (int k = 0; k < w.n; k++) if (t[w.child[k]].first_entity < 0 && area(t[w.child[k]]) > ba) { ba = area(t[w.child[k]]); best = k; }
    if (best < 0) break;
    int c = w.child[best];
    w.child[best] = t[c].left; w.child[w.n++] = t[c].right;
  }
  wide[i] = w;
  for (int k = 0; k < w.n; k++) if (t[w.child[k]].first_entity < 0) collapse(t, w.child[k]);
}
struct St { double visits=0, boxes=0, leaves=0, rays=0, trips=0; };
static int sphere_test(const std::vector<rtb_sphere>& sph, const rtb_bvh_node& lf, V o, V d, float a, float* best_t, int best, St& st) {
  for (int i = 0; i < lf.entity_count; i++) {
    st.leaves++;
    const rtb_sphere& s = sph[lf.first_entity + i];
    const V oc = o - V{s.center[0], s.center[1], s.center[2]};
    const float b = dot(oc, d), c = dot(oc, oc) - s.radius * s.radius, disc = b * b - a * c;
    if (disc > 0) { const float sq = std::sqrt(disc); float t = (-b - sq) / a; if (!(t > 1e-4f)) t = (-b + sq) / a;
      if (t > 1e-4f && t < *best_t) { *best_t = t; best = lf.first_entity + i; } }
  }
  return best;
}
static int walk2(const std::vector<rtb_bvh_node>& nodes, const std::vector<rtb_sphere>& sph, V o, V d, float* best_t, St& st) {
  const V inv{1.0f / d.x, 1.0f / d.y, 1.0f / d.z}; int best = -1; *best_t = INFINITY; int stack[128], top = 0, cur = 0; const float a = dot(d, d);
  for (;;) { st.trips++; bool pop = false; const rtb_bvh_node& nd = nodes[cur];
    if (nd.first_entity < 0) { st.visits++; st.boxes += 2; float tl, xl, tr, xr; slab(nodes[nd.left], o, inv, &tl, &xl); slab(nodes[nd.right], o, inv, &tr, &xr);
      const float limit = *best_t * 1.0005f; const bool hl = tl < std::fmin(xl, limit), hr = tr < std::fmin(xr, limit);
      if (hl && hr) { const bool lf = tl <= tr; stack[top++] = lf ? nd.right : nd.left; cur = lf ? nd.left : nd.right; } else if (hl || hr) cur = hl ? nd.left : nd.right; else pop = true; }
    if (!pop && nodes[cur].first_entity >= 0) { best = sphere_test(sph, nodes[cur], o, d, a, best_t, best, st); pop = true; }
    if (pop) { if (top == 0) break; cur = stack[--top]; } }
  st.rays++; return best;
}
static int walk4(const std::vector<rtb_bvh_node>& nodes, const std::vector<rtb_sphere>& sph, V o, V d, float* best_t, St& st) {
  const V inv{1.0f / d.x, 1.0f / d.y, 1.0f / d.z}; int best = -1; *best_t = INFINITY; int stack[256], top = 0, cur = 0; const float a = dot(d, d);
  for (;;) { st.trips++; bool pop = false;
    if (nodes[cur].first_entity < 0) { st.visits++; const Wide& w = wide[cur]; st.boxes += 4; float t[4]; int c[4]; int n = 0; const float limit = *best_t * 1.0005f;
      for (int k = 0; k < w.n; k++) { float tn, tx; slab(nodes[w.child[k]], o, inv, &tn, &tx); if (tn < std::fmin(tx, limit)) { t[n] = tn; c[n] = w.child[k]; n++; } }
      for (int i = 1; i < n; i++) for (int j = i; j > 0 && t[j] < t[j-1]; j--) { std::swap(t[j], t[j-1]); std::swap(c[j], c[j-1]); }
      if (n == 0) pop = true; else { for (int k = n - 1; k >= 1; k--) stack[top++] = c[k]; cur = c[0]; } }
    if (!pop && nodes[cur].first_entity >= 0) { best = sphere_test(sph, nodes[cur], o, d, a, best_t, best, st); pop = true; }
    if (pop) { if (top == 0) break; cur = stack[--top]; } }
  st.rays++; return best;
}
int main(int argc, char** argv) {
  const int scene_id = atoi(argv[1]); const uint32_t target = atoi(argv[2]); const int n_paths = atoi(argv[3]);
  rtbh_scene_info info; rtbh_scene_generate(scene_id, 700, target, nullptr, 0, nullptr, 0, &info);
  std::vector<rtb_sphere> spheres(info.sphere_count), ordered(info.sphere_count); std::vector<rtb_material> mats(info.material_count);
  rtbh_scene_generate(scene_id, 700, target, spheres.data(), spheres.size(), mats.data(), mats.size(), &info);
  std::vector<rtb_bvh_node> ref(2 * spheres.size() + 1); size_t nn = 0;
  rtbh_build_bvh(spheres.data(), spheres.size(), 16, ordered.data(), ordered.size(), ref.data(), ref.size(), &nn); ref.resize(nn);
  std::vector<rtb_bvh_node> sah; rtb_retree::retree(ref.data(), ref.size(), 60, sah, 8);
  wide.resize(sah.size()); collapse(sah, 0);
  rtb_view view; float focus; rtbh_view_from_camera(&info.camera, 16.0f / 9.0f, ref.data(), ref.size(), ordered.data(), ordered.size(), 1.0f, &view, &focus);
  std::mt19937 rng(1); std::uniform_real_distribution<float> U(0.0f, 1.0f);
  auto unit = [&]() { for (;;) { V v{2 * U(rng) - 1, 2 * U(rng) - 1, 2 * U(rng) - 1}; const float l = dot(v, v); if (l > 1e-4f && l <= 1.0f) return norm(v); } };
  St a, b; long mism = 0;
  for (int p = 0; p < n_paths; p++) {
    const float u = U(rng), v = U(rng); V o{view.origin[0], view.origin[1], view.origin[2]};
    V tp{view.lower_left_corner[0] + u * view.horizontal[0] + v * view.vertical[0], view.lower_left_corner[1] + u * view.horizontal[1] + v * view.vertical[1], view.lower_left_corner[2] + u * view.horizontal[2] + v * view.vertical[2]};
    V d = norm(tp - o);
    for (int depth = 0; depth < 50; depth++) {
      float ta, tb; const int ha = walk2(sah, ordered, o, d, &ta, a); const int hb = walk4(sah, ordered, o, d, &tb, b); if (ha != hb) mism++;
      if (ha < 0) break; const rtb_sphere& s = ordered[ha]; const V P = o + d * ta; V N = (P - V{s.center[0], s.center[1], s.center[2]}) * (1.0f / s.radius);
      const rtb_material& m = mats[s.material]; V nd;
      if (m.type == RTB_MATERIAL_DIELECTRIC) { const bool e = dot(d, N) < 0; const V n = e ? N : N * -1.0f; const float eta = e ? 1.0f / m.index_of_refraction : m.index_of_refraction; const float ci = -dot(d, n), k = 1 - eta * eta * (1 - ci * ci);
        if (k < 0 || U(rng) < 0.1f) nd = d + n * (2 * ci); else nd = d * eta + n * (eta * ci - std::sqrt(k)); N = dot(nd, N) >= 0 ? N : N * -1.0f; }
      else if (m.metallic > 0.5f) { nd = d - N * (2 * dot(d, N)) + unit() * (1.0f - m.glossiness); if (dot(nd, N) <= 0) break; }
      else nd = N + unit();
      o = P + N * 0.001f; d = norm(nd);
    }
  }
  printf("mismatch %ld\nbinary: visits %.2f boxes %.2f leaf tests %.2f trips %.2f\n4-wide: visits %.2f boxes %.2f leaf tests %.2f trips %.2f\n", mism, a.visits/a.rays, a.boxes/a.rays, a.leaves/a.rays, a.trips/a.rays, b.visits/b.rays, b.boxes/b.rays, b.leaves/b.rays, b.trips/b.rays);
}
