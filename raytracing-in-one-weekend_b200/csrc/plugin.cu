// plugin.cu — librtb.so: the C ABI of include/rtb.h over the sm_100a kernels.
//
// Replaces, behind one blocking call, what the reference's host does around the sample job
// (Unity/Raytracer.cs:671-738): fill the job struct, schedule it over W*H pixels, wait.
// Style precedent for the exports: OptixDenoiser/OptixDenoiser/OptixDenoiser.h:1-10 (extern
// "C", int status codes) and Runtime/Jobs/DenoiseJobs.cs:10-39 (a blocking native call made
// from a worker thread).  No CPU fallback: without a CUDA device every entry point fails.

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "aux_kernels.cuh"
#include "sample_kernels.cuh"
#include "volume_kernel.cuh"
#include "retree.hpp"

using namespace rtbk;

// csrc/fast_kernels.cu: the lean sphere megakernel compiled with fast arithmetic (RTB_OPT_MATH = 1)
extern "C" int rtb_fast_launch_spheres(const void* args, size_t args_bytes, int scene_in_smem, int flavor, unsigned grid, size_t smem,
                                       int max_smem_optin, void* stream);

namespace {

thread_local std::string g_thread_error;

struct DeviceBuffers {
  size_t capacity = 0;  // pixels
  float *in_color = nullptr, *in_weight = nullptr, *in_normal = nullptr, *in_albedo = nullptr;
  float *out_color = nullptr, *out_weight = nullptr, *out_normal = nullptr, *out_albedo = nullptr;
  rtb_diagnostics* diagnostics = nullptr;
};

}  // namespace

struct rtb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
  std::string last_error;
  rtb_log_fn log_fn = nullptr;
  void* log_user = nullptr;
  int sm_count = 0;
  int max_smem_optin = 0;

  // scene
  unsigned char* d_blob = nullptr;
  DevMaterial* d_materials = nullptr;
  uint16_t* d_sky = nullptr;
  int sky_w = 0, sky_h = 0;
  std::vector<DevMaterial> host_materials;   // as uploaded (before derive_materials_kernel), for rtb_upload_textures
  unsigned char* d_tex_pixels = nullptr;
  int4* d_tex_images = nullptr;
  int4* d_mat_textures = nullptr;
  float2* d_tri_uv = nullptr;
  uint32_t* d_chain_ref = nullptr;
  float4* d_chain_boxes = nullptr;
  SceneDesc scene{};
  bool has_scene = false;

  // work counters / options
  // One tile counter per launch, taken round-robin from a ring: batches enqueued on different streams (or a device
  // batch overlapping rtb_sample_batch) never share a counter.  (A counter is reused after kTileRing launches; that
  // many batches of one context are never in flight at once.)
  static constexpr uint32_t kTileRing = 256;
  uint32_t* d_tile_ring = nullptr;
  std::atomic<uint32_t> ring_next{0};
  // CancellationToken relay (BatchArgs::cancel_flag / cancel_epoch): the kernels poll a device word; the blocking call
  // watches the caller's token while it waits and, when it is set, copies the batch's epoch into that word on cancel_stream.
  uint32_t* d_cancel = nullptr;       // device word
  uint32_t* h_epoch = nullptr;        // pinned source of the cancel copy (first word of a 256-byte pinned block)
  // Pinned landing places of the per-batch read-backs.  They must NOT be pageable: an "asynchronous" device-to-host copy into
  // pageable memory blocks the calling thread until the stream reaches it — i.e. until the kernel has finished —, and the
  // thread that is to relay the CancellationToken (wait_for_streams) would only start watching it afterwards (found in
  // round 2: batches of worlds with media could not be cancelled in flight).
  uint32_t* h_status = nullptr;       // h_epoch + 4
  rtb_counters* h_counters = nullptr; // (char*)h_epoch + 64
  uint32_t cancel_epoch = 0;          // epoch of the batch in flight (never 0: the word starts as 0)
  bool cancel_sent = false;           // this batch's cancel copy was issued
  cudaStream_t cancel_stream = nullptr;
  uint32_t* d_status = nullptr;       // sticky kStatus* bits raised by kernels
  cudaEvent_t ev_done = nullptr;
  unsigned long long* d_counters = nullptr;
  cudaStream_t counters_stream = nullptr;   // stream of the last instrumented launch
  rtb_counters counters{};
  int default_kernel = 2;
  int64_t opt_counters = 0, opt_kernel = 0, opt_collapse = kDefaultCollapse, opt_walk_chains = 0, opt_host_access = 1, opt_noise = 0, opt_math = 0, opt_retree = 1;
  bool last_in_place = false;
  float last_ms = 0.0f;
  bool smem_attr_set[2][12] = {};

  DeviceBuffers buf;
  MetricsAcc* d_metrics_partial = nullptr;
  rtb_diagnostics* d_scratch_diag = nullptr;
  size_t scratch_diag_capacity = 0;
  std::map<void*, size_t> registered;
  std::mutex mu;
};

namespace {

int fail(rtb_ctx* ctx, int code, const char* fmt, ...) {
  char msg[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(msg, sizeof msg, fmt, ap);
  va_end(ap);
  if (ctx) {
    ctx->last_error = msg;
    if (ctx->log_fn) ctx->log_fn(2, msg, ctx->log_user);
  }
  g_thread_error = msg;
  return code;
}

#define RTB_CUDA(ctx, expr)                                                                           \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess)                                                                            \
      return fail(ctx, _e == cudaErrorMemoryAllocation ? RTB_ERR_OUT_OF_MEMORY : RTB_ERR_CUDA + (int)_e, \
                  "%s failed: %s", #expr, cudaGetErrorString(_e));                                    \
  } while (0)

// ---- scene flattening: rtb_bvh_node[] (reference order, root = 0) -> device blob ----------
//
// The device tree is the reference tree with its bottom COLLAPSED: a subtree that holds at most
// `collapse_k` spheres becomes one leaf whose spheres are tested directly.  Spheres are re-laid
// out in depth-first leaf order, so every subtree owns a contiguous range.  The reference accepts
// a sphere hit only if the box of every node on the path root -> leaf is hit
// (SampleBatchJob.cs:403-448); the boxes a collapsed leaf no longer tests are kept per sphere in
// the "chain" arrays, and the kernel re-applies them, exactly, to the rare accepted hit whose
// geometry does not already prove they pass (kernel_common.cuh: chain_guard).
struct HostBlob {
  std::vector<unsigned char> bytes;
  std::vector<DevMaterial> materials;
  std::vector<uint32_t> chain_ref;     // per device sphere: first chain box | count << 24
  std::vector<float> chain_boxes;      // 8 floats per box: min.xyz, -, max.xyz, -
  SceneDesc desc{};
  bool retreed = false;                // the blob's tree is retree.hpp's, not the host's topology
};

struct Flattener {
  const rtb_bvh_node* nodes;
  size_t node_count, sphere_count;      // sphere_count: length of the list the leaves index (entities, or spheres)
  const rtb_sphere* spheres;
  const rtb_entity* entities = nullptr; // nullptr: entity i is sphere i
  uint32_t collapse_k;
  static bool is_placed_type(uint32_t t) { return (t & RTB_ENTITY_PLACED) || t == RTB_ENTITY_RECT || t == RTB_ENTITY_BOX; }
  bool is_placed(size_t e) const { return entities && is_placed_type(entities[e].type); }
  bool is_triangle(size_t e) const { return entities && entities[e].type == RTB_ENTITY_TRIANGLE; }
  bool not_plain_sphere(size_t e) const { return is_triangle(e) || is_placed(e); }
  const rtb_sphere& sphere_of(size_t e) const { return spheres[entities ? entities[e].index : e]; }
  std::vector<uint8_t> visited;
  std::vector<uint8_t> entity_is_medium;   // per leaf-list entity: it wears a ProbabilisticVolume (empty: the world has no media)
  std::vector<uint8_t> subtree_media;      // per reference node: some entity below it wears a medium
  std::vector<uint32_t> subtree_spheres;   // per reference node
  std::vector<float> inner;                // 16 floats per device inner node
  std::vector<uint32_t> order;             // device sphere -> host sphere
  std::vector<uint32_t> leaf_count;        // per device sphere: count of the device leaf starting here
  std::vector<uint32_t> chain_ref;
  std::vector<float> chain_boxes;
  std::vector<int32_t> path;               // reference nodes below the current collapse root
  uint32_t max_depth = 0;
  bool collapsed_any = false;
  bool count_table_used = false;           // some leaf's count is not in its ref (16 or more entities, or none): not the lean tree build
  const char* error = nullptr;

  // pass 1: validate (tree-ness, ranges) and count spheres per subtree
  uint32_t count_pass(int32_t n, uint32_t depth) {
    if (error) return 0;
    if (n < 0 || (size_t)n >= node_count) { error = "BVH child index out of range"; return 0; }
    if (visited[n]) { error = "BVH is not a tree (node reachable twice)"; return 0; }
    if (depth > 4096) { error = "BVH deeper than 4096 levels"; return 0; }
    visited[n] = 1;
    const rtb_bvh_node& nd = nodes[n];
    uint32_t c;
    if (nd.first_entity >= 0) {
      if (nd.entity_count < 0 || (size_t)nd.first_entity + (size_t)nd.entity_count > sphere_count) {
        error = "BVH leaf range outside the sphere array";
        return 0;
      }
      c = (uint32_t)nd.entity_count;
      if (!entity_is_medium.empty())
        for (int i = 0; i < nd.entity_count; i++) subtree_media[n] |= entity_is_medium[(size_t)nd.first_entity + i];
    } else {
      if (nd.left < 0 || nd.right < 0) { error = "BVH inner node without two children"; return 0; }
      c = count_pass(nd.left, depth + 1) + count_pass(nd.right, depth + 1);
      if (!error) subtree_media[n] = subtree_media[nd.left] | subtree_media[nd.right];
    }
    subtree_spheres[n] = c;
    return c;
  }

  // The guard in the kernel certifies the skipped boxes from the sphere's geometry; that needs every
  // skipped box to contain the sphere shrunk by kChainShrink (true for any BVH built from the spheres'
  // bounds, Sphere.cs:16-23; a host could pass anything).  Otherwise the subtree is not collapsed.
  bool box_contains(const rtb_bvh_node& nd, size_t e) const {
    if (not_plain_sphere(e)) return false;     // the guard is sphere geometry: subtrees with anything else are not collapsed
    const rtb_sphere& s = sphere_of(e);
    const float r = std::fabs(s.radius) * (1.0f - kChainShrink);
    for (int k = 0; k < 3; k++)
      if (!(nd.bounds_min[k] <= s.center[k] - r && nd.bounds_max[k] >= s.center[k] + r)) return false;
    return true;
  }
  // every box strictly below `root` contains every sphere below it
  bool collapsible(int32_t n, bool is_root) {
    const rtb_bvh_node& nd = nodes[n];
    if (nd.first_entity >= 0) {
      if (!is_root)
        for (int i = 0; i < nd.entity_count; i++)
          if (!box_contains(nd, (size_t)nd.first_entity + i)) return false;
      for (int i = 0; i < nd.entity_count; i++)
        if (not_plain_sphere((size_t)nd.first_entity + i)) return false;
      return true;
    }
    if (!collapsible(nd.left, false) || !collapsible(nd.right, false)) return false;
    if (!is_root) {
      // inner boxes enclose their children's boxes in any sane tree; check against the spheres directly
      std::vector<int32_t> st{n};
      while (!st.empty()) {
        const rtb_bvh_node& x = nodes[st.back()];
        st.pop_back();
        if (x.first_entity >= 0) {
          for (int i = 0; i < x.entity_count; i++)
            if (!box_contains(nd, (size_t)x.first_entity + i)) return false;
        } else {
          st.push_back(x.left);
          st.push_back(x.right);
        }
      }
    }
    return true;
  }

  // emits the spheres of reference subtree n in depth-first order with their chains
  void emit_spheres(int32_t n, bool is_root) {
    const rtb_bvh_node& nd = nodes[n];
    if (!is_root) path.push_back(n);
    if (nd.first_entity >= 0) {
      for (int i = 0; i < nd.entity_count; i++) {
        order.push_back((uint32_t)(nd.first_entity + i));
        leaf_count.push_back(0);
        const uint32_t first_box = (uint32_t)(chain_boxes.size() / 8);
        for (size_t k = path.size(); k-- > 0;) {       // nearest (tightest) box first
          const rtb_bvh_node& b = nodes[path[k]];
          const float q[8] = {b.bounds_min[0], b.bounds_min[1], b.bounds_min[2], 0.0f, b.bounds_max[0], b.bounds_max[1], b.bounds_max[2], 0.0f};
          chain_boxes.insert(chain_boxes.end(), q, q + 8);
        }
        chain_ref.push_back(first_box | ((uint32_t)path.size() << 24));
      }
    } else {
      emit_spheres(nd.left, false);
      emit_spheres(nd.right, false);
    }
    if (!is_root) path.pop_back();
  }

  static int32_t leaf_ref(uint32_t first, uint32_t count) {
    const uint32_t code = (first << 4) | (count >= 1 && count <= 15 ? count - 1 : 15u);   // 15: count in leaf_count[first]
    return ~(int32_t)code;
  }

  // pass 2: device nodes
  int32_t ref_of(int32_t n, uint32_t depth) {
    if (error) return 0;
    max_depth = std::max(max_depth, depth);
    const rtb_bvh_node& nd = nodes[n];
    const uint32_t c = subtree_spheres[n];
    const bool is_leaf = nd.first_entity >= 0;
    if (is_leaf || (c <= collapse_k && c <= 15 && path_len_ok(n) && collapsible(n, true))) {
      if (!is_leaf) collapsed_any = true;
      if (c == 0 || c >= 16) count_table_used = true;
      const uint32_t first = (uint32_t)order.size();
      if (first + c >= (1u << 27)) { error = "scene too large"; return 0; }
      if (c == 0) {                       // empty leaf: a slot whose leaf_count is 0
        order.push_back(0xFFFFFFFFu);
        leaf_count.push_back(0);
        chain_ref.push_back(0);
        return leaf_ref(first, 0);
      }
      emit_spheres(n, true);
      leaf_count[first] = c;
      return leaf_ref(first, c);
    }
    const int32_t self = (int32_t)(inner.size() / 16);
    if (self >= (1 << 24) / (int32_t)(kNodeStride / 16)) { error = "scene too large"; return 0; }
    inner.resize(inner.size() + 16, 0.0f);
    const rtb_bvh_node& l = nodes[nd.left];
    const rtb_bvh_node& r = nodes[nd.right];
    const int32_t lref = ref_of(nd.left, depth + 1);
    const int32_t rref = ref_of(nd.right, depth + 1);
    if (error) return 0;
    float* q = &inner[(size_t)self * 16];
    // x and y of every corner as an aligned pair, the four z's as two pairs (kernel_common.cuh: node_lmin ..., aabb_range_pair)
    q[0] = l.bounds_min[0]; q[1] = l.bounds_min[1]; q[2] = l.bounds_max[0]; q[3] = l.bounds_max[1];
    q[4] = r.bounds_min[0]; q[5] = r.bounds_min[1]; q[6] = r.bounds_max[0]; q[7] = r.bounds_max[1];
    q[8] = l.bounds_min[2]; q[9] = l.bounds_max[2]; q[10] = r.bounds_min[2]; q[11] = r.bounds_max[2];
    memcpy(&q[12], &lref, 4);
    memcpy(&q[13], &rref, 4);
    // word 14: which child subtrees hold an entity that wears a medium (media.cuh: the media walks skip the others)
    const uint32_t media = (subtree_media[nd.left] ? 1u : 0u) | (subtree_media[nd.right] ? 2u : 0u);
    memcpy(&q[14], &media, 4);
    return self * (int32_t)kNodeStride; // inner refs are BYTE offsets of the node record in the blob (inner_off == 0)
  }
  // a chain holds at most 255 boxes
  bool path_len_ok(int32_t n) const {
    uint32_t d = 0;
    std::vector<std::pair<int32_t, uint32_t>> st{{n, 0}};
    while (!st.empty()) {
      auto [x, dx] = st.back();
      st.pop_back();
      d = std::max(d, dx);
      if (nodes[x].first_entity < 0) { st.push_back({nodes[x].left, dx + 1}); st.push_back({nodes[x].right, dx + 1}); }
    }
    return d < 200;
  }
};

bool almost_equals_1(float v) { return std::fabs(1.0f - v) < 1e-6f; }  // MathExtensions.cs:23-27

const char* build_blob(const rtb_entity* entities, size_t entity_count, const rtb_sphere* spheres, size_t sphere_count,
                       const rtb_triangle* triangles, size_t triangle_count, const rtb_placed_entity* placed, size_t placed_count,
                       const rtb_material* materials, size_t material_count, const rtb_bvh_node* nodes, size_t node_count,
                       uint32_t collapse_k, int retree, HostBlob* out, int* status) {
  *status = RTB_ERR_INVALID_ARGUMENT;
  for (size_t i = 0; i < entity_count; i++) {
    const uint32_t base = entities[i].type & ~(uint32_t)RTB_ENTITY_PLACED;
    if (base < RTB_ENTITY_SPHERE || base > RTB_ENTITY_TRIANGLE) {
      *status = RTB_ERR_UNSUPPORTED;
      return "unknown entity type";
    }
    if (Flattener::is_placed_type(entities[i].type)) {
      if (base == RTB_ENTITY_TRIANGLE) return "triangles are always world-space (Entity.cs:92-93): not a placed entity";
      if (entities[i].index >= placed_count) return "entity index out of range";
      if (placed[entities[i].index].type != base) return "entity type differs from its placed record";
    } else if (entities[i].index >= (base == RTB_ENTITY_SPHERE ? sphere_count : triangle_count)) {
      return "entity index out of range";
    }
  }
  for (size_t i = 0; i < placed_count; i++) {
    const rtb_placed_entity& e = placed[i];
    if (e.type != RTB_ENTITY_SPHERE && e.type != RTB_ENTITY_RECT && e.type != RTB_ENTITY_BOX) return "placed entity of an unknown type";
    if (e.material >= material_count) return "placed entity material index out of range";
    if (e.moving && e.time_range[0] == e.time_range[1]) return "time range cannot be empty for moving entities (Entity.cs:54)";
  }
  for (size_t i = 0; i < triangle_count; i++)
    if (triangles[i].material >= material_count) return "triangle material index out of range";
  const size_t leaf_list_count = entities ? entity_count : sphere_count;
  bool has_volumes = false;
  for (size_t i = 0; i < material_count; i++) {
    if (materials[i].type > RTB_MATERIAL_PROBABILISTIC_VOLUME) {
      *status = RTB_ERR_UNSUPPORTED;
      return "unknown material type";
    }
  }
  // a medium counts when some entity wears it (a host may keep unused materials in its buffer)
  auto wears_volume = [&](uint32_t m) { return m < material_count && materials[m].type == RTB_MATERIAL_PROBABILISTIC_VOLUME; };
  if (entities) {
    for (size_t i = 0; i < entity_count && !has_volumes; i++) {
      const uint32_t base = entities[i].type & ~(uint32_t)RTB_ENTITY_PLACED;
      if (Flattener::is_placed_type(entities[i].type)) has_volumes = wears_volume(placed[entities[i].index].material);
      else if (base == RTB_ENTITY_TRIANGLE) has_volumes = wears_volume(triangles[entities[i].index].material);
      else has_volumes = wears_volume(spheres[entities[i].index].material);
    }
  } else {
    for (size_t i = 0; i < sphere_count && !has_volumes; i++) has_volumes = wears_volume(spheres[i].material);
  }
  // the volume kernel collects candidates exactly like the reference: every host leaf must stay a device leaf
  if (has_volumes) collapse_k = 1;
  for (size_t i = 0; i < sphere_count; i++)
    if (spheres[i].material >= material_count) return "sphere material index out of range";

  // Another topology over the host tree's leaves when the world qualifies (retree.hpp: the reference's candidate set depends
  // on the leaf boxes only; what is walked here is a surface-area-heuristic tree over the same leaves).  Everything below —
  // validation, flattening — sees it as "the host's tree".  retree == 1: worlds of spheres and triangles, whose frames do not
  // depend on the topology (measured: one frame checksum with either tree on configs 3, 4, 5 and on the mesh world — the
  // triangle flavour resolves exact distance ties by the reference's candidate order, kernel_common.cuh: nearer); 2: also
  // worlds with placed entities, whose flavour leaves a tie to the entity visited first (1 path in 1.3e8 came out differently
  // on the Cornell box).  Media worlds keep the host's topology (no gain: their walks cannot prune boxes that hold media).
  std::vector<rtb_bvh_node> rebuilt;
  const rtb_bvh_node* const host_nodes = nodes;     // the reference's topology: its depth-first leaf order is the candidate order
  out->retreed = false;
  bool plain_spheres = true, no_placed = true;
  for (size_t i = 0; i < entity_count; i++) {
    plain_spheres = plain_spheres && entities[i].type == RTB_ENTITY_SPHERE;
    no_placed = no_placed && !Flattener::is_placed_type(entities[i].type);
  }
  if (retree && !has_volumes && (no_placed || retree >= 2) && rtb_retree::retree(nodes, node_count, kStackMax - 2, rebuilt)) {
    nodes = rebuilt.data();
    node_count = rebuilt.size();
    out->retreed = true;
    if (!plain_spheres) collapse_k = 1;              // the tie rule of the triangle / placed flavours needs one slot order (below)
  }

  Flattener f{};
  f.nodes = nodes; f.node_count = node_count; f.sphere_count = leaf_list_count; f.spheres = spheres; f.entities = entities;
  f.collapse_k = std::max<uint32_t>(1, std::min<uint32_t>(collapse_k, 15));
  f.visited.assign(node_count, 0);
  f.subtree_spheres.assign(node_count, 0);
  f.subtree_media.assign(node_count, 0);
  if (has_volumes) {
    f.entity_is_medium.assign(leaf_list_count, 0);
    for (size_t i = 0; i < leaf_list_count; i++) {
      uint32_t m;
      if (!entities) m = spheres[i].material;
      else if (Flattener::is_placed_type(entities[i].type)) m = placed[entities[i].index].material;
      else if ((entities[i].type & ~(uint32_t)RTB_ENTITY_PLACED) == RTB_ENTITY_TRIANGLE) m = triangles[entities[i].index].material;
      else m = spheres[entities[i].index].material;
      f.entity_is_medium[i] = wears_volume(m) ? 1 : 0;
    }
  }
  SceneDesc& d = out->desc;
  d = SceneDesc{};
  if (node_count > 0) {
    f.count_pass(0, 1);
    if (f.error) return f.error;
    d.has_root = f.subtree_spheres[0] > 0 ? 1 : 0;   // empty world
    if (d.has_root) {
      d.root_ref = f.ref_of(0, 1);
      if (f.error) return f.error;
      for (int k = 0; k < 3; k++) { d.root_min[k] = nodes[0].bounds_min[k]; d.root_max[k] = nodes[0].bounds_max[k]; }
    }
  }
  if (f.max_depth > (uint32_t)kStackMax) {
    *status = RTB_ERR_UNSUPPORTED;
    return "BVH deeper than 64 levels";
  }
  // Slots in the REFERENCE's depth-first leaf order whatever topology was flattened: that order is the reference's candidate
  // order, which decides between entities at exactly the same distance (kernel_common.cuh: nearer; media.cuh: visited_later).
  // The flattener laid the slots out in the depth-first order of the tree it was given; for a re-built tree they are permuted
  // here (a leaf's entities stay together: the reference visits them together too).
  if (out->retreed && !f.collapsed_any && d.has_root) {
    std::vector<uint32_t> rank(leaf_list_count, 0xFFFFFFFFu);
    uint32_t r = 0;
    std::vector<int32_t> st{0};
    while (!st.empty()) {
      const rtb_bvh_node& nd = host_nodes[st.back()];
      st.pop_back();
      if (nd.first_entity >= 0) {
        for (int i = 0; i < nd.entity_count; i++) rank[(size_t)nd.first_entity + i] = r++;
      } else {
        st.push_back(nd.right);
        st.push_back(nd.left);
      }
    }
    const size_t n_slots = f.order.size();
    bool ok = r == n_slots && f.leaf_count.size() == n_slots && f.chain_ref.size() == n_slots;
    std::vector<uint32_t> perm(n_slots);
    for (size_t s = 0; s < n_slots && ok; s++) {
      const uint32_t h = f.order[s];
      ok = h != 0xFFFFFFFFu && rank[h] != 0xFFFFFFFFu;
      if (ok) perm[s] = rank[h];
    }
    if (ok) {
      std::vector<uint32_t> order2(n_slots), count2(n_slots), chain2(n_slots);
      for (size_t s = 0; s < n_slots; s++) { order2[perm[s]] = f.order[s]; count2[perm[s]] = f.leaf_count[s]; chain2[perm[s]] = f.chain_ref[s]; }
      f.order.swap(order2); f.leaf_count.swap(count2); f.chain_ref.swap(chain2);
      auto move_leaf = [&](int32_t ref) {
        if (ref >= 0) return ref;
        const uint32_t code = (uint32_t)~ref;
        return ~(int32_t)((perm[code >> 4] << 4) | (code & 15u));
      };
      for (size_t i = 0; i < f.inner.size() / 16; i++) {
        int32_t c[2];
        memcpy(c, &f.inner[i * 16 + 12], 8);
        c[0] = move_leaf(c[0]); c[1] = move_leaf(c[1]);
        memcpy(&f.inner[i * 16 + 12], c, 8);
      }
      d.root_ref = move_leaf(d.root_ref);
    }
  }
  d.max_depth = f.max_depth;
  d.n_inner = (uint32_t)(f.inner.size() / 16);
  const size_t n_dev = f.order.size();
  d.n_spheres = (uint32_t)n_dev;
  d.n_materials = (uint32_t)material_count;
  d.has_chains = f.collapsed_any ? 1 : 0;
  d.has_big_leaves = f.count_table_used ? 1 : 0;
  d.n_triangles = (uint32_t)triangle_count;
  d.n_placed = (uint32_t)placed_count;
  d.has_volumes = has_volumes ? 1u : 0u;
  d.skip_root_test = out->retreed && d.has_root && d.root_ref >= 0 ? 1u : 0u;

  auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  size_t off = 0;
  d.inner_off = (uint32_t)off; off = align16(off + (size_t)d.n_inner * kNodeStride);
  d.sphere_off = (uint32_t)off; off = align16(off + (n_dev + 1) * 16);
  d.leaf_count_off = (uint32_t)off; off = align16(off + (n_dev + 1) * 4);
  d.mat_index_off = (uint32_t)off; off = align16(off + (n_dev + 1) * 4);
  d.tri_off = (uint32_t)off; off = align16(off + triangle_count * 80);
  d.placed_off = (uint32_t)off; off = align16(off + placed_count * 112);
  if (off == 0) off = 16;
  d.blob_bytes = (uint32_t)off;
  out->bytes.assign(off, 0);
  unsigned char* b = out->bytes.data();
  // leaf refs leave the recursion as ~(first slot << 4 | count code); now that the layout is known they become
  // ~(byte offset of the first slot in the blob | count code) (slots are 16 bytes, so the low 4 bits are free)
  if (d.sphere_off + (n_dev + 1) * 16 >= (1ull << 31)) return "scene too large";
  auto patch = [&](int32_t ref) {
    if (ref >= 0) return ref;
    const uint32_t code = (uint32_t)~ref;
    return ~(int32_t)((d.sphere_off + (code >> 4) * 16u) | (code & 15u));
  };
  for (uint32_t i = 0; i < d.n_inner; i++) {
    int32_t r[2];
    memcpy(r, &f.inner[(size_t)i * 16 + 12], 8);
    r[0] = patch(r[0]); r[1] = patch(r[1]);
    memcpy(&f.inner[(size_t)i * 16 + 12], r, 8);
  }
  if (d.has_root) d.root_ref = patch(d.root_ref);
  for (uint32_t i = 0; i < d.n_inner; i++) memcpy(b + d.inner_off + (size_t)i * kNodeStride, &f.inner[(size_t)i * 16], 64);
  for (size_t i = 0; i < n_dev; i++) {
    const uint32_t h = f.order[i];
    float s[4] = {0, 0, 0, 0};
    uint32_t mat = 0;
    if (h != 0xFFFFFFFFu && f.is_placed(h)) {
      const uint32_t pi = entities[h].index;        // slot = (placed index as bits, 1 as bits, 0, NaN)
      const uint32_t nan_bits = 0x7fc00000u, one = 1u;
      memcpy(&s[0], &pi, 4);
      memcpy(&s[1], &one, 4);
      memcpy(&s[3], &nan_bits, 4);
      mat = placed[pi].material;
    } else if (h != 0xFFFFFFFFu && f.is_triangle(h)) {
      const uint32_t ti = entities[h].index;        // slot = (triangle index as bits, 0, 0, NaN)
      const uint32_t nan_bits = 0x7fc00000u;
      memcpy(&s[0], &ti, 4);
      memcpy(&s[3], &nan_bits, 4);
      mat = triangles[ti].material;
    } else if (h != 0xFFFFFFFFu) {
      const rtb_sphere& sp = f.sphere_of(h);
      if (sp.radius != sp.radius) return "sphere with a NaN radius";
      s[0] = sp.center[0]; s[1] = sp.center[1]; s[2] = sp.center[2]; s[3] = sp.radius;
      mat = sp.material;
    }
    memcpy(b + d.sphere_off + i * 16, s, 16);
    memcpy(b + d.leaf_count_off + i * 4, &f.leaf_count[i], 4);
    if (has_volumes && wears_volume(mat)) mat |= 0x80000000u;   // media.cuh: kMediumBit (only worlds with media carry it, and only media.cuh reads those)
    memcpy(b + d.mat_index_off + i * 4, &mat, 4);
  }
  for (size_t i = 0; i < triangle_count; i++) {
    const rtb_triangle& t = triangles[i];
    const float q[20] = {t.edge2[0], t.edge2[1], t.edge2[2], t.edge1[0], t.edge1[1], t.edge1[2], t.v0[0], t.v0[1], t.v0[2],
                         t.normals[0][0], t.normals[0][1], t.normals[0][2], t.normals[1][0], t.normals[1][1], t.normals[1][2],
                         t.normals[2][0], t.normals[2][1], t.normals[2][2], 0.0f, 0.0f};
    memcpy(b + d.tri_off + i * 80, q, 80);
  }
  for (size_t i = 0; i < placed_count; i++) {
    const rtb_placed_entity& e = placed[i];
    um::rigid origin;
    origin.rot.x = e.rotation[0]; origin.rot.y = e.rotation[1]; origin.rot.z = e.rotation[2]; origin.rot.w = e.rotation[3];
    origin.pos = um::mk(e.position[0], e.position[1], e.position[2]);
    const um::rigid inv = um::inverse(origin);     // Entity ctor (Entity.cs:51-52); used by static entities only
    float c[6] = {0, 0, 0, 0, 0, 0};
    if (e.type == RTB_ENTITY_SPHERE) {
      c[0] = e.size[0];
    } else if (e.type == RTB_ENTITY_RECT) {        // Rect ctor (Rect.cs:11-15)
      c[0] = um::div(-e.size[0], 2.0f); c[1] = um::div(-e.size[1], 2.0f);
      c[2] = um::div(e.size[0], 2.0f); c[3] = um::div(e.size[1], 2.0f);
    } else {                                       // Box ctor (Box.cs:11-15)
      for (int k = 0; k < 3; k++) { c[k] = um::div(e.size[k], 2.0f); c[3 + k] = um::rcp(c[k]); }
    }
    const uint32_t flags = e.type | (e.moving ? 1u << 8 : 0u);
    float flags_f;
    memcpy(&flags_f, &flags, 4);
    const float q[28] = {origin.rot.x, origin.rot.y, origin.rot.z, origin.rot.w,
                         origin.pos.x, origin.pos.y, origin.pos.z, flags_f,
                         inv.rot.x, inv.rot.y, inv.rot.z, inv.rot.w,
                         inv.pos.x, inv.pos.y, inv.pos.z, 0.0f,
                         e.destination_offset[0], e.destination_offset[1], e.destination_offset[2], e.time_range[0],
                         e.time_range[1], c[0], c[1], c[2],
                         c[3], c[4], c[5], 0.0f};
    memcpy(b + d.placed_off + i * 112, q, 112);
  }
  out->chain_ref = std::move(f.chain_ref);
  out->chain_boxes = std::move(f.chain_boxes);
  out->materials.resize(std::max<size_t>(material_count, 1));
  for (size_t i = 0; i < material_count; i++) {
    DevMaterial m{};
    const rtb_material& s = materials[i];
    for (int k = 0; k < 3; k++) { m.albedo[k] = s.albedo[k]; m.emission[k] = s.emission[k]; }
    m.type = s.type;
    m.glossiness = s.glossiness;
    m.metallic = s.metallic;
    m.ior = s.index_of_refraction;      // Standard: replaced on the device by lerp(1.5, 1.1, metallic)
    // Material.IsPerfectSpecular (Material.cs:181-196)
    m.perfect_specular = s.type == RTB_MATERIAL_DIELECTRIC ||
                         (s.type == RTB_MATERIAL_STANDARD && almost_equals_1(s.metallic) && almost_equals_1(s.glossiness));
    out->materials[i] = m;              // roughness / alpha / r0 are filled by derive_materials_kernel
  }
  *status = RTB_OK;
  return nullptr;
}

// ---- launch -----------------------------------------------------------------------------
struct ActiveRows {
  int first_row, row_step, n_rows;
};

// Rows this batch touches: the interlace test of SampleBatchJob.cs:69 intersected with the
// [row_begin, row_end) extension.
ActiveRows active_rows(const rtb_batch_params& p, int height, int range_begin, int range_end) {
  int lo = 0, hi = height;
  if (p.row_end > p.row_begin) { lo = std::max(lo, p.row_begin); hi = std::min(hi, p.row_end); }
  if (range_end > range_begin) { lo = std::max(lo, range_begin); hi = std::min(hi, range_end); }
  ActiveRows r{0, p.slice_divider, 0};
  if (hi <= lo) return r;
  int first = lo + ((p.slice_offset - lo) % p.slice_divider + p.slice_divider) % p.slice_divider;
  if (first >= hi) return r;
  r.first_row = first;
  r.n_rows = (hi - 1 - first) / p.slice_divider + 1;
  return r;
}

int validate_params(rtb_ctx* ctx, const rtb_batch_params* p, int* width, int* height) {
  if (!p) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "params is NULL");
  if (!(p->size[0] >= 1.0f && p->size[1] >= 1.0f && p->size[0] <= 65536.0f && p->size[1] <= 65536.0f))
    return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "Size out of range");
  *width = (int)p->size[0];
  *height = (int)p->size[1];
  if ((uint64_t)*width * (uint64_t)*height >= (1ull << 31))
    return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "Size: more than 2^31 pixels");
  if (p->slice_divider < 1 || p->slice_offset < 0 || p->slice_offset >= p->slice_divider)
    return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "SliceOffset/SliceDivider invalid");
  // TraceDepth is 1..500 in the reference (Raytracer.cs:90); 0 would mean "no bounce iteration at all"
  if (p->trace_depth < 1 || p->trace_depth > 65535) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "TraceDepth out of range (1..65535)");
  if (p->sample_count_range[0] > (1u << 20) || p->sample_count_range[1] > (1u << 20))
    return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "SampleCountRange above 2^20 samples per pixel per batch");
  if (p->environment.sky_type > RTB_SKY_CUBEMAP) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "unknown sky type");
  if (p->environment.sky_type == RTB_SKY_CUBEMAP && !ctx->d_sky)
    return fail(ctx, RTB_ERR_NO_SCENE, "SkyType.CubeMap without rtb_upload_sky_cubemap");
  if (p->row_begin < 0 || p->row_end < 0 || p->row_end > *height)
    return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "row_begin/row_end out of range");
  return RTB_OK;
}

void choose_tiles(BatchArgs& a, uint32_t n_warps_full, uint32_t max_spp);

template <bool SMEM, bool COUNTERS, int FLAVOR>
int launch_mega_t(rtb_ctx* ctx, BatchArgs& a, cudaStream_t stream, uint32_t max_spp) {
  constexpr int kMegaBlock = mega_block(FLAVOR), kMegaWarps = kMegaBlock / 32;
  const size_t smem = mega_smem_bytes(a.scene.blob_bytes, SMEM, FLAVOR);
  auto kernel = sample_megakernel<SMEM, COUNTERS, FLAVOR>;
  if (!ctx->smem_attr_set[SMEM][COUNTERS + 2 * FLAVOR]) {
    RTB_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin));
    ctx->smem_attr_set[SMEM][COUNTERS + 2 * FLAVOR] = true;
  }
  int blocks_per_sm = 0;
  RTB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, kMegaBlock, smem));
  if (blocks_per_sm < 1) return fail(ctx, RTB_ERR_CUDA, "megakernel does not fit on an SM (smem %zu B)", smem);
  const uint32_t n_warps_full = (uint32_t)(ctx->sm_count * blocks_per_sm * kMegaWarps);

  choose_tiles(a, n_warps_full, max_spp);
  const uint32_t ctas_needed = (a.n_tiles + kMegaWarps - 1) / kMegaWarps;
  const uint32_t grid = std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)(ctx->sm_count * blocks_per_sm), ctas_needed));
  a.tile_counter = ctx->d_tile_ring + (ctx->ring_next.fetch_add(1) % rtb_ctx::kTileRing);
  RTB_CUDA(ctx, cudaMemsetAsync(a.tile_counter, 0, sizeof(uint32_t), stream));
  if (FLAVOR <= kFlavorChains && !COUNTERS && ctx->opt_math == 1) {
    // the same kernel source built with fast arithmetic (fast_kernels.cu); same launch geometry
    const int e = rtb_fast_launch_spheres(&a, sizeof a, SMEM ? 1 : 0, FLAVOR, grid, smem, ctx->max_smem_optin, stream);
    if (e != 0) return fail(ctx, RTB_ERR_CUDA + e, "fast-math megakernel launch failed: %s", cudaGetErrorString((cudaError_t)e));
    return RTB_OK;
  }
  kernel<<<grid, kMegaBlock, smem, stream>>>(a);
  RTB_CUDA(ctx, cudaGetLastError());
  return RTB_OK;
}

// Tile size of the persistent kernel: ~2048 samples per warp tile (measured best on B200, profiles/README.md), at least ~6 tiles per
// resident warp when the image is small, never fewer than 1024 samples per tile unless the pixel count forces it.
void choose_tiles(BatchArgs& a, uint32_t n_warps_full, uint32_t max_spp) {
  static const uint32_t target = [] { const char* e = getenv("RTB_TILE_SAMPLES"); return e ? (uint32_t)std::max(32, atoi(e)) : 2048u; }();
  // free lanes a warp collects before refilling them (sample_megakernel): worth it when the walk diverges anyway (a real
  // tree), not when the world is a linear list every lane walks in step
  static const int refill_env = [] { const char* e = getenv("RTB_REFILL_MIN"); return e ? std::min(32, std::max(1, atoi(e))) : 0; }();
  a.refill_min = refill_env ? (uint32_t)refill_env : (a.scene.n_inner >= 64 ? 8u : 1u);
  int tp = (int)std::min<uint32_t>(kHalfPixels, std::max<uint32_t>(1, (target + max_spp - 1) / std::max<uint32_t>(max_spp, 1)));
  while (tp > 1 && (a.n_active_pixels + tp - 1) / tp < 6 * n_warps_full && (uint64_t)(tp / 2) * max_spp >= 1024) tp /= 2;
  // guided self-scheduling: the last ~8 tiles per resident warp are a quarter of the size, the last ~8 after
  // those a sixteenth (never fewer than ~256 samples per tile: below that the per-tile drain costs more than the tail)
  auto shrink = [&](int t, int by) {
    int r = std::max(1, t / by);
    while (r < t && (uint64_t)r * max_spp < 256) r *= 2;
    return std::min(r, t);
  };
  const int size[3] = {tp, shrink(tp, 4), shrink(tp, 16)};
  const uint32_t P = a.n_active_pixels;
  static const int per_warp = [] { const char* e = getenv("RTB_TAIL_TILES"); return e ? std::max(0, atoi(e)) : 8; }();
  uint32_t tail2 = std::min<uint64_t>(P, (uint64_t)per_warp * n_warps_full * size[2]);
  uint32_t tail1 = std::min<uint64_t>(P - tail2, (uint64_t)per_warp * n_warps_full * size[1]);
  if (size[2] == size[1]) { tail1 += tail2; tail2 = 0; }
  if (size[1] == size[0]) { tail1 = 0; if (size[2] == size[0]) tail2 = 0; }
  uint32_t begin[4] = {0, P - tail1 - tail2, P - tail2, P};
  // phase boundaries on multiples of the phase's tile size so only the last tile of a phase is ragged
  uint32_t tile = 0;
  for (int k = 0; k < 3; k++) {
    a.phase_size[k] = size[k];
    a.phase_pixel[k] = begin[k];
    a.phase_tile[k] = tile;
    tile += (begin[k + 1] - begin[k] + (uint32_t)size[k] - 1) / (uint32_t)size[k];
  }
  a.phase_pixel[3] = P;
  a.phase_tile[3] = tile;
  a.n_tiles = tile;
}

int launch_batch(rtb_ctx* ctx, const rtb_batch_params& p, const rtb_batch_buffers& dev, int width, int height,
                 ActiveRows rows, cudaStream_t stream) {
  if (rows.n_rows <= 0) return RTB_OK;
  BatchArgs a{};
  a.p = p;
  a.b = dev;
  a.scene = ctx->scene;
  a.scene.sky_faces = ctx->d_sky;
  a.scene.tex_pixels = ctx->d_tex_pixels;
  a.scene.tex_images = ctx->d_tex_images;
  a.scene.mat_textures = ctx->d_mat_textures;
  a.scene.tri_uv = ctx->d_tri_uv;
  a.scene.sky_w = ctx->sky_w;
  a.scene.sky_h = ctx->sky_h;
  a.width = width;
  a.height = height;
  a.first_row = rows.first_row;
  a.row_step = rows.row_step;
  a.n_rows = rows.n_rows;
  a.n_active_pixels = (uint32_t)rows.n_rows * (uint32_t)width;
  // every launch gets a fresh epoch: the word the kernels poll can only equal it after send_cancel for THIS launch
  if (++ctx->cancel_epoch == 0u) ctx->cancel_epoch = 1u;
  a.cancel_flag = ctx->d_cancel;
  a.cancel_epoch = ctx->cancel_epoch;
  a.scene.status = ctx->d_status;
  const bool counters = ctx->opt_counters != 0;
  a.counters = counters ? ctx->d_counters : nullptr;
  const uint32_t max_spp = std::max(p.sample_count_range[0], p.sample_count_range[1]);

  if (counters) {
    if (!a.b.out_diagnostics) {  // the counter pass reads the per-pixel diagnostics
      const size_t need = (size_t)width * height;
      if (ctx->scratch_diag_capacity < need) {
        if (ctx->d_scratch_diag) cudaFree(ctx->d_scratch_diag);
        ctx->d_scratch_diag = nullptr;
        ctx->scratch_diag_capacity = 0;
        RTB_CUDA(ctx, cudaMalloc(&ctx->d_scratch_diag, need * sizeof(rtb_diagnostics)));
        ctx->scratch_diag_capacity = need;
      }
      a.b.out_diagnostics = ctx->d_scratch_diag;
    }
    RTB_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), stream));
    ctx->counters_stream = stream;
  }

  int kernel_kind = (int)ctx->opt_kernel;
  if (kernel_kind == 0) kernel_kind = ctx->default_kernel;
  // Worlds with ProbabilisticVolume materials (media.cuh).  The product path is the megakernel's media flavour; the collect-all
  // kernel sample_volumes (bit-identical to the CPU oracle: the reference's list, the reference's accumulation order) runs as
  // the validator (RTB_OPT_KERNEL = 1), for the reference's sequential white-noise stream, and for media worlds with image
  // textures (the media flavour has no texture code).
  const bool media_mega = ctx->scene.has_volumes && kernel_kind != 1 && !ctx->opt_noise && !ctx->d_mat_textures;
  if (media_mega) {
    const bool fits = mega_smem_bytes(ctx->scene.blob_bytes, true, kFlavorMedia) <= (size_t)ctx->max_smem_optin && ctx->scene.blob_bytes < (1u << 20);
    int rc;
    if (fits) rc = counters ? launch_mega_t<true, true, kFlavorMedia>(ctx, a, stream, max_spp) : launch_mega_t<true, false, kFlavorMedia>(ctx, a, stream, max_spp);
    else rc = counters ? launch_mega_t<false, true, kFlavorMedia>(ctx, a, stream, max_spp) : launch_mega_t<false, false, kFlavorMedia>(ctx, a, stream, max_spp);
    if (rc != RTB_OK) return rc;
  } else
  if (ctx->scene.has_volumes) {
    if (ctx->opt_noise) {                       // one thread per pixel (a sequential stream per pixel)
      const uint32_t grid = (a.n_active_pixels + 127) / 128;
      if (counters) sample_volumes<true, true><<<grid, 128, 0, stream>>>(a);
      else sample_volumes<false, true><<<grid, 128, 0, stream>>>(a);
    } else {                                    // one warp per pixel
      const uint32_t grid = (a.n_active_pixels + 3) / 4;
      if (counters) sample_volumes<true, false><<<grid, 128, 0, stream>>>(a);
      else sample_volumes<false, false><<<grid, 128, 0, stream>>>(a);
    }
    RTB_CUDA(ctx, cudaGetLastError());
  } else
  if (ctx->opt_noise && kernel_kind != 1)
    return fail(ctx, RTB_ERR_UNSUPPORTED, "RTB_OPT_NOISE = 1 (the reference's sequential white-noise stream) needs RTB_OPT_KERNEL = 1");
  else if (kernel_kind == 1) {
    const uint32_t grid = (a.n_active_pixels + 127) / 128;
    if (ctx->opt_noise) {
      if (counters) sample_simple<true, true><<<grid, 128, 0, stream>>>(a);
      else sample_simple<false, true><<<grid, 128, 0, stream>>>(a);
    } else {
      if (counters) sample_simple<true, false><<<grid, 128, 0, stream>>>(a);
      else sample_simple<false, false><<<grid, 128, 0, stream>>>(a);
    }
    RTB_CUDA(ctx, cudaGetLastError());
  } else {
    // the instrumented build and worlds with triangles take the general flavour; sphere worlds take the lean ones
    const int flavor = ctx->scene.n_placed ? (ctx->d_mat_textures ? kFlavorPlacedTextured : kFlavorPlaced)
                       : (counters || ctx->scene.n_triangles || ctx->d_mat_textures) ? kFlavorGeneral : (ctx->scene.has_chains || ctx->scene.has_big_leaves ? kFlavorChains : kFlavorSpheres);
    const bool fits = mega_smem_bytes(ctx->scene.blob_bytes, true, flavor) <= (size_t)ctx->max_smem_optin &&
                      ctx->scene.blob_bytes < (1u << 20);
    int rc;
    if (flavor == kFlavorPlaced) {
      if (fits) rc = counters ? launch_mega_t<true, true, kFlavorPlaced>(ctx, a, stream, max_spp) : launch_mega_t<true, false, kFlavorPlaced>(ctx, a, stream, max_spp);
      else rc = counters ? launch_mega_t<false, true, kFlavorPlaced>(ctx, a, stream, max_spp) : launch_mega_t<false, false, kFlavorPlaced>(ctx, a, stream, max_spp);
    } else if (flavor == kFlavorPlacedTextured) {
      if (fits) rc = counters ? launch_mega_t<true, true, kFlavorPlacedTextured>(ctx, a, stream, max_spp) : launch_mega_t<true, false, kFlavorPlacedTextured>(ctx, a, stream, max_spp);
      else rc = counters ? launch_mega_t<false, true, kFlavorPlacedTextured>(ctx, a, stream, max_spp) : launch_mega_t<false, false, kFlavorPlacedTextured>(ctx, a, stream, max_spp);
    } else
    if (fits) rc = counters ? launch_mega_t<true, true, kFlavorGeneral>(ctx, a, stream, max_spp)
                   : flavor == kFlavorGeneral ? launch_mega_t<true, false, kFlavorGeneral>(ctx, a, stream, max_spp)
                   : flavor == kFlavorChains ? launch_mega_t<true, false, kFlavorChains>(ctx, a, stream, max_spp)
                                             : launch_mega_t<true, false, kFlavorSpheres>(ctx, a, stream, max_spp);
    else rc = counters ? launch_mega_t<false, true, kFlavorGeneral>(ctx, a, stream, max_spp)
              : flavor == kFlavorGeneral ? launch_mega_t<false, false, kFlavorGeneral>(ctx, a, stream, max_spp)
              : flavor == kFlavorChains ? launch_mega_t<false, false, kFlavorChains>(ctx, a, stream, max_spp)
                                        : launch_mega_t<false, false, kFlavorSpheres>(ctx, a, stream, max_spp);
    if (rc != RTB_OK) return rc;
  }
  if (counters) {
    counters_from_diagnostics<<<ctx->sm_count * 4, 256, 0, stream>>>(a.b.out_diagnostics, a.b.out_color, a.b.in_color, a);
    RTB_CUDA(ctx, cudaGetLastError());
  }
  return RTB_OK;
}

int ensure_buffers(rtb_ctx* ctx, size_t pixels) {
  DeviceBuffers& b = ctx->buf;
  if (b.capacity >= pixels) return RTB_OK;
  float** ptrs[] = {&b.in_color, &b.in_weight, &b.in_normal, &b.in_albedo, &b.out_color, &b.out_weight, &b.out_normal, &b.out_albedo};
  for (float** p : ptrs) { if (*p) cudaFree(*p); *p = nullptr; }
  if (b.diagnostics) cudaFree(b.diagnostics);
  b.diagnostics = nullptr;
  b.capacity = 0;
  const size_t elems[] = {4, 1, 3, 3, 4, 1, 3, 3};
  for (int i = 0; i < 8; i++) RTB_CUDA(ctx, cudaMalloc(ptrs[i], pixels * elems[i] * sizeof(float)));
  RTB_CUDA(ctx, cudaMalloc(&b.diagnostics, pixels * sizeof(rtb_diagnostics)));
  b.capacity = pixels;
  return RTB_OK;
}

// Copies the active rows of one per-pixel array between host and device.
cudaError_t copy_rows(void* dst, const void* src, size_t elem_bytes, int width, ActiveRows rows, cudaMemcpyKind kind,
                      cudaStream_t stream) {
  const size_t row_bytes = (size_t)width * elem_bytes;
  const size_t offset = (size_t)rows.first_row * row_bytes;
  const size_t pitch = row_bytes * (size_t)rows.row_step;
  if (rows.row_step == 1)
    return cudaMemcpyAsync((char*)dst + offset, (const char*)src + offset, row_bytes * (size_t)rows.n_rows, kind, stream);
  return cudaMemcpy2DAsync((char*)dst + offset, pitch, (const char*)src + offset, pitch, row_bytes, (size_t)rows.n_rows, kind, stream);
}

void drop_textures(rtb_ctx* ctx) {
  void* ptrs[] = {ctx->d_tex_pixels, ctx->d_tex_images, ctx->d_mat_textures, ctx->d_tri_uv};
  for (void* p : ptrs) if (p) cudaFree(p);
  ctx->d_tex_pixels = nullptr;
  ctx->d_tex_images = nullptr;
  ctx->d_mat_textures = nullptr;
  ctx->d_tri_uv = nullptr;
}

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Tells the kernels of the context's current batch to stop: the batch's epoch goes into the word they poll.
void send_cancel(rtb_ctx* ctx) {
  if (ctx->cancel_sent) return;
  ctx->cancel_sent = true;
  DeviceGuard g(ctx->device);
  *ctx->h_epoch = ctx->cancel_epoch;
  cudaMemcpyAsync(ctx->d_cancel, ctx->h_epoch, sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->cancel_stream);
}

// cudaStreamSynchronize of every context's stream that keeps relaying the caller's CancellationToken into the flags the
// running kernels poll.
cudaError_t wait_for_streams(rtb_ctx* const* ctxs, int n, const volatile uint8_t* cancel) {
  DeviceGuard restore(ctxs[0]->device);
  if (!cancel) {
    for (int i = 0; i < n; i++) {
      cudaError_t e = cudaSetDevice(ctxs[i]->device);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctxs[i]->stream);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  }
  for (int i = 0; i < n; i++) {
    cudaError_t e = cudaSetDevice(ctxs[i]->device);
    if (e == cudaSuccess) e = cudaEventRecord(ctxs[i]->ev_done, ctxs[i]->stream);
    if (e != cudaSuccess) return e;
  }
  const auto t0 = std::chrono::steady_clock::now();
  int pending = n;
  std::vector<char> done((size_t)n, 0);
  for (uint32_t spins = 0; pending > 0; spins++) {
    if (*cancel)
      for (int i = 0; i < n; i++) send_cancel(ctxs[i]);
    for (int i = 0; i < n; i++) {
      if (done[i]) continue;
      const cudaError_t e = cudaEventQuery(ctxs[i]->ev_done);
      if (e == cudaErrorNotReady) continue;
      if (e != cudaSuccess) return e;
      done[i] = 1;
      pending--;
    }
    // short batches: spin (like the runtime's own blocking wait); long ones: yield, then nap 20 us between polls
    if (pending == 0 || spins < 256) continue;
    if (std::chrono::steady_clock::now() - t0 < std::chrono::milliseconds(2)) std::this_thread::yield();
    else std::this_thread::sleep_for(std::chrono::microseconds(20));
  }
  return cudaSuccess;
}



// One host-buffer batch of one context in three phases, so that a multi-device batch (rtb_multi_sample_batch) can start
// every device before it waits for any:  begin = validate, stage the inputs (pageable hosts only), launch;
// drain = queue the read-back of the outputs (pageable hosts only), the counters and the status word; then the caller
// waits for the stream(s); end = kernel time, cancellation and status verdicts.  The caller holds ctx->mu throughout.
struct HostBatch {
  bool empty = true, in_place = false;
  int width = 0, height = 0;
  ActiveRows all{};
  const rtb_batch_buffers* host = nullptr;
  uint32_t status = 0;
};

int host_batch_begin(rtb_ctx* ctx, const rtb_batch_params* params, const rtb_batch_buffers* host, HostBatch* st) {
  int width, height;
  int rc = validate_params(ctx, params, &width, &height);
  if (rc != RTB_OK) return rc;
  if (!ctx->has_scene) return fail(ctx, RTB_ERR_NO_SCENE, "rtb_upload_scene has not been called");
  if (!host || !host->in_color || !host->in_sample_count_weight || !host->in_normal || !host->in_albedo || !host->out_color ||
      !host->out_sample_count_weight || !host->out_normal || !host->out_albedo)
    return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "a batch buffer pointer is NULL");
  DeviceGuard g(ctx->device);
  const size_t pixels = (size_t)width * height;
  cudaStream_t s = ctx->stream;
  const ActiveRows all = active_rows(*params, height, 0, 0);
  st->empty = all.n_rows <= 0;
  st->width = width; st->height = height; st->all = all; st->host = host;
  if (st->empty) return RTB_OK;

  // Pinned host arrays (rtb_register_host_buffer, or any cudaHostAlloc'd memory) are read and written IN PLACE by the
  // kernel over PCIe: each accumulator crosses the bus once, inside the kernel, overlapped with tracing, instead of in
  // eight staged copies around it.  Pageable arrays take the staged path.
  rtb_batch_buffers dev{};
  bool in_place = ctx->opt_host_access != 0;
  if (in_place) {
    const void* hp[9] = {host->in_color, host->in_sample_count_weight, host->in_normal, host->in_albedo, host->out_color,
                         host->out_sample_count_weight, host->out_normal, host->out_albedo, host->out_diagnostics};
    void* dp[9] = {};
    for (int i = 0; i < 9 && in_place; i++) {
      if (!hp[i]) continue;             // diagnostics may be NULL
      cudaPointerAttributes at{};
      if (cudaPointerGetAttributes(&at, hp[i]) != cudaSuccess || at.type != cudaMemoryTypeHost || !at.devicePointer) {
        cudaGetLastError();
        in_place = false;
      } else {
        dp[i] = at.devicePointer;
      }
    }
    if (in_place) {
      dev.in_color = (const float*)dp[0]; dev.in_sample_count_weight = (const float*)dp[1];
      dev.in_normal = (const float*)dp[2]; dev.in_albedo = (const float*)dp[3];
      dev.out_color = (float*)dp[4]; dev.out_sample_count_weight = (float*)dp[5];
      dev.out_normal = (float*)dp[6]; dev.out_albedo = (float*)dp[7];
      dev.out_diagnostics = (rtb_diagnostics*)dp[8];
    }
  }
  ctx->last_in_place = st->in_place = in_place;
  if (!in_place) {
    if ((rc = ensure_buffers(ctx, pixels)) != RTB_OK) return rc;
    DeviceBuffers& d = ctx->buf;
    RTB_CUDA(ctx, copy_rows(d.in_color, host->in_color, 16, width, all, cudaMemcpyHostToDevice, s));
    RTB_CUDA(ctx, copy_rows(d.in_weight, host->in_sample_count_weight, 4, width, all, cudaMemcpyHostToDevice, s));
    RTB_CUDA(ctx, copy_rows(d.in_normal, host->in_normal, 12, width, all, cudaMemcpyHostToDevice, s));
    RTB_CUDA(ctx, copy_rows(d.in_albedo, host->in_albedo, 12, width, all, cudaMemcpyHostToDevice, s));
    dev.in_color = d.in_color; dev.in_sample_count_weight = d.in_weight; dev.in_normal = d.in_normal; dev.in_albedo = d.in_albedo;
    dev.out_color = d.out_color; dev.out_sample_count_weight = d.out_weight; dev.out_normal = d.out_normal; dev.out_albedo = d.out_albedo;
    dev.out_diagnostics = host->out_diagnostics ? d.diagnostics : nullptr;
  }

  // CancellationToken (SampleBatchJob.cs:61; the host flips it through a raw pointer, Raytracer.cs:189-192, then
  // Complete()s, :512-515): ONE launch; the kernel polls the context's mapped flag, wait_for_streams relays the token into it.
  ctx->cancel_sent = false;
  RTB_CUDA(ctx, cudaEventRecord(ctx->ev_start, s));
  if ((rc = launch_batch(ctx, *params, dev, width, height, all, s)) != RTB_OK) return rc;
  RTB_CUDA(ctx, cudaEventRecord(ctx->ev_stop, s));
  return RTB_OK;
}

int host_batch_drain(rtb_ctx* ctx, HostBatch* st) {
  DeviceGuard g(ctx->device);
  cudaStream_t s = ctx->stream;
  const rtb_batch_buffers* host = st->host;
  if (!st->in_place) {
    DeviceBuffers& d = ctx->buf;
    RTB_CUDA(ctx, copy_rows(host->out_color, d.out_color, 16, st->width, st->all, cudaMemcpyDeviceToHost, s));
    RTB_CUDA(ctx, copy_rows(host->out_sample_count_weight, d.out_weight, 4, st->width, st->all, cudaMemcpyDeviceToHost, s));
    RTB_CUDA(ctx, copy_rows(host->out_normal, d.out_normal, 12, st->width, st->all, cudaMemcpyDeviceToHost, s));
    RTB_CUDA(ctx, copy_rows(host->out_albedo, d.out_albedo, 12, st->width, st->all, cudaMemcpyDeviceToHost, s));
    if (host->out_diagnostics)
      RTB_CUDA(ctx, copy_rows(host->out_diagnostics, d.diagnostics, sizeof(rtb_diagnostics), st->width, st->all, cudaMemcpyDeviceToHost, s));
  }
  if (ctx->opt_counters)
    RTB_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, sizeof(rtb_counters), cudaMemcpyDeviceToHost, s));
  if (ctx->scene.has_volumes) RTB_CUDA(ctx, cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  return RTB_OK;
}

int host_batch_end(rtb_ctx* ctx, HostBatch* st, const volatile uint8_t* cancel) {
  DeviceGuard g(ctx->device);
  RTB_CUDA(ctx, cudaEventElapsedTime(&ctx->last_ms, ctx->ev_start, ctx->ev_stop));
  if (ctx->cancel_sent) cudaStreamSynchronize(ctx->cancel_stream);   // the pinned source is reused by the next batch
  if (ctx->cancel_sent || (cancel && *cancel)) return fail(ctx, RTB_ERR_CANCELLED, "cancelled");
  if (ctx->opt_counters) ctx->counters = *ctx->h_counters;
  if (ctx->scene.has_volumes) st->status = *ctx->h_status;
  if (st->status & kStatusHitListOverflow) {
    cudaMemsetAsync(ctx->d_status, 0, sizeof(uint32_t), ctx->stream);
    return fail(ctx, RTB_ERR_UNSUPPORTED, "a ray met %d or more entities in a world with participating media (the reference's hit list grows, "
                "HybridCollections.cs:65-71; this kernel's does not): outputs are not the reference's", kMaxRayHits);
  }
  return RTB_OK;
}


// ---- one host, N GPUs ------------------------------------------------------------------------
// Row tiles with near-equal modelled cost: bounds[g] = first row of device g's tile, bounds[n] = hi.  Every tile of a
// non-empty range gets at least one row while rows last.
void balance_rows(const double* cost, int lo, int hi, int n, int* bounds) {
  const int rows = std::max(0, hi - lo);
  std::vector<double> cum((size_t)rows + 1, 0.0);
  for (int r = 0; r < rows; r++) cum[(size_t)r + 1] = cum[(size_t)r] + std::max(cost ? cost[lo + r] : 1.0, 1e-12);
  bounds[0] = lo;
  for (int g = 1; g < n; g++) {
    const double target = cum[(size_t)rows] * (double)g / (double)n;
    int b = (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
    b = std::max(b, bounds[g - 1] - lo + 1);
    b = std::min(b, rows - (n - g));
    b = std::max(b, bounds[g - 1] - lo);       // more devices than rows: empty tiles at the end
    bounds[g] = lo + std::min(std::max(b, 0), rows);
  }
  bounds[n] = hi;
}

uint64_t fnv1a(uint64_t h, const void* data, size_t bytes) {
  const unsigned char* p = static_cast<const unsigned char*>(data);
  for (size_t i = 0; i < bytes; i++) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

}  // namespace

// rtb_multi: the frame of ONE sample job rendered by every GPU of the box behind one call — what the reference's single
// call site (Raytracer.cs:671-736) needs when the host has N devices.  Every pixel is independent and the Philox stream
// is keyed by the global pixel index, so device g renders rows [bounds[g], bounds[g + 1]) of the SAME
// buffers: host arrays in place (pinned) or staged per device (pageable), or device arrays on one GPU that the others
// reach over NVLink peer access.  No gather, no collective: the tiles land where the consumer reads them.
struct rtb_multi {
  std::vector<rtb_ctx*> ctx;
  std::string last_error;
  std::mutex mu;
  // tile balancer: per-row cost model (instrumented probe batch) corrected by measured kernel times
  std::vector<double> row_cost;
  uint64_t model_key = 0;
  uint64_t scene_generation = 1;
  std::vector<int> bounds;              // n + 1 row bounds of the last batch
  std::vector<float> kernel_ms;         // per device, last batch
  bool times_pending = false;           // device-buffer batches: kernel times are read when their events have completed
  std::vector<char> peer_enabled;       // [from * n + to]
  int64_t opt_balance = 1;
  std::vector<cudaEvent_t> ev_join;     // per device: end of its part of a device-buffer batch
  cudaEvent_t ev_fork = nullptr;        // on the owner's stream: inputs ready
};

namespace {

int mfail(rtb_multi* m, int code, const char* fmt, ...) {
  char msg[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(msg, sizeof msg, fmt, ap);
  va_end(ap);
  if (m) m->last_error = msg;
  g_thread_error = msg;
  return code;
}

int mforward(rtb_multi* m, int i, int rc) {     // a per-device failure becomes the multi handle's error
  if (rc != RTB_OK) m->last_error = "device " + std::to_string(m->ctx[i]->device) + ": " + m->ctx[i]->last_error;
  return rc;
}

// The cost model of the rows [lo, hi) for this (world, size, view): each device probes an equal share of the rows with
// the instrumented kernel at a few samples per pixel and reduces its diagnostics to per-row costs.
int multi_probe(rtb_multi* m, const rtb_batch_params& params, int width, int height, int lo, int hi) {
  const int n = (int)m->ctx.size();
  m->row_cost.assign((size_t)height, 0.0);
  std::vector<int> eq((size_t)n + 1);
  balance_rows(nullptr, lo, hi, n, eq.data());
  std::vector<std::vector<float>> part((size_t)n);
  std::vector<float*> d_cost((size_t)n, nullptr);
  int rc = RTB_OK;
  for (int g = 0; g < n && rc == RTB_OK; g++) {
    rtb_ctx* c = m->ctx[g];
    if (eq[g + 1] <= eq[g]) continue;
    DeviceGuard dg(c->device);
    if ((rc = ensure_buffers(c, (size_t)width * height)) != RTB_OK) break;
    rtb_batch_params p = params;
    p.seed = 12345u;
    const uint32_t k = std::max(1u, std::min(8u, std::max(params.sample_count_range[0], params.sample_count_range[1])));
    p.sample_count_range[0] = p.sample_count_range[1] = k;
    p.row_begin = eq[g];
    p.row_end = eq[g + 1];
    DeviceBuffers& d = c->buf;
    const size_t off = (size_t)eq[g] * width, cnt = (size_t)(eq[g + 1] - eq[g]) * width;
    cudaMemsetAsync(d.in_color + 4 * off, 0, cnt * 16, c->stream);
    cudaMemsetAsync(d.in_weight + off, 0, cnt * 4, c->stream);
    cudaMemsetAsync(d.in_normal + 3 * off, 0, cnt * 12, c->stream);
    cudaMemsetAsync(d.in_albedo + 3 * off, 0, cnt * 12, c->stream);
    rtb_batch_buffers dev{};
    dev.in_color = d.in_color; dev.in_sample_count_weight = d.in_weight; dev.in_normal = d.in_normal; dev.in_albedo = d.in_albedo;
    dev.out_color = d.out_color; dev.out_sample_count_weight = d.out_weight; dev.out_normal = d.out_normal; dev.out_albedo = d.out_albedo;
    dev.out_diagnostics = d.diagnostics;
    const int64_t was = c->opt_counters;
    c->opt_counters = 1;
    rc = launch_batch(c, p, dev, width, height, active_rows(p, height, 0, 0), c->stream);
    c->opt_counters = was;
    if (rc != RTB_OK) { mforward(m, g, rc); break; }
    if (cudaMalloc(&d_cost[g], (size_t)height * sizeof(float)) != cudaSuccess) { rc = mfail(m, RTB_ERR_OUT_OF_MEMORY, "out of device memory"); break; }
    row_cost_kernel<<<(unsigned)(eq[g + 1] - eq[g]), 256, 0, c->stream>>>(d.diagnostics, width, eq[g], d_cost[g]);
    part[g].resize((size_t)height);
    cudaMemcpyAsync(part[g].data() + eq[g], d_cost[g] + eq[g], (size_t)(eq[g + 1] - eq[g]) * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
  }
  for (int g = 0; g < n; g++) {
    rtb_ctx* c = m->ctx[g];
    DeviceGuard dg(c->device);
    const cudaError_t e = cudaStreamSynchronize(c->stream);
    if (d_cost[g]) cudaFree(d_cost[g]);
    if (e != cudaSuccess && rc == RTB_OK) rc = mfail(m, RTB_ERR_CUDA + (int)e, "probe batch: %s", cudaGetErrorString(e));
    if (rc == RTB_OK && !part[g].empty())
      for (int r = eq[g]; r < eq[g + 1]; r++) m->row_cost[(size_t)r] = (double)part[g][(size_t)r];
  }
  return rc;
}

// Row bounds of this batch: probe when the model does not describe this (world, size, view, slice), else reuse the model
// the last batches' kernel times corrected.
int multi_tiles(rtb_multi* m, const rtb_batch_params& params, int width, int height, int* lo_out, int* hi_out) {
  const int n = (int)m->ctx.size();
  int lo = 0, hi = height;
  if (params.row_end > params.row_begin) { lo = std::max(lo, params.row_begin); hi = std::min(hi, params.row_end); }
  *lo_out = lo; *hi_out = hi;
  m->bounds.assign((size_t)n + 1, lo);
  if (!m->opt_balance || n == 1) {
    balance_rows(nullptr, lo, hi, n, m->bounds.data());
    return RTB_OK;
  }
  uint64_t key = 1469598103934665603ull;
  key = fnv1a(key, &m->scene_generation, sizeof m->scene_generation);
  key = fnv1a(key, params.size, sizeof params.size);
  key = fnv1a(key, &params.view, sizeof params.view);
  key = fnv1a(key, &params.trace_depth, sizeof params.trace_depth);
  const int range[2] = {lo, hi};
  key = fnv1a(key, range, sizeof range);
  if (key != m->model_key || m->row_cost.size() != (size_t)height) {
    const int rc = multi_probe(m, params, width, height, lo, hi);
    if (rc != RTB_OK) return rc;
    m->model_key = key;
  }
  // rows the interlace test skips cost nothing
  std::vector<double> cost(m->row_cost);
  if (params.slice_divider > 1)
    for (int r = 0; r < height; r++)
      if (r % params.slice_divider != params.slice_offset) cost[(size_t)r] = 0.0;
  balance_rows(cost.data(), lo, hi, n, m->bounds.data());
  return RTB_OK;
}

// measured kernel time of each tile -> the model: scale the tile's rows so that they sum to the time it took
void multi_feedback(rtb_multi* m) {
  const int n = (int)m->ctx.size();
  if (!m->opt_balance || n == 1 || m->row_cost.empty()) return;
  for (int g = 0; g < n; g++) {
    double c = 0.0;
    for (int r = m->bounds[g]; r < m->bounds[g + 1]; r++) c += m->row_cost[(size_t)r];
    if (c > 0.0 && m->kernel_ms[g] > 0.0f)
      for (int r = m->bounds[g]; r < m->bounds[g + 1]; r++) m->row_cost[(size_t)r] *= (double)m->kernel_ms[g] / c;
  }
}

void multi_collect_pending_times(rtb_multi* m) {
  if (!m->times_pending) return;
  const int n = (int)m->ctx.size();
  for (int g = 0; g < n; g++)
    if (m->bounds[g + 1] > m->bounds[g] && cudaEventQuery(m->ctx[g]->ev_stop) != cudaSuccess) { cudaGetLastError(); return; }
  for (int g = 0; g < n; g++) {
    m->kernel_ms[g] = 0.0f;
    if (m->bounds[g + 1] > m->bounds[g]) cudaEventElapsedTime(&m->kernel_ms[g], m->ctx[g]->ev_start, m->ctx[g]->ev_stop);
  }
  m->times_pending = false;
  multi_feedback(m);
}

template <typename F>
int multi_each(rtb_multi* m, F f) {
  if (!m) return mfail(nullptr, RTB_ERR_INVALID_ARGUMENT, "multi handle is NULL");
  std::lock_guard<std::mutex> lock(m->mu);
  for (size_t i = 0; i < m->ctx.size(); i++) {
    const int rc = f(m->ctx[i]);
    if (rc != RTB_OK) return mforward(m, (int)i, rc);
  }
  m->scene_generation++;
  return RTB_OK;
}

}  // namespace

extern "C" {

int rtb_abi_version(void) { return RTB_ABI_VERSION; }

int rtb_create(int device, rtb_ctx** out_ctx) {
  if (!out_ctx) return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "out_ctx is NULL");
  *out_ctx = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess) return fail(nullptr, RTB_ERR_CUDA + (int)e, "no CUDA device: %s (there is no CPU fallback)", cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "device %d out of range (%d devices)", device, count);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(nullptr, RTB_ERR_CUDA + (int)e, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10) return fail(nullptr, RTB_ERR_UNSUPPORTED, "device %d is sm_%d%d; librtb is built for sm_100a only", device, prop.major, prop.minor);
  rtb_ctx* ctx = new (std::nothrow) rtb_ctx();
  if (!ctx) return fail(nullptr, RTB_ERR_OUT_OF_MEMORY, "out of host memory");
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  if (const char* k = getenv("RTB_KERNEL")) {         // experiment knob; RTB_OPT_KERNEL is the API
    const long v = strtol(k, nullptr, 10);
    if (v >= 1 && v <= 2) ctx->default_kernel = (int)v;
  }
  if (const char* k = getenv("RTB_RETREE")) ctx->opt_retree = std::max(0l, std::min(2l, strtol(k, nullptr, 10)));   // experiment knob; RTB_OPT_RETREE is the API
  if (const char* k = getenv("RTB_LEAF_SPHERES")) {   // experiment knob; RTB_OPT_LEAF_SPHERES is the API
    const long v = strtol(k, nullptr, 10);
    if (v >= 1 && v <= 15) ctx->opt_collapse = v;
  }
  DeviceGuard g(device);
  auto bail = [&](cudaError_t err, const char* what) {
    int rc = fail(nullptr, RTB_ERR_CUDA + (int)err, "%s: %s", what, cudaGetErrorString(err));
    rtb_destroy(ctx);
    return rc;
  };
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
  if ((e = cudaEventCreate(&ctx->ev_start)) != cudaSuccess) return bail(e, "cudaEventCreate");
  if ((e = cudaEventCreate(&ctx->ev_stop)) != cudaSuccess) return bail(e, "cudaEventCreate");
  if ((e = cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
  if ((e = cudaMalloc(&ctx->d_tile_ring, rtb_ctx::kTileRing * sizeof(uint32_t))) != cudaSuccess) return bail(e, "cudaMalloc");
  if ((e = cudaMalloc(&ctx->d_status, sizeof(uint32_t))) != cudaSuccess) return bail(e, "cudaMalloc");
  if ((e = cudaMemset(ctx->d_status, 0, sizeof(uint32_t))) != cudaSuccess) return bail(e, "cudaMemset");
  if ((e = cudaMalloc(&ctx->d_cancel, sizeof(uint32_t))) != cudaSuccess) return bail(e, "cudaMalloc");
  if ((e = cudaMemset(ctx->d_cancel, 0, sizeof(uint32_t))) != cudaSuccess) return bail(e, "cudaMemset");
  if ((e = cudaHostAlloc((void**)&ctx->h_epoch, 256, cudaHostAllocDefault)) != cudaSuccess) return bail(e, "cudaHostAlloc");
  memset(ctx->h_epoch, 0, 256);
  ctx->h_status = ctx->h_epoch + 4;
  ctx->h_counters = reinterpret_cast<rtb_counters*>(reinterpret_cast<char*>(ctx->h_epoch) + 64);
  static_assert(sizeof(rtb_counters) <= 192, "pinned block");
  if ((e = cudaStreamCreateWithFlags(&ctx->cancel_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
  if ((e = cudaMalloc(&ctx->d_counters, 8 * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMalloc");
  if ((e = cudaMalloc(&ctx->d_metrics_partial, 1024 * sizeof(MetricsAcc))) != cudaSuccess) return bail(e, "cudaMalloc");
  *out_ctx = ctx;
  return RTB_OK;
}

int rtb_destroy(rtb_ctx* ctx) {
  if (!ctx) return RTB_OK;
  {
    DeviceGuard g(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->registered) cudaHostUnregister(kv.first);
    DeviceBuffers& b = ctx->buf;
    void* ptrs[] = {b.in_color, b.in_weight, b.in_normal, b.in_albedo, b.out_color, b.out_weight, b.out_normal, b.out_albedo,
                    b.diagnostics, ctx->d_sky, ctx->d_tex_pixels, ctx->d_tex_images, ctx->d_mat_textures, ctx->d_tri_uv, ctx->d_blob, ctx->d_materials, ctx->d_chain_ref, ctx->d_chain_boxes, ctx->d_tile_ring, ctx->d_status, ctx->d_cancel, ctx->d_counters, ctx->d_metrics_partial, ctx->d_scratch_diag};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (ctx->cancel_stream) { cudaStreamSynchronize(ctx->cancel_stream); cudaStreamDestroy(ctx->cancel_stream); }
    if (ctx->h_epoch) cudaFreeHost(ctx->h_epoch);
    if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    if (ctx->ev_stop) cudaEventDestroy(ctx->ev_stop);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
  }
  delete ctx;
  return RTB_OK;
}

const char* rtb_last_error(const rtb_ctx* ctx) { return ctx ? ctx->last_error.c_str() : g_thread_error.c_str(); }

int rtb_set_log_callback(rtb_ctx* ctx, rtb_log_fn fn, void* user) {
  if (!ctx) return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "ctx is NULL");
  ctx->log_fn = fn;
  ctx->log_user = user;
  return RTB_OK;
}

int rtb_upload_scene(rtb_ctx* ctx, const rtb_sphere* spheres, size_t sphere_count, const rtb_material* materials,
                     size_t material_count, const rtb_bvh_node* nodes, size_t node_count) {
  return rtb_upload_world(ctx, nullptr, 0, spheres, sphere_count, nullptr, 0, materials, material_count, nodes, node_count);
}

int rtb_upload_world(rtb_ctx* ctx, const rtb_entity* entities, size_t entity_count, const rtb_sphere* spheres, size_t sphere_count,
                     const rtb_triangle* triangles, size_t triangle_count, const rtb_material* materials,
                     size_t material_count, const rtb_bvh_node* nodes, size_t node_count) {
  return rtb_upload_placed_world(ctx, entities, entity_count, spheres, sphere_count, triangles, triangle_count, nullptr, 0, materials,
                                 material_count, nodes, node_count);
}

int rtb_upload_placed_world(rtb_ctx* ctx, const rtb_entity* entities, size_t entity_count, const rtb_sphere* spheres,
                            size_t sphere_count, const rtb_triangle* triangles, size_t triangle_count,
                            const rtb_placed_entity* placed, size_t placed_count, const rtb_material* materials,
                            size_t material_count, const rtb_bvh_node* nodes, size_t node_count) {
  if (!ctx) return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "ctx is NULL");
  if ((sphere_count && !spheres) || (material_count && !materials) || (node_count && !nodes) || (entity_count && !entities) ||
      (triangle_count && !triangles) || (placed_count && !placed))
    return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "NULL array with a non-zero count");
  if ((triangle_count || placed_count) && !entity_count) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "triangles and placed entities need an entity list");
  if (sphere_count > (1u << 27) || triangle_count > (1u << 25) || placed_count > (1u << 24) || entity_count > (1u << 27) ||
      node_count > (1u << 29))
    return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "scene too large");
  std::lock_guard<std::mutex> lock(ctx->mu);
  HostBlob hb;
  int status;
  const char* err = build_blob(entity_count ? entities : nullptr, entity_count, spheres, sphere_count, triangles, triangle_count,
                               placed, placed_count, materials, material_count, nodes, node_count, (uint32_t)ctx->opt_collapse,
                               (int)ctx->opt_retree, &hb, &status);
  if (err) return fail(ctx, status, "rtb_upload_scene: %s", err);
  DeviceGuard g(ctx->device);
  RTB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->d_blob) cudaFree(ctx->d_blob);
  if (ctx->d_materials) cudaFree(ctx->d_materials);
  if (ctx->d_chain_ref) cudaFree(ctx->d_chain_ref);
  if (ctx->d_chain_boxes) cudaFree(ctx->d_chain_boxes);
  ctx->d_blob = nullptr;
  ctx->d_materials = nullptr;
  ctx->d_chain_ref = nullptr;
  ctx->d_chain_boxes = nullptr;
  ctx->has_scene = false;
  drop_textures(ctx);
  ctx->host_materials = hb.materials;
  RTB_CUDA(ctx, cudaMalloc(&ctx->d_blob, hb.bytes.size()));
  RTB_CUDA(ctx, cudaMemcpy(ctx->d_blob, hb.bytes.data(), hb.bytes.size(), cudaMemcpyHostToDevice));
  RTB_CUDA(ctx, cudaMalloc(&ctx->d_materials, hb.materials.size() * sizeof(DevMaterial)));
  RTB_CUDA(ctx, cudaMemcpy(ctx->d_materials, hb.materials.data(), hb.materials.size() * sizeof(DevMaterial), cudaMemcpyHostToDevice));
  RTB_CUDA(ctx, cudaMalloc(&ctx->d_chain_ref, std::max<size_t>(hb.chain_ref.size(), 1) * sizeof(uint32_t)));
  RTB_CUDA(ctx, cudaMalloc(&ctx->d_chain_boxes, std::max<size_t>(hb.chain_boxes.size(), 8) * sizeof(float)));
  if (!hb.chain_ref.empty())
    RTB_CUDA(ctx, cudaMemcpy(ctx->d_chain_ref, hb.chain_ref.data(), hb.chain_ref.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  if (!hb.chain_boxes.empty())
    RTB_CUDA(ctx, cudaMemcpy(ctx->d_chain_boxes, hb.chain_boxes.data(), hb.chain_boxes.size() * sizeof(float), cudaMemcpyHostToDevice));
  if (material_count) {
    derive_materials_kernel<<<(unsigned)((material_count + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_materials, (uint32_t)material_count);
    RTB_CUDA(ctx, cudaGetLastError());
    RTB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  ctx->scene = hb.desc;
  ctx->scene.blob = ctx->d_blob;
  ctx->scene.materials = ctx->d_materials;
  ctx->scene.chain_ref = ctx->d_chain_ref;
  ctx->scene.chain_boxes = ctx->d_chain_boxes;
  if (ctx->scene.has_chains && ctx->opt_walk_chains) ctx->scene.has_chains = 2u;
  ctx->has_scene = true;
  return RTB_OK;
}

int rtb_upload_textures(rtb_ctx* ctx, const rtb_image* images, size_t image_count, const rtb_material_textures* mt,
                        size_t material_count, const float* triangle_uvs, size_t triangle_count) {
  if (!ctx) return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "ctx is NULL");
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (!ctx->has_scene) return fail(ctx, RTB_ERR_NO_SCENE, "rtb_upload_textures before a world was uploaded");
  if ((image_count && !images) || (material_count && !mt)) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "NULL array with a non-zero count");
  if (material_count != ctx->scene.n_materials) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_upload_textures: material count differs from the world's");
  if (triangle_uvs && triangle_count != ctx->scene.n_triangles)
    return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_upload_textures: triangle count differs from the world's");
  if (image_count > (1u << 20)) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "too many images");
  std::vector<int4> desc(std::max<size_t>(image_count, 1));
  size_t total = 0;
  for (size_t i = 0; i < image_count; i++) {
    const rtb_image& im = images[i];
    if (!im.pixels || im.width < 1 || im.height < 1 || im.width > 32768 || im.height > 32768 || (im.pixel_stride != 3 && im.pixel_stride != 4))
      return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_upload_textures: image %zu: need pixels, a size in 1..32768 and a stride of 3 or 4", i);
    if (total > 0x7fffffffu) return fail(ctx, RTB_ERR_UNSUPPORTED, "more than 2 GB of image data");
    desc[i] = make_int4((int)total, im.width, im.height, im.pixel_stride);
    total += (size_t)im.width * im.height * im.pixel_stride;
  }
  std::vector<int4> mat(std::max<size_t>(material_count, 1) * 2);
  std::vector<DevMaterial> dm = ctx->host_materials;
  for (size_t i = 0; i < material_count; i++) {
    const int32_t idx[4] = {mt[i].albedo_image, mt[i].emission_image, mt[i].glossiness_image, mt[i].metallic_image};
    const int32_t ch[4] = {0, 0, mt[i].glossiness_channel, mt[i].metallic_channel};
    bool any = false;
    for (int k = 0; k < 4; k++) {
      if (idx[k] < 0) continue;
      if ((size_t)idx[k] >= image_count) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_upload_textures: material %zu names image %d", i, idx[k]);
      const int need = k < 2 ? 3 : ch[k] + 1;      // a colour reads bytes 0..2 (Texture.cs:88), a scalar its channel (:136)
      if (ch[k] < 0 || need > images[idx[k]].pixel_stride) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_upload_textures: material %zu reads a channel its image does not have", i);
      any = true;
    }
    mat[2 * i] = make_int4(idx[0], idx[1], idx[2], idx[3]);
    mat[2 * i + 1] = make_int4(ch[2], ch[3], 0, 0);
    dm[i].textured = any ? 1u : 0u;
    // Material.IsPerfectSpecular needs CONSTANT Metallic and Glossiness textures (Material.cs:190-192)
    if (dm[i].type == RTB_MATERIAL_STANDARD && (idx[2] >= 0 || idx[3] >= 0)) dm[i].perfect_specular = 0;
  }
  DeviceGuard g(ctx->device);
  RTB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  drop_textures(ctx);
  RTB_CUDA(ctx, cudaMalloc(&ctx->d_tex_pixels, std::max<size_t>(total, 16)));
  for (size_t i = 0; i < image_count; i++)
    RTB_CUDA(ctx, cudaMemcpy(ctx->d_tex_pixels + desc[i].x, images[i].pixels, (size_t)images[i].width * images[i].height * images[i].pixel_stride,
                             cudaMemcpyHostToDevice));
  RTB_CUDA(ctx, cudaMalloc(&ctx->d_tex_images, desc.size() * sizeof(int4)));
  RTB_CUDA(ctx, cudaMemcpy(ctx->d_tex_images, desc.data(), desc.size() * sizeof(int4), cudaMemcpyHostToDevice));
  RTB_CUDA(ctx, cudaMalloc(&ctx->d_mat_textures, mat.size() * sizeof(int4)));
  RTB_CUDA(ctx, cudaMemcpy(ctx->d_mat_textures, mat.data(), mat.size() * sizeof(int4), cudaMemcpyHostToDevice));
  if (triangle_uvs && triangle_count) {
    RTB_CUDA(ctx, cudaMalloc(&ctx->d_tri_uv, triangle_count * 6 * sizeof(float)));
    RTB_CUDA(ctx, cudaMemcpy(ctx->d_tri_uv, triangle_uvs, triangle_count * 6 * sizeof(float), cudaMemcpyHostToDevice));
  }
  if (material_count) {
    RTB_CUDA(ctx, cudaMemcpy(ctx->d_materials, dm.data(), material_count * sizeof(DevMaterial), cudaMemcpyHostToDevice));
    derive_materials_kernel<<<(unsigned)((material_count + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_materials, (uint32_t)material_count);
    RTB_CUDA(ctx, cudaGetLastError());
    RTB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return RTB_OK;
}

int rtb_upload_sky_cubemap(rtb_ctx* ctx, const uint16_t* half_rgba, int face_width, int face_height) {
  if (!ctx) return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "ctx is NULL");
  if (half_rgba && (face_width < 1 || face_height < 1 || face_width > 16384 || face_height > 16384))
    return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_upload_sky_cubemap: bad face size");
  std::lock_guard<std::mutex> lock(ctx->mu);
  DeviceGuard g(ctx->device);
  RTB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->d_sky) cudaFree(ctx->d_sky);
  ctx->d_sky = nullptr;
  ctx->sky_w = ctx->sky_h = 0;
  if (!half_rgba) return RTB_OK;
  const size_t bytes = (size_t)6 * face_width * face_height * 4 * sizeof(uint16_t);
  RTB_CUDA(ctx, cudaMalloc(&ctx->d_sky, bytes));
  RTB_CUDA(ctx, cudaMemcpy(ctx->d_sky, half_rgba, bytes, cudaMemcpyHostToDevice));
  ctx->sky_w = face_width;
  ctx->sky_h = face_height;
  return RTB_OK;
}

int rtb_retree_bvh(const rtb_bvh_node* nodes, size_t node_count, rtb_bvh_node* out_nodes, size_t capacity, size_t* out_count) {
  if (!out_count) return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "out_count is NULL");
  *out_count = 0;
  if (node_count && !nodes) return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "NULL array with a non-zero count");
  std::vector<rtb_bvh_node> rebuilt;
  if (!rtb_retree::retree(nodes, node_count, kStackMax - 2, rebuilt)) return RTB_OK;
  if (rebuilt.size() > capacity || !out_nodes) return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "rtb_retree_bvh: %zu nodes needed", rebuilt.size());
  memcpy(out_nodes, rebuilt.data(), rebuilt.size() * sizeof(rtb_bvh_node));
  *out_count = rebuilt.size();
  return RTB_OK;
}

int rtb_describe_scene(const rtb_sphere* spheres, size_t sphere_count, const rtb_material* materials, size_t material_count,
                       const rtb_bvh_node* nodes, size_t node_count, int leaf_spheres, rtb_scene_layout* out) {
  if (!out) return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "out is NULL");
  if ((sphere_count && !spheres) || (material_count && !materials) || (node_count && !nodes))
    return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "NULL array with a non-zero count");
  if (leaf_spheres < 1 || leaf_spheres > 15) return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "leaf_spheres must be 1..15");
  HostBlob hb;
  int status;
  const char* err = build_blob(nullptr, 0, spheres, sphere_count, nullptr, 0, nullptr, 0, materials, material_count, nodes, node_count,
                               (uint32_t)leaf_spheres, /*retree*/ 0, &hb, &status);   // the layout of the HOST's topology
  if (err) return fail(nullptr, status, "rtb_describe_scene: %s", err);
  *out = rtb_scene_layout{};
  out->inner_nodes = hb.desc.n_inner;
  out->device_spheres = hb.desc.n_spheres;
  out->max_depth = hb.desc.max_depth;
  out->blob_bytes = hb.desc.blob_bytes;
  out->chain_boxes = (uint32_t)(hb.chain_boxes.size() / 8);
  out->collapsed = hb.desc.has_chains;
  const uint32_t* lc = reinterpret_cast<const uint32_t*>(hb.bytes.data() + hb.desc.leaf_count_off);
  for (uint32_t i = 0; i < hb.desc.n_spheres; i++) {
    if (lc[i]) { out->leaves++; out->max_leaf_spheres = std::max(out->max_leaf_spheres, lc[i]); }
  }
  return RTB_OK;
}

int rtb_sample_batch_device(rtb_ctx* ctx, const rtb_batch_params* params, const rtb_batch_buffers* dev, void* cuda_stream) {
  if (!ctx) return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "ctx is NULL");
  int width, height;
  int rc = validate_params(ctx, params, &width, &height);
  if (rc != RTB_OK) return rc;
  if (!ctx->has_scene) return fail(ctx, RTB_ERR_NO_SCENE, "rtb_upload_scene has not been called");
  if (!dev || !dev->in_color || !dev->in_sample_count_weight || !dev->in_normal || !dev->in_albedo || !dev->out_color ||
      !dev->out_sample_count_weight || !dev->out_normal || !dev->out_albedo)
    return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "a batch buffer pointer is NULL");
  DeviceGuard g(ctx->device);
  cudaStream_t stream = (cudaStream_t)cuda_stream;   // NULL = the CUDA default stream, as in every CUDA API
  return launch_batch(ctx, *params, *dev, width, height, active_rows(*params, height, 0, 0), stream);
}

int rtb_sample_batch(rtb_ctx* ctx, const rtb_batch_params* params, const rtb_batch_buffers* host, const volatile uint8_t* cancel) {
  if (!ctx) return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "ctx is NULL");
  if (cancel && *cancel) return fail(ctx, RTB_ERR_CANCELLED, "cancelled");
  std::lock_guard<std::mutex> lock(ctx->mu);
  HostBatch st;
  int rc = host_batch_begin(ctx, params, host, &st);
  if (rc != RTB_OK || st.empty) return rc;
  // the kernel first, with the token relayed while it runs; only then the read-backs (a device-to-host copy into pageable
  // memory blocks the caller until the stream reaches it: queued before the wait, it would keep this thread from watching the token)
  rtb_ctx* one[1] = {ctx};
  RTB_CUDA(ctx, wait_for_streams(one, 1, cancel));
  if (!ctx->cancel_sent && !(cancel && *cancel)) {
    if ((rc = host_batch_drain(ctx, &st)) != RTB_OK) return rc;
    RTB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return host_batch_end(ctx, &st, cancel);
}

int rtb_register_host_buffer(rtb_ctx* ctx, void* ptr, size_t bytes) {
  if (!ctx || !ptr || !bytes) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_register_host_buffer: bad argument");
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (ctx->registered.count(ptr)) return RTB_OK;
  DeviceGuard g(ctx->device);
  RTB_CUDA(ctx, cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
  ctx->registered[ptr] = bytes;
  return RTB_OK;
}

int rtb_unregister_host_buffer(rtb_ctx* ctx, void* ptr) {
  if (!ctx || !ptr) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_unregister_host_buffer: bad argument");
  std::lock_guard<std::mutex> lock(ctx->mu);
  auto it = ctx->registered.find(ptr);
  if (it == ctx->registered.end()) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "buffer was not registered");
  DeviceGuard g(ctx->device);
  RTB_CUDA(ctx, cudaHostUnregister(ptr));
  ctx->registered.erase(it);
  return RTB_OK;
}

int rtb_combine_device(rtb_ctx* ctx, int width, int height, int debug_mode, int ldr_albedo, const float* color4,
                       const float* normal3, const float* albedo3, float* out_color3, float* out_normal3,
                       float* out_albedo3, void* cuda_stream) {
  if (!ctx) return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "ctx is NULL");
  if (width < 1 || height < 1 || !color4) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_combine_device: bad argument");
  if ((out_normal3 && !normal3) || (out_albedo3 && !albedo3)) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_combine_device: missing input");
  DeviceGuard g(ctx->device);
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  const int n = width * height;
  const int grid = std::min((n + 255) / 256, ctx->sm_count * 8);
  combine_kernel<<<grid, 256, 0, stream>>>(width, height, debug_mode, ldr_albedo, reinterpret_cast<const float4*>(color4),
                                           normal3, albedo3, out_color3, out_normal3, out_albedo3);
  RTB_CUDA(ctx, cudaGetLastError());
  return RTB_OK;
}

int rtb_finalize_device(rtb_ctx* ctx, int width, int height, const float* color3, const float* normal3, const float* albedo3,
                        uint32_t* out_color_rgba, uint32_t* out_normal_rgba, uint32_t* out_albedo_rgba, void* cuda_stream) {
  if (!ctx) return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "ctx is NULL");
  if (width < 1 || height < 1) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_finalize_device: bad size");
  if ((out_color_rgba && !color3) || (out_normal_rgba && !normal3) || (out_albedo_rgba && !albedo3))
    return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_finalize_device: missing input");
  DeviceGuard g(ctx->device);
  const int n = width * height;
  // four pixels per thread (float4 loads, uint4 stores) when every array is 16-byte aligned; the ragged end, or everything, per pixel
  auto aligned = [](const void* p) { return ((uintptr_t)p & 15u) == 0; };
  const bool x4 = aligned(color3) && aligned(normal3) && aligned(albedo3) && aligned(out_color_rgba) && aligned(out_normal_rgba) && aligned(out_albedo_rgba);
  const int n_quads = x4 ? n / 4 : 0;
  if (n_quads > 0) {
    const int grid = std::max(1, std::min((n_quads + 255) / 256, ctx->sm_count * 16));
    finalize_kernel_x4<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(n_quads, reinterpret_cast<const float4*>(color3), reinterpret_cast<const float4*>(normal3),
                                                                   reinterpret_cast<const float4*>(albedo3), reinterpret_cast<uint4*>(out_color_rgba),
                                                                   reinterpret_cast<uint4*>(out_normal_rgba), reinterpret_cast<uint4*>(out_albedo_rgba));
    RTB_CUDA(ctx, cudaGetLastError());
  }
  if (4 * n_quads < n) {
    const int rest = n - 4 * n_quads;
    const int grid = std::max(1, std::min((rest + 255) / 256, ctx->sm_count * 16));
    finalize_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(n, color3, normal3, albedo3, out_color_rgba, out_normal_rgba, out_albedo_rgba, 4 * n_quads);
    RTB_CUDA(ctx, cudaGetLastError());
  }
  return RTB_OK;
}

int rtb_reduce_metrics_device(rtb_ctx* ctx, int width, int height, const rtb_diagnostics* diagnostics, const float* color4,
                              const float* sample_count_weight, rtb_metrics* out_host, void* cuda_stream) {
  if (!ctx) return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "ctx is NULL");
  if (width < 1 || height < 1 || !color4 || !sample_count_weight || !out_host)
    return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_reduce_metrics_device: bad argument");
  std::lock_guard<std::mutex> lock(ctx->mu);
  DeviceGuard g(ctx->device);
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  const int n = width * height;
  const int grid = std::max(1, std::min({(n + 255) / 256, ctx->sm_count * 4, 1024}));
  reduce_metrics_kernel<<<grid, 256, 0, stream>>>(n, diagnostics, reinterpret_cast<const float4*>(color4), sample_count_weight,
                                                  ctx->d_metrics_partial);
  RTB_CUDA(ctx, cudaGetLastError());
  std::vector<MetricsAcc> part(grid);
  RTB_CUDA(ctx, cudaMemcpyAsync(part.data(), ctx->d_metrics_partial, grid * sizeof(MetricsAcc), cudaMemcpyDeviceToHost, stream));
  RTB_CUDA(ctx, cudaStreamSynchronize(stream));
  MetricsAcc t = part[0];
  for (int i = 1; i < grid; i++) {
    t.rays += part[i].rays;
    t.samples += part[i].samples;
    t.w_min = um::min(t.w_min, part[i].w_min);
    t.w_max = um::max(t.w_max, part[i].w_max);
    t.s_min = std::min(t.s_min, part[i].s_min);
    t.s_max = std::max(t.s_max, part[i].s_max);
  }
  out_host->total_ray_count = t.rays;
  out_host->total_samples = t.samples;
  out_host->sample_count_weight_min = t.w_min;
  out_host->sample_count_weight_max = t.w_max;
  out_host->sample_count_min = t.s_min;
  out_host->sample_count_max = t.s_max;
  return RTB_OK;
}

int rtb_get_counters(rtb_ctx* ctx, rtb_counters* out) {
  if (!ctx || !out) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_get_counters: bad argument");
  if (!ctx->opt_counters) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "counters are disabled (rtb_set_option(RTB_OPT_COUNTERS, 1))");
  DeviceGuard g(ctx->device);
  // device-buffer batches do not synchronise: fetch the counters behind the last instrumented launch, on ITS stream
  // (not a device-wide synchronisation: other streams of the process keep running)
  RTB_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, sizeof(rtb_counters), cudaMemcpyDeviceToHost, ctx->counters_stream));
  RTB_CUDA(ctx, cudaStreamSynchronize(ctx->counters_stream));
  ctx->counters = *ctx->h_counters;
  *out = ctx->counters;
  return RTB_OK;
}

int rtb_set_option(rtb_ctx* ctx, int option, int64_t value) {
  if (!ctx) return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "ctx is NULL");
  switch (option) {
    case RTB_OPT_COUNTERS: ctx->opt_counters = value ? 1 : 0; return RTB_OK;
    case RTB_OPT_KERNEL:
      if (value < 0 || value > 2) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "RTB_OPT_KERNEL must be 0..2");
      ctx->opt_kernel = value;
      return RTB_OK;
    case RTB_OPT_LEAF_SPHERES:
      if (value < 1 || value > 15) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "RTB_OPT_LEAF_SPHERES must be 1..15");
      ctx->opt_collapse = value;
      return RTB_OK;
    case RTB_OPT_RETREE:
      if (value < 0 || value > 2) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "RTB_OPT_RETREE must be 0, 1 or 2");
      ctx->opt_retree = value;
      return RTB_OK;
    case RTB_OPT_NOISE:
      if (value < 0 || value > 1) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "RTB_OPT_NOISE must be 0 or 1");
      ctx->opt_noise = value;
      return RTB_OK;
    case RTB_OPT_HOST_ACCESS:
      ctx->opt_host_access = value ? 1 : 0;
      return RTB_OK;
    case RTB_OPT_MATH:
      if (value < 0 || value > 1) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "RTB_OPT_MATH must be 0 (parity) or 1 (fast)");
      ctx->opt_math = value;
      return RTB_OK;
    case RTB_OPT_ALWAYS_WALK_CHAINS:
      ctx->opt_walk_chains = value ? 1 : 0;
      if (ctx->has_scene && ctx->scene.has_chains) ctx->scene.has_chains = value ? 2u : 1u;
      return RTB_OK;
  }
  return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "unknown option %d", option);
}

int rtb_measure_fp32_peak(rtb_ctx* ctx, int repeats, double* out_tflops) {
  if (!ctx || !out_tflops || repeats < 1) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_measure_fp32_peak: bad argument");
  std::lock_guard<std::mutex> lock(ctx->mu);
  DeviceGuard g(ctx->device);
  const int iters = 1 << 14, blocks = ctx->sm_count * 8, threads = 256;
  float* d_out = nullptr;
  RTB_CUDA(ctx, cudaMalloc(&d_out, (size_t)blocks * threads * sizeof(float)));
  double best = 0.0;
  for (int r = 0; r < repeats + 1; r++) {   // first launch is a warm-up
    cudaEventRecord(ctx->ev_start, ctx->stream);
    fp32_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d_out, iters, 1.0000001f, 1e-7f);
    cudaEventRecord(ctx->ev_stop, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { cudaFree(d_out); return fail(ctx, RTB_ERR_CUDA + (int)e, "fp32 peak kernel: %s", cudaGetErrorString(e)); }
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop);
    const double tf = (double)blocks * threads * iters * 16.0 * 2.0 / (ms * 1e-3) / 1e12;
    if (r > 0 && tf > best) best = tf;
  }
  cudaFree(d_out);
  *out_tflops = best;
  return RTB_OK;
}

int rtb_last_batch_in_place(rtb_ctx* ctx, int* out_in_place) {
  if (!ctx || !out_in_place) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_last_batch_in_place: bad argument");
  *out_in_place = ctx->last_in_place ? 1 : 0;
  return RTB_OK;
}

int rtb_last_kernel_ms(rtb_ctx* ctx, float* out_ms) {
  if (!ctx || !out_ms) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_last_kernel_ms: bad argument");
  *out_ms = ctx->last_ms;
  return RTB_OK;
}


// ---- one host, N GPUs (include/rtb.h "multi-device") -------------------------------------------

int rtb_balance_rows(const double* row_cost, int row_begin, int row_end, int device_count, int* out_bounds) {
  if (!out_bounds || device_count < 1 || row_begin < 0 || row_end < row_begin)
    return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "rtb_balance_rows: bad argument");
  balance_rows(row_cost, row_begin, row_end, device_count, out_bounds);
  return RTB_OK;
}

int rtb_multi_create(const int* devices, int device_count, rtb_multi** out) {
  if (!out) return mfail(nullptr, RTB_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  if (!devices || device_count < 1 || device_count > RTB_MULTI_MAX_DEVICES)
    return mfail(nullptr, RTB_ERR_INVALID_ARGUMENT, "rtb_multi_create: 1..%d devices", RTB_MULTI_MAX_DEVICES);
  rtb_multi* m = new (std::nothrow) rtb_multi();
  if (!m) return mfail(nullptr, RTB_ERR_OUT_OF_MEMORY, "out of host memory");
  for (int i = 0; i < device_count; i++) {
    rtb_ctx* c = nullptr;
    const int rc = rtb_create(devices[i], &c);
    if (rc != RTB_OK) {
      const std::string why = g_thread_error;
      rtb_multi_destroy(m);
      return mfail(nullptr, rc, "%s", why.c_str());
    }
    m->ctx.push_back(c);
  }
  m->kernel_ms.assign((size_t)device_count, 0.0f);
  m->bounds.assign((size_t)device_count + 1, 0);
  m->peer_enabled.assign((size_t)device_count * device_count, 0);
  m->ev_join.assign((size_t)device_count, nullptr);
  for (int i = 0; i < device_count; i++) {
    DeviceGuard g(devices[i]);
    cudaEventCreateWithFlags(&m->ev_join[(size_t)i], cudaEventDisableTiming);
  }
  *out = m;
  return RTB_OK;
}

int rtb_multi_destroy(rtb_multi* m) {
  if (!m) return RTB_OK;
  for (size_t i = 0; i < m->ctx.size(); i++) {
    DeviceGuard g(m->ctx[i]->device);
    cudaStreamSynchronize(m->ctx[i]->stream);
    if (i < m->ev_join.size() && m->ev_join[i]) cudaEventDestroy(m->ev_join[i]);
  }
  if (m->ev_fork) cudaEventDestroy(m->ev_fork);
  for (rtb_ctx* c : m->ctx) rtb_destroy(c);
  delete m;
  return RTB_OK;
}

int rtb_multi_device_count(const rtb_multi* m) { return m ? (int)m->ctx.size() : 0; }
rtb_ctx* rtb_multi_context(rtb_multi* m, int index) { return m && index >= 0 && (size_t)index < m->ctx.size() ? m->ctx[(size_t)index] : nullptr; }
const char* rtb_multi_last_error(const rtb_multi* m) { return m ? m->last_error.c_str() : g_thread_error.c_str(); }

int rtb_multi_set_option(rtb_multi* m, int option, int64_t value) {
  if (!m) return mfail(nullptr, RTB_ERR_INVALID_ARGUMENT, "multi handle is NULL");
  std::lock_guard<std::mutex> lock(m->mu);
  if (option == RTB_OPT_BALANCE_TILES) {
    m->opt_balance = value ? 1 : 0;
    m->model_key = 0;
    return RTB_OK;
  }
  for (size_t i = 0; i < m->ctx.size(); i++) {
    const int rc = rtb_set_option(m->ctx[i], option, value);
    if (rc != RTB_OK) return mforward(m, (int)i, rc);
  }
  return RTB_OK;
}

int rtb_multi_upload_scene(rtb_multi* m, const rtb_sphere* spheres, size_t sphere_count, const rtb_material* materials,
                           size_t material_count, const rtb_bvh_node* nodes, size_t node_count) {
  return multi_each(m, [&](rtb_ctx* c) { return rtb_upload_scene(c, spheres, sphere_count, materials, material_count, nodes, node_count); });
}

int rtb_multi_upload_placed_world(rtb_multi* m, const rtb_entity* entities, size_t entity_count, const rtb_sphere* spheres,
                                  size_t sphere_count, const rtb_triangle* triangles, size_t triangle_count,
                                  const rtb_placed_entity* placed, size_t placed_count, const rtb_material* materials,
                                  size_t material_count, const rtb_bvh_node* nodes, size_t node_count) {
  return multi_each(m, [&](rtb_ctx* c) {
    return rtb_upload_placed_world(c, entities, entity_count, spheres, sphere_count, triangles, triangle_count, placed, placed_count,
                                   materials, material_count, nodes, node_count);
  });
}

int rtb_multi_upload_textures(rtb_multi* m, const rtb_image* images, size_t image_count, const rtb_material_textures* mt,
                              size_t material_count, const float* triangle_uvs, size_t triangle_count) {
  return multi_each(m, [&](rtb_ctx* c) { return rtb_upload_textures(c, images, image_count, mt, material_count, triangle_uvs, triangle_count); });
}

int rtb_multi_upload_sky_cubemap(rtb_multi* m, const uint16_t* half_rgba, int face_width, int face_height) {
  return multi_each(m, [&](rtb_ctx* c) { return rtb_upload_sky_cubemap(c, half_rgba, face_width, face_height); });
}

int rtb_multi_register_host_buffer(rtb_multi* m, void* ptr, size_t bytes) {
  if (!m) return mfail(nullptr, RTB_ERR_INVALID_ARGUMENT, "multi handle is NULL");
  std::lock_guard<std::mutex> lock(m->mu);
  // one registration (portable + mapped) serves every device of the process
  return mforward(m, 0, rtb_register_host_buffer(m->ctx[0], ptr, bytes));
}

int rtb_multi_unregister_host_buffer(rtb_multi* m, void* ptr) {
  if (!m) return mfail(nullptr, RTB_ERR_INVALID_ARGUMENT, "multi handle is NULL");
  std::lock_guard<std::mutex> lock(m->mu);
  return mforward(m, 0, rtb_unregister_host_buffer(m->ctx[0], ptr));
}

int rtb_multi_sample_batch(rtb_multi* m, const rtb_batch_params* params, const rtb_batch_buffers* host, const volatile uint8_t* cancel) {
  if (!m) return mfail(nullptr, RTB_ERR_INVALID_ARGUMENT, "multi handle is NULL");
  if (cancel && *cancel) return mfail(m, RTB_ERR_CANCELLED, "cancelled");
  std::lock_guard<std::mutex> lock(m->mu);
  const int n = (int)m->ctx.size();
  int width, height, lo, hi;
  int rc = validate_params(m->ctx[0], params, &width, &height);
  if (rc != RTB_OK) return mforward(m, 0, rc);
  multi_collect_pending_times(m);
  if ((rc = multi_tiles(m, *params, width, height, &lo, &hi)) != RTB_OK) return rc;

  std::vector<std::unique_lock<std::mutex>> locks;
  for (rtb_ctx* c : m->ctx) locks.emplace_back(c->mu);
  std::vector<HostBatch> st((size_t)n);
  std::vector<rtb_ctx*> busy;
  int first_rc = RTB_OK;
  for (int g = 0; g < n; g++) {                      // every device starts before any is waited for
    m->kernel_ms[g] = 0.0f;
    if (m->bounds[g + 1] <= m->bounds[g]) continue;
    rtb_batch_params p = *params;
    p.row_begin = m->bounds[g];
    p.row_end = m->bounds[g + 1];
    rc = host_batch_begin(m->ctx[g], &p, host, &st[g]);
    if (rc != RTB_OK) { first_rc = mforward(m, g, rc); break; }
    if (!st[g].empty) busy.push_back(m->ctx[g]);
  }
  if (first_rc != RTB_OK)                            // stop what was started
    for (rtb_ctx* c : busy) send_cancel(c);
  if (!busy.empty()) {                               // the kernels first, with the token relayed while they run (see rtb_sample_batch)
    const cudaError_t e = wait_for_streams(busy.data(), (int)busy.size(), cancel);
    if (e != cudaSuccess && first_rc == RTB_OK) first_rc = mfail(m, RTB_ERR_CUDA + (int)e, "waiting for the devices: %s", cudaGetErrorString(e));
  }
  const bool cancelled = cancel && *cancel;
  for (int g = 0; g < n && first_rc == RTB_OK && !cancelled; g++)
    if (!st[g].empty && (rc = host_batch_drain(m->ctx[g], &st[g])) != RTB_OK) first_rc = mforward(m, g, rc);
  if (!busy.empty() && first_rc == RTB_OK && !cancelled) {
    const cudaError_t e = wait_for_streams(busy.data(), (int)busy.size(), nullptr);
    if (e != cudaSuccess) first_rc = mfail(m, RTB_ERR_CUDA + (int)e, "waiting for the devices: %s", cudaGetErrorString(e));
  }
  if (first_rc != RTB_OK) return first_rc;
  for (int g = 0; g < n; g++) {
    if (st[g].empty) continue;
    rc = host_batch_end(m->ctx[g], &st[g], cancel);
    m->kernel_ms[g] = m->ctx[g]->last_ms;
    if (rc != RTB_OK && first_rc == RTB_OK) first_rc = mforward(m, g, rc);
  }
  if (first_rc == RTB_OK) multi_feedback(m);
  return first_rc;
}

int rtb_multi_sample_batch_device(rtb_multi* m, const rtb_batch_params* params, const rtb_batch_buffers* dev, int owner_index,
                                  void* owner_stream) {
  if (!m) return mfail(nullptr, RTB_ERR_INVALID_ARGUMENT, "multi handle is NULL");
  std::lock_guard<std::mutex> lock(m->mu);
  const int n = (int)m->ctx.size();
  if (owner_index < 0 || owner_index >= n) return mfail(m, RTB_ERR_INVALID_ARGUMENT, "owner_index out of range");
  int width, height, lo, hi;
  int rc = validate_params(m->ctx[0], params, &width, &height);
  if (rc != RTB_OK) return mforward(m, 0, rc);
  rtb_ctx* owner = m->ctx[(size_t)owner_index];
  // the other devices reach the owner's memory over NVLink peer access
  for (int g = 0; g < n; g++) {
    if (g == owner_index || m->ctx[g]->device == owner->device || m->peer_enabled[(size_t)g * n + owner_index]) continue;
    DeviceGuard dg(m->ctx[g]->device);
    int can = 0;
    cudaDeviceCanAccessPeer(&can, m->ctx[g]->device, owner->device);
    if (!can) return mfail(m, RTB_ERR_UNSUPPORTED, "device %d cannot access device %d's memory (no peer path)", m->ctx[g]->device, owner->device);
    const cudaError_t e = cudaDeviceEnablePeerAccess(owner->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return mfail(m, RTB_ERR_CUDA + (int)e, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
    cudaGetLastError();
    m->peer_enabled[(size_t)g * n + owner_index] = 1;
  }
  multi_collect_pending_times(m);
  if ((rc = multi_tiles(m, *params, width, height, &lo, &hi)) != RTB_OK) return rc;
  cudaStream_t os = (cudaStream_t)owner_stream;
  {
    DeviceGuard dg(owner->device);
    if (!m->ev_fork) cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming);
    cudaEventRecord(m->ev_fork, os);                 // the inputs are ready when the owner's stream gets here
  }
  for (int g = 0; g < n; g++) {
    if (m->bounds[g + 1] <= m->bounds[g]) continue;
    rtb_ctx* c = m->ctx[g];
    DeviceGuard dg(c->device);
    cudaStream_t s = g == owner_index ? os : c->stream;
    if (g != owner_index) cudaStreamWaitEvent(s, m->ev_fork, 0);
    rtb_batch_params p = *params;
    p.row_begin = m->bounds[g];
    p.row_end = m->bounds[g + 1];
    cudaEventRecord(c->ev_start, s);
    rc = rtb_sample_batch_device(c, &p, dev, s);
    cudaEventRecord(c->ev_stop, s);
    if (rc != RTB_OK) return mforward(m, g, rc);
    if (g != owner_index) {
      cudaEventRecord(m->ev_join[(size_t)g], s);
      DeviceGuard og(owner->device);
      cudaStreamWaitEvent(os, m->ev_join[(size_t)g], 0);   // the owner's stream continues when every tile has landed
    }
  }
  m->times_pending = true;
  return RTB_OK;
}

int rtb_multi_get_tiles(rtb_multi* m, int* out_bounds, float* out_kernel_ms) {
  if (!m) return mfail(nullptr, RTB_ERR_INVALID_ARGUMENT, "multi handle is NULL");
  std::lock_guard<std::mutex> lock(m->mu);
  multi_collect_pending_times(m);
  const int n = (int)m->ctx.size();
  if (out_bounds) for (int g = 0; g <= n; g++) out_bounds[g] = m->bounds[(size_t)g];
  if (out_kernel_ms) for (int g = 0; g < n; g++) out_kernel_ms[g] = m->kernel_ms[(size_t)g];
  return RTB_OK;
}

// ---- device memory a peer process can map (one process per GPU: every rank writes its tile into rank 0's frame) ----

int rtb_device_alloc(rtb_ctx* ctx, size_t bytes, void** out_ptr) {
  if (!ctx || !out_ptr || !bytes) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_device_alloc: bad argument");
  DeviceGuard g(ctx->device);
  RTB_CUDA(ctx, cudaMalloc(out_ptr, bytes));
  RTB_CUDA(ctx, cudaMemset(*out_ptr, 0, bytes));
  return RTB_OK;
}

int rtb_device_free(rtb_ctx* ctx, void* ptr) {
  if (!ctx) return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "ctx is NULL");
  DeviceGuard g(ctx->device);
  if (ptr) RTB_CUDA(ctx, cudaFree(ptr));
  return RTB_OK;
}

int rtb_ipc_export(rtb_ctx* ctx, void* device_ptr, rtb_ipc_handle* out) {
  if (!ctx || !device_ptr || !out) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_ipc_export: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(rtb_ipc_handle), "handle size");
  DeviceGuard g(ctx->device);
  cudaIpcMemHandle_t h;
  RTB_CUDA(ctx, cudaIpcGetMemHandle(&h, device_ptr));
  memset(out, 0, sizeof *out);
  memcpy(out->bytes, &h, sizeof h);
  return RTB_OK;
}

int rtb_ipc_open(rtb_ctx* ctx, const rtb_ipc_handle* handle, void** out_ptr) {
  if (!ctx || !handle || !out_ptr) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_ipc_open: bad argument");
  DeviceGuard g(ctx->device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle->bytes, sizeof h);
  RTB_CUDA(ctx, cudaIpcOpenMemHandle(out_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return RTB_OK;
}

int rtb_ipc_close(rtb_ctx* ctx, void* ptr) {
  if (!ctx || !ptr) return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_ipc_close: bad argument");
  DeviceGuard g(ctx->device);
  RTB_CUDA(ctx, cudaIpcCloseMemHandle(ptr));
  return RTB_OK;
}

}  // extern "C"
