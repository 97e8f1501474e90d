// media.cuh — participating media (MaterialType.ProbabilisticVolume): the hit search of ONE bounce-loop iteration of
// SampleBatchJob.Sample (SampleBatchJob.cs:184-303) for worlds in which some entity wears such a material.
//
// The reference decides entering / leaving / being inside a medium from the SORTED LIST OF ALL HITS along the ray
// (FindHitCandidates + FindHits, :403-475), injects an extra exit hit for convex media (:462-469) and throws a backwards ray
// to learn whether the origin is inside one (DetermineVolumeContainment / AnyBackwardsVolumeEntryHit, :477-524).  Two
// builders produce the list the bookkeeping (media_step) runs on:
//
//   collect_hits   the reference's own shape — an unpruned walk, every entity of every leaf whose box chain is hit,
//                  insertion-sorted.  The validator kernel (sample_volumes, volume_kernel.cuh) uses it.
//   gather_hits    what the bookkeeping can OBSERVE of that list (the megakernel's media flavour):
//                    * every hit of every entity that wears a medium, and
//                    * the ONE nearest hit of an opaque entity (boxes pruned at the best such hit so far, unless they hold media).
//                  Opaque hits behind the nearest one are never read: the main loop stops at the first opaque record it
//                  reaches (scatter, or obstacle inside a medium), the exit scan stops at the first record of another material,
//                  and the containment test skips opaque records altogether — it only needs the first MEDIUM hit of the
//                  whole ray, which is why that walk is not cut at the opaque hit.  Ties in distance are ordered as the
//                  reference's candidate order + stable sort would (a later candidate goes first; see visited_later).
#pragma once

#include "kernel_common.cuh"

namespace rtbk {

constexpr int kMaxRayHits = 48;     // hit records kept per ray (the reference's list starts at 32 and grows, SampleBatchJob.cs:21;
                                    // HybridCollections.cs:65-71).  A ray that fills the list raises kStatusHitListOverflow and the
                                    // batch fails with RTB_ERR_UNSUPPORTED: never a silently different image

struct RayHits {                    // FindHits' sorted hitBuffer
  float t[kMaxRayHits];
  int slot[kMaxRayHits];
  int first[kMaxRayHits];           // first slot of the record's leaf (candidate order, see visited_later)
  f3 n[kMaxRayHits];
  int count;
};

// Entity.Hit (Entity.cs:57-72) of the entity in `slot` for t in (tmin, +inf): distance and world normal.
template <bool SMEM>
__device__ __noinline__ bool entity_hit(const SceneView<SMEM>& sv, int slot, f3 o, f3 d, float tmin, const RayClock& clk, float* t_out,
                                        f3* n_out) {
  const float4 prim = sv.sphere(slot);
  if (prim.w != prim.w) {
    if (__float_as_uint(prim.y) != 0u) {
      f3 n;
      if (!placed_test(sv, __float_as_uint(prim.x), o, d, clk, t_out, &n, tmin)) return false;
      *n_out = um::normalize(n);
      return true;
    }
    float u, v, t;
    if (!triangle_uvt(sv, __float_as_uint(prim.x), o, d, &u, &v, &t)) return false;
    if (t < tmin) return false;                       // HitTests.cs:141 (tMax = +inf)
    *t_out = t;
    *n_out = hit_normal<SMEM, kFlavorGeneral>(sv, prim, o, d, t, clk);
    return true;
  }
  // HitTests.Hit(this Sphere) (HitTests.cs:23-60) behind the identity-rotation transform
  const f3 oc = o + um::mk(-prim.x, -prim.y, -prim.z);
  const float a = um::dot(d, d), b = um::dot(oc, d), c = um::dot(oc, oc) - prim.w * prim.w;
  const float disc = um::fma(b, b, -(a * c));
  if (!(disc > 0.0f)) return false;
  const float sq = um::sqrt(disc);
  float t = um::div(-b - sq, a);
  if (!(t < um::INF && t > tmin)) {
    t = um::div(-b + sq, a);
    if (!(t < um::INF && t > tmin)) return false;
  }
  *t_out = t;
  *n_out = um::normalize(um::mad(d, t, oc) / prim.w);
  return true;
}

// In worlds with media, upload marks the material word of every entity that wears one (bit 31: plugin.cu, build_blob), so
// "is this entity's material a ProbabilisticVolume" costs no second, dependent load from the material table.
constexpr uint32_t kMediumBit = 0x80000000u;
template <bool SMEM>
__device__ __forceinline__ uint32_t slot_material(const SceneView<SMEM>& sv, int slot, bool* medium) {
  const uint32_t raw = sv.material_of(slot);
  *medium = (raw & kMediumBit) != 0u;
  return raw & ~kMediumBit;
}
// EntityType.IsConvexHull (Entity.cs:22-25): Sphere or Box
template <bool SMEM>
__device__ __forceinline__ bool is_convex_hull(const SceneView<SMEM>& sv, float4 prim) {
  if (prim.w == prim.w) return true;
  if (__float_as_uint(prim.y) == 0u) return false;    // triangle
  const uint32_t type = __float_as_uint(sv.placed(__float_as_uint(prim.x), 1).w) & 0xffu;
  return type == RTB_ENTITY_SPHERE || type == RTB_ENTITY_BOX;
}

// The reference's candidate order: FindHitCandidates pushes Left then Right and pops Right first, so leaves are visited from
// the LAST one of the depth-first order to the first; a leaf's entities in ascending order.  Device slots are laid out in
// depth-first order (plugin.cu: Flattener), so for two different entities: inside one leaf the higher slot comes later,
// across leaves the LOWER slot comes later.  FindHits then pops candidates from the end of that list and sorts by distance:
// among equal distances the later candidate ends up first.
__device__ __forceinline__ bool visited_later(int slot_a, int first_a, int slot_b, int first_b) {
  return first_a == first_b ? slot_a > slot_b : slot_a < slot_b;
}

// The list's insertion rule for records that arrive in candidate order: before the first record that is not nearer.
__device__ __forceinline__ void hits_insert(RayHits* hits, float t, int slot, int first, f3 n) {
  int pos = 0;
  while (pos < hits->count && hits->t[pos] < t) pos++;
  if (pos >= kMaxRayHits) return;
  const int last = hits->count < kMaxRayHits ? hits->count : kMaxRayHits - 1;
  for (int k = last; k > pos; k--) {
    hits->t[k] = hits->t[k - 1]; hits->slot[k] = hits->slot[k - 1]; hits->first[k] = hits->first[k - 1]; hits->n[k] = hits->n[k - 1];
  }
  hits->t[pos] = t; hits->slot[pos] = slot; hits->first[pos] = first; hits->n[pos] = n;
  if (hits->count < kMaxRayHits) hits->count++;
}

template <bool SMEM>
__device__ __noinline__ bool entity_records(const SceneView<SMEM>& sv, int slot, bool with_exit, f3 o, f3 d, const RayClock& clk,
                                            float* t_out, f3* n_out, float* t2_out, f3* n2_out);

// FindHitCandidates + FindHits (SampleBatchJob.cs:403-475) without pruning: every entity of every leaf whose box chain
// the ray hits, in the reference's visit order.  With a stable order among equal distances the reference's pop + sort is:
// a later candidate goes BEFORE an earlier one at the same distance — hits_insert.
// MODE 0: fill `hits`.  MODE 1 (AnyBackwardsVolumeEntryHit, :508-524): is there a volume entity the ray enters?
// MEDIA_ONLY: subtrees without media are skipped (node word 14: bit 0 = the left subtree holds one, bit 1 = the right) and only
// entities that wear a medium are tested — the order among those is the reference's.  MODE 1 never looks at anything else.
template <bool SMEM, bool COUNTERS, bool MEDIA_ONLY>
__device__ __noinline__ bool collect_hits(const int MODE, const SceneView<SMEM>& sv, const SceneDesc& sd, f3 o, f3 d, const RayClock& clk,
                                          RayHits* hits, WorkCounters& wc) {
  if (MODE == 0) hits->count = 0;
  if (!sd.has_root) return false;
  f3 inv = um::rcp(d);
  inv = um::mk(um::isnan(inv.x) ? um::INF : inv.x, um::isnan(inv.y) ? um::INF : inv.y, um::isnan(inv.z) ? um::INF : inv.z);
  float t_enter;
  if (COUNTERS) wc.node_tests++;
  if (!aabb_hit(v3(sd.root_min), v3(sd.root_max), o, inv, &t_enter)) return false;
  int stack[kStackMax + 2];
  int sp = 0;
  stack[sp++] = sv.root(sd);
  while (sp > 0) {
    const int cur = stack[--sp];
    if (cur >= 0) {
      const float4 q0 = sv.node(cur, 0), q1 = sv.node(cur, 1), q2 = sv.node(cur, 2), q3 = sv.node(cur, 3);
      float tl, tr;
      bool hl = aabb_hit(node_lmin(q0, q1, q2), node_lmax(q0, q1, q2), o, inv, &tl);
      bool hr = aabb_hit(node_rmin(q0, q1, q2), node_rmax(q0, q1, q2), o, inv, &tr);
      if (COUNTERS) wc.node_tests += 2;
      if (MEDIA_ONLY) {
        const uint32_t media = __float_as_uint(q3.z);
        hl = hl && (media & 1u);
        hr = hr && (media & 2u);
      }
      if (hl) stack[sp++] = __float_as_int(q3.x);
      if (hr) stack[sp++] = __float_as_int(q3.y);
      continue;
    }
    const uint32_t code = (uint32_t)~cur;
    const int first = (int)(code & ~15u);
    int count = (int)(code & 15u) + 1;
    if (count == 16) count = (int)sv.leaf_count(first);
    if (COUNTERS) wc.sphere_tests += count;
    for (int i = 0; i < count; i++) {
      const int slot = first + 16 * i;
      bool medium;
      slot_material(sv, slot, &medium);
      if ((MODE == 1 || MEDIA_ONLY) && !medium) continue;
      float t;
      f3 n;
      if (MEDIA_ONLY) {                               // (the media flavour's one copy of the entity tests: same arithmetic)
        float t2;
        f3 n2;
        if (!entity_records(sv, slot, false, o, d, clk, &t, &n, &t2, &n2)) continue;
      } else if (!entity_hit(sv, slot, o, d, 0.0f, clk, &t, &n)) {
        continue;
      }
      if (MODE == 1) {
        if (um::dot(n, d) > 0) return true;           // (the caller passes the backwards ray)
        continue;
      }
      // Inject exit hits for probabilistic convex hulls (:462-469); the pair is pushed entry first, so at equal
      // distances the exit must end up behind the entry: insert it first
      if (medium && is_convex_hull(sv, sv.sphere(slot))) {
        float t2;
        f3 n2;
        if (entity_hit(sv, slot, o, d, t + 0.001f, clk, &t2, &n2)) hits_insert(hits, t2, slot, first, n2);
      }
      hits_insert(hits, t, slot, first, n);
    }
  }
  return false;
}

// Entity.Hit (Entity.cs:57-72) of the entity in `slot` for t in (0, +inf) and — with_exit, for convex media — the hit FindHits
// injects behind it (:462-469: Entity.Hit again from distance + 0.001), both from ONE evaluation of the entity's transform and
// quadratic: the second call of the reference recomputes the same numbers, finds the first root again (not beyond distance +
// 0.001) and then the second, so "entry was the first root and the second lies beyond distance + 0.001" is its outcome.
// *t2 < 0: no exit record.
template <bool SMEM>
__device__ __noinline__ bool entity_records(const SceneView<SMEM>& sv, int slot, bool with_exit, f3 o, f3 d, const RayClock& clk,
                                            float* t_out, f3* n_out, float* t2_out, f3* n2_out) {
  const float4 prim = sv.sphere(slot);
  *t2_out = -1.0f;
  if (prim.w != prim.w) {
    if (__float_as_uint(prim.y) != 0u) {
      f3 n, n2 = um::mk(0.0f);
      float t2;
      if (!placed_test_core<SMEM, true>(sv, __float_as_uint(prim.x), o, d, clk, t_out, &n, 0.0f, &t2, &n2)) return false;
      *n_out = um::normalize(n);
      if (with_exit && t2 >= 0.0f) { *t2_out = t2; *n2_out = um::normalize(n2); }
      return true;
    }
    float u, v, t;
    if (!triangle_uvt(sv, __float_as_uint(prim.x), o, d, &u, &v, &t)) return false;
    *t_out = t;                                       // (t >= 0 == tMin: HitTests.cs:141)
    *n_out = hit_normal<SMEM, kFlavorGeneral>(sv, prim, o, d, t, clk);
    return true;
  }
  const f3 oc = o + um::mk(-prim.x, -prim.y, -prim.z);
  const float a = um::dot(d, d), b = um::dot(oc, d), c = um::dot(oc, oc) - prim.w * prim.w;
  const float disc = um::fma(b, b, -(a * c));
  if (!(disc > 0.0f)) return false;
  const float sq = um::sqrt(disc);
  float t = um::div(-b - sq, a);
  bool first_root = true;
  if (!(t < um::INF && t > 0.0f)) {
    t = um::div(-b + sq, a);
    first_root = false;
    if (!(t < um::INF && t > 0.0f)) return false;
  }
  *t_out = t;
  *n_out = um::normalize(um::mad(d, t, oc) / prim.w);
  if (with_exit && first_root) {
    const float t2 = um::div(-b + sq, a);
    if (t2 < um::INF && t2 > t + 0.001f) {
      *t2_out = t2;
      *n2_out = um::normalize(um::mad(d, t2, oc) / prim.w);
    }
  }
  return true;
}

// Insertion by the list's ORDER instead of by arrival: nearer first; at equal distances the later candidate of the
// reference's walk first (visited_later), an entity's entry before its own injected exit.
__device__ __forceinline__ void hits_insert_ranked(RayHits* hits, float t, int slot, int first, f3 n, bool is_exit) {
  int pos = 0;
  while (pos < hits->count &&
         (hits->t[pos] < t ||
          (hits->t[pos] == t && (hits->slot[pos] == slot ? is_exit : visited_later(hits->slot[pos], hits->first[pos], slot, first))))) pos++;
  if (pos >= kMaxRayHits) return;
  const int last = hits->count < kMaxRayHits ? hits->count : kMaxRayHits - 1;
  for (int k = last; k > pos; k--) {
    hits->t[k] = hits->t[k - 1]; hits->slot[k] = hits->slot[k - 1]; hits->first[k] = hits->first[k - 1]; hits->n[k] = hits->n[k - 1];
  }
  hits->t[pos] = t; hits->slot[pos] = slot; hits->first[pos] = first; hits->n[pos] = n;
  if (hits->count < kMaxRayHits) hits->count++;
}

// The list the bookkeeping can observe (see the top of the file), from ONE walk in rounds: [walk until kCandBatch candidate
// entities are known] -> [intersect them, every lane of the warp at its k-th candidate together] -> [walk on, with the boxes
// now pruned at the nearest OPAQUE hit so far] ...  The walk descends into a box that is hit and either starts before that
// limit or holds an entity that wears a medium (node word 14) — media records are kept whatever their distance (the
// containment test reads the first one of the whole ray) —, nearer child first.  Intersecting inside the walk instead
// (closest_hit's way) ran the entity tests — transforms, IEEE divisions, a three-way type switch — with 2 of 32 lanes: each
// lane reaches its leaves at other trips (ncu: 45 % of the kernel's warp instructions at 7 % lane use).
constexpr int kCandBatch = 8;
template <bool SMEM, bool COUNTERS>
__device__ __forceinline__ void gather_hits(const SceneView<SMEM>& sv, const SceneDesc& sd, f3 o, f3 d, const RayClock& clk,
                                            RayHits* hits, WorkCounters& wc) {
  hits->count = 0;
  if (!sd.has_root) return;
  f3 inv = um::rcp(d);
  inv = um::mk(um::isnan(inv.x) ? um::INF : inv.x, um::isnan(inv.y) ? um::INF : inv.y, um::isnan(inv.z) ? um::INF : inv.z);
  float t_enter;
  if (COUNTERS) wc.node_tests++;
  if (!aabb_hit(v3(sd.root_min), v3(sd.root_max), o, inv, &t_enter)) return;
  float best_t = um::INF;
  int best_slot = -1, best_first = 0;
  f3 best_n = um::mk(0.0f);
  int stack[kStackMax + 2];
  int cand[kCandBatch], cand_first[kCandBatch];
  int sp = 0;
  int cur = sv.root(sd);
  bool have_cur = true;
  int leaf_first = 0, leaf_i = 0, leaf_n = 0;       // the leaf being unpacked into candidates
  const RayPairs rp{pack2(o.x, o.y), pack2(o.z, o.z), pack2(inv.x, inv.y), pack2(inv.z, inv.z)};
  for (;;) {
    int nc = 0;
    while (nc < kCandBatch) {
      if (leaf_i < leaf_n) {
        cand[nc] = leaf_first + 16 * leaf_i;
        cand_first[nc] = leaf_first;
        nc++;
        leaf_i++;
        continue;
      }
      if (!have_cur) {
        if (sp == 0) break;
        cur = stack[--sp];
      }
      have_cur = false;
      if (cur >= 0) {
        const float4 q0 = sv.node(cur, 0), q1 = sv.node(cur, 1), q2 = sv.node(cur, 2), q3 = sv.node(cur, 3);
        float tl, tr, xl, xr;
        aabb_range_pair(q0, q1, q2, rp, &tl, &xl, &tr, &xr);
        const float limit = best_t * kPruneMargin;
        const uint32_t media = __float_as_uint(q3.z);
        const bool hl = tl < xl && (tl < limit || (media & 1u));
        const bool hr = tr < xr && (tr < limit || (media & 2u));
        if (COUNTERS) wc.node_tests += 2;
        const int left = __float_as_int(q3.x), right = __float_as_int(q3.y);
        if (hl && hr) {
          const bool left_first = tl <= tr;
          stack[sp++] = left_first ? right : left;
          cur = left_first ? left : right;
          have_cur = true;
        } else if (hl || hr) {
          cur = hl ? left : right;
          have_cur = true;
        }
      } else {
        const uint32_t code = (uint32_t)~cur;
        leaf_first = (int)(code & ~15u);
        leaf_n = (int)(code & 15u) + 1;
        if (leaf_n == 16) leaf_n = (int)sv.leaf_count(leaf_first);
        leaf_i = 0;
        if (COUNTERS) wc.sphere_tests += leaf_n;
      }
    }
    if (nc == 0) break;
    for (int k = 0; k < nc; k++) {
      const int slot = cand[k], first = cand_first[k];
      bool medium;
      slot_material(sv, slot, &medium);
      float t, t2;
      f3 n, n2;
      if (!entity_records(sv, slot, medium, o, d, clk, &t, &n, &t2, &n2)) continue;
      if (medium) {
        if (t2 >= 0.0f) hits_insert_ranked(hits, t2, slot, first, n2, true);
        hits_insert_ranked(hits, t, slot, first, n, false);
      } else if (t < best_t || (t == best_t && best_slot >= 0 && visited_later(slot, first, best_slot, best_first))) {
        best_t = t; best_slot = slot; best_first = first; best_n = n;
      }
    }
  }
  if (best_slot >= 0) hits_insert_ranked(hits, best_t, best_slot, best_first, best_n, false);
}

// What one bounce-loop iteration found (SampleBatchJob.cs:184-303): the record to scatter on, or nothing (the sky ends the path).
struct MediaStep {
  bool hit;
  bool medium_hit;               // the record was made INSIDE a medium: HitRecord(distance, point, -direction), TexCoords 0
  float t;
  f3 n;
  int slot;                      // the record's entity (not meaningful for medium_hit)
  uint32_t material;
  float events;                  // rng.RandomEvents after this iteration's Material.ProbabilisticHit draws
  int current_volume;            // currentProbabilisticVolumeMaterial after the iteration (material index, -1 = null)
};

// The iteration's hit search and volume bookkeeping.  `current_volume` is the path's state coming in.  PRUNED selects the list
// builder (see the top of the file); the bookkeeping below is the same code for both, and is the reference's, statement by
// statement.  ProbabilisticHit's k-th draw of the iteration is Philox block 2 + k / 4, word k % 4 (or the next float of the
// white-noise stream).
template <bool SMEM, bool COUNTERS, bool PRUNED, bool WHITE>
__device__ __forceinline__ MediaStep media_step(const SceneView<SMEM>& sv, const SceneDesc& sd, f3 ro, f3 rd, const RayClock& clk,
                                               int current_volume, uint32_t index, uint32_t s, uint32_t depth, uint32_t seed,
                                               WhiteNoise& white, WorkCounters& wc, RayHits& hits) {
  MediaStep out;
  out.hit = false;
  out.medium_hit = false;
  out.t = 0;
  out.n = um::mk(0.0f);
  out.slot = 0;
  out.material = 0;
  uint32_t volume_draws = 0;
  float events = 0;                               // rng.RandomEvents of this iteration
  if (PRUNED) {
    gather_hits<SMEM, COUNTERS>(sv, sd, ro, rd, clk, &hits, wc);
  } else {
    collect_hits<SMEM, COUNTERS, false>(0, sv, sd, ro, rd, clk, &hits, wc);
  }
  // a FULL list may have lost records (checked here, once per ray: anything in the insert path itself — an atomic, even a
  // flag store — cost the validator kernel 20-100 %): the batch then fails with RTB_ERR_UNSUPPORTED
  if (hits.count >= kMaxRayHits && sd.status) atomicOr(sd.status, kStatusHitListOverflow);
  if (current_volume < 0) {                       // DetermineVolumeContainment (:477-506)
    for (int i = 0; i < hits.count; i++) {
      bool medium;
      const uint32_t hm = slot_material(sv, hits.slot[i], &medium);
      if (!medium) continue;
      if (um::dot(hits.n[i], rd) < 0) break;      // entry hit: not inside
      // (the backwards ray does not depend on i, and neither does its answer)
      if (collect_hits<SMEM, COUNTERS, PRUNED>(1, sv, sd, ro, -rd, clk, &hits, wc)) current_volume = (int)hm;
      break;
    }
  }

  int hit_index = 0;
  while (hit_index < hits.count) {
    float rec_t = hits.t[hit_index];
    f3 rec_n = hits.n[hit_index];
    int rec_slot = hits.slot[hit_index];
    bool rec_medium;
    uint32_t mi = slot_material(sv, rec_slot, &rec_medium);
    bool medium_hit = false;

    if (current_volume >= 0 || rec_medium) {
      const bool is_entry_hit = current_volume < 0;
      if (current_volume < 0) current_volume = (int)mi;
      int exit_index = hit_index, last_exit = -1, same_entries = 0;
      while (exit_index < hits.count) {
        if ((int)(sv.material_of(hits.slot[exit_index]) & ~kMediumBit) == current_volume) {
          if (um::dot(hits.n[exit_index], rd) < 0) {
            same_entries++;
          } else {
            same_entries--;
            last_exit = exit_index;
          }
          if (same_entries <= 0) break;
        } else {
          break;
        }
        exit_index++;
      }
      if (same_entries > 0 && last_exit != -1) exit_index = last_exit;

      if (exit_index < hits.count) {
        float distance_in_volume = hits.t[exit_index];
        float entry_distance = 0;
        if (is_entry_hit) {
          entry_distance = rec_t;
          distance_in_volume -= rec_t;
        }
        // Material.ProbabilisticHit (Material.cs:48-65); Density = the material's parameter
        const float density = __ldg(reinterpret_cast<const float*>(sd.materials + current_volume) + 9);
        events += 1.0f;
        float u;
        if (WHITE) {
          u = white.next_float();
        } else {
          const uint4 r = philox4x32_10(index, s, depth, 2u + (volume_draws >> 2), seed, kPhiloxKey1);
          const uint32_t w = volume_draws & 3u;
          u = u2f(w == 0 ? r.x : w == 1 ? r.y : w == 2 ? r.z : r.w);
          volume_draws++;
        }
        const float volume_hit_distance = -um::div(1.0f, um::max(density, 1.1920928955078125e-7f)) * um::log_unit(u);
        if (volume_hit_distance < distance_in_volume) {
          // we hit inside the volume: the record becomes (distance, point, -direction), the material the medium's
          rec_t = entry_distance + volume_hit_distance;
          rec_n = -rd;
          mi = (uint32_t)current_volume;
          medium_hit = true;
        } else {
          bool exit_medium;
          const uint32_t exit_material = slot_material(sv, hits.slot[exit_index], &exit_medium);
          current_volume = -1;
          if (exit_medium && um::dot(hits.n[exit_index], rd) > 0) {
            hit_index = exit_index + 1;           // volume exit: move to the next hit
            continue;
          }
          rec_t = hits.t[exit_index];             // obstacle: scatter on the exit hit
          rec_n = hits.n[exit_index];
          rec_slot = hits.slot[exit_index];
          mi = exit_material;
        }
      } else {
        break;                                    // no more surfaces to hit (the volume has holes): every record is dropped
      }
    }
    out.hit = true;
    out.medium_hit = medium_hit;
    out.t = rec_t;
    out.n = rec_n;
    out.slot = rec_slot;
    out.material = mi;
    break;
  }
  out.events = events;
  out.current_volume = current_volume;
  return out;
}

}  // namespace rtbk
